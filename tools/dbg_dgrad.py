"""Times ngb_conv_dgrad on the LeNet-B conv2 shape (events + CUPTI kernel durations); used for phase ablations of the
row-decomposed kernel (the phase times were additive, i.e. the kernel is issue-bound)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neograd_b200 import _backend as B
B.ensure_runtime()
P = lambda t: ctypes.c_void_p(t.data_ptr())
N, C, H, W, O, KH, KW = 4096, 6, 12, 12, 16, 5, 5
ug = torch.randn(N, O, 8, 8, device='cuda'); w = torch.randn(O, C, KH, KW, device='cuda'); dx = torch.empty(N, C, H, W, device='cuda')
st = ctypes.c_void_p(B.stream())
def run(): B.check(B.lib.ngb_conv_dgrad(P(ug), P(w), P(dx), N, C, H, W, O, KH, KW, 0, 1, 0, st))
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(10):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print(os.environ.get('NGB_DBG', '0'), 'us:', round(min(ts), 1), round(sum(ts) / len(ts), 1))
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  for _ in range(5): run()
  torch.cuda.synchronize()
d = [e.time_range.end - e.time_range.start for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and 'conv_dgrad' in e.name]
print('  kernel us (CUPTI):', [round(x, 1) for x in d])
