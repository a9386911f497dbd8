"""Host-side profile of the eagerly dispatched LeNet-B step (no graph capture): where the Python time per step goes."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neograd_b200 as ng
import bench

wl = bench.make_workload(ng, sys.argv[1] if len(sys.argv) > 1 else 'lenet_b', 4096, 0)
wl.setup_optim()
step = wl.make_step([ng.tensor(a) for a in wl.host_inputs])
for _ in range(20):
  step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
  step()
t_issue = (time.perf_counter() - t0) / 200
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 200
print(f'host issue time per step {t_issue * 1e3:.3f} ms, wall per step {t_all * 1e3:.3f} ms')
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
  step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(45)
print(s.getvalue()[:9000])
