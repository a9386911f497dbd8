"""Kernel timeline of ONE graph-replayed training step (CUPTI via torch.profiler): start offset, duration and stream of
every kernel, so overlap between the backward lanes and the real (warm-cache, un-serialised) durations are visible.
  python tools/timeline.py [--workload lenet_b] [--batch 4096] [--lanes 2] > gpurun_out/timeline.txt"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--workload', default='lenet_b')
  ap.add_argument('--batch', type=int, default=0)
  ap.add_argument('--lanes', type=int, default=3)
  ap.add_argument('--flush', type=int, default=1)
  args = ap.parse_args()
  import torch
  from torch.profiler import ProfilerActivity, profile
  import neograd_b200 as ng
  from neograd_b200.capture import CapturedStep
  import bench
  ng.set_backward_lanes(args.lanes)
  wl = bench.make_workload(ng, args.workload, args.batch or bench.BATCH[args.workload], 0)
  wl.setup_optim()
  step = CapturedStep(wl.make_step([ng.tensor(a) for a in wl.host_inputs]), warmup=3)
  flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)
  for _ in range(3):
    step()
  torch.cuda.synchronize()
  with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
      if args.flush:
        flush.zero_()
      step()
    torch.cuda.synchronize()
  evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
  evs.sort(key=lambda e: e.time_range.start)
  # last replay = everything after the last flush kernel (or the last third)
  idx = [i for i, e in enumerate(evs) if 'FillFunctor' in e.name]
  start = idx[-1] + 1 if idx else 2 * len(evs) // 3
  ks = evs[start:]
  t0 = ks[0].time_range.start
  end = max(e.time_range.end for e in ks)
  print(f'# {len(ks)} kernels, span {end - t0:.1f} us, sum of durations {sum(e.time_range.end - e.time_range.start for e in ks):.1f} us')
  print('# start_us  dur_us  stream  kernel')
  for e in ks:
    name = e.name.replace('void ', '').replace('ngb::', '')
    stream = getattr(e, 'stream', None)
    if stream is None:
      stream = getattr(e, 'device_resource_id', '?')
    print(f'{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:7.1f} {str(stream):>6}  {name[:100]}')


if __name__ == '__main__':
  main()
