"""Data-parallel path ON GPUS (skips below 2 devices): launches tools/dp_check.py under torchrun on 2 GPUs of this node.
  * the gradient a DP step applies == the full-batch gradient (fp32 mode, <= 2e-6);
  * replicas are bit-identical after 3 Adam steps, and after 200 graph-replayed steps;
  * a rank that stalls longer than the exchange's timeout makes EVERY rank raise (no update is applied past a failed
    exchange, nothing hangs) -- the fused exchange + update kernel of csrc/xgpu.cu.
The round's recorded runs (2 and 4 GPUs) are under profiles/r02_dp_check_*.txt."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _run(extra, env_extra, timeout):
  import torch
  if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs')
  env = dict(os.environ, **env_extra)
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
         '--master-port', str(_free_port()), os.path.join(ROOT, 'tools', 'dp_check.py')] + extra
  r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
  recs = [json.loads(l) for l in r.stdout.splitlines() if l.startswith('{')]
  assert r.returncode == 0 and recs, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
  return {rec['check']: rec for rec in recs}


@pytest.mark.parametrize('p2p', [True, False])
def test_dp_equivalence_two_gpus(p2p):
  recs = _run([], {} if p2p else {'NGB_NO_P2P': '1'}, 300)
  g = recs['dp_gradient_equals_full_batch']
  assert g['mode'] == ('p2p' if p2p else 'nccl') and g['ok'] and g['max_dist'] < 2e-6, g
  assert recs['replicas_bit_identical_after_3_adam_steps']['ok'], recs
  c = recs['captured_step']
  assert c['replicas_identical'] and c['xgpu_check'] == 0, c


def test_dp_stalled_rank_raises_everywhere():
  recs = _run(['--timeout-test'], {'NGB_XGPU_TIMEOUT_S': '4'}, 300)
  t = recs['timeout']
  assert t['ok'] and all('timed out' in r['raised'] for r in t['results']), t
