"""Multi-GPU test (needs >= 2 GPUs on the node; skipped on a single-GPU box): the fused gradient exchange
(bbb_adam_step_peer through PeerShardedAdam: reduce-scatter + Adam + all-gather over NVLink peer memory) against the
NCCL all-reduce + FusedAdam path, one process per GPU via torchrun.  The single-GPU suite covers the same kernel with
two ranks on one device (tests/test_gpu_step.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_peer_sharded_adam_matches_allreduce_then_adam():
    n = min(torch.cuda.device_count(), 4)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}',
                          '--master-addr', '127.0.0.1', '--master-port', '29577',
                          os.path.join(ROOT, 'tools', 'check_peer_adam.py')],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count('identical across ranks: True') == n
