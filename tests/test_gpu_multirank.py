"""Data-parallel parity ON HARDWARE: N NCCL ranks, global batch sharded over
the batch axis, ONE parrot_comm_allreduce per optimizer step, against one process on the whole batch.  The reference
cost is a masked mean over the GLOBAL batch (/root/reference/model.py:784): parameters after two optimizer steps must
agree to 1e-5, the global cost to 1e-6 relative.  Needs >= 2 GPUs (run with ``gpurun --gpus 2`` / 8); skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _world_sizes():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return sorted({w for w in (2, n) if 2 <= w <= n and 16 % w == 0})


@pytest.mark.parametrize('world', [2, 4, 8])
def test_nccl_ranks_match_single_process(world, tmp_path):
    if world not in _world_sizes():
        pytest.skip('needs %d GPUs on this box' % world)
    sys.path.insert(0, ROOT)
    from tests import dp_worker
    out = str(tmp_path / ('dp%d.npz' % world))
    env = dict(os.environ)
    env.setdefault('NCCL_DEBUG', 'WARN')
    port = 29620 + world
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'dp_worker.py'), out]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    dp = np.load(out)
    assert int(dp['info'][0]) == world            # the C-ABI communicator really spans `world` NCCL ranks
    # single process, whole batch, same seeds
    torch.cuda.set_device(0)
    model = dp_worker.make_model(torch.device('cuda', 0))
    costs, params, stats = dp_worker.run(model, False, 1, 0)
    assert np.allclose(dp['costs'], costs, rtol=1e-6, atol=0), (dp['costs'], costs)
    diff = np.abs(dp['params'] - params).max()
    print('world %d: max |param difference| after %d steps %.2e, costs %s' % (world, dp_worker.STEPS, diff, costs))
    assert diff < 1e-5, diff
    assert abs(dp['stats'][0] - stats[0]) <= 1e-4 * abs(stats[0])     # global gradient norm seen by the clip
