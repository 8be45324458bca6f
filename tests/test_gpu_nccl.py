"""Multi-GPU parity through the product's own data plane: torchrun-launched NCCL ranks,
one per GPU, run mcg_comm_* / mcg_query_batch_sharded / mcg_db_remove_overpopulated_dist and rank 0 compares with
the oracle at the same number of shards (tests/nccl_worker.py).  Needs as many GPUs as ranks (NCCL refuses two
ranks on one device); skipped -- loudly -- on smaller boxes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_pipeline_nccl(world):
    import torch
    have = torch.cuda.device_count()
    if have < world:
        pytest.skip("needs %d GPUs, this box has %d (run: gpurun --gpus %d -- python -m pytest tests/test_gpu_nccl.py -m gpu)"
                    % (world, have, world))
    port = 29600 + world + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    out = r.stdout.decode()
    print(out[-4000:])
    assert r.returncode == 0 and ("NCCL_WORKER_OK world=%d" % world) in out, out[-6000:]
