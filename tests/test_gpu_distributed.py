"""GPU, world size N > 1: the one collective of the path (engine.gather_rollout, an NCCL all-gather of rollout
buffers: the leading device axis pmap returns in JAXTrainer.simulate, hironaka/jax/jax_trainer.py:316-320) under
torchrun, one rank per GPU.  Skipped on boxes with a single GPU (the gloo twin runs on CPU in
tests/test_distributed_cpu.py)."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gather_rollout_nccl():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tests", "dist_gather_worker.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert f"gather ok {world}" in out.stdout
