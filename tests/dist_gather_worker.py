"""Worker of tests/test_gpu_distributed.py: one rank per GPU under torchrun (NCCL).  Each rank plays a short
rollout of its own shard with the fused observation, all ranks all-gather the rollout buffers
(engine.gather_rollout) and every rank checks every slice against what that rank must have produced
(the shards are regenerated from their seeds)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hironaka_b200 import constants as C, ops  # noqa: E402
from hironaka_b200.engine import gather_rollout, shard_range  # noqa: E402


def play(rank, world, total, N, d, T):
    lo, hi = shard_range(total, rank, world)
    rng = np.random.default_rng(1000 + rank)
    x = torch.from_numpy(rng.integers(0, 15, size=(hi - lo, N, d), dtype=np.int32)).cuda()
    ha = torch.from_numpy(rng.integers(0, 2 ** d - d - 1, size=(T, hi - lo), dtype=np.int32)).cuda()
    ax = torch.from_numpy(rng.integers(0, d, size=(T, hi - lo), dtype=np.int32)).cuda()
    obs = []
    for t in range(T):
        r = ops.step(x, ha[t], ax[t], ops=C.HK_OP_SHIFT | C.HK_OP_REPOSITION | C.HK_OP_NEWTON,
                     flags=C.HK_F_ACT_DISCRETE | C.HK_F_OBS_SORT_LEX | C.HK_F_OBS_RESCALE, inplace=True, want_obs=True)
        obs.append(r.obs)
    return torch.stack(obs, dim=1).reshape(-1, N * d)  # [games * T, N * d]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    total, N, d, T = 4096 * world, 20, 3, 6
    mine = play(rank, world, total, N, d, T)
    value = torch.full((mine.shape[0],), float(rank), device="cuda")
    g_obs, g_val = gather_rollout((mine, value))
    assert g_obs.shape == (world, mine.shape[0], N * d) and g_val.shape == (world, mine.shape[0])
    for r in range(world):
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        want = play(r, world, total, N, d, T)
        assert torch.equal(g_obs[r], want), f"rank {rank}: slice {r} differs"
        assert bool((g_val[r] == float(r)).all())
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("gather ok", world)


if __name__ == "__main__":
    main()
