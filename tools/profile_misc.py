"""Small drivers for the ncu captures kept under profiles/ (one mode per run):
    rollout_c2   hk_rollout, 1 Mi games (20,3), 20 steps in one launch (thread-per-game kernel, T = 20)
    rollout_c5   hk_rollout, 256 Ki games (64,5), 20 steps in one launch (warp-per-game kernel)
    obs_c2       one random-play rollout of 1 Mi games (20,3) with the fused host observation, one launch per step
    seeded_c2    hk_rollout_seeded: the same 20 steps with both players drawn in the kernel (no action streams)
Prints the device time(s) it measured with CUDA events."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hironaka_b200 import constants as C  # noqa: E402
from hironaka_b200._lib import lib  # noqa: E402

mode = sys.argv[1]
L = lib()
dev = torch.device("cuda")
stream = torch.cuda.current_stream().cuda_stream
B, N, d, T = ((1 << 18, 64, 5, 20) if mode == "rollout_c5" else (1 << 20, 20, 3, 20))
rng = np.random.default_rng(9)
x0 = torch.from_numpy(rng.integers(0, 20, size=(B, N, d), dtype=np.int32)).to(dev)
ha = torch.from_numpy(rng.integers(0, 2 ** d - d - 1, size=(T, B), dtype=np.int32)).to(dev)
ax = torch.from_numpy(rng.integers(0, d, size=(T, B), dtype=np.int32)).to(dev)
OPS = C.HK_OP_SHIFT | C.HK_OP_REPOSITION | C.HK_OP_NEWTON
assert L.hk_step(x0.data_ptr(), x0.data_ptr(), None, None, None, None, None, None, None, None, B, N, d, C.HK_DTYPE_I32,
                 C.HK_OP_NEWTON | C.HK_OP_REPOSITION, 0, -1.0, 1e8, stream) == 0
x = x0.clone()
dcount = torch.zeros(T, dtype=torch.int32, device=dev)


def timed(fn, reps=3):
    out = []
    for _ in range(reps):
        x.copy_(x0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return out


if mode in ("rollout_c2", "rollout_c5"):
    def run():
        assert L.hk_rollout(x.data_ptr(), x.data_ptr(), ha.data_ptr(), ax.data_ptr(), None, None, dcount.data_ptr(), None, B, N,
                            d, T, C.HK_DTYPE_I32, OPS, C.HK_F_ACT_DISCRETE, -1.0, stream) == 0
    print(json.dumps({mode + "_ms_per_rollout": timed(run)}))
elif mode == "seeded_c2":
    def run():
        assert L.hk_rollout_seeded(x.data_ptr(), x.data_ptr(), None, None, None, None, dcount.data_ptr(), None, B, N, d, T,
                                   C.HK_DTYPE_I32, OPS, C.HK_F_HOST_RANDOM | C.HK_F_AGENT_RANDOM, -1.0, 12345, 0, stream) == 0
    print(json.dumps({mode + "_ms_per_rollout": timed(run)}))
elif mode == "obs_c2":
    obs = torch.empty((B, N * d), dtype=torch.float32, device=dev)
    done = torch.empty(B, dtype=torch.uint8, device=dev)
    rew = torch.empty(B, dtype=torch.float32, device=dev)
    fl = C.HK_F_ACT_DISCRETE | C.HK_F_OBS_SORT_LEX | C.HK_F_OBS_RESCALE

    def run():
        for t in range(T):
            assert L.hk_step(x.data_ptr(), x.data_ptr(), ha[t].data_ptr(), ax[t].data_ptr(), done.data_ptr(), rew.data_ptr(),
                             None, obs.data_ptr(), None, None, B, N, d, C.HK_DTYPE_I32, OPS, fl, -1.0, 1e8, stream) == 0
    print(json.dumps({mode + "_ms_per_20_steps": timed(run)}))
elif mode == "obs_census_c2":
    obs = torch.empty((B, N * d), dtype=torch.float32, device=dev)
    done = torch.empty(B, dtype=torch.uint8, device=dev)
    rew = torch.empty(B, dtype=torch.float32, device=dev)
    census = torch.zeros(L.hk_census_bytes(B, N, d), dtype=torch.uint8, device=dev)
    fl = C.HK_F_ACT_DISCRETE | C.HK_F_OBS_SORT_LEX | C.HK_F_OBS_RESCALE
    per = []

    def run():
        census.zero_()
        assert L.hk_step_census(x.data_ptr(), None, None, None, None, None, None, census.data_ptr(), None, None, B, N, d,
                                C.HK_DTYPE_I32, C.HK_OP_NEWTON | C.HK_OP_REPOSITION, 0, -1.0, 1e8, stream) == 0
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(T + 1)]
        ev[0].record()
        for t in range(T):
            assert L.hk_step_census_obs(x.data_ptr(), ha[t].data_ptr(), ax[t].data_ptr(), done.data_ptr(), None, rew.data_ptr(),
                                        None, obs.data_ptr(), None, census.data_ptr(), None, None, B, N, d, C.HK_DTYPE_I32, OPS,
                                        fl, -1.0, 1e8, stream) == 0
            ev[t + 1].record()
        torch.cuda.synchronize()
        per.append([round(ev[t].elapsed_time(ev[t + 1]), 4) for t in range(T)])
    tt = timed(run)
    print(json.dumps({mode + "_ms_per_20_steps(+root)": tt, "by_step": per[-1], "mean_ms": float(np.mean(per[-1]))}))
else:
    raise SystemExit("unknown mode")
