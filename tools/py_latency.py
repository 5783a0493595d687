"""Per-call latency of the Python-level step APIs at MCTS batch sizes (tuning aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hironaka_b200 import functional as F, TensorPoints, constants as C, ops

def bench(fn, n=500, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6

rng = np.random.default_rng(0)
for B in (10, 100, 512):
    N, d = 20, 3
    x = torch.from_numpy(rng.integers(0, 20, (B, N, d)).astype(np.float32)).cuda()
    ops.step(x, ops=C.HK_OP_NEWTON | C.HK_OP_REPOSITION, inplace=True)
    ha = torch.from_numpy(rng.integers(0, 4, B).astype(np.int32)).cuda()
    ax = torch.from_numpy(rng.integers(0, 3, B).astype(np.int32)).cuda()
    env_step = F.get_env_step("host", (N, d))
    ta = F.get_take_actions("host", (N, d))
    coords = F.get_batch_decode(d)(ha).float()
    obs = x.reshape(B, -1)
    print(f"B={B}: get_env_step {bench(lambda: env_step(x, ha, ax)):.1f} us | take_actions {bench(lambda: ta(obs, coords, ax)):.1f} us | "
          f"ops.step inplace {bench(lambda: ops.step(x, ha, ax, ops=7, flags=C.HK_F_ACT_DISCRETE)):.1f} us", end="")
    if hasattr(F, "GraphedEnvStep"):
        g = F.GraphedEnvStep("host", (N, d), B)
        g.points.copy_(x.to(torch.int32) if g.points.dtype == torch.int32 else x)
        def call():
            g.host_action.copy_(ha); g.axis.copy_(ax); g()
        print(f" | GraphedEnvStep (2 copies + replay) {bench(call):.1f} us | replay only {bench(g):.1f} us", end="")
    print()
