"""C3 (MCTS node expansion): latency of one fused step + features per call at small batch sizes, thread-per-game
against warp-per-game (hk_debug_force_generic), stream launch and CUDA-graph replay."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hironaka_b200 import constants as C
from hironaka_b200._lib import lib
L = lib(); dev = torch.device("cuda"); stream = torch.cuda.current_stream().cuda_stream
out = {}
for fam in (0, 1):
    L.hk_debug_force_generic(fam)
    for B, N in ((10, 20), (100, 20), (512, 20), (2048, 20), (10, 5), (512, 5)):
        d, T = 3, 20
        rng = np.random.default_rng(3)
        x = torch.from_numpy(rng.integers(0, 20, size=(B, N, d), dtype=np.int32)).to(dev)
        ha = torch.from_numpy(rng.integers(0, 4, size=(T, B), dtype=np.int32)).to(dev)
        ax = torch.from_numpy(rng.integers(0, d, size=(T, B), dtype=np.int32)).to(dev)
        done = torch.empty(B, dtype=torch.uint8, device=dev); rew = torch.empty(B, dtype=torch.float32, device=dev)
        obs = torch.empty((B, N * d), dtype=torch.float32, device=dev)
        fl = C.HK_F_ACT_DISCRETE | C.HK_F_OBS_SORT_LEX | C.HK_F_OBS_RESCALE
        def one(i, s=stream):
            assert L.hk_step(x.data_ptr(), x.data_ptr(), ha[i % T].data_ptr(), ax[i % T].data_ptr(), done.data_ptr(), rew.data_ptr(),
                             None, obs.data_ptr(), None, None, B, N, d, C.HK_DTYPE_I32, C.HK_OP_SHIFT | C.HK_OP_REPOSITION | C.HK_OP_NEWTON,
                             fl, -1.0, 1e8, s) == 0
        def timed(fn, n=2000):
            for i in range(10): fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n): fn(i)
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n * 1e3
        s_us = timed(one)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            one(0, torch.cuda.current_stream().cuda_stream)
        g_us = timed(lambda i: g.replay())
        out[f"{'warp' if fam else 'thread'}-per-game B={B},N={N}"] = {"stream_us": round(s_us, 2), "graph_us": round(g_us, 2)}
L.hk_debug_force_generic(0)
print(json.dumps(out, indent=0))
