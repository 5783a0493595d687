"""Latency of one hk_step call by batch size for both kernel families (tuning aid, not product)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hironaka_b200 import constants as C, ops
from hironaka_b200._lib import lib

L = lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
N, d = int(sys.argv[1]) if len(sys.argv) > 1 else 20, 3
rng = np.random.default_rng(0)
print(f"N={N} d={d}: us per call (after root filter + 2 steps, typical live counts)")
for B in (10, 100, 512, 2048, 8192, 32768, 131072, 524288):
    x0 = torch.from_numpy(rng.integers(0, 20, (B, N, d), dtype=np.int32)).to(dev)
    ha = torch.from_numpy(rng.integers(0, 4, (4, B), dtype=np.int32)).to(dev)
    ax = torch.from_numpy(rng.integers(0, 3, (4, B), dtype=np.int32)).to(dev)
    done = torch.empty(B, dtype=torch.uint8, device=dev); rew = torch.empty(B, dtype=torch.float32, device=dev)
    obs = torch.empty((B, N * d), dtype=torch.float32, device=dev)
    opb = C.HK_OP_SHIFT | C.HK_OP_REPOSITION | C.HK_OP_NEWTON
    row = [f"B={B:7d}"]
    for gen in (False, True):
        ops.force_generic(gen)
        for with_obs in (False, True):
            x = x0.clone()
            L.hk_step(x.data_ptr(), x.data_ptr(), None, None, None, None, None, None, None, None, B, N, d, 0, C.HK_OP_NEWTON | C.HK_OP_REPOSITION, 0, -1.0, 1e8, stream)
            for t in range(2):
                L.hk_step(x.data_ptr(), x.data_ptr(), ha[t].data_ptr(), ax[t].data_ptr(), None, None, None, None, None, None, B, N, d, 0, opb, C.HK_F_ACT_DISCRETE, -1.0, 1e8, stream)
            snap = x.clone()
            flags = C.HK_F_ACT_DISCRETE | (C.HK_F_OBS_SORT_LEX | C.HK_F_OBS_RESCALE if with_obs else 0)
            def call():
                rc = L.hk_step(snap.data_ptr(), x.data_ptr(), ha[2].data_ptr(), ax[2].data_ptr(), done.data_ptr(), rew.data_ptr(), None,
                               obs.data_ptr() if with_obs else None, None, None, B, N, d, 0, opb, flags, -1.0, 1e8, stream)
                assert rc == 0
            for _ in range(5): call()
            torch.cuda.synchronize()
            n = 300 if B <= 32768 else 50
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n): call()
            e1.record(); torch.cuda.synchronize()
            row.append(f"{'gen' if gen else 'thr'}{'+obs' if with_obs else '    '} {e0.elapsed_time(e1) / n * 1e3:8.1f}")
    print("  ".join(row))
ops.force_generic(False)
