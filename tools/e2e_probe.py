"""Where the end-to-end step time goes: hk_session_rollout_ex on a state whose games are all at rest (the kernel
is ~14 us), with and without the per-game done read-back, with and without graph replay."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hironaka_b200 import HostSession, constants as C
from hironaka_b200._lib import lib

B, N, d, T = 1 << 20, 20, 3, 20
rng = np.random.default_rng(0)
x = -np.ones((B, N, d), np.int32)
x[:, 0] = 0  # every game: a lone point at the origin
ha = HostSession.pack_actions(rng.integers(0, 4, (T, B)), rng.integers(0, 3, (T, B)))
ha_pin = torch.from_numpy(ha).pin_memory().numpy()
done_pin = torch.empty((T, B), dtype=torch.uint8).pin_memory().numpy()
ops = C.HK_OP_SHIFT | C.HK_OP_REPOSITION | C.HK_OP_NEWTON
flags = C.HK_F_ACT_DISCRETE | C.HK_F_ACT_PACKED
out = {}
for graphs in (0, 1):
    lib().hk_debug_set_session_graphs(graphs)
    for want_done in (False, True):
        s = HostSession(x)
        s.step(None, None, C.HK_OP_NEWTON | C.HK_OP_REPOSITION, 0)
        ts = []
        for rep in range(8):
            t0 = time.perf_counter()
            s.rollout(ha_pin, None, ops, flags, done=done_pin if want_done else None)
            ts.append((time.perf_counter() - t0) * 1e6 / T)
        out[f"graphs={graphs},done={want_done}"] = [round(t, 1) for t in ts]
        s.close()
print(json.dumps(out))
