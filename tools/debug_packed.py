"""Debug helper: the packed-tier border inputs of tests/test_gpu_parity.py, mismatching games printed."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cport  # noqa: E402
from oracle import hk_oracle as O  # noqa: E402
from hironaka_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B, d = 4099, 3
rng = np.random.default_rng(255 + N)
ncls = 4
ops_bits, flags = O.OP_SHIFT | O.OP_REPOSITION | O.OP_NEWTON, O.F_ACT_DISCRETE
x = np.empty((B, N, d), np.int32)
scale = np.repeat(rng.choice([3, 20, 85, 86, 128, 255, 256, 300, 70000], (B + 31) // 32), 32)[:B]
x[:] = rng.integers(0, scale[:, None, None] + 1, (B, N, d))
x[rng.random(B) < 0.3] //= 7
hi = rng.random(B) < 0.15
x[hi, : max(1, N // 2)] = scale[hi, None, None]
dup = rng.integers(0, N, (B, 4))
for k in range(3):
    x[np.arange(B), dup[:, k + 1]] = x[np.arange(B), dup[:, 0]]
dead = rng.random((B, N)) < rng.choice([0.0, 0.3, 0.7, 0.95], B)[:, None]
x[dead] = -1
x[dead & (rng.random((B, N)) < 0.2)] = -9
x[::97] = -1
neg = np.arange(B) % 53 == 7
x[neg, 0, 0] = 4
x[neg, 0, 1] = -2
ha = rng.integers(0, ncls, B).astype(np.int32)
ax = rng.integers(0, d, B).astype(np.int32)
exp = cport.step(x, ha, ax, ops_bits, flags)
for inplace in (False, True):
    g = torch.from_numpy(x.copy()).cuda()
    r = ops.step(g, torch.from_numpy(ha).cuda(), torch.from_numpy(ax).cuda(), ops=ops_bits, flags=flags, inplace=inplace,
                 want_done=True, want_num_points=True)
    got = (g if inplace else r.state).cpu().numpy()
    bad = np.nonzero((got != exp[0]).any(axis=(1, 2)))[0]
    print("inplace", inplace, "mismatching games", len(bad), bad[:10], "neg among them", int(neg[bad].sum()))
    for b in bad[:3]:
        print("game", b, "scale", scale[b], "ha", ha[b], "ax", ax[b], "\nin\n", x[b], "\ngot\n", got[b], "\nexp\n", exp[0][b])
