"""Per-step launch times of a 20-step random-play rollout for the three thread-per-game shapes
(N = 5, 10, 20; d = 3) at 1 Mi games, int32 state, in place; algorithmic bytes vs the HBM peak."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hironaka_b200 import constants as C
from hironaka_b200._lib import lib

L = lib(); dev = torch.device("cuda", 0); stream = torch.cuda.current_stream(dev).cuda_stream
B, d, T = 1 << 20, 3, 20
if os.environ.get("HK_PDL"):
    L.hk_debug_set_pdl(1)  # programmatic dependent launch (A/B)
for N in (5, 10, 20):
    rng = np.random.default_rng(N)
    x0 = torch.from_numpy(rng.integers(0, 20, (B, N, d), dtype=np.int32)).to(dev)
    ha = torch.from_numpy(rng.integers(0, 4, (T, B), dtype=np.int32)).to(dev)
    ax = torch.from_numpy(rng.integers(0, d, (T, B), dtype=np.int32)).to(dev)
    done = torch.empty(B, dtype=torch.uint8, device=dev); rew = torch.empty(B, dtype=torch.float32, device=dev)
    assert L.hk_step(x0.data_ptr(), x0.data_ptr(), None, None, None, None, None, None, None, None, B, N, d, 0,
                     C.HK_OP_NEWTON | C.HK_OP_REPOSITION, 0, -1.0, 1e8, stream) == 0
    per = np.zeros(T)
    for rep in range(4):
        x = x0.clone()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(T + 1)]
        torch.cuda.synchronize(); ev[0].record()
        for t in range(T):
            assert L.hk_step(x.data_ptr(), x.data_ptr(), ha[t].data_ptr(), ax[t].data_ptr(), done.data_ptr(), rew.data_ptr(),
                             None, None, None, None, B, N, d, 0, C.HK_OP_SHIFT | C.HK_OP_REPOSITION | C.HK_OP_NEWTON,
                             C.HK_F_ACT_DISCRETE, -1.0, 1e8, stream) == 0
            ev[t + 1].record()
        torch.cuda.synchronize()
        if rep: per += np.array([ev[t].elapsed_time(ev[t + 1]) for t in range(T)])
    per /= 3
    byt = 8 * N * d + 13
    print(f"N={N:2d}: mean {per.mean()*1e3:6.1f} us/step = {B/(per.mean()*1e-3):.3e} game-steps/s, "
          f"{B*byt/(per.mean()*1e-3)/6543.7e9:.2f} of HBM by algorithmic bytes ({byt} B) | t0..t3 {np.round(per[:4]*1e3,1)} late {per[-1]*1e3:.1f}")
