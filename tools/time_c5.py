"""C5 workload (N=64, d=5, 256K games): per-step launch times of a 20-step random-play rollout and
the root filter, device-timed.  Tuning aid for hk_generic_kernel.  argv[1] = extra flag bits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hironaka_b200 import constants as C
from hironaka_b200._lib import lib

L = lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev).cuda_stream
B, N, d, T = 1 << 18, 64, 5, 20
rng = np.random.default_rng(5)
x0 = torch.from_numpy(rng.integers(0, 20, (B, N, d), dtype=np.int32)).to(dev)
ha = torch.from_numpy(rng.integers(0, 26, (T, B), dtype=np.int32)).to(dev)
ax = torch.from_numpy(rng.integers(0, d, (T, B), dtype=np.int32)).to(dev)
done = torch.empty(B, dtype=torch.uint8, device=dev); rew = torch.empty(B, dtype=torch.float32, device=dev)
extra = int(sys.argv[1]) if len(sys.argv) > 1 else 0
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1e8
per, root = [], []
for rep in range(4):
    x = x0.clone()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(T + 2)]
    torch.cuda.synchronize()
    ev[0].record()
    assert L.hk_step(x.data_ptr(), x.data_ptr(), None, None, None, None, None, None, None, None, B, N, d, 0,
                     C.HK_OP_NEWTON | C.HK_OP_REPOSITION, extra, -1.0, 1e8, stream) == 0
    ev[1].record()
    for t in range(T):
        assert L.hk_step(x.data_ptr(), x.data_ptr(), ha[t].data_ptr(), ax[t].data_ptr(), done.data_ptr(), rew.data_ptr(),
                         None, None, None, None, B, N, d, 0, C.HK_OP_SHIFT | C.HK_OP_REPOSITION | C.HK_OP_NEWTON,
                         C.HK_F_ACT_DISCRETE | extra, -1.0, thr, stream) == 0
        ev[t + 2].record()
    torch.cuda.synchronize()
    if rep:
        root.append(ev[0].elapsed_time(ev[1]))
        per.append([ev[t + 1].elapsed_time(ev[t + 2]) for t in range(T)])
per = np.mean(np.array(per), axis=0)
print("flags", extra, "root filter %.3f ms | mean step %.4f ms = %.3e game-steps/s | hbm frac %.3f" % (
    np.mean(root), per.mean(), B / (per.mean() * 1e-3), B * 2573 / (per.mean() * 1e-3) / 6543.7e9))
print(" per step:", " ".join("%.3f" % v for v in per), "| done", int(done.sum()), "checksum", int(x.sum()))

# one-launch rollout (hk_rollout): the 20 steps in a single launch, state read and written once
dcount = torch.zeros(T, dtype=torch.int32, device=dev)
ms = []
for rep in range(4):
    x = x0.clone()
    assert L.hk_step(x.data_ptr(), x.data_ptr(), None, None, None, None, None, None, None, None, B, N, d, 0,
                     C.HK_OP_NEWTON | C.HK_OP_REPOSITION, 0, -1.0, 1e8, stream) == 0
    dcount.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    assert L.hk_rollout(x.data_ptr(), x.data_ptr(), ha.data_ptr(), ax.data_ptr(), None, None, dcount.data_ptr(), None,
                        B, N, d, T, 0, C.HK_OP_SHIFT | C.HK_OP_REPOSITION | C.HK_OP_NEWTON, C.HK_F_ACT_DISCRETE | extra,
                        -1.0, stream) == 0
    e1.record(); torch.cuda.synchronize()
    if rep: ms.append(e0.elapsed_time(e1))
print("one-launch rollout of %d steps: %.3f ms = %.3e game-steps/s | finished per step %s | checksum %d" % (
    T, np.mean(ms), B * T / (np.mean(ms) * 1e-3), dcount.tolist()[-3:], int(x.sum())))
