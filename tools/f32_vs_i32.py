"""C2 rollout, per-step launches: float32 state (the reference's storage) vs int32 state."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hironaka_b200 import constants as C
from hironaka_b200._lib import lib
L = lib(); dev = torch.device("cuda", 0); stream = torch.cuda.current_stream(dev).cuda_stream
B, N, d, T = 1 << 20, 20, 3, 20
rng = np.random.default_rng(0)
x0 = rng.integers(0, 20, (B, N, d)).astype(np.int32)
ha = torch.from_numpy(rng.integers(0, 4, (T, B), dtype=np.int32)).to(dev)
ax = torch.from_numpy(rng.integers(0, 3, (T, B), dtype=np.int32)).to(dev)
done = torch.empty(B, dtype=torch.uint8, device=dev); rew = torch.empty(B, dtype=torch.float32, device=dev)
for name, dt, code, extra_ops, root_extra in (("int32", torch.int32, 0, 0, 0), ("float32", torch.float32, 1, 0, 0),
                                             ("float32+rescale (DQN)", torch.float32, 1, C.HK_OP_RESCALE, C.HK_OP_RESCALE),
                                             ("float32, rescaled ROOT only", torch.float32, 1, 0, C.HK_OP_RESCALE),
                                             ("float32, rescale in steps only", torch.float32, 1, C.HK_OP_RESCALE, 0)):
    root = torch.from_numpy(x0).to(dev).to(dt)
    assert L.hk_step(root.data_ptr(), root.data_ptr(), None, None, None, None, None, None, None, None, B, N, d, code, C.HK_OP_NEWTON | C.HK_OP_REPOSITION | root_extra, 0, -1.0, 1e8, stream) == 0
    per = np.zeros(T)
    for rep in range(4):
        x = root.clone()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(T + 1)]
        torch.cuda.synchronize(); ev[0].record()
        for t in range(T):
            assert L.hk_step(x.data_ptr(), x.data_ptr(), ha[t].data_ptr(), ax[t].data_ptr(), done.data_ptr(), rew.data_ptr(), None, None, None, None, B, N, d, code, C.HK_OP_SHIFT | C.HK_OP_REPOSITION | C.HK_OP_NEWTON | extra_ops, C.HK_F_ACT_DISCRETE, -1.0, 1e8, stream) == 0
            ev[t + 1].record()
        torch.cuda.synchronize()
        if rep: per += np.array([ev[t].elapsed_time(ev[t + 1]) for t in range(T)])
    per /= 3
    npts = torch.empty(B, dtype=torch.int32, device=dev)
    L.hk_dones(x.data_ptr(), done.data_ptr(), npts.data_ptr(), B, N, d, code, stream)
    print(f"{name:32s} live after rollout {npts.float().mean().item():.3f}  mean {per.mean()*1e3:7.1f} us/step  t0..t3 {np.round(per[:4]*1e3,1)}  late {per[-1]*1e3:.1f}")
