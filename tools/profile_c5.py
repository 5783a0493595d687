"""C5 workload (N=64, d=5, 256K games): root filter + one 20-step random-play rollout.
Used for the ncu captures of hk_generic_kernel committed under profiles/."""
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
x = torch.from_numpy(rng.integers(0, 20, (B, N, d), dtype=np.int32)).to(dev)
ha = torch.from_numpy(rng.integers(0, 26, (T, B), dtype=np.int32)).to(dev)
ax = torch.from_numpy(rng.integers(0, d, (T, B), dtype=np.int32)).to(dev)
done = torch.empty(B, dtype=torch.uint8, device=dev); rew = torch.empty(B, dtype=torch.float32, device=dev)
assert L.hk_step(x.data_ptr(), x.data_ptr(), None, None, None, None, None, None, None, None, B, N, d, 0,
                 C.HK_OP_NEWTON | C.HK_OP_REPOSITION, 0, -1.0, 1e8, stream) == 0
for t in range(T):
    assert L.hk_step(x.data_ptr(), x.data_ptr(), ha[t].data_ptr(), ax[t].data_ptr(), done.data_ptr(), rew.data_ptr(), None,
                     None, None, None, B, N, d, 0, C.HK_OP_SHIFT | C.HK_OP_REPOSITION | C.HK_OP_NEWTON,
                     C.HK_F_ACT_DISCRETE, -1.0, 1e8, stream) == 0
torch.cuda.synchronize()
print("done", int(done.sum()))
