"""The reference arm of bench.py: the UNMODIFIED reference (honglu2875/hironaka 0.0.1, installed into
baseline/_ref/ by baseline/install_ref.sh) driven through its own public API on the host's cores.

What runs: `hironaka.core.TensorPoints` (hironaka/core/tensor_points.py:11-126) over the torch ops of
`hironaka/src/_torch_ops.py:8-146` — per step `HostActionEncoder.decode_tensor` -> `TensorPoints.shift` ->
`TensorPoints.reposition` -> `TensorPoints.get_newton_polytope` -> `ended_batch_in_tensor` -> reward, i.e.
the C2 step composition of `get_take_actions` (hironaka/jax/util.py:117-123) through the reference's torch
implementation.  The reference's JAX implementation cannot be timed: `jax` is not installed in this image
(`jax`, `jaxlib`, `chex`, `gym`, `treelib` are stubbed with MagicMock only so that `import hironaka.core`
succeeds; the torch path never touches them).  Timed the reference's own way (`time.perf_counter`,
hironaka/trainer/timer.py:19-32) with torch using every host thread.

None of this repository's kernels or engine is on this path.
"""
from __future__ import annotations

import os
import sys
import time
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "hironaka"))


def import_reference():
    if not available():
        raise RuntimeError("baseline/_ref is missing: run baseline/install_ref.sh where /root/reference exists")
    for m in ("jax", "jax.numpy", "jaxlib", "jaxlib.xla_extension", "chex", "gym", "gym.spaces", "treelib"):
        sys.modules.setdefault(m, MagicMock())
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import torch
    from hironaka.core import TensorPoints
    from hironaka.src import HostActionEncoder
    return torch, TensorPoints, HostActionEncoder


def root_states(TensorPoints, torch, raw: np.ndarray, reposition: bool):
    """generate_pts' composition (hironaka/jax/util.py:385-392: newton -> reposition) through TensorPoints."""
    tp = TensorPoints(torch.from_numpy(raw.astype(np.float32)))
    tp.get_newton_polytope()
    if reposition:
        tp.reposition()
    return tp


def play(tp, enc, torch, ha: np.ndarray, ax: np.ndarray, reposition: bool, record=None):
    """T steps of random play on a TensorPoints batch.  ha/ax: [T, B] int arrays.  Returns the number of
    finished games after each step; `record` (a list) receives (state, done, reward) per step."""
    counts = []
    prev_done = tp.ended_batch_in_tensor
    for t in range(ha.shape[0]):
        coords = enc.decode_tensor(torch.from_numpy(ha[t]).long())
        tp.shift(coords, torch.from_numpy(ax[t]).float())
        if reposition:
            tp.reposition()
        tp.get_newton_polytope()
        done = tp.ended_batch_in_tensor
        reward = (done & ~prev_done).float()
        counts.append(int(done.sum()))
        if record is not None:
            record.append((tp.points.clone().numpy(), done.numpy().copy(), reward.numpy().copy()))
        prev_done = done
    return counts


def rate(make_inputs, games: int, steps: int, warmup: int, rollout_len: int, seed: int = 1234,
         reposition: bool = True, shape=(20, 3), threads: int | None = None):
    """game-steps/s of the reference on `games` games: `warmup` untimed steps, then EXACTLY `steps` timed
    steps walking through independent `rollout_len`-step rollouts (the workload bench.py gives the GPU arm).
    Returns (rate, cores, torch_threads, seconds)."""
    torch, TensorPoints, HostActionEncoder = import_reference()
    cores = os.cpu_count() or 1
    torch.set_num_threads(threads or cores)
    n_roll = max(1, -(-(steps + warmup) // rollout_len))
    pts, ha, ax = make_inputs(seed, games, n_roll)
    enc = HostActionEncoder(shape[1])
    batches = [root_states(TensorPoints, torch, pts[r], reposition) for r in range(n_roll)]
    k, t0 = 0, None
    for i in range(warmup + steps):
        if i == warmup:
            t0 = time.perf_counter()
        r, t = divmod(k, rollout_len)
        play(batches[r % n_roll], enc, torch, ha[r % n_roll, t:t + 1], ax[r % n_roll, t:t + 1], reposition)
        k += 1
    dt = time.perf_counter() - t0
    return games * steps / dt, cores, torch.get_num_threads(), dt
