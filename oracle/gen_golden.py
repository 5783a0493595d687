"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/ref_*.npz by running the REAL
reference (honglu2875/hironaka, imported unmodified from /root/reference) on seeded inputs.

Runs only in the build container (the reference tree does not exist on the GPU box); the
resulting small fixtures are committed.  JAX is not installed, so `jax`, `jaxlib`, `chex`
are stubbed with MagicMock exactly as SURVEY.md section 8c describes — the torch path does not
touch them.

    python oracle/gen_golden.py            # rewrites tests/golden/ref_*.npz
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

REF = os.environ.get("HIRONAKA_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    for m in ["jax", "jax.numpy", "jaxlib", "jaxlib.xla_extension", "chex", "gym", "gym.spaces", "treelib"]:
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, REF)
    from hironaka.core import TensorPoints  # noqa
    from hironaka.src import (HostActionEncoder, get_newton_polytope_torch, remove_repeated, reposition_torch,
                              rescale_torch, shift_torch)
    return dict(TensorPoints=TensorPoints, HostActionEncoder=HostActionEncoder,
                newton=get_newton_polytope_torch, remove_repeated=remove_repeated, reposition=reposition_torch,
                rescale=rescale_torch, shift=shift_torch)


def rollout_torch(ref, seed, B, N, d, T, max_value, agent, reposition, pad=-1.0):
    """The FusedGame composition (fused_game.py:150-162) with fixed players, through
    TensorPoints: random host (uniform discrete id) vs ChooseFirst / uniform-random-axis agent."""
    g = torch.Generator().manual_seed(seed)
    pts = torch.randint(0, max_value + 1, (B, N, d), generator=g).float()  # trainer.py:592-600 (inclusive)
    tp = ref["TensorPoints"](pts.clone(), padding_value=pad)
    tp.get_newton_polytope()
    enc = ref["HostActionEncoder"](d)
    ncls = 2 ** d - d - 1
    states = [tp.points.clone().numpy()]
    host_ids, axes, dones, npts, feats = [], [], [tp.ended_batch_in_tensor.numpy().copy()], [tp.get_num_points().numpy().copy()], []
    for t in range(T):
        hid = torch.randint(0, ncls, (B,), generator=g)
        coords = enc.decode_tensor(hid)
        if agent == "first":
            ax = coords.argmax(1)
        else:  # uniform over all d axes: invalid actions occur (no-op in the torch path)
            ax = torch.randint(0, d, (B,), generator=g)
        tp.shift(coords, ax.float())
        if reposition:
            tp.reposition()
        tp.get_newton_polytope()
        states.append(tp.points.clone().numpy())
        host_ids.append(hid.numpy().astype(np.int32))
        axes.append(ax.numpy().astype(np.int32))
        dones.append(tp.ended_batch_in_tensor.numpy().copy())
        npts.append(tp.get_num_points().numpy().copy())
        feats.append(tp.get_features().numpy().copy())
    return dict(init=pts.numpy(), states=np.stack(states), host_ids=np.stack(host_ids), axes=np.stack(axes),
                dones=np.stack(dones), num_points=np.stack(npts).astype(np.int32), features=np.stack(feats),
                meta=np.array([seed, B, N, d, T, max_value, int(agent == "first"), int(reposition)]))


def per_op(ref, seed, B, N, d, max_value, pad=-1.0):
    g = torch.Generator().manual_seed(seed)
    pts = torch.randint(0, max_value + 1, (B, N, d), generator=g).float()
    # sprinkle dead rows and exact duplicates
    dead = torch.rand((B, N), generator=g) < 0.25
    pts[dead] = pad
    for b in range(0, B, 3):
        i, j = int(torch.randint(0, N, (1,), generator=g)), int(torch.randint(0, N, (1,), generator=g))
        pts[b, j] = pts[b, i]
    enc = ref["HostActionEncoder"](d)
    ncls = 2 ** d - d - 1
    hid = torch.randint(0, ncls, (B,), generator=g)
    coords = enc.decode_tensor(hid)
    ax = torch.randint(0, d, (B,), generator=g)
    out = dict(points=pts.numpy(), host_ids=hid.numpy().astype(np.int32), axes=ax.numpy().astype(np.int32),
               coords=coords.numpy())
    out["shift_ignore_ended"] = ref["shift"](pts.clone(), coords, ax, inplace=False, padding_value=pad).numpy()
    out["shift_force_ended"] = ref["shift"](pts.clone(), coords, ax, inplace=False, padding_value=pad,
                                            ignore_ended_games=False).numpy()
    out["newton"] = ref["newton"](pts.clone(), inplace=False, padding_value=pad).numpy()
    rr = pts.clone()
    ref["remove_repeated"](rr, padding_value=pad)
    out["remove_repeated"] = rr.numpy()
    out["reposition"] = ref["reposition"](pts.clone(), inplace=False, padding_value=pad).numpy()
    out["rescale"] = ref["rescale"](pts.clone(), inplace=False, padding_value=pad).numpy()
    nw = ref["newton"](pts.clone(), inplace=False, padding_value=pad)
    out["rescale_after_newton"] = ref["rescale"](nw.clone(), inplace=False, padding_value=pad).numpy()
    tp = ref["TensorPoints"](pts.clone(), padding_value=pad)
    out["num_points"] = tp.get_num_points().numpy().astype(np.int32)
    out["ended"] = tp.ended_batch_in_tensor.numpy()
    return out


def list_points_goldens(ref, seed, B, N, d, max_value):
    """ListPoints order: get_newton_polytope_approx_lst (hironaka/src/_list_ops.py:9-45) on ragged lists."""
    from hironaka.src import get_newton_polytope_approx_lst
    rng = np.random.default_rng(seed)
    games, padded = [], -np.ones((B, N, d), np.float32)
    for b in range(B):
        n = int(rng.integers(1, N + 1))
        pts = rng.integers(0, max_value, (n, d)).astype(float).tolist()
        games.append(pts)
        padded[b, :n] = np.array(pts, np.float32)
    out = get_newton_polytope_approx_lst([[list(p) for p in g] for g in games], inplace=False)
    res = -np.ones((B, N, d), np.float32)
    counts = np.zeros(B, np.int32)
    for b in range(B):
        counts[b] = len(out[b])
        res[b, : counts[b]] = np.array(out[b], np.float32)
    return dict(points=padded, newton_list_order=res, counts=counts)


def tables(ref):
    out = {}
    for d in range(2, 8):
        enc = ref["HostActionEncoder"](d)
        ncls = 2 ** d - d - 1
        tab = enc.decode_tensor(torch.arange(ncls)).numpy()
        out[f"decode_{d}"] = tab
        out[f"encode_{d}"] = enc.encode_tensor(torch.tensor(tab)).numpy()
    return out


def main():
    ref = import_reference()
    os.makedirs(OUT, exist_ok=True)
    cases = [
        # name, seed, B, N, d, T, max_value, agent, reposition
        ("c1_10x3", 0, 64, 10, 3, 10, 20, "first", False),   # BASELINE config 1 shape (torch semantics)
        ("c2_20x3", 1, 64, 20, 3, 12, 20, "random", False),  # config 2 shape, invalid actions as no-ops
        ("c2_20x3_repos", 2, 48, 20, 3, 12, 20, "random", True),
        ("c4_5x3", 3, 64, 5, 3, 8, 20, "first", False),
        ("g_16x4", 4, 32, 16, 4, 8, 9, "random", True),
        ("c5_64x5", 5, 8, 64, 5, 8, 20, "random", False),    # config 5 shape
        ("g_7x2", 6, 32, 7, 2, 6, 5, "first", True),
    ]
    for name, seed, B, N, d, T, mv, agent, rep in cases:
        np.savez_compressed(os.path.join(OUT, f"ref_rollout_{name}.npz"), **rollout_torch(ref, seed, B, N, d, T, mv, agent, rep))
    for name, seed, B, N, d, mv in [("10x3", 10, 48, 10, 3, 6), ("20x3", 11, 48, 20, 3, 8), ("6x4", 12, 32, 6, 4, 4),
                                    ("64x5", 13, 6, 64, 5, 6), ("33x3", 14, 8, 33, 3, 6)]:
        np.savez_compressed(os.path.join(OUT, f"ref_ops_{name}.npz"), **per_op(ref, seed, B, N, d, mv))
    np.savez_compressed(os.path.join(OUT, "ref_tables.npz"), **tables(ref))
    for name, seed, B, N, d, mv in [("20x3", 21, 64, 20, 3, 6), ("10x4", 22, 32, 10, 4, 4), ("40x2", 23, 16, 40, 2, 9)]:
        np.savez_compressed(os.path.join(OUT, f"ref_list_{name}.npz"), **list_points_goldens(ref, seed, B, N, d, mv))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
