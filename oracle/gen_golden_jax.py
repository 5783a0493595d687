"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/ref_jax_*.npz by EXECUTING THE
REFERENCE'S OWN JAX-FLAVOUR SOURCES, unmodified, where they lie under /root/reference:

    hironaka/src/_jax_ops.py                  (shift_jax, reposition_jax, get_newton_polytope_jax, rescale_jax)
    hironaka/jax/host_action_preprocess.py    (decode tables, get_batch_decode, batch_encode)
    hironaka/jax/util.py                      (get_take_actions, get_dones, get_reward_fn, get_feature_fn,
                                               make_agent_obs, get_done_from_flatten, generate_pts,
                                               calculate_value_using_reward_fn, select_sample_after_sim)
    hironaka/jax/players.py                   (all_coord_host_fn, zeillinger_fn, choose_first/last_agent_fn)

`jax` is not installed in this image, so the files run against oracle/jax_numpy_shim.py (a NumPy
stand-in for jnp / vmap / jit / lax with float32 narrowing).  The package __init__ files of the
reference (which import flax, optax, mctx, ...) are bypassed: each file is loaded by path under its
own module name.  Runs only in the build container; the small fixtures it writes are committed.

    python oracle/gen_golden_jax.py        # rewrites tests/golden/ref_jax_*.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
REF = os.environ.get("HIRONAKA_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "..", "tests", "golden")


def load_reference():
    from oracle import jax_numpy_shim
    jax_numpy_shim.install()

    def pkg(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    pkg("hironaka")
    src = pkg("hironaka.src")
    pkg("hironaka.jax")
    ops = load("hironaka.src._jax_ops", "hironaka/src/_jax_ops.py")
    for n in ("get_newton_polytope_jax", "rescale_jax", "shift_jax", "reposition_jax", "remove_repeated_jax"):
        setattr(src, n, getattr(ops, n))
    hap = load("hironaka.jax.host_action_preprocess", "hironaka/jax/host_action_preprocess.py")
    load("hironaka.jax.loss", "hironaka/jax/loss.py")
    util = load("hironaka.jax.util", "hironaka/jax/util.py")
    players = load("hironaka.jax.players", "hironaka/jax/players.py")
    return ops, hap, util, players


def rollout(ops, hap, util, players, seed, B, N, d, T, max_value, role, reposition, rescale_points, with_players=False):
    """compute_rho's loop (hironaka/jax/jax_trainer.py:497-534) with recorded uniform random players:
    host = uniform discrete id, agent = uniform over ALL d axes (players.py:28-39,142-153: invalid
    actions occur and the JAX step applies them; a lone point keeps shifting)."""
    rng = np.random.default_rng(seed)
    spec = (N, d)
    raw = rng.integers(0, max_value, size=(1, B, N, d)).astype(np.float32)
    # generate_pts (util.py:385-392) is pmapped over a leading device axis; its random draw is replaced by `raw`
    pts = ops.get_newton_polytope_jax(raw[0])
    if reposition:
        pts = ops.reposition_jax(pts)
    if rescale_points:
        pts = ops.rescale_jax(pts)
    take_actions = util.get_take_actions(role, spec, rescale_points, reposition)
    reward_fn = util.get_reward_fn(role)
    feat_host = util.get_feature_fn("host", spec, True)
    feat_host_raw = util.get_feature_fn("host", spec, False)
    feat_agent = util.get_feature_fn("agent", spec, True)
    decode = hap.get_batch_decode(d)
    ncls = 2 ** d - d - 1
    states, dones, rewards = [pts], [util.get_dones(pts)], []
    hids, axes, fh, fhr, fa, dflat = [], [], [], [], [], []
    zeil, first, last = [], [], []
    for t in range(T):
        hid = rng.integers(0, ncls, size=B).astype(np.int32)
        ax = rng.integers(0, d, size=B).astype(np.int32)
        coords = decode(hid).astype(np.float32)
        cur = states[-1]
        if with_players:
            zeil.append(np.argmax(players.zeillinger_fn(cur), axis=1).astype(np.int32))
            aobs0 = util.make_agent_obs(cur, coords)
            first.append(np.argmax(players.choose_first_agent_fn(aobs0, spec), axis=1).astype(np.int32))
            last.append(np.argmax(players.choose_last_agent_fn(aobs0, spec), axis=1).astype(np.int32))
        if role == "host":
            nxt = take_actions(util.flatten(cur), coords, ax)
        else:
            nxt = take_actions(util.make_agent_obs(cur, coords), ax, ax)
        nxt = nxt.reshape(B, N, d)
        dn = util.get_dones(nxt)
        rewards.append(reward_fn(dn, dones[-1]))
        states.append(nxt)
        dones.append(dn)
        hids.append(hid)
        axes.append(ax)
        fh.append(feat_host(util.flatten(nxt)))
        fhr.append(feat_host_raw(util.flatten(nxt)))
        aobs = util.make_agent_obs(nxt, coords)
        fa.append(feat_agent(aobs))
        dflat.append(np.stack([util.get_done_from_flatten(util.flatten(nxt), "host", d),
                               util.get_done_from_flatten(aobs, "agent", d)]))
    out = dict(raw=raw[0], states=np.stack(states), dones=np.stack(dones), rewards=np.stack(rewards).astype(np.float32),
               host_ids=np.stack(hids), axes=np.stack(axes), feat_host=np.stack(fh), feat_host_raw=np.stack(fhr),
               feat_agent=np.stack(fa), done_from_flatten=np.stack(dflat),
               meta=np.array([seed, B, N, d, T, max_value, int(role == "agent"), int(reposition), int(rescale_points)]))
    if with_players:
        out.update(zeillinger_id=np.stack(zeil), choose_first=np.stack(first), choose_last=np.stack(last),
                   all_coord_id=np.argmax(players.all_coord_host_fn(states[0]), axis=1).astype(np.int32))
    return out


def per_op(ops, seed, B, N, d, max_value):
    """The four JAX ops one at a time on states with dead rows, duplicates, zero columns and tiny maxima."""
    rng = np.random.default_rng(seed)
    pts = rng.integers(0, max_value, size=(B, N, d)).astype(np.float32)
    dead = rng.random((B, N)) < 0.3
    pts[dead] = -1.0
    for b in range(0, B, 3):
        i, j = rng.integers(0, N, 2)
        pts[b, j] = pts[b, i]
    pts[1::5, :, 0] = np.where(pts[1::5, :, 0] >= 0, 0.0, -1.0)  # a zero column: reposition leaves it alone
    coord = rng.integers(0, 2, size=(B, d)).astype(np.float32)
    axis = rng.integers(0, d, size=B).astype(np.int32)
    out = dict(points=pts, coord=coord, axes=axis)
    out["shift"] = ops.shift_jax(pts, coord, axis)
    out["reposition"] = ops.reposition_jax(pts)
    out["remove_repeated"] = ops.remove_repeated_jax(pts)
    out["newton"] = ops.get_newton_polytope_jax(pts)
    out["rescale"] = ops.rescale_jax(pts)
    tiny = pts.copy()
    tiny[tiny > 0] *= np.float32(1e-10)  # maxima <= 1e-8: calculate_rescale returns the points unchanged
    tiny[0] = np.where(pts[0] >= 0, 0.0, -1.0)
    out["tiny_points"] = tiny
    out["tiny_rescale"] = ops.rescale_jax(tiny)
    return out


def value_targets(util, seed, B, T):
    rng = np.random.default_rng(seed)
    start = rng.integers(1, 9, size=(B, 1))
    drops = np.cumsum(rng.integers(0, 3, size=(B, T)), axis=1)
    num_points = np.clip(start + 6 - drops, 0, None).astype(np.int32)
    prior = rng.standard_normal((B, T)).astype(np.float32)
    out = dict(num_points=num_points)
    for role in ("host", "agent"):
        for unified in (False, True):
            est = util.get_value_est_fn(role)
            rew = util.get_reward_fn("agent" if (unified or role == "agent") else "host")
            v = util.calculate_value_using_reward_fn(prior, num_points, np.float32(0.9), rew, est, unified)
            out[f"value_{role}_{int(unified)}"] = np.asarray(v, dtype=np.float32)
    return out


def select_after_sim(util, seed, S, N, d):
    """select_sample_after_sim (util.py:351-382): the deterministic part (unfinished states) and the count
    contract of the random part (|selected| >= |undone|, selected is a superset of undone)."""
    rng = np.random.default_rng(seed)
    pts = -np.ones((S, N, d), np.float32)
    live = rng.integers(0, 4, size=S)
    for s in range(S):
        pts[s, : live[s]] = rng.integers(0, 9, size=(live[s], d))
    coords = rng.integers(0, 2, size=(S, d)).astype(np.float32)
    host_obs = pts.reshape(S, -1)
    agent_obs = np.concatenate([host_obs, coords], axis=1)
    dummy = (np.zeros((S, 4), np.float32), np.zeros(S, np.float32))
    out = dict(host_obs=host_obs, agent_obs=agent_obs)
    out["undone_host"] = util.select_sample_after_sim("host", (host_obs,) + dummy, d, False, key=np.array([0, 1], np.uint32))
    out["undone_agent"] = util.select_sample_after_sim("agent", (agent_obs,) + dummy, d, False, key=np.array([0, 1], np.uint32))
    return out


def main():
    ops, hap, util, players = load_reference()
    os.makedirs(OUT, exist_ok=True)
    cases = [
        # name, seed, B, N, d, T, max_value, role, reposition, rescale_points, players
        ("c2_20x3", 101, 96, 20, 3, 20, 20, "host", True, False, True),     # the headline configuration
        ("c2_20x3_agent", 102, 48, 20, 3, 12, 20, "agent", True, False, False),
        ("c2_20x3_norepos", 103, 48, 20, 3, 12, 20, "host", False, False, False),
        ("c2_20x3_rescaled", 104, 48, 20, 3, 12, 20, "host", True, True, False),
        ("c1_10x3", 105, 64, 10, 3, 10, 21, "host", True, False, True),
        ("c4_5x3", 106, 64, 5, 3, 8, 20, "host", True, False, True),
        ("c5_64x5", 107, 6, 64, 5, 10, 20, "host", True, False, False),      # config 5 shape
        ("g_16x4", 108, 24, 16, 4, 8, 9, "host", True, False, True),
    ]
    for name, *a in cases:
        np.savez_compressed(os.path.join(OUT, f"ref_jax_rollout_{name}.npz"), **rollout(ops, hap, util, players, *a))
        print("rollout", name, flush=True)
    for name, seed, B, N, d, mv in [("20x3", 111, 40, 20, 3, 8), ("10x3", 112, 40, 10, 3, 6), ("64x5", 113, 4, 64, 5, 6),
                                    ("7x2", 114, 32, 7, 2, 5)]:
        np.savez_compressed(os.path.join(OUT, f"ref_jax_ops_{name}.npz"), **per_op(ops, seed, B, N, d, mv))
    tabs = {}
    for d in range(2, 8):
        ncls = 2 ** d - d - 1
        tab = np.asarray(hap.get_batch_decode(d)(np.arange(ncls)))
        tabs[f"decode_{d}"] = tab
        tabs[f"encode_{d}"] = np.asarray(hap.batch_encode(tab)).astype(np.int32)
        tabs[f"from_one_hot_{d}"] = np.asarray(hap.get_batch_decode_from_one_hot(d)(np.eye(ncls, dtype=np.float32)))
    np.savez_compressed(os.path.join(OUT, "ref_jax_tables.npz"), **tabs)
    np.savez_compressed(os.path.join(OUT, "ref_jax_value_targets.npz"), **value_targets(util, 121, 32, 12))
    np.savez_compressed(os.path.join(OUT, "ref_jax_select_after_sim.npz"), **select_after_sim(util, 131, 200, 5, 3))
    print("wrote", sorted(f for f in os.listdir(OUT) if f.startswith("ref_jax_")))


if __name__ == "__main__":
    main()
