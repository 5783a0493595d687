"""Pins the oracle's JAX flavour (NumPy restatement + C port) against outputs of the reference's OWN
JAX sources (hironaka/src/_jax_ops.py, hironaka/jax/util.py, host_action_preprocess.py, players.py),
executed unmodified over a NumPy stand-in for jax in the build container
(oracle/gen_golden_jax.py -> tests/golden/ref_jax_*.npz).  CPU only."""
import glob
import os

import numpy as np
import pytest

from oracle import cport
from oracle import hk_oracle as O


def jax_rollout_files(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "ref_jax_rollout_*.npz")))


def jax_case(g):
    seed, B, N, d, T, mv, agent, repos, resc = g["meta"].tolist()
    ops = O.OP_SHIFT | O.OP_NEWTON | (O.OP_REPOSITION if repos else 0) | (O.OP_RESCALE if resc else 0)
    flags = O.F_ACT_DISCRETE | O.F_RESCALE_EPS | (O.F_ROLE_AGENT if agent else 0)
    root_ops = O.OP_NEWTON | (O.OP_REPOSITION if repos else 0) | (O.OP_RESCALE if resc else 0)
    return B, N, d, T, ops, flags, root_ops, bool(resc)


def test_jax_golden_files_present(golden_dir):
    assert len(jax_rollout_files(golden_dir)) >= 8
    for f in ("ref_jax_tables.npz", "ref_jax_value_targets.npz", "ref_jax_select_after_sim.npz"):
        assert os.path.exists(os.path.join(golden_dir, f))
    assert len(glob.glob(os.path.join(golden_dir, "ref_jax_ops_*.npz"))) >= 4


@pytest.mark.parametrize("impl", ["numpy", "c_f32", "c_i32"])
def test_jax_golden_rollouts(golden_dir, impl):
    """take_actions -> get_dones -> reward_fn -> feature_fn over whole random-play rollouts: invalid
    actions are applied, lone points keep shifting (util.py:117-123, _jax_ops.py:76-90)."""
    for path in jax_rollout_files(golden_dir):
        g = np.load(path)
        B, N, d, T, ops, flags, root_ops, resc = jax_case(g)
        if resc and impl == "c_i32":
            continue  # rescaled states are fractions
        x = g["raw"].astype(np.int32) if impl == "c_i32" else g["raw"]
        if impl == "numpy":
            x = O.step(x, None, None, root_ops, flags)[0]
        else:
            x = cport.step(x, None, None, root_ops, flags)[0]
        assert np.array_equal(x.astype(np.float32), g["states"][0]), path
        for t in range(T):
            hid, ax = g["host_ids"][t], g["axes"][t]
            if impl == "numpy":
                x, done, rew, npts, _ = O.step(x, hid, ax, ops, flags)
                fh = O.feature_fn("host", (N, d), x.reshape(B, -1), True)
                fr = O.feature_fn("host", (N, d), x.reshape(B, -1), False)
                fa = O.feature_fn("agent", (N, d), O.make_agent_obs(x, O.decode_table(d)[hid].astype(np.float32)), True)
            else:
                x, done, rew, npts = cport.step(x, hid, ax, ops, flags)
                ff = O.F_OBS_SORT_LEX | O.F_RESCALE_EPS
                fh = cport.features(x, ff | O.F_OBS_RESCALE)
                fr = cport.features(x, ff)
                fa = cport.features(x, ff | O.F_OBS_RESCALE | O.F_ACT_DISCRETE, obs_coord=hid)
            assert np.array_equal(x.astype(np.float32), g["states"][t + 1]), (path, t)
            assert np.array_equal(done.astype(bool), g["dones"][t + 1]), (path, t)
            assert np.array_equal(rew, g["rewards"][t]), (path, t)
            assert np.array_equal(fh, g["feat_host"][t]), (path, t)
            assert np.array_equal(fr, g["feat_host_raw"][t]), (path, t)
            assert np.array_equal(fa, g["feat_agent"][t]), (path, t)
            if impl == "numpy":
                dff = g["done_from_flatten"][t]
                assert np.array_equal(O.get_done_from_flatten(x.reshape(B, -1), "host", d), dff[0])
                assert np.array_equal(O.get_done_from_flatten(fa, "agent", d), dff[1])
        # the workload has the behaviour the JAX flavour is about: games end, and (without the reposition
        # that sends a lone point to the origin) ended games keep moving
        if "c2_20x3.npz" in path:
            assert g["dones"][-1].mean() > 0.9
        if "c2_20x3_norepos.npz" in path:
            ended_moves = sum(int((g["dones"][t][:, None, None] & (g["states"][t] != g["states"][t + 1])).any())
                              for t in range(T))
            assert ended_moves > 0


def test_jax_golden_fixed_players(golden_dir):
    """zeillinger_fn / choose_first / choose_last / all_coord (players.py:42-105,156-212) on rollout states."""
    seen = 0
    for path in jax_rollout_files(golden_dir):
        g = np.load(path)
        if "zeillinger_id" not in g.files:
            continue
        B, N, d, T, *_ = jax_case(g)
        assert np.array_equal(np.argmax(O.all_coord_host_fn(g["states"][0]), axis=1), g["all_coord_id"])
        for t in range(T):
            st = g["states"][t]
            assert np.array_equal(np.argmax(O.zeillinger_fn(st), axis=1), g["zeillinger_id"][t]), (path, t)
            obs = O.make_agent_obs(st, O.decode_table(d)[g["host_ids"][t]].astype(np.float32))
            assert np.array_equal(np.argmax(O.choose_first_agent_fn(obs, (N, d)), axis=1), g["choose_first"][t])
            assert np.array_equal(np.argmax(O.choose_last_agent_fn(obs, (N, d)), axis=1), g["choose_last"][t])
            seen += 1
    assert seen >= 20


@pytest.mark.parametrize("impl", ["numpy", "c_f32", "c_i32"])
def test_jax_golden_ops(golden_dir, impl):
    for path in sorted(glob.glob(os.path.join(golden_dir, "ref_jax_ops_*.npz"))):
        g = np.load(path)
        x = g["points"]
        B, N, d = x.shape
        cm = (g["coord"].astype(np.int64) * (1 << np.arange(d))).sum(1).astype(np.int32)
        ax = g["axes"]
        if impl == "numpy":
            assert np.array_equal(O.shift_jax(x, g["coord"], ax), g["shift"]), path
            assert np.array_equal(O.reposition_jax(x), g["reposition"]), path
            assert np.array_equal(O.remove_repeated_jax(x), g["remove_repeated"]), path
            assert np.array_equal(O.get_newton_polytope_jax(x), g["newton"]), path
            assert np.array_equal(O.rescale_jax(x), g["rescale"]), path
            assert np.array_equal(O.rescale_jax(g["tiny_points"]), g["tiny_rescale"]), path
            assert np.array_equal(O.step(g["tiny_points"], None, None, O.OP_RESCALE, O.F_RESCALE_EPS)[0], g["tiny_rescale"])
        else:
            xx = x.astype(np.int32) if impl == "c_i32" else x
            cast = lambda a: a.astype(np.float32)
            assert np.array_equal(cast(cport.step(xx, cm, ax, O.OP_SHIFT, 0)[0]), g["shift"]), path
            assert np.array_equal(cast(cport.step(xx, None, None, O.OP_REPOSITION, 0)[0]), g["reposition"]), path
            assert np.array_equal(cast(cport.step(xx, None, None, O.OP_DEDUPE, 0)[0]), g["remove_repeated"]), path
            assert np.array_equal(cast(cport.step(xx, None, None, O.OP_NEWTON, 0)[0]), g["newton"]), path
            if impl == "c_f32":
                assert np.array_equal(cport.step(x, None, None, O.OP_RESCALE, O.F_RESCALE_EPS)[0], g["rescale"]), path
                assert np.array_equal(cport.step(g["tiny_points"], None, None, O.OP_RESCALE, O.F_RESCALE_EPS)[0],
                                      g["tiny_rescale"]), path
                assert np.array_equal(cport.features(g["tiny_points"], O.F_OBS_RESCALE | O.F_RESCALE_EPS).reshape(B, N, d),
                                      g["tiny_rescale"]), path
                # without the flag the torch rule divides (only a maximum of exactly 0 is replaced)
                assert not np.array_equal(cport.step(g["tiny_points"], None, None, O.OP_RESCALE, 0)[0], g["tiny_rescale"])


def test_jax_golden_tables(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_jax_tables.npz"))
    for d in range(2, 8):
        assert np.array_equal(O.decode_table(d), g[f"decode_{d}"])
        assert np.array_equal(O.decode_table(d), g[f"from_one_hot_{d}"])
        assert np.array_equal(O.encode(g[f"decode_{d}"]), g[f"encode_{d}"])


def test_jax_golden_value_targets(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_jax_value_targets.npz"))
    for role in ("host", "agent"):
        for unified in (False, True):
            reward_sign = -1 if (unified or role == "agent") else 1
            v = O.calculate_value_using_reward_fn(g["num_points"], 0.9, reward_sign, 1 if role == "host" else -1, unified)
            assert np.allclose(v, g[f"value_{role}_{int(unified)}"], rtol=1e-5, atol=1e-6), (role, unified)


def test_jax_golden_select_after_sim(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_jax_select_after_sim.npz"))
    d = 3
    assert np.array_equal((g["host_obs"] >= 0).sum(-1) > d, g["undone_host"])
    assert np.array_equal((g["agent_obs"] >= 0).sum(-1) > 2 * d, g["undone_agent"])
    assert 0 < g["undone_host"].sum() < len(g["undone_host"])
