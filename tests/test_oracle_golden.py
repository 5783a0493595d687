"""Pins the oracle (NumPy restatement + C port) against
  (1) the reference's own known-answer vectors (tests/kat.py), and
  (2) outputs of the real reference run in the build container (tests/golden/ref_*.npz).
CPU only."""
import glob
import os

import numpy as np
import pytest

from oracle import cport
from oracle import hk_oracle as O
from tests import kat as K

OPS_TORCH_FLAGS = O.F_NOOP_INVALID | O.F_FREEZE_ENDED


# ---------------------------------------------------------------- KATs: torch flavour


def test_kat_newton_shift_reposition_rescale_torch():
    p = O.get_newton_polytope_torch(K.R_IN)
    assert np.array_equal(p, K.R)
    p = O.shift_torch(p, K.R_COORD_LISTS, K.R_AXIS)
    assert np.array_equal(p, K.R2)
    p = O.reposition_torch(p)
    assert np.array_equal(p, K.R3)
    p = O.rescale_torch(p)
    assert np.array_equal(p, K.RS)  # the reference compares str(); exact float32 equality is stronger


def test_kat_invalid_and_ended_torch():
    p = O.get_newton_polytope_torch(K.R_IN)
    q = O.shift_torch(p, K.INVALID_COORD_LISTS, K.INVALID_AXIS)
    assert np.array_equal(q, K.R)
    assert np.array_equal(O.shift_torch(K.ENDED_P, [[0, 1]], [1], ignore_ended_games=True), K.ENDED_P)
    assert np.array_equal(O.shift_torch(K.ENDED_P, [[0, 1]], [1], ignore_ended_games=False), K.ENDED_Q)


def test_kat_remove_repeated_and_origin20():
    assert np.array_equal(O.remove_repeated(K.REP_IN), K.REP_OUT)
    assert np.array_equal(O.get_newton_polytope_torch(K.ORIGIN20_IN), K.ORIGIN20_OUT)
    assert np.isfinite(O.rescale_torch(K.RESCALE0_IN)).all()


# ---------------------------------------------------------------- KATs: JAX flavour


def test_kat_jax_ops():
    assert np.array_equal(O.get_newton_polytope_jax(K.R_IN), K.R)
    assert np.array_equal(O.get_newton_polytope_jax(K.EXTREME_IN), K.EXTREME_OUT)
    p = O.shift_jax(K.R_IN, K.R_COORD_BIN, K.R_AXIS)
    assert np.array_equal(O.get_newton_polytope_jax(p), K.R2)
    assert np.array_equal(O.rescale_jax(K.RESCALE_IN), K.RESCALE_OUT)
    assert np.array_equal(O.reposition_jax(O.get_newton_polytope_jax(p)), K.R3)
    assert np.allclose(O.rescale_jax(K.R3), K.RS)
    assert np.array_equal(O.shift_jax(K.FEAT_IN, np.array([[0, 1, 1]]), np.array([1])), K.FSHIFT_OUT)


def test_kat_jax_features():
    assert np.array_equal(O.feature_fn("host", (6, 3), K.FEAT_IN.reshape(1, -1)), K.FEAT_SORTED)
    assert np.array_equal(O.feature_fn("agent", (3, 3), K.AGENT_FEAT_IN, scale_observation=False), K.AGENT_FEAT_NOSCALE)
    assert np.array_equal(O.feature_fn("agent", (3, 3), K.AGENT_FEAT_IN, scale_observation=True), K.AGENT_FEAT_SCALE)


def test_kat_action_tables():
    assert np.array_equal(O.decode_table(3), K.DECODE_3)
    assert np.array_equal(O.encode(K.ENCODE_IN), K.ENCODE_OUT)
    assert np.array_equal(O.encode_one_hot(K.ENCODE_IN), K.ENCODE_ONE_HOT_OUT)
    assert np.array_equal(O.decode_table(3)[K.ENCODE_OUT], K.ENCODE_IN)
    # HostActionEncoder examples, test/testUtil.py:226-236 style: [0,1] <-> 0 in dimension 3
    assert O.encode(np.array([[1, 1, 0]]))[0] == 0


def test_kat_take_actions_composition():
    ones = np.ones(2, dtype=np.float32)
    out = O.take_actions("host", (4, 3), K.TA_HOST_OBS.reshape(2, -1), K.TA_COORDS, ones, rescale_points=True,
                         reposition=False)
    expect = O.rescale_jax(O.get_newton_polytope_jax(O.shift_jax(K.TA_HOST_OBS, K.TA_COORDS, ones))).reshape(2, -1)
    assert np.array_equal(out, expect)
    combined = O.make_agent_obs(K.TA_HOST_OBS, K.TA_COORDS)
    assert combined.shape == (2, 15)
    out = O.take_actions("agent", (4, 3), combined, ones, ones, rescale_points=False, reposition=False)
    expect = O.get_newton_polytope_jax(O.shift_jax(K.TA_HOST_OBS, K.TA_COORDS, ones)).reshape(2, -1)
    assert np.array_equal(out, expect)
    assert O.make_agent_obs(np.ones((32, 20, 3), np.float32), np.ones((32, 3), np.float32)).shape == (32, 63)


def test_kat_fixed_players():
    assert np.array_equal(O.all_coord_host_fn(K.HOSTS_OBS), np.array([[0, 0, 0, 1], [0, 0, 0, 1]], np.float32))
    assert np.array_equal(O.zeillinger_fn_slice(K.ZEIL_PTS), K.ZEIL_OUT)
    assert np.array_equal(O.zeillinger_fn(K.ZEIL_OBS2), O.encode_one_hot(K.ZEIL_OBS2_MB))
    assert np.allclose(O.zeillinger_fn(K.ZEIL_PTS3), K.ZEIL_OUT)
    obs = O.make_agent_obs(K.HOSTS_OBS, K.AGENT_COORDS)
    assert np.array_equal(O.choose_first_agent_fn(obs, (2, 3)), K.CHOOSE_FIRST_OUT)
    assert np.array_equal(O.choose_last_agent_fn(obs, (2, 3)), K.CHOOSE_LAST_OUT)


def test_kat_dones_rewards():
    d0 = O.get_dones(K.R)
    assert d0.tolist() == [False, False]
    assert O.get_dones(K.ORIGIN20_OUT).tolist() == [True]
    assert O.reward_fn("host", np.array([True, True, False]), np.array([False, True, False])).tolist() == [1, 0, 0]
    assert O.reward_fn("agent", np.array([True, True, False]), np.array([False, True, False])).tolist() == [-1, 0, 0]
    assert O.default_reward("agent", np.array([True, False])).tolist() == [-1, 0]
    assert O.get_done_from_flatten(K.AGENT_FEAT_IN, "agent", 3).tolist() == [False, False]


def test_kat_value_targets():
    """calculate_value_using_reward_fn / rollout_postprocess vs test/testJAXTrainer.py:330-389."""
    for npts, role, unified, expect in K.VALUE_KATS:
        reward_sign = -1 if (unified or role == "agent") else 1
        v = O.calculate_value_using_reward_fn(npts, 0.99, reward_sign, 1 if role == "host" else -1, unified)
        assert np.allclose(v, expect, rtol=1e-5, atol=1e-8)
    pol, val = np.zeros((1, 4, 3), np.float32), np.zeros((1, 4), np.float32)
    assert np.allclose(O.rollout_postprocess(K.VALUE_OBS_AGENT, pol, val, "agent", 3, 0.99, True)[2], [-1, 1, -1, 1])
    o, p_, v = O.rollout_postprocess(K.VALUE_OBS_HOST, pol, val, "host", 3, 0.99, True)
    assert np.allclose(v, [0.99, -1, 1, -1]) and o.shape == (4, 18) and p_.shape == (4, 3)


# ---------------------------------------------------------------- KATs through the C port


@pytest.mark.parametrize("dtype", [np.float32, np.int32])
def test_kat_cport(dtype):
    x = K.R_IN.astype(dtype)
    n, *_ = cport.step(x, None, None, O.OP_NEWTON, 0)
    assert np.array_equal(n, K.R.astype(dtype))
    s, *_ = cport.step(n, K.R_COORD_MASK, K.R_AXIS, O.OP_SHIFT, OPS_TORCH_FLAGS)
    assert np.array_equal(s, K.R2.astype(dtype))
    r, *_ = cport.step(s, None, None, O.OP_REPOSITION, 0)
    assert np.array_equal(r, K.R3.astype(dtype))
    # fused: shift + reposition on the newton output in one call
    f, done, rew, npts = cport.step(n, K.R_COORD_MASK, K.R_AXIS, O.OP_SHIFT | O.OP_REPOSITION | O.OP_NEWTON, 0)
    assert np.array_equal(f, K.R3.astype(dtype))
    assert done.tolist() == [0, 0] and rew.tolist() == [0, 0] and npts.tolist() == [3, 2]
    # invalid action: torch = no-op, JAX = applied
    inv_mask = np.array([0b0010, 0b1101], np.int32)
    q, *_ = cport.step(n, inv_mask, np.array(K.INVALID_AXIS, np.int32), O.OP_SHIFT, OPS_TORCH_FLAGS)
    assert np.array_equal(q, K.R.astype(dtype))
    e, *_ = cport.step(K.ENDED_P.astype(dtype), np.array([0b11], np.int32), np.array([1], np.int32), O.OP_SHIFT,
                       OPS_TORCH_FLAGS)
    assert np.array_equal(e, K.ENDED_P.astype(dtype))
    e, *_ = cport.step(K.ENDED_P.astype(dtype), np.array([0b11], np.int32), np.array([1], np.int32), O.OP_SHIFT,
                       O.F_NOOP_INVALID)
    assert np.array_equal(e, K.ENDED_Q.astype(dtype))
    o, done, rew, npts = cport.step(K.ORIGIN20_IN.astype(dtype), None, None, O.OP_NEWTON, 0)
    assert np.array_equal(o, K.ORIGIN20_OUT.astype(dtype))
    assert done.tolist() == [1] and rew.tolist() == [1.0] and npts.tolist() == [1]
    assert np.array_equal(cport.step(K.REP_IN.astype(dtype), None, None, O.OP_NEWTON, 0)[0], K.REP_OUT.astype(dtype))


def test_kat_cport_float_only():
    assert np.array_equal(cport.rescale(K.R3), K.RS)
    assert np.array_equal(cport.rescale(K.RESCALE_IN), K.RESCALE_OUT)
    assert np.isfinite(cport.rescale(K.RESCALE0_IN)).all()
    f = cport.features(K.FEAT_IN, O.F_OBS_SORT_LEX | O.F_OBS_RESCALE)
    assert np.array_equal(f, K.FEAT_SORTED)
    s, *_ = cport.step(K.FEAT_IN, np.array([0b110], np.int32), np.array([1], np.int32), O.OP_SHIFT, 0)
    assert np.array_equal(s, K.FSHIFT_OUT)
    pts = K.AGENT_FEAT_IN[:, :9].reshape(2, 3, 3).copy()
    cm = np.array([0b110, 0b101], np.int32)
    assert np.array_equal(cport.features(pts, O.F_OBS_SORT_LEX, obs_coord=cm), K.AGENT_FEAT_NOSCALE)
    assert np.array_equal(cport.features(pts, O.F_OBS_SORT_LEX | O.F_OBS_RESCALE, obs_coord=cm), K.AGENT_FEAT_SCALE)
    assert np.array_equal(cport.features(pts.astype(np.int32), O.F_OBS_SORT_LEX | O.F_OBS_RESCALE, obs_coord=cm),
                          K.AGENT_FEAT_SCALE)
    # discrete ids through the decode table
    s2, *_ = cport.step(K.TA_HOST_OBS, np.array([2, 3], np.int32), np.array([1, 1], np.int32),
                        O.OP_SHIFT | O.OP_NEWTON | O.OP_RESCALE, O.F_ACT_DISCRETE)
    ones = np.ones(2, np.float32)
    assert np.array_equal(s2.reshape(2, -1),
                          O.take_actions("host", (4, 3), K.TA_HOST_OBS.reshape(2, -1), K.TA_COORDS, ones, True, False))


# ---------------------------------------------------------------- real-reference goldens


def _rollout_files(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "ref_rollout_*.npz")))


def _ops_files(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "ref_ops_*.npz")))


def test_golden_files_present(golden_dir):
    assert len(_rollout_files(golden_dir)) >= 7 and len(_ops_files(golden_dir)) >= 5
    assert os.path.exists(os.path.join(golden_dir, "ref_tables.npz"))


@pytest.mark.parametrize("impl", ["numpy", "c_f32", "c_i32"])
def test_golden_rollouts(golden_dir, impl):
    for path in _rollout_files(golden_dir):
        g = np.load(path)
        seed, B, N, d, T, mv, first, repos = g["meta"].tolist()
        ops = O.OP_SHIFT | O.OP_NEWTON | (O.OP_REPOSITION if repos else 0)
        flags = OPS_TORCH_FLAGS | O.F_ACT_DISCRETE
        x = g["init"]
        if impl == "numpy":
            x = O.get_newton_polytope_torch(x)
        else:
            x = x.astype(np.int32) if impl == "c_i32" else x
            x, done, _, npts = cport.step(x, None, None, O.OP_NEWTON, 0)
            assert np.array_equal(done.astype(bool), g["dones"][0]) and np.array_equal(npts, g["num_points"][0])
        assert np.array_equal(x.astype(np.float32), g["states"][0]), path
        for t in range(T):
            prev_done = g["dones"][t]
            if impl == "numpy":
                x, done, rew, npts, _ = O.step(x, g["host_ids"][t], g["axes"][t], ops, flags)
            else:
                x, done, rew, npts = cport.step(x, g["host_ids"][t], g["axes"][t], ops, flags)
            assert np.array_equal(x.astype(np.float32), g["states"][t + 1]), (path, t)
            assert np.array_equal(done.astype(bool), g["dones"][t + 1]), (path, t)
            assert np.array_equal(npts, g["num_points"][t + 1]), (path, t)
            assert np.array_equal(rew, (g["dones"][t + 1] & ~prev_done).astype(np.float32)), (path, t)
            # TensorPoints.get_features: compare modulo ties on coordinate 0 (reference argsort is unstable);
            # coordinate-0 columns must agree exactly, full rows must agree wherever the key is unique.
            feat = O.get_features_torch(x.astype(np.float32)) if impl == "numpy" else \
                cport.features(x, O.F_OBS_SORT_COORD0).reshape(B, N, d)
            ref_feat = g["features"][t]
            assert np.array_equal(feat[:, :, 0], ref_feat[:, :, 0]), (path, t)
            if N <= 16:  # torch's argsort coincides with the stable order up to 16 elements
                assert np.array_equal(feat, ref_feat), (path, t)


@pytest.mark.parametrize("impl", ["numpy", "c_f32", "c_i32"])
def test_golden_ops(golden_dir, impl):
    for path in _ops_files(golden_dir):
        g = np.load(path)
        x = g["points"]
        B, N, d = x.shape
        hid, ax = g["host_ids"], g["axes"]
        if impl == "numpy":
            coords = O.decode_table(d)[hid].astype(np.float32)
            assert np.array_equal(coords, g["coords"])
            assert np.array_equal(O.shift_torch(x, coords, ax), g["shift_ignore_ended"]), path
            assert np.array_equal(O.shift_torch(x, coords, ax, ignore_ended_games=False), g["shift_force_ended"]), path
            assert np.array_equal(O.remove_repeated(x), g["remove_repeated"]), path
            assert np.array_equal(O.get_newton_polytope_torch(x), g["newton"]), path
            assert np.array_equal(O.get_newton_polytope_jax(x), g["newton"]), path
            assert np.array_equal(O.reposition_torch(x), g["reposition"]), path
            assert np.array_equal(O.reposition_jax(x), g["reposition"]), path
            assert np.array_equal(O.rescale_torch(x), g["rescale"]), path
            assert np.array_equal(O.rescale_jax(x), g["rescale"]), path
            assert np.array_equal(O.rescale_torch(g["newton"]), g["rescale_after_newton"]), path
            assert np.array_equal(O.get_num_points(x), g["num_points"]) and np.array_equal(O.ended_batch(x), g["ended"])
        else:
            xx = x.astype(np.int32) if impl == "c_i32" else x
            cast = lambda a: a.astype(np.float32)
            fl = O.F_ACT_DISCRETE
            assert np.array_equal(cast(cport.step(xx, hid, ax, O.OP_SHIFT, fl | OPS_TORCH_FLAGS)[0]), g["shift_ignore_ended"])
            assert np.array_equal(cast(cport.step(xx, hid, ax, O.OP_SHIFT, fl | O.F_NOOP_INVALID)[0]), g["shift_force_ended"])
            n, done, _, npts = cport.step(xx, None, None, O.OP_NEWTON, 0)
            assert np.array_equal(cast(n), g["newton"]), path
            assert np.array_equal(cast(cport.step(xx, None, None, O.OP_REPOSITION, 0)[0]), g["reposition"]), path
            _, done0, _, npts0 = cport.step(xx, None, None, 0, 0)
            assert np.array_equal(npts0, g["num_points"]) and np.array_equal(done0.astype(bool), g["ended"])
            if impl == "c_f32":
                assert np.array_equal(cport.rescale(x), g["rescale"]), path
                assert np.array_equal(cport.step(x, None, None, O.OP_NEWTON | O.OP_RESCALE, 0)[0], g["rescale_after_newton"])
                assert np.array_equal(cport.features(n, O.F_OBS_RESCALE).reshape(B, N, d), g["rescale_after_newton"])
            else:
                assert np.array_equal(cport.features(n, O.F_OBS_RESCALE).reshape(B, N, d), g["rescale_after_newton"])


def test_golden_list_points_order(golden_dir):
    """ListPoints order (get_newton_polytope_approx_lst, _list_ops.py:9-45): survivors sorted
    descending with coordinate 0 primary and compacted."""
    files = sorted(glob.glob(os.path.join(golden_dir, "ref_list_*.npz")))
    assert len(files) >= 3
    for path in files:
        g = np.load(path)
        for x in (g["points"], g["points"].astype(np.int32)):
            n, _, _, npts = cport.step(x, None, None, O.OP_NEWTON, 0)
            assert np.array_equal(npts, g["counts"]), path
            assert np.array_equal(cport.features(n, 1 << 12).reshape(x.shape), g["newton_list_order"]), path


def test_golden_tables(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_tables.npz"))
    for d in range(2, 8):
        assert np.array_equal(O.decode_table(d), g[f"decode_{d}"].astype(np.int32))
        assert np.array_equal(O.encode(g[f"decode_{d}"]), g[f"encode_{d}"])
        assert np.array_equal(O.encode(O.decode_table(d)), np.arange(2 ** d - d - 1))


def test_philox_known_answers():
    """The counter-based generator of the in-kernel random players (include/hironaka_b200.h, HK_F_HOST_RANDOM):
    Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors: philox4x32 10)."""
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kats:
        got = O.philox4x32_10(*[np.array([v], np.uint64) for v in c], *k)
        assert tuple(int(g[0]) for g in got) == want
    ha, ax = O.random_player_actions(100000, 3, 4, 42)
    assert set(np.unique(ha)) == {0, 1, 2, 3} and set(np.unique(ax)) == {0, 1, 2}
    assert abs(np.bincount(ha.ravel()) / ha.size - 0.25).max() < 0.01  # uniform over the four coordinate sets
    assert np.array_equal(O.random_player_actions(100000, 3, 2, 42, step_offset=2)[0], ha[2:])
