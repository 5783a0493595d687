"""GPU: the vectorised environments (hironaka_b200.vec_env) against fixtures from the REAL reference
environments and, on larger batches, against the NumPy restatement that those fixtures pin
(tests/env_restatement.py, tests/test_gym_env_cpu.py)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from tests import env_restatement as E  # noqa: E402
from tests.test_gym_env_cpu import AGENT_SETS, HOST_SETS  # noqa: E402

pytestmark = pytest.mark.gpu


def bits(mask, d):
    return (mask[:, None] >> np.arange(d)) & 1


@pytest.mark.parametrize("name", sorted(AGENT_SETS))
def test_vec_agent_env_against_reference_fixture(golden_dir, name):
    from hironaka_b200 import VecHironakaAgentEnv
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = AGENT_SETS[name]
    B, N, d = g["points"].shape
    discrete = g["actions"].ndim == 2
    env = VecHironakaAgentEnv(B, agent="choose_first", use_discrete_actions_for_host=discrete, dimension=d,
                              max_num_points=N, **cfg)
    ref = E.AgentEnv(N, d, **cfg)
    exact = not cfg["scale_observation"]
    o = env.reset(torch.from_numpy(g["points"])).cpu().numpy()
    assert np.array_equal(o, ref.reset(g["points"]))
    if exact:
        assert np.array_equal(o, g["obs0"])
    for t in range(g["actions"].shape[0]):
        a = g["actions"][t]
        obs, rew, stop, _ = env.step(torch.from_numpy(a))
        robs, rrew, rstop = ref.step(a if discrete else E.multibinary_to_mask(a))
        assert np.array_equal(obs.cpu().numpy(), robs), (name, t)
        assert np.array_equal(rew.cpu().numpy().astype(np.float64), rrew), (name, t)
        assert np.array_equal(stop.cpu().numpy(), rstop), (name, t)
        if exact:  # and therefore the real environment, bit for bit
            assert np.array_equal(obs.cpu().numpy(), g["obs"][t]) and np.array_equal(rrew, g["reward"][t])
            assert np.array_equal(rstop, g["stopped"][t])


@pytest.mark.parametrize("name", sorted(HOST_SETS))
def test_vec_host_env_against_reference_fixture(golden_dir, name):
    from hironaka_b200 import VecHironakaHostEnv
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = dict(HOST_SETS[name])
    host = cfg.pop("host")
    B, N, d = g["points"].shape
    env = VecHironakaHostEnv(B, host={"Zeillinger": "zeillinger", "AllCoordHost": "all_coord"}[host], dimension=d,
                             max_num_points=N, **cfg)
    ref = E.HostEnv(N, d, host=host, **cfg)
    exact = not cfg["scale_observation"]
    o = env.reset(torch.from_numpy(g["points"]))
    robs, rcoords = ref.reset(g["points"])
    assert np.array_equal(o["points"].cpu().numpy(), robs)
    assert np.array_equal(o["coords"].cpu().numpy(), bits(rcoords, d))
    if exact:
        assert np.array_equal(o["points"].cpu().numpy(), g["obs0"]) and np.array_equal(o["coords"].cpu().numpy(), g["coords0"])
    alive = np.ones(B, bool)
    for t in range(g["actions"].shape[0]):
        o, rew, stop, _ = env.step(torch.from_numpy(g["actions"][t]))
        robs, rcoords, rrew, rstop = ref.step(g["actions"][t])
        assert np.array_equal(o["points"].cpu().numpy(), robs), (name, t)
        assert np.array_equal(o["coords"].cpu().numpy(), bits(rcoords, d)), (name, t)
        assert np.allclose(rew.cpu().numpy(), rrew), (name, t)
        assert np.array_equal(stop.cpu().numpy(), rstop), (name, t)
        if exact:
            assert np.array_equal(o["points"].cpu().numpy()[alive], g["obs"][t][alive])
            assert np.array_equal(o["coords"].cpu().numpy()[alive], g["coords"][t][alive])
            alive &= ~g["stopped"][t]


@pytest.mark.parametrize("shape", [(3000, 20, 3), (1000, 10, 3), (500, 16, 4)], ids=lambda s: "x".join(map(str, s)))
def test_vec_envs_on_large_batches(shape):
    """Both environments against the restatement on batches far beyond what the one-game-at-a-time
    reference plays in reasonable time, with thresholds that trigger."""
    from hironaka_b200 import VecHironakaAgentEnv, VecHironakaHostEnv
    B, N, d = shape
    rng = np.random.default_rng(B + N)
    pts = rng.integers(0, 12, size=(B, N, d)).astype(np.int32)
    cfg = dict(scale_observation=True, step_threshold=7, value_threshold=400, reward_based_on_point_reduction=True)
    env = VecHironakaAgentEnv(B, agent="choose_first", dimension=d, max_num_points=N, **cfg)
    ref = E.AgentEnv(N, d, **cfg)
    assert np.array_equal(env.reset(torch.from_numpy(pts)).cpu().numpy(), ref.reset(pts))
    for t in range(9):
        a = rng.integers(0, 2, size=(B, d)).astype(np.int32)
        obs, rew, stop, _ = env.step(torch.from_numpy(a))
        robs, rrew, rstop = ref.step(E.multibinary_to_mask(a))
        assert np.array_equal(obs.cpu().numpy(), robs) and np.array_equal(rew.cpu().numpy().astype(np.float64), rrew)
        assert np.array_equal(stop.cpu().numpy(), rstop)
    hcfg = dict(scale_observation=True, value_threshold=400, invalid_move_penalty=-0.25)
    henv = VecHironakaHostEnv(B, host="zeillinger", dimension=d, max_num_points=N, **hcfg)
    href = E.HostEnv(N, d, host="Zeillinger", **hcfg)
    o = henv.reset(torch.from_numpy(pts))
    robs, rcoords = href.reset(pts)
    assert np.array_equal(o["points"].cpu().numpy(), robs) and np.array_equal(o["coords"].cpu().numpy(), bits(rcoords, d))
    for t in range(9):
        a = rng.integers(0, d, size=B).astype(np.int32)
        o, rew, stop, _ = henv.step(torch.from_numpy(a))
        robs, rcoords, rrew, rstop = href.step(a)
        assert np.array_equal(o["points"].cpu().numpy(), robs), t
        assert np.array_equal(o["coords"].cpu().numpy(), bits(rcoords, d)), t
        assert np.allclose(rew.cpu().numpy(), rrew) and np.array_equal(stop.cpu().numpy(), rstop)


def test_host_policy_agrees_with_fused_player():
    """hk_host_policy reports the set the fused Zeillinger host plays: stepping with the reported mask
    equals stepping with HK_F_HOST_ZEILLINGER."""
    from hironaka_b200 import constants as C, ops
    rng = np.random.default_rng(3)
    for (B, N, d) in ((777, 20, 3), (300, 10, 3), (200, 16, 4), (64, 64, 5)):
        x = rng.integers(0, 9, size=(B, N, d)).astype(np.int32)
        x[rng.random((B, N)) < 0.4] = -1
        st = torch.from_numpy(x).cuda()
        ops.step(st, ops=C.HK_OP_NEWTON, inplace=True)
        mask = ops.host_policy(st, "zeillinger")
        assert torch.equal(ops.host_policy(st, "all_coord"), torch.full_like(mask, (1 << d) - 1) * (ops.dones(st, True)[1] > 0))
        a, b = st.clone(), st.clone()
        ops.step(a, None, None, ops=C.HK_OP_SHIFT | C.HK_OP_NEWTON, flags=C.HK_F_HOST_ZEILLINGER | C.HK_F_AGENT_FIRST, inplace=True)
        ops.step(b, mask, None, ops=C.HK_OP_SHIFT | C.HK_OP_NEWTON, flags=C.HK_F_AGENT_FIRST, inplace=True)
        assert torch.equal(a, b)


def test_vec_agent_env_random_agent_and_generated_points():
    """RandomAgent: the axis is drawn uniformly among the chosen coordinates (fewer than two chosen:
    no move); reset() without points draws them like generate_points (hironaka/src/_fn.py:184-185)."""
    from hironaka_b200 import VecHironakaAgentEnv
    B, N, d = 20000, 10, 3
    g = torch.Generator(device="cuda").manual_seed(5)
    env = VecHironakaAgentEnv(B, agent="random", dimension=d, max_num_points=N, max_value=15, scale_observation=False,
                              generator=g)
    obs = env.reset()
    assert obs.shape == (B, N, d) and float(obs.max()) < 15 and float(obs.min()) >= -1
    before = env.points.clone()
    # single coordinate chosen: nothing may move
    a = torch.zeros((B, d), dtype=torch.int32)
    a[:, 1] = 1
    env.step(a)
    assert torch.equal(env.points, before)
    # two coordinates chosen: exactly one of them receives the sum, each about half of the time
    a[:, 2] = 1
    live = before[:, :, 0] >= 0
    env.step(a)
    # undo the filter's effect by recomputing the shift on the old state for both candidate axes
    s = before[:, :, 1] + before[:, :, 2]
    moved1 = before.clone(); moved1[:, :, 1] = torch.where(live, s, moved1[:, :, 1])
    moved2 = before.clone(); moved2[:, :, 2] = torch.where(live, s, moved2[:, :, 2])
    from hironaka_b200 import constants as C, ops
    for m in (moved1, moved2):
        ops.step(m, ops=C.HK_OP_NEWTON, inplace=True)
    is1 = (env.points == moved1).reshape(B, -1).all(1)
    is2 = (env.points == moved2).reshape(B, -1).all(1)
    assert bool((is1 | is2).all())
    frac = float((is1 & ~is2).float().sum() / ((is1 ^ is2).float().sum() + 1e-9))
    assert 0.45 < frac < 0.55, frac


def test_overflow_flags_per_game():
    """hk_overflow: per-game form of exceed_threshold (>= for TensorPoints, > for ListPoints), both dtypes and
    kernel-independent shapes, against numpy."""
    from hironaka_b200 import ops
    rng = np.random.default_rng(9)
    for (B, N, d) in [(1000, 20, 3), (77, 64, 5), (5, 1, 3)]:
        x = rng.integers(-1, 50, size=(B, N, d)).astype(np.int32)
        x[x < 0] = -1
        for dtype in (np.int32, np.float32):
            xv = torch.from_numpy(x.astype(dtype)).cuda()
            for thr in (49.0, 48.0, 1e8):
                mx = x.reshape(B, -1).max(1)
                assert np.array_equal(ops.overflow(xv, thr).cpu().numpy(), mx >= thr)
                assert np.array_equal(ops.overflow(xv, thr, strict=True).cpu().numpy(), mx > thr)


@pytest.mark.parametrize("name", sorted(HOST_SETS))
def test_vec_host_env_captured_step(golden_dir, name):
    """VecHironakaHostEnv.capture_step: one CUDA-graph replay per environment step, against the eager step on
    the same action sequence (observation, coordinates, reward, stop flags, and the environment's own state)."""
    from hironaka_b200 import VecHironakaHostEnv
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = dict(HOST_SETS[name])
    host = {"Zeillinger": "zeillinger", "AllCoordHost": "all_coord"}[cfg.pop("host")]
    B, N, d = g["points"].shape
    cfg["value_threshold"] = 300.0  # exercise the overflow flags too
    eager = VecHironakaHostEnv(B, host=host, dimension=d, max_num_points=N, **cfg)
    graphed = VecHironakaHostEnv(B, host=host, dimension=d, max_num_points=N, **cfg)
    eager.reset(torch.from_numpy(g["points"]))
    graphed.reset(torch.from_numpy(g["points"]))
    step = graphed.capture_step()
    assert torch.equal(graphed.points, eager.points) and torch.equal(graphed.current_step, eager.current_step)
    rng = np.random.default_rng(3)
    for t in range(8):
        a = torch.from_numpy(rng.integers(0, d, B).astype(np.int32)).cuda()
        eo, er, es, _ = eager.step(a)
        go, gr, gs, _ = step(a)
        assert torch.equal(go["points"], eo["points"]) and torch.equal(go["coords"], eo["coords"]), (name, t)
        assert torch.equal(gr, er) and torch.equal(gs, es), (name, t)
        assert torch.equal(graphed.points, eager.points) and torch.equal(graphed.exceed_threshold, eager.exceed_threshold)
