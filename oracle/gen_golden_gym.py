"""
ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/ref_gym_*.npz by running the REAL
reference gym environments (hironaka/gym_env/hironaka_agent_env.py, hironaka_host_env.py, imported
unmodified from /root/reference) one game at a time, with deterministic players
(ChooseFirstAgent, Zeillinger / AllCoordHost) and seeded action streams.

`gym` is not installed: a minimal stand-in module (Env base class, inert spaces) is put in
sys.modules before the import; the environments only use gym for their space declarations.
Runs only in the build container; the fixtures are committed.

    python oracle/gen_golden_gym.py
"""
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

REF = os.environ.get("HIRONAKA_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    gym = types.ModuleType("gym")
    gym.__version__ = "0.21.0"

    class Env:  # noqa: D401 - stand-in for gym.Env
        pass

    class _Space:
        def __init__(self, *a, **k):
            pass

    spaces = types.ModuleType("gym.spaces")
    for name in ("Box", "Dict", "MultiBinary", "Discrete", "Space"):
        setattr(spaces, name, type(name, (_Space,), {}))
    gym.Env, gym.spaces = Env, spaces
    gym.envs = MagicMock()
    sys.modules["gym"], sys.modules["gym.spaces"] = gym, spaces
    sys.modules["gym.envs"], sys.modules["gym.envs.registration"] = gym.envs, MagicMock()
    for m in ["jax", "jax.numpy", "jaxlib", "jaxlib.xla_extension", "chex", "treelib"]:
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, REF)
    from hironaka.agent import ChooseFirstAgent
    from hironaka.gym_env.hironaka_agent_env import HironakaAgentEnv
    from hironaka.gym_env.hironaka_host_env import HironakaHostEnv
    from hironaka.host import AllCoordHost, Zeillinger
    return dict(AgentEnv=HironakaAgentEnv, HostEnv=HironakaHostEnv, ChooseFirstAgent=ChooseFirstAgent,
                Zeillinger=Zeillinger, AllCoordHost=AllCoordHost)


def agent_env_set(ref, seed, B, N, d, T, max_value, action_mode, **cfg):
    """B independent HironakaAgentEnv games (fixed ChooseFirstAgent, host actions from a seeded
    stream).  Stopped environments keep being stepped (the reference allows it); the vectorised
    environment must agree step by step."""
    rng = np.random.default_rng(seed)
    pts = rng.integers(0, max_value, size=(B, N, d))
    ncls = 2 ** d - d - 1
    if action_mode == "multibinary":
        acts = rng.integers(0, 2, size=(T, B, d))
    elif action_mode == "compressed":
        acts = rng.integers(0, ncls, size=(T, B))
    else:  # raw binary code of the coordinate set
        acts = rng.integers(0, 2 ** d, size=(T, B))
    obs0 = np.zeros((B, N, d), np.float32)
    obs = np.zeros((T, B, N, d), np.float32)
    rew = np.zeros((T, B), np.float64)
    stop = np.zeros((T, B), bool)
    for b in range(B):
        env = ref["AgentEnv"](ref["ChooseFirstAgent"](), dimension=d, max_num_points=N, max_value=max_value,
                              use_discrete_actions_for_host=(action_mode != "multibinary"),
                              compressed_host_output=(action_mode == "compressed"), **cfg)
        if action_mode == "compressed":  # the reference decodes ids with decode_action (binary code): feed the code
            pass
        obs0[b] = env.reset(points=[pts[b].tolist()])
        for t in range(T):
            a = acts[t, b]
            o, r, s, _ = env.step(np.array(a) if action_mode == "multibinary" else int(a))
            obs[t, b], rew[t, b], stop[t, b] = o, r, s
    return dict(points=pts.astype(np.int32), actions=acts.astype(np.int32), obs0=obs0, obs=obs, reward=rew, stopped=stop)


def host_env_set(ref, seed, B, N, d, T, max_value, host, **cfg):
    """B independent HironakaHostEnv games (fixed Zeillinger / AllCoord host, agent axes from a seeded
    stream: a uniformly random axis, so invalid moves occur)."""
    rng = np.random.default_rng(seed)
    pts = rng.integers(0, max_value, size=(B, N, d))
    acts = rng.integers(0, d, size=(T, B))
    obs0 = np.zeros((B, N, d), np.float32)
    coords0 = np.zeros((B, d), np.int8)
    obs = np.zeros((T, B, N, d), np.float32)
    coords = np.zeros((T, B, d), np.int8)
    rew = np.zeros((T, B), np.float64)
    stop = np.zeros((T, B), bool)
    for b in range(B):
        env = ref["HostEnv"](ref[host](), dimension=d, max_num_points=N, max_value=max_value, **cfg)
        o = env.reset(points=[pts[b].tolist()])
        obs0[b], coords0[b] = o["points"], o["coords"]
        for t in range(T):
            o, r, s, _ = env.step(int(acts[t, b]))
            obs[t, b], coords[t, b], rew[t, b], stop[t, b] = o["points"], o["coords"], r, s
            if s:  # the reference host refuses ended games (assert not points.ended): freeze the record
                obs[t + 1:, b], coords[t + 1:, b], rew[t + 1:, b], stop[t + 1:, b] = o["points"], 0, np.nan, True
                break
    return dict(points=pts.astype(np.int32), actions=acts.astype(np.int32), obs0=obs0, coords0=coords0, obs=obs,
                coords=coords, reward=rew, stopped=stop)


def main():
    ref = import_reference()
    os.makedirs(OUT, exist_ok=True)
    sets = {
        "ref_gym_agent_mb_noscale": agent_env_set(ref, 1, 48, 10, 3, 12, 10, "multibinary", scale_observation=False,
                                                  reward_based_on_point_reduction=True),
        "ref_gym_agent_code_scale": agent_env_set(ref, 2, 48, 10, 3, 12, 10, "code", scale_observation=True),
        "ref_gym_agent_mb_threshold": agent_env_set(ref, 3, 48, 8, 4, 10, 20, "multibinary", scale_observation=False,
                                                    step_threshold=6, value_threshold=300,
                                                    fixed_penalty_crossing_threshold=-5),
        "ref_gym_agent_mb_steppenalty": agent_env_set(ref, 4, 32, 6, 3, 8, 12, "multibinary", scale_observation=False,
                                                      step_threshold=5),
        "ref_gym_host_zeillinger_noscale": host_env_set(ref, 5, 48, 10, 3, 12, 10, "Zeillinger", scale_observation=False),
        "ref_gym_host_zeillinger_scale": host_env_set(ref, 6, 48, 10, 3, 12, 10, "Zeillinger", scale_observation=True),
        "ref_gym_host_allcoord_stop": host_env_set(ref, 7, 32, 8, 4, 10, 10, "AllCoordHost", scale_observation=False,
                                                   stop_after_invalid_move=True, invalid_move_penalty=-0.5,
                                                   value_threshold=500),
    }
    for name, data in sets.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
        print(name, {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
