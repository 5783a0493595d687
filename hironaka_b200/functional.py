"""The functional env-step API of ``hironaka/jax/util.py`` on torch CUDA tensors.

Same factory names and call signatures as the reference (file:line in each docstring) so that a
``recurrent_fn`` (hironaka/jax/recurrent_fn.py:84-121) or ``compute_rho`` loop
(hironaka/jax/jax_trainer.py:497-534) reads the same; arrays are torch CUDA float32 tensors
instead of ``jnp`` arrays and every game operation is one launch of the sm_100a kernels.
JAX semantics: shift is applied unconditionally (invalid axis included, ended games included),
order is shift -> reposition -> newton -> rescale, padding is -1.

``get_env_step`` is the fused form the reference composes by hand in ``recurrent_fn``:
decode action -> step -> dones -> reward -> features in ONE launch.
"""
from __future__ import annotations

import functools
from typing import Callable, Optional, Tuple

import torch

from . import constants as C
from . import ops as _ops
from .host_action import (batch_encode, batch_encode_one_hot, decode_table, get_batch_decode,  # noqa: F401
                          get_batch_decode_from_one_hot)


def flatten(x: torch.Tensor) -> torch.Tensor:
    """vmap(jnp.ravel) (util.py:19)."""
    return x.reshape(x.shape[0], -1)


def make_agent_obs(pts: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """Flattened points concatenated with the host's coordinate set (util.py:22-31)."""
    return torch.cat([flatten(pts), coords.to(pts.dtype)], dim=1)


def get_dones(pts: torch.Tensor) -> torch.Tensor:
    """done iff fewer than 2 live rows (util.py:34-35)."""
    return _ops.dones(pts.contiguous())[0]


def get_done_from_flatten(obs: torch.Tensor, role: str, dimension: int) -> torch.Tensor:
    """util.py:38-39: #(entries >= 0) <= d (+ d for the agent's appended 0/1 coordinates), i.e.
    at most one live row in the points part."""
    width = obs.shape[-1] - (dimension if role == "agent" else 0)
    pts = obs[..., :width].reshape(obs.shape[0], width // dimension, dimension)
    return _ops.dones(pts.contiguous())[0]


@functools.lru_cache()
def get_preprocess_fns(role: str, spec: Tuple[int, int]) -> Tuple[Callable, Callable]:
    """(obs_preprocess, coords_preprocess) for a role (util.py:42-79)."""
    n, d = spec
    if role == "host":
        def obs_preprocess(observations):
            return observations.reshape(-1, n, d)

        def coords_preprocess(observations, actions):
            return actions
    elif role == "agent":
        def obs_preprocess(observations):
            return observations[:, : n * d].reshape(-1, n, d)

        def coords_preprocess(observations, actions):
            return observations[:, n * d: n * d + d]
    else:
        raise ValueError(f"role must be either host or agent. Got {role}.")
    return obs_preprocess, coords_preprocess


def _state_of(points: torch.Tensor) -> torch.Tensor:
    if points.dtype not in (torch.float32, torch.int32):
        points = points.to(torch.float32)
    return points.contiguous()


@functools.lru_cache()
def get_take_actions(role: str, spec: Tuple[int, int], rescale_points: bool = False,
                     reposition: bool = True) -> Callable:
    """Factory of ``take_actions(observations, actions, axis) -> [B, N*d]`` (util.py:82-125)."""
    obs_preprocess, coords_preprocess = get_preprocess_fns(role, spec)
    n, d = spec
    op_bits = C.HK_OP_SHIFT | C.HK_OP_NEWTON | (C.HK_OP_REPOSITION if reposition else 0) | \
        (C.HK_OP_RESCALE if rescale_points else 0)

    def take_actions(observations: torch.Tensor, actions: torch.Tensor, axis: torch.Tensor) -> torch.Tensor:
        points = _state_of(obs_preprocess(observations))
        coords = coords_preprocess(observations, actions)
        mask = _ops.coords_to_mask(coords, d, points.device)
        r = _ops.step(points, mask, axis, ops=op_bits, flags=C.JAX_SEMANTICS | C.HK_F_RESCALE_EPS, inplace=False)
        return r.state.reshape(-1, n * d)

    return take_actions


@functools.lru_cache()
def get_reward_fn(role: str) -> Callable:
    """host: +1 on the step the game ends; agent: -1 (util.py:128-149)."""
    if role not in ("host", "agent"):
        raise ValueError(f"role must be either host or agent. Got {role}.")
    sign = 1.0 if role == "host" else -1.0

    def reward_fn(dones: torch.Tensor, prev_dones: torch.Tensor) -> torch.Tensor:
        return (dones & ~prev_dones).to(torch.float32) * sign

    return reward_fn


@functools.lru_cache()
def get_feature_fn(role: str, spec: Tuple, scale_observation: bool = True) -> Callable:
    """Feature function: (rescale) + stable descending lexsort of rows, last coordinate primary;
    the agent keeps its d coordinates appended unchanged (util.py:172-214)."""
    assert len(spec) == 2
    n, d = spec
    flags = C.HK_F_OBS_SORT_LEX | C.HK_F_RESCALE_EPS | (C.HK_F_OBS_RESCALE if scale_observation else 0)
    if role == "host":
        def feature_fn(observations: torch.Tensor) -> torch.Tensor:
            pts = _state_of(observations.reshape(-1, n, d))
            return _ops.features(pts, flags=flags)
    elif role == "agent":
        def feature_fn(observations: torch.Tensor) -> torch.Tensor:
            pts = _state_of(observations[:, : n * d].reshape(-1, n, d))
            coords = observations[:, n * d: n * d + d]
            mask = _ops.coords_to_mask(coords, d, pts.device)
            return _ops.features(pts, flags=flags, obs_coord=mask)
    else:
        raise ValueError(f"role must be either host or agent. Got {role}.")
    return feature_fn


def generate_pts(generator: Optional[torch.Generator], shape: Tuple[int, int, int], max_value: int,
                 dtype=torch.float32, rescale: bool = True, reposition: bool = True, device="cuda") -> torch.Tensor:
    """Root states: randint[0, max_value) -> newton -> (reposition) -> (rescale) (util.py:385-392).
    The random draw is torch's (the reference uses jax.random); everything after it is one launch."""
    pts = torch.randint(0, max_value, shape, generator=generator, device=device).to(dtype)
    op_bits = C.HK_OP_NEWTON | (C.HK_OP_REPOSITION if reposition else 0) | (C.HK_OP_RESCALE if rescale else 0)
    _ops.step(pts, ops=op_bits, flags=C.HK_F_RESCALE_EPS, inplace=True)
    return pts


@functools.lru_cache()
def get_env_step(role: str, spec: Tuple[int, int], rescale_points: bool = False, reposition: bool = True,
                 scale_observation: bool = True, discrete_host_action: bool = True,
                 with_features: bool = True) -> Callable:
    """Fused step for MCTS node expansion / rollouts: ONE launch does what recurrent_fn
    composes from take_actions + get_dones + reward_fn + feature_fn (recurrent_fn.py:84-121).

    Returns ``env_step(points[B,N,d], host_action[B], axis[B], next_coord=None) ->
    (next_points[B,N,d], dones[B] bool, rewards[B] f32, features or None)``.  ``host_action`` holds
    discrete ids (mctx actions) when ``discrete_host_action`` else coordinate bitmasks; the
    reward sign follows ``role``; features are the role's network input (agent: next_coord
    appended)."""
    n, d = spec
    op_bits = C.HK_OP_SHIFT | C.HK_OP_NEWTON | (C.HK_OP_REPOSITION if reposition else 0) | \
        (C.HK_OP_RESCALE if rescale_points else 0)
    flags = (C.HK_F_ACT_DISCRETE if discrete_host_action else 0) | (C.HK_F_ROLE_AGENT if role == "agent" else 0) | \
        C.HK_F_OBS_SORT_LEX | C.HK_F_RESCALE_EPS | (C.HK_F_OBS_RESCALE if scale_observation else 0)

    def env_step(points: torch.Tensor, host_action: torch.Tensor, axis: torch.Tensor, next_coord=None,
                 inplace: bool = False):
        r = _ops.step(_state_of(points), host_action, axis, ops=op_bits, flags=flags, inplace=inplace,
                      want_done=True, want_reward=True, want_obs=with_features, obs_coord=next_coord)
        return r.state, r.done, r.reward, r.obs

    return env_step


class GraphedEnvStep:
    """``get_env_step`` with static buffers, captured once in a CUDA graph: the form for the
    latency-bound caller (MCTS node expansion at eval_batch_size 10-512, recurrent_fn.py:84-121),
    where a replay costs a few microseconds instead of a Python call with four allocations.

    Write the inputs into ``points`` (in place state, [B,N,d]), ``host_action`` and ``axis``
    (int32 [B]; ``next_coord`` for the agent role), call the object, read ``done`` (bool [B]),
    ``reward`` (float [B]) and ``obs`` (float [B, N*d (+d)]).  The device entry points neither
    allocate nor synchronise, which is what makes the capture legal."""

    def __init__(self, role: str, spec: Tuple[int, int], batch_size: int, device="cuda", dtype=torch.int32,
                 rescale_points: bool = False, reposition: bool = True, scale_observation: bool = True,
                 discrete_host_action: bool = True, with_features: bool = True):
        n, d = spec
        self.role, self.spec, self.batch_size = role, spec, batch_size
        dev = torch.device(device)
        self.points = torch.full((batch_size, n, d), -1, dtype=dtype, device=dev)
        self.host_action = torch.zeros(batch_size, dtype=torch.int32, device=dev)
        self.axis = torch.zeros(batch_size, dtype=torch.int32, device=dev)
        self.next_coord = torch.zeros(batch_size, dtype=torch.int32, device=dev) if role == "agent" else None
        self._done = torch.zeros(batch_size, dtype=torch.uint8, device=dev)
        self.done = self._done.view(torch.bool)
        self.reward = torch.zeros(batch_size, dtype=torch.float32, device=dev)
        self.obs = torch.zeros((batch_size, n * d + (d if role == "agent" else 0)), dtype=torch.float32, device=dev) \
            if with_features else None
        self._ops = C.HK_OP_SHIFT | C.HK_OP_NEWTON | (C.HK_OP_REPOSITION if reposition else 0) | \
            (C.HK_OP_RESCALE if rescale_points else 0)
        self._flags = (C.HK_F_ACT_DISCRETE if discrete_host_action else 0) | (C.HK_F_ROLE_AGENT if role == "agent" else 0) | \
            C.HK_F_OBS_SORT_LEX | C.HK_F_RESCALE_EPS | (C.HK_F_OBS_RESCALE if scale_observation else 0)
        self._dt = C.HK_DTYPE_I32 if dtype == torch.int32 else C.HK_DTYPE_F32
        from ._lib import check, lib
        self._lib, self._check = lib(), check
        with torch.cuda.device(dev):
            self._launch(torch.cuda.current_stream(dev).cuda_stream)  # warm-up: shared-memory opt-in happens here
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._launch(torch.cuda.current_stream(dev).cuda_stream)

    def _launch(self, stream):
        n, d = self.spec
        p = lambda t: None if t is None else t.data_ptr()
        self._check(self._lib.hk_step(p(self.points), p(self.points), p(self.host_action), p(self.axis), p(self._done),
                                      p(self.reward), None, p(self.obs), p(self.next_coord), None, self.batch_size, n, d,
                                      self._dt, self._ops, self._flags, -1.0, 1e8, stream), "hk_step")

    def __call__(self):
        self.graph.replay()
        return self.points, self.done, self.reward, self.obs


def calculate_value_using_reward_fn(value_prior: torch.Tensor, num_points: torch.Tensor, discount: float,
                                    reward_role: str, est_role: str, use_unified_tree: bool) -> torch.Tensor:
    """Value targets from per-step point counts (util.py:261-284).  The reference takes the reward and
    estimate functions as callables; here they are named by role ("host"/"agent"), which is all the
    reference ever passes (jax_trainer.py:579-580).  value_prior only provides shape and dtype."""
    from ._lib import check, lib
    B, T = value_prior.shape
    npts = num_points.to(torch.int32).contiguous()
    out = torch.empty((B, T), dtype=torch.float32, device=npts.device)
    with torch.cuda.device(npts.device):
        rc = lib().hk_value_targets(None, npts.data_ptr(), None, out.data_ptr(), B, T, 0, 1, 0, float(discount),
                                    1 if est_role == "host" else -1, 1 if reward_role == "host" else -1,
                                    1 if use_unified_tree else 0, torch.cuda.current_stream(npts.device).cuda_stream)
    check(rc, "hk_value_targets")
    return out.to(value_prior.dtype)


def rollout_postprocess(rollouts, role: str, dimension: int, discount: float, use_unified_tree: bool = True):
    """(obs[b,T,input_dim], policy[b,T,A], value[b,T]) -> flattened rollouts with the value prior
    replaced by the ground-truth discounted value (JAXTrainer.rollout_postprocess,
    jax_trainer.py:558-592); point counts are recovered from the observations in the kernel."""
    from ._lib import check, lib
    obs, policy, value = rollouts
    B, T, W = obs.shape
    o = obs.to(torch.float32).contiguous()
    out = torch.empty((B, T), dtype=torch.float32, device=o.device)
    offset = 1 if use_unified_tree or role == "agent" else 0
    reward_role = "agent" if use_unified_tree else role
    with torch.cuda.device(o.device):
        rc = lib().hk_value_targets(o.data_ptr(), None, None, out.data_ptr(), B, T, W, dimension, offset, float(discount),
                                    1 if role == "host" else -1, 1 if reward_role == "host" else -1,
                                    1 if use_unified_tree else 0, torch.cuda.current_stream(o.device).cuda_stream)
    check(rc, "hk_value_targets")
    return obs.reshape(-1, W), policy.reshape(-1, policy.shape[2]), out.reshape(-1).to(value.dtype)


# ---- policy-side glue of util.py (tensor plumbing around the nets; no game arithmetic) ------------------

@functools.lru_cache()
def get_value_est_fn(role: str) -> Callable:
    """Value estimate of an unfinished game from its point count: +-1 / max(num_points, 1)
    (get_value_est_fn, util.py:152-169)."""
    sign = 1 if role == "host" else -1

    def est_fn(last_values: torch.Tensor, num_points: torch.Tensor) -> torch.Tensor:
        return 1.0 / torch.clamp(num_points.to(torch.float32), min=1.0) * sign
    return est_fn


def apply_agent_action_mask(agent_policy: Callable, dimension: int) -> Callable:
    """Masks an agent policy with the host's coordinate choice, read from the last `dimension` entries
    of the observation (apply_agent_action_mask, util.py:287-305): masked-out logits become -inf."""

    def masked_agent_policy(x: torch.Tensor, *args, **kwargs):
        mask = x[..., x.shape[-1] - dimension:] > 0.5
        policy_prior, value_prior = agent_policy(x, *args, **kwargs)
        return torch.where(mask, policy_prior, torch.full_like(policy_prior, float("-inf"))), value_prior

    masked_agent_policy.__name__ = getattr(agent_policy, "__name__", type(agent_policy).__name__)
    return masked_agent_policy


def action_wrapper(policy_value_fn: Callable, dimension: Optional[int] = None) -> Callable:
    """(policy_logits, value) function -> one-hot action function, with the agent's action mask when
    `dimension` is given (action_wrapper, util.py:308-327)."""
    masked_action = policy_value_fn if dimension is None else apply_agent_action_mask(policy_value_fn, dimension)

    def wrapped_action_fn(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        out, _ = masked_action(x, *args, **kwargs)
        return torch.nn.functional.one_hot(out.argmax(dim=-1), out.shape[-1]).to(out.dtype)

    wrapped_action_fn.__name__ = getattr(policy_value_fn, "__name__", type(policy_value_fn).__name__)
    return wrapped_action_fn


def get_dynamic_policy_fn(spec: Tuple[int, int], host_fn: Callable, agent_fn: Callable) -> Callable:
    """One policy function for a unified MC tree (get_dynamic_policy_fn, util.py:218-258): the input is
    always [B, (N+1)*d]; a host state has its last d entries padded with 0, an agent state carries the
    host's coordinate set there.  Like the reference (`lax.cond` on ONE predicate for the whole batch,
    :231-233) the batch is a host batch iff SOME row has an all-zero coordinate block; the agent's
    logits are padded with -inf to the host's 2^d - d - 1 actions (:235-238)."""
    n, d = spec
    extra_action_dim = 2 ** d - 2 * d - 1

    def dynamic_policy_fn(state: torch.Tensor, host_and_agent_args, *args, **kwargs):
        host_args, agent_args = host_and_agent_args
        coord = state[:, n * d: n * d + d]
        use_host = bool(torch.isclose(coord, torch.zeros((), dtype=coord.dtype, device=coord.device)).all(dim=-1).any())
        if use_host:
            return host_fn(state, *host_args, *args, **kwargs)
        policy, value = agent_fn(state, *agent_args, *args, **kwargs)
        pad = torch.full((policy.shape[0], extra_action_dim), float("-inf"), dtype=policy.dtype, device=policy.device)
        return torch.cat([policy, pad], dim=1), value

    return dynamic_policy_fn


def select_sample_after_sim(role: str, rollout, dimension: int, mix_random_terminal_states: bool = True,
                            generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Mask of the rollout states kept for training (select_sample_after_sim, util.py:351-382): every
    state of an unfinished game (more than d, resp. 2d for the agent, non-negative entries), plus —
    with `mix_random_terminal_states` — the states whose value in a uniform random permutation of
    0..size-1 is below the number of unfinished states (:375-377), i.e. about as many random extra states
    as there are unfinished ones.
    RNG contract: the permutation comes from torch (``torch.randperm`` with `generator`), not from
    jax.random's threefry stream, so the random part has the reference's distribution but not its bits;
    the deterministic part (`mix_random_terminal_states=False`) is identical."""
    obs = rollout[0]
    size = obs.shape[0]
    offset = dimension if role == "agent" else 0
    undone = (obs >= 0).sum(dim=-1) > (dimension + offset)
    if not mix_random_terminal_states:
        return undone
    random_idx = torch.randperm(size, generator=generator, device=obs.device if generator is None or
                                generator.device.type != "cpu" else "cpu").to(obs.device)
    return undone | (random_idx < undone.sum())


def rollout_sanity_tests(rollout, spec: Tuple[int, int]) -> bool:
    """The reference's check of a rollout (obs, policy, value) (rollout_sanity_tests, util.py:395-423):
    masked-out actions of agent states must carry -inf logits, and the policy must not already be a
    softmax (rows that sum to 1 with entries in [0, 1])."""
    obs, policy, value = rollout
    n, d = spec
    if obs.shape[-1] == (n + 1) * d:
        mask = obs[..., -d:] > 0.5
        is_host = (~mask).all(dim=-1, keepdim=True)  # host observations are padded with zeros there
        mask = torch.cat([mask | is_host, is_host.expand(*is_host.shape[:-1], policy.shape[-1] - d)], dim=-1)
        if bool((policy[~mask] != float("-inf")).any()):
            return False
    soft = torch.isclose(policy.sum(dim=-1), torch.ones((), dtype=policy.dtype, device=policy.device)).all() and \
        ((policy <= 1.0) & (policy >= 0.0)).all()
    return not bool(soft)
