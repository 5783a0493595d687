"""Vectorised Hironaka environments: B games of the reference's gym environments per call.

``VecHironakaAgentEnv`` / ``VecHironakaHostEnv`` follow ``HironakaAgentEnv`` / ``HironakaHostEnv``
(hironaka/gym_env/hironaka_agent_env.py:13-83, hironaka_host_env.py:11-81, hironaka_base.py:17-140)
rule for rule — rewards, thresholds, invalid-move penalty, the reset quirks — with every game
operation a kernel launch on the whole batch (hk_step with the fixed agent inside, hk_features in
``ListPoints`` order, hk_host_policy).  ``gym`` itself is not a dependency: the classes expose
``reset`` / ``step`` with arrays in place of the reference's per-game objects
(observation [B, N, d] float32 in ListPoints order: rows sorted descending, compacted, padded).

Differences that are deliberate: the state is kept as exact numbers (int32 for the agent
environment, integer-valued float32 in list order for the host environment) and only the
OBSERVATION is rescaled; the reference rescales its float64 state in place every step, which plays
the same game (every rule is invariant under a positive scaling) but makes `value_threshold`
meaningless once `scale_observation` is on.  Here the threshold always applies to the unscaled values.
Players with their own randomness (RandomAgent, RandomHost) draw from a torch generator instead of
Python's `random` / `numpy.random`.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Union

import torch

from . import constants as C
from . import ops as _ops


class _VecHironakaBase:
    def __init__(self, num_envs: int, dimension: int = 3, max_num_points: int = 10, max_value: int = 10,
                 padding_value: float = -1.0, value_threshold: Optional[float] = None, step_threshold: int = 1000,
                 fixed_penalty_crossing_threshold: Optional[float] = None, stop_at_threshold: bool = True,
                 improve_efficiency: bool = False, scale_observation: bool = True,
                 reward_based_on_point_reduction: bool = False, device="cuda",
                 generator: Optional[torch.Generator] = None, **kwargs):
        self.num_envs = num_envs
        self.dimension = dimension
        self.max_num_points = max_num_points
        self.max_value = max_value
        self.padding_value = padding_value
        self.value_threshold = value_threshold
        self.step_threshold = step_threshold
        self.fixed_penalty_crossing_threshold = fixed_penalty_crossing_threshold
        self.stop_at_threshold = stop_at_threshold
        self.improve_efficiency = improve_efficiency
        self.scale_observation = scale_observation
        self.reward_based_on_point_reduction = reward_based_on_point_reduction
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _ops.HironakaB200Error("the vectorised environments have no CPU path: device must be a CUDA device")
        self.generator = generator
        self._state: Optional[torch.Tensor] = None
        self.current_step = torch.zeros(num_envs, dtype=torch.int32, device=self.device)
        self.exceed_threshold = torch.zeros(num_envs, dtype=torch.bool, device=self.device)

    # ---- pieces shared by both environments -----------------------------------------------
    def _initial_points(self, points) -> torch.Tensor:
        B, N, d = self.num_envs, self.max_num_points, self.dimension
        if points is None:  # generate_points: randint(0, max_value) (hironaka/src/_fn.py:184-185)
            pts = torch.randint(0, self.max_value, (B, N, d), generator=self.generator, device=self.device,
                                dtype=torch.int32)
        else:
            pts = torch.as_tensor(points).to(self.device)
            if pts.dim() == 2:
                pts = pts.unsqueeze(0)
            if pts.shape[0] != B or pts.shape[2] != d or pts.shape[1] > N:
                raise ValueError(f"points must be [{B}, <= {N}, {d}]; got {tuple(pts.shape)}")
            if pts.shape[1] < N:  # pad up to max_num_points
                pad = torch.full((B, N - pts.shape[1], d), -1, dtype=pts.dtype, device=self.device)
                pts = torch.cat([pts, pad], dim=1)
        return pts.contiguous()

    def _observe(self, state: torch.Tensor) -> torch.Tensor:
        """Padded points in ListPoints order (sorted descending, compacted), rescaled if asked:
        _get_padded_points / get_features of the reference environments."""
        flags = C.HK_F_OBS_SORT_LEX_FIRST | (C.HK_F_OBS_RESCALE if self.scale_observation else 0)
        obs = _ops.features(state, flags=flags, padding_value=self.padding_value)
        return obs.view(self.num_envs, self.max_num_points, self.dimension)

    def _exceeds(self, state: torch.Tensor) -> torch.Tensor:
        if self.value_threshold is None:
            return torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        return _ops.overflow(state, self.value_threshold, strict=True)  # ListPoints.exceed_threshold: strict

    @property
    def points(self) -> torch.Tensor:
        return self._state


class VecHironakaAgentEnv(_VecHironakaBase):
    """B copies of HironakaAgentEnv: the environment holds a fixed agent and receives the HOST's
    coordinate choice.  agent: "choose_first" (ChooseFirstAgent, hironaka/agent.py:91-96),
    "choose_last", or "random" (RandomAgent: uniform over the chosen coordinates, torch generator).

    step(action): action is [B, d] multi-binary, or [B] integer codes when
    use_discrete_actions_for_host (the binary code of the coordinate set, as decode_action reads it,
    hironaka/src/_fn.py:156-170).  Returns (observation [B,N,d] f32, reward [B] f32, stopped [B] bool, info)."""

    def __init__(self, num_envs: int, agent: str = "choose_first", use_discrete_actions_for_host: bool = False,
                 compressed_host_output: bool = True, **kwargs):
        super().__init__(num_envs, **kwargs)
        if agent not in ("choose_first", "choose_last", "random"):
            raise ValueError(f"unknown fixed agent {agent!r}")
        self.agent = agent
        self.use_discrete_actions_for_host = use_discrete_actions_for_host
        self.compressed_host_output = compressed_host_output

    def reset(self, points=None):
        self._state = self._initial_points(points).to(torch.int32)
        _ops.step(self._state, ops=C.HK_OP_NEWTON, inplace=True)  # reset(): get_newton_polytope (hironaka_base.py:97)
        self.current_step.zero_()
        self.exceed_threshold.zero_()
        self.last_action_taken = None
        return self._observe(self._state)

    def _coord_mask(self, action) -> torch.Tensor:
        a = torch.as_tensor(action, device=self.device)
        if self.use_discrete_actions_for_host:
            if a.shape != (self.num_envs,):
                raise ValueError(f"discrete host actions must have shape ({self.num_envs},)")
            return a.to(torch.int32).contiguous()  # the code IS the bitmask (bit k <=> coordinate k)
        return _ops.coords_to_mask(a, self.dimension, self.device)

    def step(self, action):
        B = self.num_envs
        self.current_step += 1
        before = _ops.dones(self._state, want_num_points=True)[1]
        mask = self._coord_mask(action)
        flags = C.HK_F_NOOP_INVALID  # shift_lst: axis not in coords -> nothing moves; ended games are NOT frozen
        axis = None
        if self.agent == "choose_first":
            flags |= C.HK_F_AGENT_FIRST
        elif self.agent == "choose_last":
            flags |= C.HK_F_AGENT_LAST
        else:  # uniform over the chosen coordinates; fewer than two chosen -> no move (agent.py:83-88)
            bits = ((mask.unsqueeze(1) >> torch.arange(self.dimension, device=self.device)) & 1).float()
            r = torch.rand((B, self.dimension), generator=self.generator, device=self.device) + 1.0
            axis = (r * bits).argmax(1).to(torch.int32)
            axis = torch.where(bits.sum(1) > 1, axis, torch.full_like(axis, -1))
        r = _ops.step(self._state, mask, axis, ops=C.HK_OP_SHIFT | C.HK_OP_NEWTON, flags=flags, inplace=True,
                      want_done=True, want_num_points=True)
        ended = r.done
        after = r.num_points
        reward = torch.zeros(B, dtype=torch.float32, device=self.device)
        stopped = ended.clone()
        self.exceed_threshold = self._exceeds(self._state)
        if self.stop_at_threshold:
            hit = (self.current_step >= self.step_threshold) | self.exceed_threshold
            stopped |= hit
            penalty = float(-self.step_threshold if self.fixed_penalty_crossing_threshold is None
                            else self.fixed_penalty_crossing_threshold)
            reward += hit.float() * penalty
        if self.reward_based_on_point_reduction:
            reward += (before - after).float()
        reward += ended.float()
        return self._observe(self._state), reward, stopped, {}


class VecHironakaHostEnv(_VecHironakaBase):
    """B copies of HironakaHostEnv: the environment holds a fixed host ("zeillinger" | "all_coord") and
    receives the AGENT's axis.  Observations are {"points": [B,N,d] f32, "coords": [B,d] int8 multi-binary}.
    The state is kept in ListPoints order, which is the order hironaka/host.py scans pairs in."""

    def __init__(self, num_envs: int, host: str = "zeillinger", invalid_move_penalty: float = -1e-3,
                 stop_after_invalid_move: bool = False, **kwargs):
        super().__init__(num_envs, **kwargs)
        if host not in ("zeillinger", "all_coord"):
            raise ValueError(f"unknown fixed host {host!r}")
        self.host = host
        self.invalid_move_penalty = invalid_move_penalty
        self.stop_after_invalid_move = stop_after_invalid_move
        self._coords = torch.zeros(num_envs, dtype=torch.int32, device=self.device)

    def _list_order(self, state: torch.Tensor) -> torch.Tensor:
        obs = _ops.features(state, flags=C.HK_F_OBS_SORT_LEX_FIRST, padding_value=-1.0)
        return obs.view(self.num_envs, self.max_num_points, self.dimension)

    def _coords_multi_bin(self) -> torch.Tensor:
        return ((self._coords.unsqueeze(1) >> torch.arange(self.dimension, device=self.device)) & 1).to(torch.int8)

    def _obs(self) -> Dict[str, torch.Tensor]:
        return {"points": self._observe(self._state), "coords": self._coords_multi_bin()}

    def reset(self, points=None):
        st = self._initial_points(points).to(torch.float32)
        _ops.step(st, ops=C.HK_OP_NEWTON, inplace=True)
        self._state = self._list_order(st)
        self.current_step.zero_()
        self.exceed_threshold.zero_()
        self._coords.zero_()
        # _post_reset_update: the reference takes a step with action None to get the host's first choice
        # (hironaka_host_env.py:38-39); its reward is dropped but its side effects are not: the step counter
        # starts at 1 and, with stop_after_invalid_move, the environment comes out of reset already stopped.
        self.step(None)
        return self._obs()

    def step(self, action):
        B = self.num_envs
        self.current_step += 1
        if action is None:
            valid = torch.zeros(B, dtype=torch.bool, device=self.device)
            axis = torch.full((B,), -1, dtype=torch.int32, device=self.device)
        else:
            axis = torch.as_tensor(action, device=self.device).to(torch.int32).reshape(B)
            valid = ((self._coords >> axis.clamp(0, 31)) & 1).bool() & (axis >= 0) & (axis < self.dimension)
        # an invalid move leaves the points alone (the filter is idempotent on a filtered state)
        r = _ops.step(self._state, self._coords, axis, ops=C.HK_OP_SHIFT | C.HK_OP_NEWTON, flags=C.HK_F_NOOP_INVALID,
                      inplace=True, want_done=True)
        self._state.copy_(self._list_order(self._state))  # (in place: the buffer stays put, so a step can be captured)
        ended = r.done
        reward = torch.where(valid, (~ended).float(), torch.full((B,), float(self.invalid_move_penalty),
                                                                  device=self.device))
        stopped = ended.clone()
        if self.stop_after_invalid_move:
            stopped |= ~valid
        self.exceed_threshold.copy_(self._exceeds(self._state))
        stopped |= self.exceed_threshold
        choice = _ops.host_policy(self._state, self.host)
        self._coords.copy_(torch.where(stopped, torch.zeros_like(choice), choice))
        self.last_action_taken = self._coords
        return self._obs(), reward, stopped, {}

    def capture_step(self, warmup: int = 2) -> "GraphedHostEnvStep":
        """One `step` of all B environments as a single CUDA-graph replay (the eager step is a dozen small
        launches: the fused move, the ListPoints re-sort, the overflow flags, the host's next choice, the
        observation, and the reward / stop bookkeeping).  Call after `reset`; see GraphedHostEnvStep."""
        return GraphedHostEnvStep(self, warmup)


class GraphedHostEnvStep:
    """`VecHironakaHostEnv.step` captured once: write the agents' axes into `action` (int32 [B]), call the
    object, read `points` [B,N,d] f32, `coords` [B,d] int8, `reward` [B] f32, `stopped` [B] bool (static
    tensors, overwritten by every replay).  The environment object stays in sync (its state, coordinates,
    step counter and overflow flags are the buffers the graph updates)."""

    def __init__(self, env: VecHironakaHostEnv, warmup: int = 2):
        self.env = env
        dev = env.device
        self.action = torch.zeros(env.num_envs, dtype=torch.int32, device=dev)
        saved = (env._state.clone(), env._coords.clone(), env.current_step.clone(), env.exceed_threshold.clone())
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                env.step(self.action)
            self._restore(saved)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            obs, self.reward, self.stopped, _ = env.step(self.action)
            self.points, self.coords = obs["points"], obs["coords"]
        self._restore(saved)  # the capture itself does not run the step

    def _restore(self, saved):
        env = self.env
        env._state.copy_(saved[0])
        env._coords.copy_(saved[1])
        env.current_step.copy_(saved[2])
        env.exceed_threshold.copy_(saved[3])

    def __call__(self, action=None):
        if action is not None:
            self.action.copy_(torch.as_tensor(action, device=self.env.device).to(torch.int32).reshape(-1))
        self.graph.replay()
        return {"points": self.points, "coords": self.coords}, self.reward, self.stopped, {}
