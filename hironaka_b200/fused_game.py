"""``FusedGame`` on B200: drop-in for ``hironaka.trainer.fused_game.FusedGame``
(hironaka/trainer/fused_game.py:11-182).

Same constructor and ``step / host_move / agent_move`` signatures and the same experience
tuple ``(obs, actions[:,None] int32, rewards[:,None] f32, dones[:,None] bool, next_obs)`` filtered
by ``~done`` with order preserved (dtypes pinned by test/testTrainer.py:105-118).  The only
change is below the surface: the reference's three point operations per move
(``shift -> get_newton_polytope -> rescale``, :150-162, ~45 ATen launches and a dozen
``[B,N,N,d]`` temporaries) are ONE launch of the fused step kernel.

The policy networks are the caller's (torch modules); only the env step is this package's.

Interface glue (constructor arguments, the order in which the torch RNG is consumed by exploration, the experience
tuple) is adapted from the reference file cited above so that a trainer sees identical behaviour; the point
operations, ``step_into`` and ``graphed_step_into`` are new.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Callable, Optional, Tuple, Type, Union

import torch

from . import constants as C
from . import ops as _ops
from .host_action import HostActionEncoder
from .tensor_points import TensorPoints


class FusedGame:
    def __init__(self, host_net: torch.nn.Module, agent_net: torch.nn.Module,
                 device: Optional[Union[str, torch.device]] = "cuda", log_time: Optional[bool] = True,
                 reward_func: Optional[Callable] = None, dtype: Optional[Union[Type, torch.dtype]] = torch.float32):
        self.device = torch.device(device)
        self.host_net = host_net.to(self.device)
        self.agent_net = agent_net.to(self.device)
        self.log_time = log_time
        if reward_func is None:
            self._rewards = self._default_reward
        else:
            assert isinstance(reward_func, Callable), f"reward_function must be callable. Got {type(reward_func)}."
            self._rewards = reward_func
        self.dtype = dtype
        self._make_type_for_nets(self.dtype)
        self.host_action_encoder = None
        self.time_log = dict()

    def step(self, points: TensorPoints, sample_for: str, masked=True, scale_observation=True, exploration_rate=0.2):
        """Progress the game one move and return
        (observations, actions, rewards, dones, next_observations) for the games that were not
        already over (fused_game.py:54-102)."""
        assert sample_for in ["host", "agent"], f"sample_for must be one of 'host' and 'agent'. Got {sample_for}."
        if points.dtype != self.dtype:
            points.type(self.dtype)
        observations = points.get_features()
        done = points.ended_batch_in_tensor
        e_r = exploration_rate if sample_for == "host" else 0.0
        host_move, chosen_actions = self.host_move(points, exploration_rate=e_r)
        e_r = exploration_rate if sample_for == "agent" else 0.0
        agent_move = self.agent_move(points, host_move, masked=masked, scale_observation=scale_observation,
                                     inplace=True, exploration_rate=e_r)
        next_done = points.ended_batch_in_tensor
        next_observations = points.get_features()
        keep = ~done
        if sample_for == "host":
            output_obs = observations[keep].clone()
            output_actions = chosen_actions[keep].clone()
            next_observations = next_observations[keep].clone()
        else:
            next_host_move, _ = self.host_move(points, exploration_rate=exploration_rate)
            output_obs = {"points": observations[keep].clone(), "coords": host_move[keep].clone()}
            output_actions = agent_move[keep].clone()
            next_observations = {"points": next_observations[keep].clone(), "coords": next_host_move[keep].clone()}
        next_done = next_done[keep].clone()
        return (output_obs, output_actions.reshape(-1, 1),
                self._rewards(sample_for, output_obs, next_observations, next_done).reshape(-1, 1),
                next_done.reshape(-1, 1), next_observations)

    def step_into(self, buffer, points: TensorPoints, sample_for: str, masked=True, scale_observation=True,
                  exploration_rate=0.2) -> None:
        """``step`` followed by ``ReplayBuffer.add`` with NO host round-trip: the `[~done]` filter of
        the reference (a sync per masked tensor, fused_game.py:82-94) becomes an order-preserving
        device-side compaction straight into the circular buffer (ReplayBuffer.add_masked)."""
        assert sample_for in ["host", "agent"]
        if points.dtype != self.dtype:
            points.type(self.dtype)
        observations = points.get_features()
        done = points.ended_batch_in_tensor
        host_move, chosen_actions = self.host_move(points, exploration_rate=exploration_rate if sample_for == "host" else 0.0)
        agent_move = self.agent_move(points, host_move, masked=masked, scale_observation=scale_observation, inplace=True,
                                     exploration_rate=exploration_rate if sample_for == "agent" else 0.0)
        next_done = points.ended_batch_in_tensor
        next_observations = points.get_features()
        # the reward function sees the UNFILTERED per-game tensors (a custom `reward_func` must be row-wise,
        # as the default is: fused_game.py:175-182); rows of games that were over are dropped by the writer
        if sample_for == "host":
            reward = self._rewards(sample_for, observations, next_observations, next_done)
            buffer.add_masked(done, observations, chosen_actions, reward, next_done, next_observations)
        else:
            next_host_move, _ = self.host_move(points, exploration_rate=exploration_rate)
            obs_d = {"points": observations, "coords": host_move}
            next_d = {"points": next_observations, "coords": next_host_move}
            reward = self._rewards(sample_for, obs_d, next_d, next_done)
            buffer.add_masked(done, obs_d, agent_move, reward, next_done, next_d)

    def graphed_step_into(self, buffer, points: TensorPoints, sample_for: str, masked=True, scale_observation=True,
                          exploration_rate=0.2, warmup: int = 2) -> "torch.cuda.CUDAGraph":
        """``step_into`` captured ONCE in a CUDA graph; ``graph.replay()`` then plays one move of every
        game and appends the experiences, with no Python between the launches.  This is the form for
        replay-buffer generation at DQN batch sizes (a few thousand games), where the step is a few
        dozen small launches and the eager version is bound by Python and launch latency
        (BASELINE config 4).  Legal because no entry point of the library allocates or synchronises and
        the buffer's write position lives on the device; the nets' inference and the exploration
        noise (torch RNG, graph-safe) are captured with it.  `points` and `buffer` must stay the same
        objects (the graph holds their storage); the warm-up moves are undone before the capture."""
        dev = points.points.device
        saved = points.points.clone()
        counters = [t.clone() for t in (buffer._pos, buffer._full, buffer._appended)]
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # allocator, shared-memory opt-in, lazy module state
                self.step_into(buffer, points, sample_for, masked, scale_observation, exploration_rate)
            points.points.copy_(saved)
            for t, c in zip((buffer._pos, buffer._full, buffer._appended), counters):
                t.copy_(c)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.step_into(buffer, points, sample_for, masked, scale_observation, exploration_rate)
        return graph

    def host_move(self, points: TensorPoints, masked=True, exploration_rate=0.0) -> Tuple[torch.Tensor, torch.Tensor]:
        """Host net -> argmax (or noise) -> multi-binary host move (fused_game.py:104-122)."""
        with torch.inference_mode():
            output = self.host_net(points.get_features().to(self.device))
        if self.host_action_encoder is None:
            self.host_action_encoder = HostActionEncoder(points.dimension)
        noise = torch.rand(output.shape, device=self.device, dtype=output.dtype)
        random_mask = torch.rand(output.shape[0], 1, device=self.device).le(exploration_rate)
        output = output * ~random_mask + noise * random_mask
        chosen_actions = torch.argmax(output, dim=1).type(torch.int32)
        host_move_binary = self.host_action_encoder.decode_tensor(chosen_actions, dtype=self.dtype)
        return host_move_binary, chosen_actions

    def agent_move(self, points: TensorPoints, host_moves: torch.Tensor, masked: Optional[bool] = True,
                   scale_observation: Optional[bool] = True, inplace: Optional[bool] = True,
                   exploration_rate: Optional[float] = 0.0) -> torch.Tensor:
        """Agent net -> (masked) argmax (or noise) -> ONE fused launch of shift -> newton -> (rescale)
        on the points (fused_game.py:124-163)."""
        with torch.inference_mode():
            action_prob = self.agent_net({"points": points.get_features().to(self.device),
                                          "coords": host_moves.to(self.device)})
        if masked:
            minimum = torch.finfo(action_prob.dtype).min
            action_prob = action_prob * host_moves + (1 - host_moves) * minimum
        actions = torch.argmax(action_prob, dim=1)
        noise = torch.randint(0, action_prob.shape[1], actions.shape, device=actions.device, dtype=actions.dtype)
        random_mask = torch.rand(actions.shape[0], device=actions.device).le(exploration_rate)
        actions = actions * ~random_mask + noise * random_mask
        if inplace:
            state = points.points
            if state.dtype not in (torch.float32, torch.int32):
                raise _ops.HironakaB200Error(f"FusedGame point operations need float32 or int32 points; got {state.dtype}")
            op_bits = C.HK_OP_SHIFT | C.HK_OP_NEWTON | (C.HK_OP_RESCALE if scale_observation else 0)
            mask = _ops.coords_to_mask(host_moves, points.dimension, state.device)
            _ops.step(state, mask, actions, ops=op_bits, flags=C.TORCH_SEMANTICS, padding_value=points.padding_value,
                      inplace=True)
        return actions

    def _make_type_for_nets(self, dtype: torch.dtype):
        for role in ["host", "agent"]:
            net = getattr(self, f"{role}_net")
            param = next(net.parameters(), None)
            if param is not None and param.dtype != dtype:
                setattr(self, f"{role}_net", deepcopy(net).type(dtype))

    @staticmethod
    def _default_reward(sample_for: str, obs, next_obs, next_done: torch.Tensor) -> torch.Tensor:
        """host: +1 when the move ends the game; agent: -1 (fused_game.py:175-182)."""
        if sample_for == "host":
            return next_done.type(torch.float32).clone()
        elif sample_for == "agent":
            return (-next_done.type(torch.float32)).clone()
