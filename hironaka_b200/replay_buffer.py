"""Circular tensor replay buffer with a device-resident write position: drop-in for
``hironaka.trainer.replay_buffer.ReplayBuffer`` (hironaka/trainer/replay_buffer.py:6-176) plus
``add_masked``, the form that needs no host round-trip.

``add(obs, action, reward, done, next_obs)`` takes already-filtered experiences exactly like the
reference (FusedGame.step output).  ``add_masked(skip, ...)`` takes UNFILTERED per-game tensors
and a ``skip`` mask (the games that were over before the move): the kept rows are compacted in
order and written with wrap-around by one stream-ordered call of ``hk_experience_append``; the
host never learns how many rows were kept (``pos``/``full`` are device tensors; ``pos_host()``
reads them when a caller does want the number).
"""
from __future__ import annotations

from typing import Dict, Tuple, Type, Union

import torch

from ._lib import check, lib


def _ptr(t):
    return None if t is None else t.data_ptr()


class ReplayBuffer:
    def __init__(self, input_shape: Union[Dict, Tuple], output_dim: int, buffer_size: int, device: torch.device,
                 dtype: Union[Type, torch.dtype] = torch.float32, **kwargs):
        self.input_shape = input_shape
        self.output_dim = output_dim
        self.buffer_size = buffer_size
        self.device = torch.device(device)
        self.dtype = dtype
        if self.device.type != "cuda":
            raise RuntimeError("hironaka_b200.ReplayBuffer lives on a CUDA device (no CPU path)")
        if dtype != torch.float32:
            raise TypeError("observations are stored as float32 (the feature dtype of the engine)")

        def alloc(shape):
            return torch.zeros((buffer_size, *shape), device=self.device, dtype=self.dtype)

        if isinstance(input_shape, dict):
            self.observations = {k: alloc(v) for k, v in input_shape.items()}
            self.next_observations = {k: alloc(v) for k, v in input_shape.items()}
        else:
            self.observations = alloc(input_shape)
            self.next_observations = alloc(input_shape)
        self.actions = torch.zeros((buffer_size, 1), device=self.device, dtype=torch.int32)
        self.rewards = torch.zeros((buffer_size, 1), device=self.device, dtype=torch.float32)
        self.dones = torch.zeros((buffer_size, 1), device=self.device, dtype=torch.bool)
        self._pos = torch.zeros(1, device=self.device, dtype=torch.int64)
        self._full = torch.zeros(1, device=self.device, dtype=torch.int32)
        self._appended = torch.zeros(1, device=self.device, dtype=torch.int32)
        self._scratch = None

    # ---- the reference's host-visible counters (a device->host read each) --------------------
    @property
    def pos(self) -> int:
        return int(self._pos.item())

    @property
    def full(self) -> bool:
        return bool(self._full.item())

    def reset(self):
        self._pos.zero_()
        self._full.zero_()

    # ---- writes ---------------------------------------------------------------------------------
    def add_masked(self, skip: torch.Tensor, obs, action: torch.Tensor, reward: torch.Tensor, done: torch.Tensor,
                   next_obs) -> None:
        """Append the rows with skip == False, in order; no host synchronisation.
        obs / next_obs: [B, ...] float32 tensor, or dict {"points": [B, N, d], "coords": [B, d]};
        action [B] or [B,1] int32, reward [B] or [B,1] float32, done [B] or [B,1] bool."""
        B = skip.shape[0]
        # B is a host-known upper bound of the rows kept: the reference asserts buffer_size > length
        # (replay_buffer.py:75); more kept rows than slots would make several rows race for one slot
        if B >= self.buffer_size:
            raise ValueError(f"{B} samples are more than the buffer size ({self.buffer_size}).")
        skip_u8 = skip.contiguous().view(torch.uint8) if skip.dtype == torch.bool else skip.to(torch.uint8).contiguous()
        act = action.reshape(B).to(torch.int32).contiguous()
        rew = reward.reshape(B).to(torch.float32).contiguous()
        dn = done.reshape(B)
        dn = dn.contiguous().view(torch.uint8) if dn.dtype == torch.bool else dn.to(torch.uint8).contiguous()
        if isinstance(self.observations, dict):
            o, no = obs["points"], next_obs["points"]
            c, nc = obs["coords"].to(torch.float32).contiguous(), next_obs["coords"].to(torch.float32).contiguous()
            bo, bno = self.observations["points"], self.next_observations["points"]
            bc, bnc = self.observations["coords"], self.next_observations["coords"]
            cw = c.reshape(B, -1).shape[1]
        else:
            o, no, c, nc, bc, bnc, cw = obs, next_obs, None, None, None, None, 0
            bo, bno = self.observations, self.next_observations
        o = o.reshape(B, -1).to(torch.float32).contiguous()
        no = no.reshape(B, -1).to(torch.float32).contiguous()
        ow = o.shape[1]
        assert bo[0].numel() == ow, f"observation width {ow} does not match the buffer ({bo[0].numel()})"
        need = lib().hk_experience_scratch_words(B)
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, device=self.device, dtype=torch.int32)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = lib().hk_experience_append(
                _ptr(skip_u8), _ptr(o), _ptr(no), ow, _ptr(c), _ptr(nc), cw, _ptr(act), _ptr(rew), _ptr(dn), _ptr(bo),
                _ptr(bno), _ptr(bc), _ptr(bnc), _ptr(self.actions), _ptr(self.rewards), _ptr(self.dones.view(torch.uint8)),
                self.buffer_size, _ptr(self._pos), _ptr(self._full), _ptr(self._appended), _ptr(self._scratch), B, stream)
        check(rc, "hk_experience_append")

    def add(self, obs, action: torch.Tensor, reward: torch.Tensor, done: torch.Tensor, next_obs, clone=True):
        """The reference's signature: experiences that are already filtered (every row is kept)."""
        assert action.shape[1:] == torch.Size([1]) and reward.shape[1:] == torch.Size([1]) and done.shape[1:] == torch.Size([1])
        length = action.shape[0]
        assert self.buffer_size > length, f"{length} samples are more than the buffer size."
        if length == 0:
            return
        keep_all = torch.zeros(length, device=self.device, dtype=torch.uint8)
        self.add_masked(keep_all, obs, action, reward, done, next_obs)

    # ---- reads ----------------------------------------------------------------------------------
    def sample(self, batch_size: int, device: torch.device = None, clone: bool = True, check_empty: bool = True) -> Tuple:
        """Uniform sample over the filled part.  Like the reference (replay_buffer.py:139) an empty buffer is
        an error; that check reads the write position back (one device->host sync).  Pass
        `check_empty=False` inside a CUDA graph or a sync-free loop: the bound is then used on the device
        only, and an empty buffer yields row 0 (zeros)."""
        if check_empty and not self.full and self.pos == 0:
            raise AssertionError("cannot sample from an empty replay buffer")
        bound = torch.where(self._full > 0, torch.full_like(self._pos, self.buffer_size), self._pos)
        u = torch.rand(batch_size, device=self.device, dtype=torch.float64)
        idx = torch.clamp((u * bound.to(torch.float64)).to(torch.int64), max=self.buffer_size - 1)

        def take(t):
            r = t[idx]
            return r.to(device) if device is not None else r

        out = []
        for data in (self.observations, self.actions, self.rewards, self.dones, self.next_observations):
            out.append({k: take(v) for k, v in data.items()} if isinstance(data, dict) else take(data))
        return tuple(out)
