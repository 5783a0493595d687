"""Host action encoding: discrete id 0..2^d-d-2  <->  coordinate subsets with >= 2 elements.

Mirrors ``HostActionEncoder`` (hironaka/src/_fn.py:241-325) and the JAX table helpers
(hironaka/jax/host_action_preprocess.py:8-103).  The id-th action is the id-th integer in
3..2^d-1 that is not a power of two, read as a bitmask (bit k <=> coordinate k); the kernels
evaluate the same bijection in closed form (csrc/hk_common.cuh: decode_host_action), so no table
ever travels to the device on the step path.  These helpers exist for callers that want the
multi-binary form on the host side of the API.
"""
from __future__ import annotations

import functools
import math
from typing import Callable, List

import numpy as np
import torch

_MAX_DIM = 11  # hironaka/jax/host_action_preprocess.py:27


def action_masks(dimension: int) -> List[int]:
    """Bitmasks of the legal host actions in id order."""
    return [m for m in range(1, 2 ** dimension) if m & (m - 1)]


def decode_id(action: int) -> int:
    """Closed form used by the kernels: mask = t + floor(log2(t + floor(log2 t))), t = id + 2."""
    t = action + 2
    return t + int(math.floor(math.log2(t + int(math.floor(math.log2(t))))))


def encode_mask(mask: int) -> int:
    """id = mask - floor(log2 mask) - 2  (_fn.py:285-296)."""
    return mask - (mask.bit_length() - 1) - 2


@functools.lru_cache()
def _table_np(dimension: int) -> np.ndarray:
    masks = action_masks(dimension)
    return np.array([[(m >> k) & 1 for k in range(dimension)] for m in masks], dtype=np.int32).reshape(-1, dimension)


def decode_table(dimension: int, device=None) -> torch.Tensor:
    """[2^d-d-1, d] int32 multi-binary table (host_action_preprocess.py:8-24)."""
    return torch.as_tensor(_table_np(dimension), device=device)


class HostActionEncoder:
    """Same surface as the reference class: encode / encode_tensor / decode / decode_tensor."""

    def __init__(self, dim: int = 3):
        self.dim = dim
        self.action_translate = action_masks(dim)
        self.binary_table = _table_np(dim).astype(np.float64)
        self.cached_binary_tables = {}

    def encode(self, coords: List[int]) -> int:
        assert len(coords) > 1
        mask = 0
        for c in coords:
            mask += 2 ** c
        return int(encode_mask(mask))

    def encode_tensor(self, coords: torch.Tensor) -> torch.Tensor:
        assert len(coords.shape) == 2
        weights = 2 ** torch.arange(self.dim, device=coords.device)
        masks = torch.sum(weights * coords.type(torch.int32), dim=1)
        return masks - torch.log2(masks).type(torch.int32) - 2

    def decode(self, action: int) -> List[int]:
        assert (action < 2 ** self.dim - self.dim - 1) and (action >= 0)
        mask = self.action_translate[action]
        return [k for k in range(self.dim) if (mask >> k) & 1]

    def decode_tensor(self, actions: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        assert len(actions.shape) == 1
        key = (str(actions.device), dtype)
        if key not in self.cached_binary_tables:
            self.cached_binary_tables[key] = torch.tensor(self.binary_table, device=actions.device, dtype=dtype)
        return self.cached_binary_tables[key][actions.long()]


@functools.lru_cache()
def get_batch_decode(dimension: int) -> Callable:
    """ids [B] -> multi-binary [B, d] (host_action_preprocess.py:58-65)."""
    if dimension >= _MAX_DIM:
        raise ValueError(f"Dimension is capped at {_MAX_DIM}. Got {dimension}.")
    cache = {}

    def batch_decode(ids: torch.Tensor) -> torch.Tensor:
        tab = cache.get(ids.device)
        if tab is None:
            tab = cache[ids.device] = decode_table(dimension, ids.device)
        return tab[ids.long()]

    return batch_decode


@functools.lru_cache()
def get_batch_decode_from_one_hot(dimension: int) -> Callable:
    """one-hot [B, 2^d-d-1] -> multi-binary [B, d] (host_action_preprocess.py:68-75)."""
    dec = get_batch_decode(dimension)

    def batch_decode_from_one_hot(one_hot: torch.Tensor) -> torch.Tensor:
        return dec(torch.argmax(one_hot, dim=-1))

    return batch_decode_from_one_hot


def batch_encode(multi_binary: torch.Tensor) -> torch.Tensor:
    """multi-binary [B, d] -> ids [B] (host_action_preprocess.py:78-87,102)."""
    d = multi_binary.shape[-1]
    masks = (multi_binary.to(torch.int64) * (2 ** torch.arange(d, device=multi_binary.device))).sum(-1)
    return masks - torch.floor(torch.log2(masks.to(torch.float64))).to(torch.int64) - 2


def batch_encode_one_hot(multi_binary: torch.Tensor) -> torch.Tensor:
    """multi-binary [B, d] -> one-hot float32 [B, 2^d-d-1] (host_action_preprocess.py:90-103)."""
    d = multi_binary.shape[-1]
    cls = 2 ** d - d - 1
    ids = batch_encode(multi_binary)
    return (torch.arange(cls, device=multi_binary.device)[None, :] == ids[:, None]).to(torch.float32)
