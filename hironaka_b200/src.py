"""Drop-in for the torch op surface of ``hironaka.src`` (hironaka/src/_torch_ops.py,
hironaka/src/_fn.py), backed by the sm_100a kernels.

Same names, argument meaning and in-place convention as the reference:
``op(points, ..., inplace=True, padding_value=-1.0)`` mutates ``points`` and returns ``None``,
or returns a new tensor when ``inplace=False``.  ``points`` must be a CUDA tensor of dtype
float32 (the reference's storage) or int32 (the engine's native storage).

Differences, all deliberate and documented in DESIGN.md:
  * no CPU path (a CPU tensor raises);
  * the well-formedness ``assert`` of shift_torch (_torch_ops.py:80) is not evaluated — it is a
    device->host sync in the reference; rows are live iff coordinate 0 is >= 0;
  * no ``[B,N,N,d]`` temporaries, no host syncs, one kernel launch per call.
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
import torch

from . import constants as C
from . import ops as _ops
from .host_action import HostActionEncoder  # noqa: F401  (re-exported like hironaka.src)


def batched_coord_list_to_binary(f: List[List[int]], dimension: int) -> np.ndarray:
    """hironaka/src/_fn.py:99-106."""
    out = np.zeros((len(f), dimension))
    for b, c in enumerate(f):
        out[b][list(c)] = 1
    return out


def _run(points: torch.Tensor, op_bits: int, inplace: bool, padding_value: float, host_action=None, axis=None,
         flags: int = 0) -> Optional[torch.Tensor]:
    assert len(points.shape) == 3
    if points.is_contiguous():
        r = _ops.step(points, host_action, axis, ops=op_bits, flags=flags, padding_value=padding_value,
                      inplace=inplace)
        return None if inplace else r.state
    work = points.contiguous()
    _ops.step(work, host_action, axis, ops=op_bits, flags=flags, padding_value=padding_value, inplace=True)
    if inplace:
        points.copy_(work)
        return None
    return work


def remove_repeated(points: torch.Tensor, padding_value: Optional[float] = -1.0):
    """Later copies of identical rows become padding; always in place (_fn.py:192-213)."""
    _run(points, C.HK_OP_DEDUPE, True, padding_value)
    return None


def get_newton_polytope_approx_torch(points: torch.Tensor, inplace: Optional[bool] = True,
                                     padding_value: Optional[float] = -1.0):
    """Dedupe + dominance filter, slots preserved (_torch_ops.py:8-39)."""
    return _run(points, C.HK_OP_NEWTON, inplace, padding_value)


def get_newton_polytope_torch(points: torch.Tensor, inplace: Optional[bool] = True,
                              padding_value: Optional[float] = -1.0):
    return get_newton_polytope_approx_torch(points, inplace=inplace, padding_value=padding_value)


def shift_torch(points: torch.Tensor, coord: Union[torch.Tensor, List[List[int]]],
                axis: Union[torch.Tensor, List[int]], inplace: Optional[bool] = True,
                padding_value: Optional[float] = -1.0, ignore_ended_games: Optional[bool] = True):
    """x_axis <- sum of the chosen coordinates (_torch_ops.py:46-110).

    `coord` is a list of index lists or a [B, d] multi-binary tensor (not the same encoding, see
    the reference's note); `axis` a list or [B] tensor (float accepted, as FusedGame passes it).
    Invalid actions are no-ops; so are ended games when `ignore_ended_games`."""
    assert len(points.shape) == 3
    batch_size, _, dimension = points.shape
    if not isinstance(coord, (list, torch.Tensor)):
        raise Exception(f"unsupported input type for coord. Got {type(coord)}.")
    if not isinstance(axis, (list, torch.Tensor)):
        raise Exception(f"unsupported input type for axis. Got {type(axis)},")
    mask = _ops.coords_to_mask(coord, dimension, points.device)
    assert mask.shape == (batch_size,)
    if isinstance(axis, list):
        axis = torch.tensor(axis, device=points.device)
    assert axis.shape == (batch_size,)
    flags = C.HK_F_NOOP_INVALID | (C.HK_F_FREEZE_ENDED if ignore_ended_games else 0)
    return _run(points, C.HK_OP_SHIFT, inplace, padding_value, host_action=mask, axis=axis, flags=flags)


def reposition_torch(points: torch.Tensor, inplace: Optional[bool] = True, padding_value: Optional[float] = -1.0):
    """Per game and coordinate subtract the min over live rows (_torch_ops.py:113-133)."""
    return _run(points, C.HK_OP_REPOSITION, inplace, padding_value)


def rescale_torch(points: torch.Tensor, inplace: Optional[bool] = True, padding_value: Optional[float] = -1.0):
    """Divide live entries by the game max, IEEE float32 (_torch_ops.py:136-146).  float32 only."""
    return _run(points, C.HK_OP_RESCALE, inplace, padding_value)


__all__ = [
    "batched_coord_list_to_binary", "remove_repeated", "get_newton_polytope_approx_torch",
    "get_newton_polytope_torch", "shift_torch", "reposition_torch", "rescale_torch", "HostActionEncoder",
]
