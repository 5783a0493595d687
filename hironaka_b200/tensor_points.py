"""``TensorPoints`` on B200: the ``point_cls=`` drop-in for ``hironaka.core.TensorPoints``
(hironaka/core/tensor_points.py:11-126).

``DQNTrainer(..., point_cls=hironaka_b200.TensorPoints)``, ``FusedGame.step`` and
``Trainer.get_rollout`` use: ``.points .dtype .device .dimension .type() .get_features()
.ended .ended_batch_in_tensor .shift(coords, axis) .get_newton_polytope() .rescale()
.reposition() .get_num_points() .copy()`` — all present with the reference's meaning.  The
point tensor stays a float32 (or int32) CUDA tensor owned by the object and mutated in place;
every operation is one kernel launch through the C-ABI.

Constructor argument handling is adapted from the reference file cited above (same names, defaults and error
behaviour); everything behavioural dispatches to the CUDA kernels.
"""
from __future__ import annotations

from typing import List, Optional, Type, Union

import numpy as np
import torch

from . import constants as C
from . import ops as _ops
from . import src as _src
from .points_base import PointsBase


def _padded(points: List[List[List[float]]], new_length: int, constant_value: float) -> np.ndarray:
    """Nested list with ragged point axis -> [B, new_length, d] (get_batched_padded_array, _fn.py:78-87)."""
    out = []
    for game in points:
        g = np.array(game, dtype=float).reshape(len(game), -1)
        out.append(np.pad(g, ((0, new_length - g.shape[0]), (0, 0)), mode="constant", constant_values=constant_value))
    return np.stack(out, axis=0)


class TensorPoints(PointsBase):
    subcls_config_keys = ["value_threshold", "device", "padding_value", "dtype"]
    running_attributes = ["distinguished_points"]

    def __init__(self, points: Union[torch.Tensor, List[List[List[float]]], np.ndarray],
                 value_threshold: Optional[float] = 1e8, device: Optional[Union[str, torch.device]] = "cuda",
                 padding_value: Optional[float] = -1.0, distinguished_points: Optional[List[int]] = None,
                 dtype: Optional[Union[Type, torch.dtype]] = torch.float32, **kwargs):
        self.value_threshold = value_threshold
        assert padding_value <= 0.0, f"'padding_value' must be a non-positive number. Got {padding_value} instead."
        self.dtype = dtype
        self.device = torch.device(device) if isinstance(device, str) else device
        if "device_key" in kwargs:  # legacy parameter of the reference
            self.device = torch.device(kwargs["device_key"])
        if self.device.type != "cuda":
            raise _ops.HironakaB200Error(
                f"hironaka_b200.TensorPoints has no CPU path; got device={self.device}. Pass a CUDA device.")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())

        if isinstance(points, list):
            points = torch.tensor(_padded(points, kwargs["max_num_points"], padding_value), device=self.device,
                                  dtype=self.dtype)
        elif isinstance(points, np.ndarray):
            points = torch.tensor(points, device=self.device, dtype=self.dtype)
        elif isinstance(points, torch.Tensor):
            points = points.type(self.dtype).to(self.device)
        else:
            raise Exception(f"Input must be a Tensor, a numpy array or a nested list. Got {type(points)}.")
        if points.dim() == 3 and not points.is_contiguous():
            points = points.contiguous()

        self.padding_value = padding_value
        self.distinguished_points = distinguished_points
        super().__init__(points, **kwargs)

    # ---- reference surface ---------------------------------------------------------------
    def exceed_threshold(self) -> bool:
        """Whether the maximal value reached the threshold (tensor_points.py:57-63)."""
        if self.value_threshold is None:
            return False
        flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        _ops.step(self.points, ops=0, write_state=False, exceed_flag=flag, value_threshold=float(self.value_threshold))
        return bool(flag.item())

    def get_num_points(self) -> torch.Tensor:
        return _ops.dones(self.points, want_num_points=True)[1].to(torch.int64)

    def get_features(self) -> torch.Tensor:
        """Rows sorted by coordinate 0, descending, STABLE (tensor_points.py:72-74); a new tensor."""
        f = _ops.features(self.points, flags=C.HK_F_OBS_SORT_COORD0, padding_value=self.padding_value)
        f = f.view(self.batch_size, self.max_num_points, self.dimension)
        return f if self.points.dtype == torch.float32 else f.to(self.points.dtype)

    def to_list_points(self) -> List[List[List[float]]]:
        """The state in ``ListPoints`` form (hironaka/core/list_points.py): per game the live points
        only, sorted descending lexicographically with coordinate 0 primary — the order
        ``get_newton_polytope_approx_lst`` leaves them in (hironaka/src/_list_ops.py:9-45).  One
        features launch, then one device->host copy."""
        f = _ops.features(self.points, flags=C.HK_F_OBS_SORT_LEX_FIRST, padding_value=self.padding_value)
        f = f.view(self.batch_size, self.max_num_points, self.dimension).cpu()
        n = _ops.dones(self.points, want_num_points=True)[1].cpu().tolist()
        return [f[b, : n[b]].tolist() for b in range(self.batch_size)]

    def type(self, t: Union[Type, torch.dtype]):
        self.dtype = t
        self.points = self.points.type(t)
        self.config["dtype"] = t

    @property
    def ended_batch_in_tensor(self) -> torch.Tensor:
        return _ops.dones(self.points)[0]

    # ---- the four game operations: one launch each through hironaka_b200.src ----------------
    def _apply(self, op: str, *args, inplace: bool = True, ignore_ended_games: bool = True, **kwargs):
        pad = self.padding_value
        if op == "shift":
            coords, axis = args
            return _src.shift_torch(self.points, coords, axis, inplace=inplace, padding_value=pad,
                                    ignore_ended_games=ignore_ended_games)
        fn = {"get_newton_polytope": _src.get_newton_polytope_torch, "reposition": _src.reposition_torch,
              "rescale": _src.rescale_torch}[op]
        return fn(self.points, inplace=inplace, padding_value=pad)

    @property
    def ended_batch(self) -> torch.Tensor:
        return _ops.dones(self.points)[0]

    @property
    def ended(self) -> bool:
        # one device->host read instead of the reference's Python all() over a tensor
        return bool(_ops.dones(self.points)[0].all().item())

    def __repr__(self) -> str:
        return str(self.points)

    def __hash__(self) -> int:
        return hash(self.points.detach().cpu().numpy().round(8).tobytes())


CudaTensorPoints = TensorPoints
