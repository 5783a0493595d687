"""The part of the reference's ``PointsBase`` contract (hironaka/core/points_base.py:7-279) that the
``point_cls=`` seam needs, written for one concrete storage: a ``[batch, max_num_points, dimension]``
CUDA tensor whose removed slots hold a non-positive padding value.

What callers of the reference rely on and this keeps: ``config`` (the constructor keywords that
``copy()`` replays, listed by the subclass in ``subcls_config_keys`` plus ``max_num_points``),
``running_attributes`` (carried over by ``copy()``), the four game operations mutating in place and
returning ``self`` — or a fresh object when ``inplace=False`` (:90-140) — and ``ended`` /
``ended_batch`` / ``__getitem__``.  The list-of-lists storage variants of the reference are not
rebuilt; a subclass supplies ``_apply(op, ...)`` for its storage.
"""
from __future__ import annotations

import copy as _copy
import logging
from typing import Any, List


class PointsBase:
    subcls_config_keys: List[str] = []
    running_attributes: List[str] = []
    _GAME_OPS = ("shift", "reposition", "get_newton_polytope", "rescale")

    def __init__(self, points: Any, max_num_points: int | None = None, **_ignored):
        self.logger = logging.getLogger(type(self).__name__)
        if points.dim() == 2:  # a single game: the reference adds the batch axis with a warning
            self.logger.warning("Points are 3-dimensional: batch, max_num_points, coordinates. "
                                "A batch dimension is automatically added.")
            points = points.unsqueeze(0)
        if points.dim() != 3:
            raise ValueError("Input dimension must be 2 or 3.")
        self.points = points
        self.batch_size, self.max_num_points, self.dimension = points.shape
        if max_num_points is not None:
            if max_num_points < self.max_num_points:
                self.logger.warning("Specified max_num_points is smaller than the one in input. Ignored.")
            else:
                self.max_num_points = max_num_points
        missing = [k for k in self.subcls_config_keys if not hasattr(self, k)]
        if missing:
            raise Exception(f"Must initialize keys in 'subcls_config_keys' before calling super().__init__: {missing}")
        self.config = {k: getattr(self, k) for k in (*self.subcls_config_keys, "max_num_points")}

    def copy(self, points=None) -> "PointsBase":
        """A new object on a clone of the points (or on `points`), same config, running attributes deep-copied."""
        twin = type(self)(self.points.clone().detach() if points is None else points, **self.config)
        for name in self.running_attributes:
            if not hasattr(self, name):
                raise Exception(f"Attribute {name} is not initialized.")
            setattr(twin, name, _copy.deepcopy(getattr(self, name)))
        return twin

    def _run(self, op: str, *args, inplace: bool = True, **kwargs) -> "PointsBase":
        result = self._apply(op, *args, inplace=inplace, **kwargs)
        return self if inplace else self.copy(points=result)

    def shift(self, coords, axis, inplace=True, **kwargs):
        return self._run("shift", coords, axis, inplace=inplace, **kwargs)

    def reposition(self, inplace=True, **kwargs):
        return self._run("reposition", inplace=inplace, **kwargs)

    def get_newton_polytope(self, inplace=True, **kwargs):
        return self._run("get_newton_polytope", inplace=inplace, **kwargs)

    def rescale(self, inplace=True, **kwargs):
        return self._run("rescale", inplace=inplace, **kwargs)

    def _apply(self, op: str, *args, inplace: bool, **kwargs):
        raise NotImplementedError("a storage-specific subclass runs the game operations")

    @property
    def ended_batch(self):
        raise NotImplementedError

    @property
    def ended(self) -> bool:
        return bool(self.ended_batch.all())

    def get_features(self):
        return self.points

    def __getitem__(self, item: int):
        return self.points[item]
