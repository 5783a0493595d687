"""Host-buffer sessions: the NumPy / ctypes calling style of the reference's own C binding
(hironaka/src/_np_ops.py:6-15,57-84) on top of ``hk_session_*``.  All arrays are HOST arrays; the
session owns the device state of one shard of games, and every call includes its H2D / D2H
copies.  No torch involved."""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import constants as C
from ._lib import check, lib


class HostSession:
    def __init__(self, points: np.ndarray, device: int = 0, padding_value: float = -1.0):
        if points.dtype == np.int32:
            self.dtype = C.HK_DTYPE_I32
        elif points.dtype == np.float32:
            self.dtype = C.HK_DTYPE_F32
        else:
            raise TypeError("points must be int32 or float32")
        if points.ndim != 3:
            raise ValueError("points must be [B, N, d]")
        self.B, self.N, self.d = points.shape
        self._np_dtype = points.dtype
        self._h = ctypes.c_void_p()
        check(lib().hk_session_create(ctypes.byref(self._h), device, self.B, self.N, self.d, self.dtype,
                                      float(padding_value)), "hk_session_create")
        self.set_state(points)

    def set_state(self, points: np.ndarray) -> None:
        p = np.ascontiguousarray(points, dtype=self._np_dtype)
        assert p.shape == (self.B, self.N, self.d)
        check(lib().hk_session_set_state(self._h, p.ctypes.data), "hk_session_set_state")

    def get_state(self) -> np.ndarray:
        out = np.empty((self.B, self.N, self.d), dtype=self._np_dtype)
        check(lib().hk_session_get_state(self._h, out.ctypes.data), "hk_session_get_state")
        return out

    def step(self, host_action: Optional[np.ndarray], axis: Optional[np.ndarray], ops: int, flags: int,
             done: Optional[np.ndarray] = None, reward: Optional[np.ndarray] = None, want_done_count: bool = True):
        """H2D(actions) -> one fused step -> D2H(done, reward, done_count); blocking.
        `done` (uint8 [B]) / `reward` (float32 [B]) are caller-provided output buffers or None."""
        packed = bool(flags & (C.HK_F_ACT_PACKED | C.HK_F_ACT_NIBBLE))  # the element size of the action arrays follows the flags
        want = np.uint8 if flags & (C.HK_F_ACT_U8 | C.HK_F_ACT_PACKED | C.HK_F_ACT_NIBBLE) else np.int32
        ha = None if host_action is None else np.ascontiguousarray(host_action, dtype=want)
        ax = None if (axis is None or packed) else np.ascontiguousarray(axis, dtype=want)
        nact = (self.B + 1) // 2 if flags & C.HK_F_ACT_NIBBLE else self.B
        for a in (ha, ax):
            if a is not None and a.shape != (nact,):
                raise ValueError(f"actions must have shape ({nact},)")
        cnt = ctypes.c_int32(0)
        rc = lib().hk_session_step(self._h, None if ha is None else ha.ctypes.data,
                                   None if ax is None else ax.ctypes.data,
                                   None if done is None else done.ctypes.data,
                                   None if reward is None else reward.ctypes.data,
                                   ctypes.addressof(cnt) if want_done_count else None, ops, flags)
        check(rc, "hk_session_step")
        return cnt.value if want_done_count else None

    def rollout(self, host_actions: np.ndarray, axes: np.ndarray, ops: int, flags: int,
                done: Optional[np.ndarray] = None, done_bits: Optional[np.ndarray] = None) -> np.ndarray:
        """T steps from host action streams [T, B] (int32, or uint8 with HK_F_ACT_U8): uploads are
        double-buffered against the running step, the finished-game count of every step is read
        back; returns int32 [T].  `done` (uint8 [T, B], caller-provided, ideally pinned) also receives
        every step's per-game done flags, read back on a third stream; `done_bits` (uint32 [T, ceil(B/32)])
        receives them as bit masks instead (an eighth of the bytes; `unpack_done_bits` expands them).
        HK_F_ACT_NIBBLE streams are uint8 [T, ceil(B/2)] (`pack_actions_nibble`).  Pass pinned arrays
        (e.g. torch.empty(..., pin_memory=True).numpy()) for full PCIe speed."""
        packed = bool(flags & (C.HK_F_ACT_PACKED | C.HK_F_ACT_NIBBLE))  # both players in one stream; `axes` unused
        want = np.uint8 if flags & (C.HK_F_ACT_U8 | C.HK_F_ACT_PACKED | C.HK_F_ACT_NIBBLE) else np.int32
        ha = np.ascontiguousarray(host_actions, dtype=want)
        ax = None if packed else np.ascontiguousarray(axes, dtype=want)
        width = (self.B + 1) // 2 if flags & C.HK_F_ACT_NIBBLE else self.B
        if ha.ndim != 2 or ha.shape[1] != width or (ax is not None and ax.shape != ha.shape):
            raise ValueError(f"host_actions and axes must be [T, {width}]")
        counts = np.zeros(ha.shape[0], dtype=np.int32)
        if done is not None and (done.dtype != np.uint8 or done.shape != (ha.shape[0], self.B) or not done.flags.c_contiguous):
            raise ValueError("done must be a contiguous uint8 [T, B] array")
        if done_bits is not None:
            if done is not None:
                raise ValueError("ask for done or done_bits, not both")
            if done_bits.dtype != np.uint32 or done_bits.shape != (ha.shape[0], (self.B + 31) // 32) or \
                    not done_bits.flags.c_contiguous:
                raise ValueError("done_bits must be a contiguous uint32 [T, ceil(B/32)] array")
            check(lib().hk_session_rollout_bits(self._h, ha.ctypes.data, None if ax is None else ax.ctypes.data, ha.shape[0],
                                                counts.ctypes.data, done_bits.ctypes.data, ops, flags),
                  "hk_session_rollout_bits")
            return counts
        check(lib().hk_session_rollout_ex(self._h, ha.ctypes.data, None if ax is None else ax.ctypes.data, ha.shape[0],
                                          counts.ctypes.data, None if done is None else done.ctypes.data, ops, flags),
              "hk_session_rollout_ex")
        return counts

    @staticmethod
    def pack_actions(host_actions: np.ndarray, axes: np.ndarray) -> np.ndarray:
        """HK_F_ACT_PACKED encoding of two action streams: host action (< 32) | axis (< 8) << 5, uint8."""
        ha, ax = np.asarray(host_actions), np.asarray(axes)
        if ha.max(initial=0) > 31 or ax.max(initial=0) > 7 or ha.min(initial=0) < 0 or ax.min(initial=0) < 0:
            raise ValueError("packed actions need host actions < 32 and axes < 8")
        return (ha.astype(np.uint8) | (ax.astype(np.uint8) << 5)).astype(np.uint8)

    @staticmethod
    def pack_actions_nibble(host_actions: np.ndarray, axes: np.ndarray) -> np.ndarray:
        """HK_F_ACT_NIBBLE encoding: [..., B] discrete host ids (< 4) and axes (< 4) -> uint8 [..., ceil(B/2)], game 2i
        in the low nibble and 2i+1 in the high nibble of byte i, nibble = id | axis << 2."""
        ha, ax = np.asarray(host_actions), np.asarray(axes)
        if ha.max(initial=0) > 3 or ax.max(initial=0) > 3 or ha.min(initial=0) < 0 or ax.min(initial=0) < 0:
            raise ValueError("nibble-packed actions need host ids < 4 and axes < 4")
        nib = (ha.astype(np.uint8) | (ax.astype(np.uint8) << 2)).astype(np.uint8)
        if nib.shape[-1] % 2:
            nib = np.concatenate([nib, np.zeros(nib.shape[:-1] + (1,), np.uint8)], axis=-1)
        return (nib[..., 0::2] | (nib[..., 1::2] << 4)).astype(np.uint8)

    @staticmethod
    def unpack_done_bits(done_bits: np.ndarray, B: int) -> np.ndarray:
        """uint32 [..., ceil(B/32)] bit masks -> bool [..., B] (bit g % 32 of word g / 32)."""
        bits = np.unpackbits(np.ascontiguousarray(done_bits).view(np.uint8), axis=-1, bitorder="little")
        return bits[..., :B].astype(bool)

    def close(self) -> None:
        if self._h:
            lib().hk_session_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
