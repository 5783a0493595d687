"""
ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/hk_oracle.c (the C restatement).

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
Never imported by the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhk_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hk_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "hironaka_b200.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_SO) for f in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        p = ctypes.c_void_p
        for name in ("hk_oracle_step_i32", "hk_oracle_step_f32"):
            f = getattr(L, name)
            f.argtypes = [p, p, p, p, p, p, p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                          ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float]
            f.restype = ctypes.c_int
        for name in ("hk_oracle_features_i32", "hk_oracle_features_f32"):
            f = getattr(L, name)
            f.argtypes = [p, p, p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint32, ctypes.c_float]
            f.restype = ctypes.c_int
        L.hk_oracle_rescale_f32.argtypes = [p, p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_float]
        L.hk_oracle_rescale_f32.restype = ctypes.c_int
        L.hk_oracle_threads.restype = ctypes.c_int
        L.hk_oracle_set_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def threads() -> int:
    return lib().hk_oracle_threads()


def set_threads(n: int) -> None:
    lib().hk_oracle_set_threads(n)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def step(state: np.ndarray, host_action, axis, ops: int, flags: int, padding_value: float = -1.0,
         want=("done", "reward", "num_points"), out: np.ndarray | None = None):
    """One game-step on a contiguous int32 or float32 [B,N,d] array.
    Returns (new_state, done u8 | None, reward f32 | None, num_points i32 | None)."""
    assert state.dtype in (np.int32, np.float32) and state.ndim == 3
    state = np.ascontiguousarray(state)
    B, N, d = state.shape
    new = np.empty_like(state) if out is None else out
    ha = None if host_action is None else np.ascontiguousarray(host_action, dtype=np.int32)
    ax = None if axis is None else np.ascontiguousarray(axis, dtype=np.int32)
    done = np.empty(B, np.uint8) if "done" in want else None
    rew = np.empty(B, np.float32) if "reward" in want else None
    npts = np.empty(B, np.int32) if "num_points" in want else None
    fn = lib().hk_oracle_step_i32 if state.dtype == np.int32 else lib().hk_oracle_step_f32
    rc = fn(_ptr(state), _ptr(new), _ptr(ha), _ptr(ax), _ptr(done), _ptr(rew), _ptr(npts), B, N, d, ops, flags,
            float(padding_value))
    if rc != 0:
        raise RuntimeError(f"hk_oracle_step failed: {rc}")
    return new, done, rew, npts


def features(state: np.ndarray, flags: int, obs_coord=None, padding_value: float = -1.0) -> np.ndarray:
    assert state.dtype in (np.int32, np.float32) and state.ndim == 3
    state = np.ascontiguousarray(state)
    B, N, d = state.shape
    oc = None if obs_coord is None else np.ascontiguousarray(obs_coord, dtype=np.int32)
    obs = np.empty((B, N * d + (d if oc is not None else 0)), np.float32)
    fn = lib().hk_oracle_features_i32 if state.dtype == np.int32 else lib().hk_oracle_features_f32
    rc = fn(_ptr(state), _ptr(obs), _ptr(oc), B, N, d, flags, float(padding_value))
    if rc != 0:
        raise RuntimeError(f"hk_oracle_features failed: {rc}")
    return obs


def rescale(state: np.ndarray, padding_value: float = -1.0) -> np.ndarray:
    state = np.ascontiguousarray(state, dtype=np.float32)
    B, N, d = state.shape
    out = np.empty_like(state)
    lib().hk_oracle_rescale_f32(_ptr(state), _ptr(out), B, N, d, float(padding_value))
    return out
