"""
ORACLE — TEST INFRASTRUCTURE ONLY.  A NumPy stand-in for the few names of `jax`, `jax.numpy`,
`jax.lax`, `jax.nn`, `chex` and `flax.core` that the reference's JAX env-step sources touch
(hironaka/src/_jax_ops.py, hironaka/jax/util.py, hironaka/jax/host_action_preprocess.py,
hironaka/jax/players.py), so that oracle/gen_golden_jax.py can EXECUTE THOSE SOURCES UNMODIFIED
in the build container, where `jax` is not installed.

What the stand-in keeps of JAX's semantics (x64 disabled, as the reference runs):
  * float results are float32, integer results int32 (every jnp function result is narrowed);
  * `vmap(fun, in_axes, out_axes)` maps by a Python loop over the mapped axis and stacks;
  * `jit`, `pmap` (over a leading device axis) are plain Python;
  * `lexsort` is NumPy's (stable, last key primary — the same contract as jnp.lexsort).
Nothing under hironaka_b200/ imports this module.
"""
from __future__ import annotations

import sys
import types

import numpy as np


def _narrow(x):
    if isinstance(x, tuple):
        return tuple(_narrow(v) for v in x)
    if isinstance(x, (np.ndarray, np.generic)):
        if x.dtype == np.float64:
            return x.astype(np.float32)
        if x.dtype == np.int64:
            return x.astype(np.int32)
    return x


class _Jnp(types.ModuleType):
    """jax.numpy: NumPy functions with float64 -> float32 / int64 -> int32 narrowing of results."""

    ndarray = np.ndarray
    float32, int32, uint32, bool_ = np.float32, np.int32, np.uint32, np.bool_
    inf = np.float32(np.inf)
    pi = np.pi

    def __getattr__(self, name):
        fn = getattr(np, name)
        if not callable(fn) or isinstance(fn, type):
            return fn

        def wrapped(*a, **k):
            if name == "clip" and "a_min" in k:  # jnp.clip(x, a_min=, a_max=)
                k = dict(k)
                k["min"] = k.pop("a_min")
                if "a_max" in k:
                    k["max"] = k.pop("a_max")
            return _narrow(fn(*a, **k))
        wrapped.__name__ = name
        return wrapped

    @staticmethod
    def array(obj, dtype=None):
        return _narrow(np.array(obj, dtype=dtype))

    @staticmethod
    def asarray(obj, dtype=None):
        return _narrow(np.asarray(obj, dtype=dtype))


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args, **kw):
        axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args), (axes, len(args))
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = np.shape(a)[ax]
                break
        assert n is not None
        outs = []
        for i in range(n):
            sl = [a if ax is None else np.take(a, i, axis=ax) for a, ax in zip(args, axes)]
            outs.append(fun(*sl, **kw))
        if isinstance(outs[0], tuple):
            return tuple(_narrow(np.stack([o[k] for o in outs], axis=out_axes)) for k in range(len(outs[0])))
        return _narrow(np.stack([np.asarray(o) for o in outs], axis=out_axes))
    mapped.__name__ = getattr(fun, "__name__", "vmapped")
    return mapped


def jit(fun=None, **_kw):
    if fun is None:
        return lambda f: f
    return fun


def pmap(fun, static_broadcasted_argnums=(), **_kw):
    """Maps the non-static arguments over their leading (device) axis."""
    static = set(static_broadcasted_argnums)

    def mapped(*args):
        n = None
        for i, a in enumerate(args):
            if i not in static:
                n = np.shape(a)[0]
                break
        outs = [fun(*[a if i in static else a[k] for i, a in enumerate(args)]) for k in range(n)]
        return _narrow(np.stack(outs, axis=0))
    return mapped


def _dynamic_slice(operand, start_indices, slice_sizes):
    idx = tuple(slice(int(s), int(s) + int(n)) for s, n in zip(start_indices, slice_sizes))
    return np.asarray(operand)[idx]


def _cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def _one_hot(x, num_classes, dtype=np.float32, axis=-1):
    x = np.asarray(x)
    return (x[..., None] == np.arange(num_classes)).astype(dtype)


class _SeededRandom(types.ModuleType):
    """jax.random stand-in: NOT bit-compatible with threefry.  A key is any integer array; draws come
    from np.random.default_rng(sum of the key words), so a golden script can regenerate them."""

    @staticmethod
    def PRNGKey(seed):
        return np.array([0, int(seed) & 0xFFFFFFFF], dtype=np.uint32)

    @staticmethod
    def _rng(key):
        return np.random.default_rng(int(np.asarray(key, dtype=np.uint64).sum()))

    @classmethod
    def randint(cls, key, shape, minval, maxval, dtype=np.int32):
        return cls._rng(key).integers(minval, maxval, size=shape).astype(dtype)

    @classmethod
    def choice(cls, key, a, shape=(), replace=True):
        return cls._rng(key).choice(np.asarray(a), size=shape, replace=replace)

    @classmethod
    def split(cls, key, num=2):
        base = int(np.asarray(key, dtype=np.uint64).sum())
        return np.array([[k + 1, base] for k in range(num)], dtype=np.uint32)


def install():
    """Registers the stand-in modules in sys.modules (only names that are not importable for real)."""
    jnp = _Jnp("jax.numpy")
    lax = types.ModuleType("jax.lax")
    lax.mul = lambda a, b: _narrow(np.multiply(a, b))
    lax.add = lambda a, b: _narrow(np.add(a, b))
    lax.dynamic_slice = _dynamic_slice
    lax.cond = _cond
    nn = types.ModuleType("jax.nn")
    nn.one_hot = _one_hot
    rnd = _SeededRandom("jax.random")
    jax = types.ModuleType("jax")
    jax.numpy, jax.lax, jax.nn, jax.random = jnp, lax, nn, rnd
    jax.vmap, jax.jit, jax.pmap = vmap, jit, pmap
    ex = types.ModuleType("jax.example_libraries")
    opt = types.ModuleType("jax.example_libraries.optimizers")
    opt.l2_norm = lambda tree: np.sqrt(sum(np.vdot(x, x) for x in tree))
    ex.optimizers = opt
    jax.example_libraries = ex
    chex = types.ModuleType("chex")
    flax = types.ModuleType("flax")
    core = types.ModuleType("flax.core")
    core.FrozenDict = dict
    flax.core = core
    for name, mod in {"jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.nn": nn, "jax.random": rnd,
                      "jax.example_libraries": ex, "jax.example_libraries.optimizers": opt, "chex": chex,
                      "flax": flax, "flax.core": core}.items():
        sys.modules[name] = mod
    return jax
