"""
ORACLE — TEST INFRASTRUCTURE ONLY.  NumPy restatement of the reference env-step path.

Nothing under ``hironaka_b200/`` may import this module.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg use it,
and only as the checker or as the timed CPU baseline.

Every function restates one function of honglu2875/hironaka (file:line relative to the
reference root) in NumPy, keeping the reference's order of operations so that float32
results are bit-identical, and keeping the two flavours of the reference apart:

  * ``*_torch``  follow hironaka/src/_torch_ops.py + hironaka/src/_fn.py (TensorPoints path)
  * ``*_jax``    follow hironaka/src/_jax_ops.py + hironaka/jax/util.py   (JAX path)

Parity pinning: both flavours reproduce the reference's known-answer vectors (tests/kat.py) and
outputs of the reference's OWN code run in the build container: the torch flavour imported as is
(tests/golden/ref_*.npz, oracle/gen_golden.py), the JAX flavour by executing the reference's JAX
source files unmodified over a NumPy stand-in for jax.numpy / vmap / jit, since jax is not
installed here (tests/golden/ref_jax_*.npz, oracle/gen_golden_jax.py + oracle/jax_numpy_shim.py).

The [B,N,N,d] temporaries of the reference are kept on purpose (this is the restatement, not
the fast path); use oracle/hk_oracle.c through ``oracle.cport`` for large batches.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# torch flavour
# --------------------------------------------------------------------------------------


def remove_repeated(points: np.ndarray, padding_value: float = -1.0) -> np.ndarray:
    """hironaka/src/_fn.py:192-213 — later copies of identical rows become padding."""
    B, N, d = points.shape
    difference = points[:, :, None, :] - points[:, None, :, :]
    lower = ~np.triu(np.ones((N, N), dtype=bool), k=0)  # strict lower triangle: j < i
    repeated = ((difference == 0).all(3) & lower[None]).any(2)
    out = points.copy()
    out[repeated] = padding_value
    return out


def get_newton_polytope_torch(points: np.ndarray, padding_value: float = -1.0) -> np.ndarray:
    """hironaka/src/_torch_ops.py:8-39 — dedupe, then drop rows dominated by another live row."""
    points = remove_repeated(points, padding_value)
    B, N, d = points.shape
    available = points >= 0
    filter_matrix = available[:, :, None, :] & available[:, None, :, :]
    difference = points[:, :, None, :] - points[:, None, :, :]
    diag_filter = ~np.eye(N, dtype=bool)[None, :, :, None]
    to_remove = ((difference >= 0) & diag_filter & filter_matrix).all(3).any(2)
    out = points.copy()
    out[to_remove] = padding_value
    return out


def coord_list_to_binary(coords, dimension: int) -> np.ndarray:
    """hironaka/src/_fn.py:99-106."""
    out = np.zeros((len(coords), dimension), dtype=np.float32)
    for b, c in enumerate(coords):
        out[b, list(c)] = 1
    return out


def shift_torch(points: np.ndarray, coord, axis, padding_value: float = -1.0,
                ignore_ended_games: bool = True) -> np.ndarray:
    """hironaka/src/_torch_ops.py:46-110.

    coord: list of index lists or [B,d] 0/1 array; axis: [B].  Invalid action (axis not in the
    coordinate set, :90-91) and, with ignore_ended_games, games with < 2 live rows (:92-93) are
    left unchanged.  Dead rows are rewritten with padding_value (:104)."""
    B, N, d = points.shape
    dt = points.dtype
    if isinstance(coord, list):
        coord = coord_list_to_binary(coord, d)
    coord = np.asarray(coord).astype(dt)
    axis = np.asarray(axis).astype(np.int64)
    assert coord.shape == (B, d) and axis.shape == (B,)
    available = points >= 0
    assert (available.all(2) == available.any(2)).all()
    axis_binary = np.zeros((B, d), dtype=dt)
    axis_binary[np.arange(B), axis] = 1
    valid = ((axis_binary - coord) <= 0).all(1)
    axis_binary = axis_binary * valid[:, None].astype(dt)
    if ignore_ended_games:
        axis_binary = axis_binary * ((points[:, :, 0] >= 0).sum(1) >= 2)[:, None].astype(dt)
    trans = (axis_binary[:, :, None] * coord[:, None, :] + np.eye(d, dtype=dt)[None]
             - axis_binary[:, :, None] * axis_binary[:, None, :])
    # torch.matmul([B,N,d,d],[B,N,d,1]); sequential accumulation over j in float32
    transformed = np.zeros_like(points)
    for j in range(d):
        transformed = transformed + trans[:, None, :, j] * points[:, :, j:j + 1]
    out = np.where(available, transformed, np.asarray(padding_value, dtype=dt))
    return out.astype(dt)


def reposition_torch(points: np.ndarray, padding_value: float = -1.0) -> np.ndarray:
    """hironaka/src/_torch_ops.py:113-133."""
    dt = points.dtype
    available = points >= 0
    maximum = points.max()
    pre = np.where(available, points, (maximum + 1).astype(dt))
    cmin = pre.min(axis=1)
    unfiltered = points - cmin[:, None, :]
    return np.where(available, unfiltered, np.asarray(padding_value, dtype=dt)).astype(dt)


def rescale_torch(points: np.ndarray, padding_value: float = -1.0) -> np.ndarray:
    """hironaka/src/_torch_ops.py:136-146 — divide live entries by the per-game max (0 -> 1)."""
    dt = points.dtype
    available = points >= 0
    max_val = points.max(axis=(1, 2))
    max_val = max_val + (max_val == 0).astype(dt)
    r = (points * available.astype(dt)) / max_val[:, None, None] + np.asarray(padding_value, dt) * (~available).astype(dt)
    return r.astype(dt)


def get_num_points(points: np.ndarray) -> np.ndarray:
    """hironaka/core/tensor_points.py:65-70."""
    return (points[:, :, 0] >= 0).sum(1)


def ended_batch(points: np.ndarray) -> np.ndarray:
    """hironaka/core/tensor_points.py:113-120 — a game has ended iff <= 1 live row."""
    return get_num_points(points) <= 1


def get_features_torch(points: np.ndarray) -> np.ndarray:
    """hironaka/core/tensor_points.py:72-74 — rows sorted by coordinate 0, descending.

    The reference's argsort is not stable (tie order is implementation-defined for N >= 17,
    SURVEY.md section 7 hard part 2); the engine defines the STABLE order (lowest index first)
    and this restatement does the same."""
    order = np.argsort(-points[:, :, 0], axis=1, kind="stable")
    return np.take_along_axis(points, order[:, :, None], axis=1).copy()


def default_reward(sample_for: str, next_done: np.ndarray) -> np.ndarray:
    """hironaka/trainer/fused_game.py:175-182."""
    r = next_done.astype(np.float32)
    return r if sample_for == "host" else -r


def fused_game_point_ops(points: np.ndarray, host_moves: np.ndarray, actions: np.ndarray,
                         scale_observation: bool, padding_value: float = -1.0) -> np.ndarray:
    """hironaka/trainer/fused_game.py:150-162 — shift -> newton -> (rescale)."""
    p = shift_torch(points, host_moves, actions, padding_value)
    p = get_newton_polytope_torch(p, padding_value)
    if scale_observation:
        p = rescale_torch(p, padding_value)
    return p


# --------------------------------------------------------------------------------------
# host action table
# --------------------------------------------------------------------------------------


def decode_table(dimension: int) -> np.ndarray:
    """hironaka/jax/host_action_preprocess.py:8-24 and HostActionEncoder.__init__
    (hironaka/src/_fn.py:255-269): row i = multi-binary vector of the i-th integer in
    1..2^d-1 that is not a power of two; bit k <=> coordinate k."""
    rows = []
    for i in range(2 ** dimension):
        if i == 0 or i & (i - 1) == 0:
            continue
        rows.append([(i >> k) & 1 for k in range(dimension)])
    return np.array(rows, dtype=np.int32)


def encode(multi_binary: np.ndarray) -> np.ndarray:
    """hironaka/jax/host_action_preprocess.py:78-87, _fn.py:285-296: id = m - floor(log2 m) - 2."""
    mb = np.asarray(multi_binary)
    d = mb.shape[-1]
    m = (mb.astype(np.int64) * (2 ** np.arange(d))).sum(-1)
    return (m - np.floor(np.log2(m)).astype(np.int64) - 2).astype(np.int64)


def encode_one_hot(multi_binary: np.ndarray) -> np.ndarray:
    """hironaka/jax/host_action_preprocess.py:90-99."""
    d = np.asarray(multi_binary).shape[-1]
    cls = 2 ** d - d - 1
    return (np.arange(cls)[None, :] == encode(multi_binary)[..., None]).astype(np.float32)


# --------------------------------------------------------------------------------------
# JAX flavour
# --------------------------------------------------------------------------------------


def remove_repeated_jax(points: np.ndarray, padding_value: float = -1.0) -> np.ndarray:
    """hironaka/src/_jax_ops.py:15-40."""
    B, N, d = points.shape
    eq = (points[:, :, None, :] == points[:, None, :, :]).all(3)
    rep = (eq & ~np.triu(np.ones((N, N), dtype=bool), k=0)[None]).any(2)
    mask = ~rep
    dt = points.dtype
    return (points * mask[:, :, None].astype(dt) + ((~mask) * padding_value).astype(dt)[:, :, None]).astype(dt)


def get_newton_polytope_jax(points: np.ndarray, padding_value: float = -1.0) -> np.ndarray:
    """hironaka/src/_jax_ops.py:43-73 (get_interior + get_newton_polytope_approx_jax)."""
    points = remove_repeated_jax(points)
    B, N, d = points.shape
    dt = points.dtype
    available = (points >= 0).all(2)
    amask = available[:, :, None] & available[:, None, :]
    diff = points[:, :, None, :] - points[:, None, :, :]
    res = ~(((diff >= 0).all(3)) & (~np.eye(N, dtype=bool))[None] & amask).any(2)
    return (points * res[:, :, None].astype(dt) + ((~res) * padding_value).astype(dt)[:, :, None]).astype(dt)


def shift_jax(points: np.ndarray, coord: np.ndarray, axis: np.ndarray, padding_value: float = -1.0) -> np.ndarray:
    """hironaka/src/_jax_ops.py:76-90 — x_a <- sum_j c_j x_j UNCONDITIONALLY (no validity or
    ended-game check); live row = any(x >= 0)."""
    B, N, d = points.shape
    dt = points.dtype
    coord = np.asarray(coord).astype(dt)
    axis = np.asarray(axis)
    axis_binary = np.arange(d)[None, :] == axis[:, None]
    prod = points * coord[:, None, :]
    s = np.zeros((B, N), dtype=dt)
    for j in range(d):  # jnp.sum over the last axis, sequential in float32
        s = s + prod[:, :, j]
    shifted = s[:, :, None] * axis_binary[:, None, :].astype(dt) + points * (~axis_binary)[:, None, :].astype(dt)
    available = (points >= 0).any(2)
    return (shifted * available[:, :, None].astype(dt) + ((~available) * padding_value).astype(dt)[:, :, None]).astype(dt)


def rescale_jax(points: np.ndarray, padding_value: float = -1.0) -> np.ndarray:
    """hironaka/src/_jax_ops.py:93-111 — per game x/max unless max <= 1e-8."""
    dt = points.dtype
    available = (points >= 0).any(2)
    maximum = points.max(axis=(1, 2), keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        raw = np.where(maximum <= 1e-8, points, points / maximum)
    return (raw * available[:, :, None].astype(dt) + ((~available) * padding_value).astype(dt)[:, :, None]).astype(dt)


def reposition_jax(points: np.ndarray, padding_value: float = -1.0) -> np.ndarray:
    """hironaka/src/_jax_ops.py:114-123 — per game and coordinate subtract the min over live
    entries; column unchanged when that min is <= 0."""
    dt = points.dtype
    available = points >= 0
    colmax = points.max(axis=1, keepdims=True)
    modified = points * available.astype(dt) + ((~available) * colmax).astype(dt)
    minimal = modified.min(axis=1, keepdims=True)
    moved = (points - minimal) * available.astype(dt) + ((~available) * padding_value).astype(dt)
    return np.where(minimal <= 0.0, points, moved).astype(dt)


def make_agent_obs(pts: np.ndarray, coords: np.ndarray) -> np.ndarray:
    """hironaka/jax/util.py:22-31."""
    return np.concatenate([pts.reshape(pts.shape[0], -1), coords], axis=1)


def get_dones(pts: np.ndarray) -> np.ndarray:
    """hironaka/jax/util.py:34-35."""
    return (pts[:, :, 0] >= 0).sum(1) < 2


def get_done_from_flatten(obs: np.ndarray, role: str, dimension: int) -> np.ndarray:
    """hironaka/jax/util.py:38-39."""
    return (obs >= 0).sum(-1) <= dimension + (role == "agent") * dimension


def take_actions(role: str, spec, observations: np.ndarray, actions: np.ndarray, axis: np.ndarray,
                 rescale_points: bool = False, reposition: bool = True) -> np.ndarray:
    """hironaka/jax/util.py:82-125 — shift -> (reposition) -> newton -> (rescale), flattened."""
    N, d = spec
    if role == "host":
        points = observations.reshape(-1, N, d)
        coords = actions
    elif role == "agent":
        points = observations[:, : N * d].reshape(-1, N, d)
        coords = observations[:, N * d: N * d + d]
    else:
        raise ValueError(role)
    p = shift_jax(points, coords, axis)
    if reposition:
        p = reposition_jax(p)
    p = get_newton_polytope_jax(p)
    if rescale_points:
        p = rescale_jax(p)
    return p.reshape(-1, N * d)


def reward_fn(role: str, dones: np.ndarray, prev_dones: np.ndarray) -> np.ndarray:
    """hironaka/jax/util.py:128-149."""
    r = (dones & ~prev_dones).astype(np.float32)
    return r if role == "host" else -r


def order_and_rescale(x: np.ndarray, spec, scale_observation: bool = True) -> np.ndarray:
    """hironaka/jax/util.py:186-196 — (rescale) then stable lexsort of rows, descending, last
    coordinate primary."""
    N, d = spec
    xr = x.reshape(-1, N, d)
    if scale_observation:
        xr = rescale_jax(xr)
    out = np.empty_like(xr)
    for b in range(xr.shape[0]):
        keys = tuple(-xr[b, :, k] for k in range(d))  # np.lexsort: LAST key is primary
        idx = np.lexsort(keys)
        out[b] = xr[b, idx]
    return out.reshape(-1, N * d)


def feature_fn(role: str, spec, observations: np.ndarray, scale_observation: bool = True) -> np.ndarray:
    """hironaka/jax/util.py:172-214."""
    N, d = spec
    if role == "host":
        return order_and_rescale(observations, spec, scale_observation)
    pts = order_and_rescale(observations[:, : N * d], spec, scale_observation)
    coords = observations[:, N * d: N * d + d]
    return np.concatenate([pts, coords], axis=1)


def generate_pts(rng: np.random.Generator, shape, max_value: int, dtype=np.float32, rescale: bool = True,
                 reposition: bool = True) -> np.ndarray:
    """hironaka/jax/util.py:385-392 (random source replaced by a NumPy generator)."""
    pts = rng.integers(0, max_value, size=shape).astype(dtype)
    pts = get_newton_polytope_jax(pts)
    if reposition:
        pts = reposition_jax(pts)
    if rescale:
        pts = rescale_jax(pts)
    return pts


def calculate_value_using_reward_fn(num_points: np.ndarray, discount: float, reward_sign: int, est_sign: int,
                                    use_unified_tree: bool) -> np.ndarray:
    """hironaka/jax/util.py:261-284 with reward_fn = +-(dones & ~prev_dones) (:128-149) and
    est_fn = sign / clip(num_points, 1) (:152-169).  float32 like the reference."""
    num_points = np.asarray(num_points)
    B, T = num_points.shape
    done = num_points <= 1
    next_done = np.concatenate([done[:, 1:], np.zeros((B, 1), dtype=bool)], axis=1)
    reward = (next_done & ~done).astype(np.float32) * np.float32(reward_sign)
    diff = np.arange(T).reshape(1, -1) - np.arange(T).reshape(-1, 1)
    g = np.float32(-discount if use_unified_tree else discount)
    with np.errstate(over="ignore"):
        table = np.clip(np.power(g, diff.astype(np.float32)), -1, 1).astype(np.float32)
    discounted = reward @ table.T  # vmap(matmul)(table, reward[b])
    sign = (-1) ** (T + 1) if use_unified_tree else 1
    est = (np.float32(1) / np.clip(num_points[:, -1], 1, None).astype(np.float32)) * np.float32(est_sign) * np.float32(sign)
    unfinished = ((~done[:, -1:]) * est[:, None]) * np.power(g, np.arange(T)[::-1].astype(np.float32))[None, :]
    return (discounted + unfinished).astype(np.float32)


def rollout_postprocess(obs: np.ndarray, policy: np.ndarray, value: np.ndarray, role: str, dimension: int,
                        discount: float, use_unified_tree: bool = True):
    """hironaka/jax/jax_trainer.py:558-592."""
    offset = 1 if use_unified_tree or role == "agent" else 0
    num_points = (obs >= 0).sum(-1) // dimension - offset
    reward_sign = -1 if (use_unified_tree or role == "agent") else 1
    est_sign = 1 if role == "host" else -1
    v = calculate_value_using_reward_fn(num_points, discount, reward_sign, est_sign, use_unified_tree)
    return obs.reshape(-1, obs.shape[2]), policy.reshape(-1, policy.shape[2]), v.ravel().astype(value.dtype)


# ---- fixed players (hironaka/jax/players.py) ------------------------------------------


def all_coord_host_fn(pts: np.ndarray) -> np.ndarray:
    """players.py:42-52."""
    B, N, d = pts.shape
    cls = 2 ** d - d - 1
    out = np.zeros((B, cls), dtype=np.float32)
    out[:, cls - 1] = 1
    return out


def choose_first_agent_fn(obs: np.ndarray, spec) -> np.ndarray:
    """players.py:156-183."""
    N, d = spec
    coords = obs[:, N * d: N * d + d]
    return (np.arange(d)[None] == coords.argmax(1)[:, None]).astype(np.float32)


def choose_last_agent_fn(obs: np.ndarray, spec) -> np.ndarray:
    """players.py:186-212."""
    N, d = spec
    coords = obs[:, N * d: N * d + d].astype(np.float32)
    a = (coords + np.arange(d, dtype=np.float32)[None] * np.float32(1e-5)).argmax(1)
    return (np.arange(d)[None] == a[:, None]).astype(np.float32)


def zeillinger_fn_slice(pts: np.ndarray) -> np.ndarray:
    """players.py:55-105 — Zeillinger host on one game [N,d] -> one-hot host action."""
    n, d = pts.shape
    pts = pts.astype(np.float32)
    char = np.full((n, n, 2), np.inf, dtype=np.float32)
    for i in range(n):
        for j in range(n):
            v1, v2 = pts[i], pts[j]
            diff = v1 - v2
            mx, mn = diff.max(), diff.min()
            if (v1 < 0).any() or (v2 < 0).any() or np.isclose(mx, mn):
                continue
            mc, nc = (diff == mx).sum(), (diff == mn).sum()
            char[i, j] = (mx - mn, mc if mx == mn else mc + nc)
    cv = char.reshape(-1, 2)
    min_index = np.lexsort((cv[:, 1], cv[:, 0]))[0]
    diffs = (pts[:, None, :] - pts[None, :, :]).reshape(-1, d)
    v = diffs[min_index]
    amin, amax = int(v.argmin()), int(v.argmax())
    cls = 2 ** d - d - 1
    if amin == amax:
        out = np.zeros(cls, dtype=np.float32)
        out[0] = 1
        return out
    mb = ((np.arange(d) == amin) | (np.arange(d) == amax)).astype(np.int64)
    return encode_one_hot(mb[None])[0]


def zeillinger_fn(pts: np.ndarray) -> np.ndarray:
    return np.stack([zeillinger_fn_slice(p) for p in pts]).astype(np.float32)


# ---- the in-kernel random players (the library's RNG contract, include/hironaka_b200.h) ---------


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) on uint32
    arrays; returns the four output words.  Known answers of the Random123 distribution are checked in
    tests/test_oracle_golden.py."""
    c = [np.asarray(x, dtype=np.uint64) & np.uint64(0xFFFFFFFF) for x in (c0, c1, c2, c3)]
    k = [np.uint64(int(k0) & 0xFFFFFFFF), np.uint64(int(k1) & 0xFFFFFFFF)]
    M0, M1, LO, S = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF), np.uint64(32)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> S) ^ c[1] ^ k[0], p1 & LO, (p0 >> S) ^ c[3] ^ k[1], p0 & LO]
        k = [(k[0] + np.uint64(0x9E3779B9)) & LO, (k[1] + np.uint64(0xBB67AE85)) & LO]
    return [x.astype(np.uint32) for x in c]


def random_player_actions(B: int, d: int, T: int, seed: int, step_offset: int = 0):
    """Action streams [T, B] of the library's in-kernel random players (HK_F_HOST_RANDOM / HK_F_AGENT_RANDOM):
    counter (game low, game high, step_offset + t, 0), key (seed low, seed high); host id = floor(word0 * ncls /
    2^32) over the 2^d - d - 1 coordinate sets (random_host_fn, hironaka/jax/players.py:28-39), axis =
    floor(word1 * d / 2^32) over all d axes (random_agent_fn, players.py:142-153)."""
    g = np.arange(B, dtype=np.uint64)
    ncls = 2 ** d - d - 1
    ha = np.empty((T, B), np.int32)
    ax = np.empty((T, B), np.int32)
    for t in range(T):
        r = philox4x32_10(g & np.uint64(0xFFFFFFFF), g >> np.uint64(32), np.full(B, step_offset + t, np.uint64),
                          np.zeros(B, np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        ha[t] = ((r[0].astype(np.uint64) * np.uint64(ncls)) >> np.uint64(32)).astype(np.int32)
        ax[t] = ((r[1].astype(np.uint64) * np.uint64(d)) >> np.uint64(32)).astype(np.int32)
    return ha, ax


# --------------------------------------------------------------------------------------
# unified step used by the parity tests (same flag vocabulary as include/hironaka_b200.h)
# --------------------------------------------------------------------------------------

OP_SHIFT, OP_REPOSITION, OP_NEWTON, OP_RESCALE, OP_DEDUPE = 1, 2, 4, 8, 16
F_NOOP_INVALID, F_FREEZE_ENDED, F_ACT_DISCRETE, F_ROLE_AGENT = 1, 2, 4, 8
F_OBS_RESCALE, F_OBS_SORT_COORD0, F_OBS_SORT_LEX = 16, 32, 64
F_RESCALE_EPS = 1 << 15


def masks_to_binary(mask: np.ndarray, d: int) -> np.ndarray:
    return ((np.asarray(mask)[:, None] >> np.arange(d)[None, :]) & 1).astype(np.float32)


def step(points: np.ndarray, host_action, axis, ops: int, flags: int, padding_value: float = -1.0,
         obs_coord=None, want_obs: bool = False):
    """One game-step with the C-ABI's op/flag vocabulary, composed from the restated reference
    functions above: torch flavour when NOOP_INVALID/FREEZE_ENDED are set, JAX flavour (the `_jax`
    restatements throughout) otherwise; F_RESCALE_EPS selects calculate_rescale's `max <= 1e-8 ->
    unchanged` rule (_jax_ops.py:93-98) for the rescale op and the rescaled observation.
    Works on a float32 copy (the reference's storage) and returns
    (new_points[f32], done, reward, num_points, obs or None)."""
    p = np.asarray(points).astype(np.float32)
    B, N, d = p.shape
    prev_done = get_dones(p)
    if ops & OP_SHIFT:
        ha = np.asarray(host_action)
        if flags & F_ACT_DISCRETE:
            coord = decode_table(d)[ha].astype(np.float32)
        else:
            coord = masks_to_binary(ha, d)
        ax = np.asarray(axis)
        if flags & (F_NOOP_INVALID | F_FREEZE_ENDED):
            if flags & F_NOOP_INVALID:
                p = shift_torch(p, coord, ax, padding_value, ignore_ended_games=bool(flags & F_FREEZE_ENDED))
            else:  # freeze ended games only: JAX shift on the games that are still running
                q = shift_jax(p, coord, ax, padding_value)
                p = np.where(prev_done[:, None, None], np.where(p >= 0, p, np.float32(padding_value)), q)
        else:
            p = shift_jax(p, coord, ax, padding_value)
    jax_flavour = not (flags & (F_NOOP_INVALID | F_FREEZE_ENDED))
    rescale = (lambda q: rescale_jax(q, padding_value)) if flags & F_RESCALE_EPS else \
        (lambda q: rescale_torch(q, padding_value))
    if jax_flavour and padding_value == -1.0:
        if ops & OP_REPOSITION:
            p = reposition_jax(p)
        if ops & OP_DEDUPE:
            p = remove_repeated_jax(p)
        if ops & OP_NEWTON:
            p = get_newton_polytope_jax(p)
    else:
        if ops & OP_REPOSITION:
            p = reposition_torch(p, padding_value)
        if ops & OP_DEDUPE:
            p = remove_repeated(p, padding_value)
        if ops & OP_NEWTON:
            p = get_newton_polytope_torch(p, padding_value)
    if ops & OP_RESCALE:
        p = rescale(p)
    done = get_dones(p)
    rew = reward_fn("agent" if flags & F_ROLE_AGENT else "host", done, prev_done)
    obs = None
    if want_obs:
        f = p
        if flags & F_OBS_RESCALE:
            f = rescale(f)
        if flags & F_OBS_SORT_COORD0:
            f = get_features_torch(f)
        elif flags & F_OBS_SORT_LEX:
            f = order_and_rescale(f.reshape(B, -1), (N, d), scale_observation=False).reshape(B, N, d)
        obs = f.reshape(B, N * d)
        if obs_coord is not None:
            oc = np.asarray(obs_coord)
            cb = decode_table(d)[oc].astype(np.float32) if flags & F_ACT_DISCRETE else masks_to_binary(oc, d)
            obs = np.concatenate([obs, cb], axis=1)
    return p, done, rew, get_num_points(p), obs
