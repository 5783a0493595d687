"""TEST INFRASTRUCTURE: NumPy restatement of the reference's gym environments on a batch
(hironaka/gym_env/hironaka_agent_env.py:44-83, hironaka_host_env.py:38-81), built on the oracle's
step / features.  It is checked against fixtures generated from the REAL environments
(oracle/gen_golden_gym.py -> tests/golden/ref_gym_*.npz) and then serves as the checker of the CUDA
environments on batches the reference cannot play in reasonable time."""
import numpy as np

from oracle import cport
from oracle import hk_oracle as O

F_LEX_FIRST = 1 << 12


def observe(state, scale):
    B, N, d = state.shape
    return cport.features(state, F_LEX_FIRST | (O.F_OBS_RESCALE if scale else 0)).reshape(B, N, d)


def multibinary_to_mask(a):
    return (a.astype(np.int64) << np.arange(a.shape[1])).sum(1).astype(np.int32)


def zeillinger_mask(listed):
    """hironaka/host.py:50-92 on one game given in ListPoints order (rows with coordinate 0 >= 0)."""
    pts = [p for p in listed.tolist() if p[0] >= 0]
    if len(pts) <= 1:
        return 0
    best = None
    for i in range(len(pts)):
        for j in range(i + 1, len(pts)):
            v = [pts[i][k] - pts[j][k] for k in range(len(pts[i]))]
            mx, mn = max(v), min(v)
            key = (mx - mn, sum(x == mx for x in v) + sum(x == mn for x in v))
            if best is None or key < best[0]:
                best = (key, v)
    r = [int(np.argmin(best[1])), int(np.argmax(best[1]))]
    return (1 << r[0]) | (1 << r[1]) if r[0] != r[1] else 0b11


class AgentEnv:
    def __init__(self, N, d, scale_observation=True, value_threshold=None, step_threshold=1000,
                 fixed_penalty_crossing_threshold=None, stop_at_threshold=True, reward_based_on_point_reduction=False):
        self.__dict__.update(locals())

    def reset(self, points):
        self.state = cport.step(np.ascontiguousarray(points, dtype=np.int32), None, None, O.OP_NEWTON, 0)[0]
        self.cur = np.zeros(len(points), np.int64)
        return observe(self.state, self.scale_observation)

    def step(self, mask):
        self.cur += 1
        before = (self.state[:, :, 0] >= 0).sum(1)
        nbits = np.array([bin(int(m)).count("1") for m in mask])
        first = np.array([(int(m) & -int(m)).bit_length() - 1 if m else 0 for m in mask])
        axis = np.where(nbits > 1, first, -1).astype(np.int32)  # ChooseFirstAgent: min(coord) if len > 1 else None
        self.state = cport.step(self.state, mask.astype(np.int32), axis, O.OP_SHIFT | O.OP_NEWTON, O.F_NOOP_INVALID)[0]
        after = (self.state[:, :, 0] >= 0).sum(1)
        ended = after <= 1
        reward = np.zeros(len(mask))
        stopped = ended.copy()
        exceed = np.zeros(len(mask), bool) if self.value_threshold is None else \
            self.state.reshape(len(mask), -1).max(1) > self.value_threshold
        if self.stop_at_threshold:
            hit = (self.cur >= self.step_threshold) | exceed
            stopped |= hit
            reward += hit * (-self.step_threshold if self.fixed_penalty_crossing_threshold is None
                             else self.fixed_penalty_crossing_threshold)
        if self.reward_based_on_point_reduction:
            reward += before - after
        reward += ended
        return observe(self.state, self.scale_observation), reward, stopped


class HostEnv:
    def __init__(self, N, d, host="Zeillinger", scale_observation=True, value_threshold=None, invalid_move_penalty=-1e-3,
                 stop_after_invalid_move=False):
        self.__dict__.update(locals())

    def _relist(self):
        self.state = observe(self.state, False).astype(np.int32)

    def reset(self, points):
        self.state = cport.step(np.ascontiguousarray(points, dtype=np.int32), None, None, O.OP_NEWTON, 0)[0]
        self._relist()
        self.coords = np.zeros(len(points), np.int32)
        self.step(None)
        return observe(self.state, self.scale_observation), self.coords.copy()

    def step(self, action):
        B = len(self.state)
        if action is None:
            valid, axis = np.zeros(B, bool), np.full(B, -1, np.int32)
        else:
            axis = np.asarray(action, np.int32)
            valid = ((self.coords >> axis) & 1).astype(bool)
        self.state = cport.step(self.state, self.coords, axis, O.OP_SHIFT | O.OP_NEWTON, O.F_NOOP_INVALID)[0]
        self._relist()
        ended = (self.state[:, :, 0] >= 0).sum(1) <= 1
        reward = np.where(valid, (~ended).astype(float), self.invalid_move_penalty)
        stopped = ended.copy()
        if self.stop_after_invalid_move:
            stopped |= ~valid
        if self.value_threshold is not None:
            stopped |= self.state.reshape(B, -1).max(1) > self.value_threshold
        if self.host == "Zeillinger":
            choice = np.array([zeillinger_mask(s) for s in self.state], np.int32)
        else:
            choice = np.full(B, (1 << self.d) - 1, np.int32)
        self.coords = np.where(stopped, 0, choice).astype(np.int32)
        return observe(self.state, self.scale_observation), self.coords.copy(), reward, stopped
