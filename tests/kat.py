"""Known-answer vectors of the reference's own unit tests, restated as NumPy data.

Sources (paths relative to the reference root):
  R_IN, R, R2, R3, RS       test/testTensorPoints.py:10-46,48-82  ==  test/testJAX.py:42-101
  ORIGIN20_IN / _OUT        test/testTensorPoints.py:113-172
  FEAT_IN / FEAT_SORTED     test/testJAX.py:102-131 (host feature lexsort)
  AGENT_FEAT_*              test/testJAX.py:147-153
  DECODE_3, ENCODE_*        test/testJAX.py:205-219, test/testUtil.py:226-236
"""
import numpy as np

f32 = np.float32

R_IN = np.array([[[1, 2, 3, 4], [2, 3, 4, 5], [4, 1, 2, 3], [1, 6, 7, 3]],
                 [[0, 1, 3, 5], [1, 1, 1, 1], [9, 8, 2, 1], [-1, -1, -1, -1]]], dtype=f32)
R = np.array([[[1, 2, 3, 4], [-1, -1, -1, -1], [4, 1, 2, 3], [1, 6, 7, 3]],
              [[0, 1, 3, 5], [1, 1, 1, 1], [-1, -1, -1, -1], [-1, -1, -1, -1]]], dtype=f32)
# shift with coords [[1,2],[0,2,3]] (multi-binary [[0,1,1,0],[1,0,1,1]]) and axis [1,3]
R_COORD_LISTS = [[1, 2], [0, 2, 3]]
R_COORD_BIN = np.array([[0, 1, 1, 0], [1, 0, 1, 1]], dtype=f32)
R_COORD_MASK = np.array([0b0110, 0b1101], dtype=np.int32)
R_AXIS = np.array([1, 3], dtype=np.int32)
R2 = np.array([[[1, 5, 3, 4], [-1, -1, -1, -1], [4, 3, 2, 3], [1, 13, 7, 3]],
               [[0, 1, 3, 8], [1, 1, 1, 3], [-1, -1, -1, -1], [-1, -1, -1, -1]]], dtype=f32)
R3 = np.array([[[0, 2, 1, 1], [-1, -1, -1, -1], [3, 0, 0, 0], [0, 10, 5, 0]],
               [[0, 0, 2, 5], [1, 0, 0, 0], [-1, -1, -1, -1], [-1, -1, -1, -1]]], dtype=f32)
RS = np.array([[[0.0, 0.2, 0.1, 0.1], [-1, -1, -1, -1], [0.3, 0, 0, 0], [0.0, 1.0, 0.5, 0.0]],
               [[0.0, 0.0, 0.4, 1.0], [0.2, 0, 0, 0], [-1, -1, -1, -1], [-1, -1, -1, -1]]], dtype=f32)

# invalid actions / ended games: test/testTensorPoints.py:84-103
INVALID_COORD_LISTS = [[1], [0, 2, 3]]
INVALID_AXIS = [0, 1]
ENDED_P = np.array([[[1, 0, 0, 1], [-1, -1, -1, -1]]], dtype=f32)
ENDED_Q = np.array([[[1, 1, 0, 1], [-1, -1, -1, -1]]], dtype=f32)

# remove_repeated: test/testTensorPoints.py:105-111; duplicate rows: test/testJAX.py:87-97
REP_IN = np.array([[[0, 0, 0], [0, 0, 0]]], dtype=f32)
REP_OUT = np.array([[[0, 0, 0], [-1, -1, -1]]], dtype=f32)
EXTREME_IN = np.array([[[1, 1, 1], [1, 1, 1]]], dtype=f32)
EXTREME_OUT = np.array([[[1, 1, 1], [-1, -1, -1]]], dtype=f32)

ORIGIN20_IN = np.array([[[4, 2, 4], [4, 0, 3], [3, 2, 4], [3, 3, 3], [3, 4, 2], [3, 0, 0], [3, 1, 2], [3, 0, 1],
                         [3, 3, 3], [3, 0, 3], [2, 0, 1], [2, 1, 3], [2, 1, 1], [1, 4, 2], [1, 4, 1], [1, 4, 3],
                         [1, 0, 4], [1, 3, 4], [0, 0, 0], [0, 0, 0]]], dtype=f32)
ORIGIN20_OUT = np.full((1, 20, 3), -1, dtype=f32)
ORIGIN20_OUT[0, 18] = 0

# rescale: test/testJAX.py:88-100
RESCALE_IN = np.array([[[1, 2, 3], [2, 3, 4], [-1, -1, -1]], [[0, 0, 0], [-1, -1, -1], [-1, -1, -1]]], dtype=f32)
RESCALE_OUT = np.array([[[0.25, 0.5, 0.75], [0.5, 0.75, 1.0], [-1, -1, -1]],
                        [[0, 0, 0], [-1, -1, -1], [-1, -1, -1]]], dtype=f32)
# rescale by zero stays finite: test/testTensorPoints.py:174-183
RESCALE0_IN = np.array([[[0, 0, 0, 0], [-1, -1, -1, -1], [-1, -1, -1, -1], [-1, -1, -1, -1]],
                        [[0, 1, 3, 5], [1, 1, 1, 1], [9, 8, 2, 1], [-1, -1, -1, -1]]], dtype=f32)

FEAT_IN = np.array([[[0.10526316, 0.2631579, 0.0], [0.0, 0.8947368, 0.15789473], [-1, -1, -1],
                     [0.10526316, 0.21052632, 1.0], [-1, -1, -1], [0.84210527, 0.05263158, 0.47368422]]], dtype=f32)
FEAT_SORTED = np.array([[0.10526316, 0.21052632, 1.0, 0.84210527, 0.05263158, 0.47368422, 0.0, 0.8947368,
                         0.15789473, 0.10526316, 0.2631579, 0.0, -1, -1, -1, -1, -1, -1]], dtype=f32)
# float shift: test/testJAX.py:132-144 (coords [[0,1,1]], axis [1])
FSHIFT_OUT = np.array([[[0.10526316, 0.2631579, 0.0], [0.0, 1.0526316, 0.15789473], [-1, -1, -1],
                        [0.10526316, 1.2105263, 1.0], [-1, -1, -1], [0.84210527, 0.5263158, 0.47368422]]], dtype=f32)

AGENT_FEAT_IN = np.array([[1, 2, 3, -1, -1, -1, 2, 3, 4, 0, 1, 1], [-1, -1, -1, 0, 0, 1, 0, 1, 1, 1, 0, 1]], dtype=f32)
AGENT_FEAT_NOSCALE = np.array([[2, 3, 4, 1, 2, 3, -1, -1, -1, 0, 1, 1], [0, 1, 1, 0, 0, 1, -1, -1, -1, 1, 0, 1]], dtype=f32)
AGENT_FEAT_SCALE = np.array([[0.5, 0.75, 1.0, 0.25, 0.5, 0.75, -1, -1, -1, 0, 1, 1],
                             [0, 1, 1, 0, 0, 1, -1, -1, -1, 1, 0, 1]], dtype=f32)

DECODE_3 = np.array([[1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]], dtype=np.int32)
ENCODE_IN = np.array([[1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.int32)
ENCODE_OUT = np.array([1, 3, 2])
ENCODE_ONE_HOT_OUT = np.array([[0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=f32)

# take_actions composition: test/testJAX.py:461-488
TA_HOST_OBS = np.array([[[1, 2, 3], [2, 3, 4], [0, 9, 0], [-1, -1, -1]],
                        [[4, 2, 2], [-1, -1, -1], [0, 0, 1], [-1, -1, -1]]], dtype=f32)
TA_COORDS = np.array([[0, 1, 1], [1, 1, 1]], dtype=f32)
TA_COORD_MASK = np.array([0b110, 0b111], dtype=np.int32)

# fixed players: test/testJAX.py:232-276
HOSTS_OBS = np.array([[[1, 2, 3], [2, 3, 4]], [[0, 1, 2], [-1, -1, -1]]], dtype=f32)
ZEIL_PTS = np.array([[0, 0, 4], [5, 0, 1], [1, 5, 1], [0, 25, 0]], dtype=f32)
ZEIL_OUT = np.array([0, 1, 0, 0], dtype=f32)
ZEIL_OBS2 = np.array([[[19, 15, 0, 10], [12, 0, 14, 9], [8, 14, 8, 18], [3, 18, 17, 12], [19, 6, 1, 13]],
                      [[17, 3, 6, 9], [19, 1, 13, 12], [14, 0, 6, 7], [2, 15, 3, 16], [0, 16, 1, 5]],
                      [[19, 0, 8, 6], [8, 9, 17, 1], [2, 3, 7, 14], [6, 19, 9, 12], [0, 19, 19, 14]]], dtype=f32)
ZEIL_OBS2_MB = np.array([[0, 1, 0, 1], [1, 0, 1, 0], [1, 0, 1, 0]], dtype=np.int32)
ZEIL_PTS3 = np.full((1, 10, 3), -1, dtype=f32)
ZEIL_PTS3[0, 2] = [259, 5, 5]
ZEIL_PTS3[0, 4] = [841, 17, 0]
ZEIL_PTS3[0, 9] = [147, 3, 12]
AGENT_COORDS = np.array([[1, 1, 0], [0, 1, 1]], dtype=f32)
CHOOSE_FIRST_OUT = np.array([[1, 0, 0], [0, 1, 0]], dtype=f32)
CHOOSE_LAST_OUT = np.array([[0, 1, 0], [0, 0, 1]], dtype=f32)

# rollout value targets: test/testJAXTrainer.py:91-389 (discount 0.99, dimension 3, N 5).  Point counts per
# step are what rollout_postprocess recovers from the observations (jax_trainer.py:584).
VALUE_KATS = [
    # (num_points [B,T], role, use_unified_tree, expected values)
    (np.array([[2, 2, 2, 2, 1], [2, 2, 2, 2, 2]]), "agent", False,
     np.array([[-0.970299, -0.9801, -0.98999995, -1.0, -1.0], [-0.480298, -0.4851495, -0.49005002, -0.495, -0.5]], f32)),
    (np.array([[3, 3, 3, 3, 2]]), "agent", True, np.array([[-0.480298, 0.4851495, -0.49005002, 0.495, -0.5]], f32)),
    (np.array([[3, 3, 3, 2]]), "host", True, np.array([[0.4851495, -0.49005002, 0.495, -0.5]], f32)),
    (np.array([[2, 1, 1, 1]]), "agent", True, np.array([[-1, 1, -1, 1]], f32)),
    (np.array([[2, 2, 1, 1]]), "host", True, np.array([[0.99, -1, 1, -1]], f32)),
]
# the observation form of the last two (unified tree: host rows are zero-padded with d entries)
VALUE_OBS_AGENT = np.array([[[-1, -1, -1, -1, -1, -1, 1, 1, 14, -1, -1, -1, 4, 1, 3, 1, 1, 0],
                             [-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 4, 5, 3, 0, 0, 0],
                             [-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 4, 5, 3, 1, 0, 1],
                             [-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 4, 5, 3, 0, 0, 0]]], dtype=f32)
VALUE_OBS_HOST = np.array([[[-1, -1, -1, -1, -1, -1, 1, 1, 14, -1, -1, -1, 4, 1, 3, 0, 0, 0],
                            [-1, -1, -1, -1, -1, -1, 1, 1, 14, -1, -1, -1, 4, 5, 3, 1, 1, 1],
                            [-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 4, 5, 3, 0, 0, 0],
                            [-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 4, 5, 3, 1, 1, 1]]], dtype=f32)

# BASELINE config 4: the six surface-singularity (ADE) start configurations of hironaka/jax/search.py:123-128
ADE_STARTS = np.array([
    [[3, 0, 0], [0, 5, 0], [0, 0, 2]],
    [[2, 0, 0], [0, 3, 0], [0, 0, 3]],
    [[2, 0, 0], [0, 3, 0], [0, 0, 4]],
    [[2, 0, 0], [0, 2, 1], [0, 0, 5]],
    [[2, 0, 0], [0, 2, 0], [0, 0, 4]],
    [[3, 0, 0], [0, 5, 0], [0, 2, 2]],
], dtype=f32)
