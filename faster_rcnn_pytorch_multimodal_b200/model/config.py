"""The slice of the reference's global ``cfg`` that the detection glue reads.

Mirrors ``lib/model/config.py`` (line numbers cited per key) with the same attribute
paths (``cfg.TRAIN.RPN_NMS_THRESH`` ...), so code written against the reference's
``from model.config import cfg`` keeps working.  Only keys on the hot path exist here.
"""
import numpy as np


class AttrDict(dict):
    """Minimal attribute-access dict (the reference uses easydict.EasyDict, config.py:9)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


__C = AttrDict()
cfg = __C

__C.NET_TYPE = 'lidar'                       # config.py:54
__C.DB_NAME = ''                             # config.py:57; the CLIs set it (waymo / kitti / cadc / nuscenes)
__C.USE_FPN = False                          # config.py:51
__C.RNG_SEED = 3                             # config.py:346
__C.POOLING_MODE = 'align'                   # config.py:364
__C.POOLING_SIZE = 7                         # config.py:367
# RoIAlign samples per bin and axis.  Not a key of the reference's config: its value lives in the missing
# lib/nets/network.py; 2 is the upstream default the fork descends from.  <= 0 = adaptive.
__C.POOLING_SAMPLING_RATIO = 2
__C.ANCHOR_SCALES = [2, 4, 8, 16, 32]        # config.py:373
__C.ANCHOR_RATIOS = [0.5, 0.75, 1, 1.25, 2]  # config.py:378

__C.TRAIN = AttrDict()
__C.TRAIN.USE_GT = False                     # config.py:108
__C.TRAIN.ROI_BATCH_SIZE = 256               # config.py:123
__C.TRAIN.FG_FRACTION = 0.25                 # config.py:126
__C.TRAIN.FG_THRESH = 0.6                    # config.py:129
__C.TRAIN.DC_THRESH = 0.5                    # config.py:130
__C.TRAIN.BG_THRESH_HI = 0.5                 # config.py:133
__C.TRAIN.BG_THRESH_LO = 0.0                 # config.py:134
__C.TRAIN.BBOX_INSIDE_WEIGHTS = (1.0, 1.0, 1.0, 1.0)          # config.py:157
__C.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED = True           # config.py:161
__C.TRAIN.RPN_POSITIVE_OVERLAP = 0.7         # config.py:174
__C.TRAIN.RPN_NEGATIVE_OVERLAP = 0.3         # config.py:177
__C.TRAIN.RPN_CLOBBER_POSITIVES = False      # config.py:180
__C.TRAIN.RPN_FG_FRACTION = 0.5              # config.py:183
__C.TRAIN.RPN_BATCHSIZE = 256                # config.py:186
__C.TRAIN.RPN_NMS_THRESH = 0.7               # config.py:189
__C.TRAIN.RPN_PRE_NMS_TOP_N = 12000          # config.py:192
__C.TRAIN.RPN_POST_NMS_TOP_N = 2000          # config.py:195
__C.TRAIN.RPN_BBOX_INSIDE_WEIGHTS = (1.0, 1.0, 1.0, 1.0)      # config.py:198
__C.TRAIN.RPN_POSITIVE_WEIGHT = -1.0         # config.py:203
__C.TRAIN.IGNORE_DC = False                  # config.py:210
__C.TRAIN.LIDAR = AttrDict()
__C.TRAIN.IMAGE = AttrDict()
__C.TRAIN.LIDAR.BBOX_NORMALIZE_MEANS = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)   # config.py:219
__C.TRAIN.LIDAR.BBOX_NORMALIZE_STDS = (0.1, 0.1, 0.1, 0.2, 0.2, 0.2, 1.0)    # config.py:220
__C.TRAIN.IMAGE.BBOX_NORMALIZE_MEANS = (0.0, 0.0, 0.0, 0.0)                  # config.py:222
__C.TRAIN.IMAGE.BBOX_NORMALIZE_STDS = (0.1, 0.1, 0.2, 0.2)                   # config.py:223

__C.TEST = AttrDict()
__C.TEST.NMS_THRESH = 0.6                    # config.py:234
__C.TEST.RPN_NMS_THRESH = 0.7                # config.py:250
__C.TEST.RPN_PRE_NMS_TOP_N = 6000            # config.py:253
__C.TEST.RPN_POST_NMS_TOP_N = 300            # config.py:256
__C.TEST.MODE = 'nms'                        # config.py:263
__C.TEST.RPN_TOP_N = 5000                    # config.py:266
__C.TEST.IGNORE_DC = False                   # config.py:268

__C.UC = AttrDict()
__C.UC.EN_BBOX_ALEATORIC = False             # config.py:39
__C.UC.EN_CLS_ALEATORIC = False              # config.py:40
__C.UC.EN_BBOX_EPISTEMIC = False             # config.py:41
__C.UC.EN_CLS_EPISTEMIC = False              # config.py:43
__C.UC.E_NUM_SAMPLE = 10                     # config.py:46
__C.UC.SORT_TYPE = ''                        # config.py:47

__C.DEBUG = AttrDict()
__C.DEBUG.EN_TEST_MSG = False                # config.py:31 (reference default True; it only prints)

__C.LIDAR = AttrDict()
__C.LIDAR.X_RANGE = [0, 70]                  # config.py:397
__C.LIDAR.Y_RANGE = [-40, 40]                # config.py:398
__C.LIDAR.Z_RANGE = [-3, 3]                  # config.py:399
__C.LIDAR.VOXEL_LEN = 0.1                    # config.py:400
__C.LIDAR.VOXEL_HEIGHT = 0.5                 # config.py:401
__C.LIDAR.NUM_SLICES = 12                    # config.py:402
__C.LIDAR.NUM_META_CHANNEL = 3               # config.py:403
__C.LIDAR.NUM_CHANNEL = __C.LIDAR.NUM_SLICES + __C.LIDAR.NUM_META_CHANNEL   # config.py:404
__C.LIDAR.MAX_PTS_PER_VOXEL = 32             # config.py:405
__C.LIDAR.MAX_NUM_VOXEL = 25000              # config.py:406
__C.LIDAR.ANCHORS = np.array([[4.73, 2.08, 1.77]])            # config.py:421
__C.LIDAR.ANCHOR_SCALES = np.array([[1]])    # config.py:422
__C.LIDAR.ANCHOR_ANGLES = np.array([0, np.pi / 2])            # config.py:423
__C.LIDAR.NUM_BBOX_ELEM = 7                  # config.py:425
__C.IMAGE = AttrDict()
__C.IMAGE.NUM_BBOX_ELEM = 4                  # config.py:429

# Fork-independent switch (not in the reference): 'strict' reproduces the reference as it
# runs (never any background RoI); 'intended' samples [LO, HI) background.
__C.TRAIN.BG_MODE = 'strict'
