"""Hot-path constants of fast_rcnn/config.py (reference lines cited per entry).

Only the values read on the hot path are mirrored; the reference's yaml/CLI override
machinery (config.py:384-412) is a control-plane feature and out of scope.  `cfg` is an
attribute dict so `cfg.TEST.RPN_NMS_THRESH` and `cfg['TEST'].RPN_NMS_THRESH` both work,
as with easydict.
"""


class _AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


cfg = _AttrDict()
cfg.TRAIN = _AttrDict(
    IMS_PER_BATCH=1,                 # config.py:115
    WS_IMS_PER_BATCH=2,              # config.py:49
    BATCH_SIZE=128,                  # config.py:118
    FG_FRACTION=0.25,                # config.py:121
    FG_THRESH=0.5,                   # config.py:124
    BG_THRESH_HI=0.5,                # config.py:128
    BG_THRESH_LO=0.0,                # config.py:129
    BBOX_NORMALIZE_TARGETS_PRECOMPUTED=False,
    BBOX_NORMALIZE_MEANS=(0.0, 0.0, 0.0, 0.0),
    BBOX_NORMALIZE_STDS=(0.1, 0.1, 0.2, 0.2),
    BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0),
    RPN_POSITIVE_OVERLAP=0.7,        # config.py:196
    RPN_NEGATIVE_OVERLAP=0.3,        # config.py:198
    RPN_CLOBBER_POSITIVES=False,     # config.py:200
    RPN_FG_FRACTION=0.5,             # config.py:202
    RPN_BATCHSIZE=256,               # config.py:204
    RPN_NMS_THRESH=0.7,              # config.py:206
    RPN_PRE_NMS_TOP_N=12000,         # config.py:208
    RPN_POST_NMS_TOP_N=2000,         # config.py:210
    RPN_MIN_SIZE=16,                 # config.py:212
    RPN_BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0),
    RPN_POSITIVE_WEIGHT=-1.0,
)
cfg.TEST = _AttrDict(
    NMS=0.3,                         # config.py:238
    CLS_AGNOSTIC_NMS=False,
    RPN_NMS_THRESH=0.7,              # config.py:257
    RPN_PRE_NMS_TOP_N=6000,          # config.py:259
    RPN_POST_NMS_TOP_N=300,          # config.py:261
    RPN_MIN_SIZE=16,                 # config.py:265
)
cfg.MAX_GT_PER_IMAGE = 20            # config.py:92
cfg.RNG_SEED = 3                     # config.py:290
cfg.EPS = 1e-14
cfg.USE_GPU_NMS = False              # config.py:321 -- selects cpu_nms *semantics* (>=, f64)
cfg.GPU_ID = 0
