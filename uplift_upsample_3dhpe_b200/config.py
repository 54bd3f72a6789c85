"""Config contract of the hot path.

Accepts the reference's ``config/*.json`` files unchanged: every key is set as an
attribute over class-level defaults, exactly like the reference loader
(reference: common/utils/config.py:49-94, defaults
common/net/uplift_upsample_transformer_config.py:13-106).  Only the keys the
hot path consumes carry defaults here; unknown keys are kept verbatim so a
reference JSON round-trips through :meth:`dump`.

``PRESETS`` restates the hyper-parameters of the four shipped reference configs
(config/h36m_81.json, h36m_351.json, h36m_351_pt.json, amass_351.json) as diffs
against the defaults so tests and benches do not need the reference tree.
"""
from __future__ import annotations

import copy
import json
import os
from typing import Any, Dict

# Defaults of the keys the forward pass / training step read
# (common/net/uplift_upsample_transformer_config.py:13-106).
_DEFAULTS: Dict[str, Any] = dict(
    GPU_ID=0,
    BATCH_SIZE=256,
    ARCH="UpliftUpsampleTransformer",
    SPATIAL_EMBED_DIM=32,
    TEMPORAL_EMBED_DIM=348,
    MLP_RATIO=2,
    NUM_HEADS=8,
    SPATIAL_TRANSFORMER_BLOCKS=4,
    TEMPORAL_TRANSFORMER_BLOCKS=4,
    STRIDES=[3, 3, 3],
    PADDINGS=None,
    QKV_BIAS=True,
    DROP_PATH_RATE=[0.1, 0.1, 0.0],
    DROP_RATE=0.0,
    ATTENTION_DROP_RATE=0.0,
    OUTPUT_BN=False,
    USE_REFINE=False,
    TOKEN_MASK_RATE=0.0,
    LEARNABLE_MASKED_TOKEN=False,
    NUM_KEYPOINTS=17,
    SEQUENCE_LENGTH=27,
    SEQUENCE_STRIDE=1,
    MASK_STRIDE=None,
    STRIDE_MASK_RAND_SHIFT=False,
    FIRST_STRIDED_TOKEN_ATTENTION_LAYER=0,
    LOSS_WEIGHT_SEQUENCE=1.0,
    LOSS_WEIGHT_CENTER=1.0,
    ROOT_KEYTPOINT=6,  # (sic) the reference's spelling
    OPTIMIZER="Adam",
    OPTIMIZER_PARAMS={"amsgrad": True, "epsilon": 1e-08},
    SCHEDULE="ExponentialDecayWithSteps",
    SCHEDULE_PARAMS={
        "initial_learning_rate": 1e-3,
        "decay_steps": 12000,
        "decay_rate": 0.95,
        "large_decay_steps": 60000,
        "large_decay_rate": 0.5,
    },
    WEIGHT_DECAY=None,
    EMA_ENABLED=False,
    EMA_DECAY=None,
    # evaluation glue (eval.py:154-222)
    PADDING_TYPE="copy",
    TEST_STRIDED_EVAL=True,
    EVAL_FLIP=True,
    # our-17-point order, left/right swapped (config/h36m_351.json:4-22, identical in the four shipped configs)
    AUGM_FLIP_KEYPOINT_ORDER=[5, 4, 3, 2, 1, 0, 6, 7, 8, 9, 10, 16, 15, 14, 13, 12, 11],
)

_COMMON_351_81 = dict(
    TEMPORAL_EMBED_DIM=384,
    FIRST_STRIDED_TOKEN_ATTENTION_LAYER=1,
    LOSS_WEIGHT_CENTER=0.5,
    LOSS_WEIGHT_SEQUENCE=0.5,
    OPTIMIZER="AdamW",
    OPTIMIZER_PARAMS={},
    SCHEDULE="ExponentialDecay",
    SCHEDULE_PARAMS={"decay_rate": 0.99, "decay_steps": 6000,
                     "initial_learning_rate": 4e-05, "staircase": True},
    STRIDE_MASK_RAND_SHIFT=True,
    WEIGHT_DECAY=4e-06,
)

PRESETS: Dict[str, Dict[str, Any]] = {
    # config/h36m_351.json (N=351 frames: 71 tokens x stride 5)
    "h36m_351": dict(_COMMON_351_81, BATCH_SIZE=512, MASK_STRIDE=[5, 10, 20],
                     PADDINGS=[[0, 0], [0, 0], [0, 0]], SEQUENCE_LENGTH=71,
                     SEQUENCE_STRIDE=5, STRIDES=[3, 10, 3]),
    # config/h36m_81.json (N=81 frames: 41 tokens x stride 2)
    "h36m_81": dict(_COMMON_351_81, BATCH_SIZE=256, MASK_STRIDE=[4, 10, 20],
                    PADDINGS=[[1, 1], [0, 0], [0, 0]], SEQUENCE_LENGTH=41,
                    SEQUENCE_STRIDE=2, STRIDES=[4, 4, 3], EMA_ENABLED=True, EMA_DECAY=0.999),
}
# config/amass_351.json: architecture and optimiser identical to h36m_351.
PRESETS["amass_351"] = dict(PRESETS["h36m_351"])
# config/h36m_351_pt.json: fine-tuning schedule (lr 2e-5, wd 2e-6).
PRESETS["h36m_351_pt"] = dict(
    PRESETS["h36m_351"], WEIGHT_DECAY=2e-06,
    SCHEDULE_PARAMS={"decay_rate": 0.99, "decay_steps": 6000,
                     "initial_learning_rate": 2e-05, "staircase": True})


class UpliftUpsampleConfig:
    """Attribute bag with the reference's key names.

    ``UpliftUpsampleConfig("config/h36m_351.json")`` loads a reference JSON (or
    the reference's "KEY json-value" txt format); ``UpliftUpsampleConfig.preset(name)``
    builds one of the shipped configurations without the file.
    """

    def __init__(self, config_file: str | None = None, file_mode: str | None = None, **overrides):
        for k, v in _DEFAULTS.items():
            setattr(self, k, copy.deepcopy(v))
        if config_file is not None:
            self.load(config_file, file_mode)
        for k, v in overrides.items():
            setattr(self, k, v)

    @classmethod
    def preset(cls, name: str, **overrides) -> "UpliftUpsampleConfig":
        if name not in PRESETS:
            raise KeyError(f"unknown preset {name!r}; have {sorted(PRESETS)}")
        cfg = cls()
        for k, v in PRESETS[name].items():
            setattr(cfg, k, copy.deepcopy(v))
        for k, v in overrides.items():
            setattr(cfg, k, v)
        return cfg

    def load(self, config_file: str, file_mode: str | None = None) -> None:
        if not os.path.exists(config_file):
            raise FileNotFoundError(config_file)
        if file_mode is None:
            ext = os.path.splitext(config_file)[1]
            if ext not in (".txt", ".json"):
                raise ValueError(f"config extension must be .json or .txt, got {ext!r}")
            file_mode = "txt" if ext == ".txt" else "json"
        if file_mode == "txt":
            with open(config_file, "r") as f:
                for line in f:
                    line = line.strip("\r\n ")
                    if not line or line.startswith("#"):
                        continue
                    parts = line.split(" ", 1)
                    if len(parts) < 2 or not parts[1].strip():
                        continue
                    setattr(self, parts[0], json.loads(parts[1].strip().replace("'", '"')))
        else:
            with open(config_file, "r") as f:
                for key, value in json.load(f).items():
                    setattr(self, key, value)

    def as_dict(self) -> Dict[str, Any]:
        return {k: v for k, v in vars(self).items() if not k.startswith("_")}

    def dump(self, config_file: str) -> None:
        with open(config_file, "w") as f:
            json.dump(self.as_dict(), f, indent=4, sort_keys=True)

    def copy(self) -> "UpliftUpsampleConfig":
        new = self.__class__()
        for k, v in self.as_dict().items():
            setattr(new, k, copy.deepcopy(v))
        return new
