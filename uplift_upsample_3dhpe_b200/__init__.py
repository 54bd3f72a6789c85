"""B200-native uplift-and-upsample 3-D human pose transformer (hot path of
goldbricklemon/uplift-upsample-3dhpe: common/net forward pass + training step).

Host side mirrors the reference's Python interface (config keys, call contract,
weight layout); all arithmetic runs in hand-written sm_100a CUDA kernels behind
the C-ABI declared in include/uu3d.h.
"""
from .config import UpliftUpsampleConfig, PRESETS            # noqa: F401
from .spec import ModelSpec, spec_from_config, forward_macs   # noqa: F401
from . import stride_mask, weights, h5lite                    # noqa: F401

__version__ = "0.1.0"
