"""The C-ABI library loads and exports every symbol include/uu3d.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from uplift_upsample_3dhpe_b200 import _lib, stride_mask

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "uu3d.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(uu_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/uu3d.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.uu_version() >= 100
    # stride-mask rule through the C ABI is bit-exact with the numpy rule, negative indices included
    from uplift_upsample_3dhpe_b200.model import host_stride_mask
    for n_tok, so, si in ((71, 5, 5), (71, 5, 10), (71, 5, 20), (41, 2, 4), (41, 2, 10), (41, 2, 20)):
        for shift in (-1003, -7, -1, 0, 1, 3, 5, 23, 20000):
            assert np.array_equal(host_stride_mask(n_tok, so, si, shift),
                                  stride_mask.stride_mask(n_tok, so, si, center_frame=shift))
    with pytest.raises(_lib.UUError):
        host_stride_mask(71, 5, 4, 0)


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig
    from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer
    with pytest.raises(_lib.UUError, match="no CUDA device"):
        build_uplift_upsample_transformer(UpliftUpsampleConfig.preset("h36m_81"))
