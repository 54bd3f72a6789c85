"""Stride-mask rule of the data generators — the bit-exact integer contract.

reference: common/dataset/uplifiting_dataset.py:329-339 (mask stride selection),
:377-394 (index arithmetic; numpy floor-mod on negative indices).

The mask is 1 (True) on tokens that carry a 2-D pose.  ``s_out`` is
SEQUENCE_STRIDE and ``s_in`` the absolute MASK_STRIDE, both in video frames.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np


def check_strides(s_out: int, s_in: int) -> None:
    """uplifiting_dataset.py:252-254."""
    if s_in < s_out or s_in % s_out != 0:
        raise ValueError(f"MASK_STRIDE {s_in} must be a multiple of SEQUENCE_STRIDE {s_out}")


def stride_mask(n_tok: int, s_out: int, s_in: int, *, center_frame: int | None = None,
                shift_tokens: int = 0) -> np.ndarray:
    """One window's mask.

    center_frame: global-alignment mode (eval; :381-384) — the window is centred on
        video frame ``center_frame`` and the mask is aligned to global frame indices.
    shift_tokens: training mode (:386-392) — mask shifted by ``shift_tokens * s_out`` frames.
    """
    check_strides(s_out, s_in)
    idx = (np.arange(n_tok, dtype=np.int64) - n_tok // 2) * s_out
    if center_frame is not None:
        idx = idx + int(center_frame)
    else:
        idx = idx + int(shift_tokens) * s_out
    return np.equal(idx % s_in, 0)          # numpy % is floor-mod


def rand_shift_range(mask_stride_tokens: int):
    """(low, high, endpoint) of the training-time random shift (:388-390)."""
    max_shift = int(math.ceil((mask_stride_tokens - 1) / 2))
    return -max_shift, max_shift, mask_stride_tokens % 2 != 0


def batch_stride_masks_eval(n_tok: int, s_out: int, s_in: int, center_frames: Sequence[int]) -> np.ndarray:
    """Globally aligned masks for windows centred on the given frames -> bool (B, n_tok)."""
    return np.stack([stride_mask(n_tok, s_out, s_in, center_frame=int(i)) for i in center_frames])


def batch_stride_masks_train(n_tok: int, s_out: int, mask_strides: Sequence[int], batch: int, seed: int,
                             rand_shift: bool = True) -> np.ndarray:
    """Training-mode masks: per window draw an absolute mask stride from the list and a random
    shift, with two PCG64 streams seeded alike as the generator does (:318-319, :335-337, :389-391)."""
    shift_rng = np.random.default_rng(seed=seed)
    stride_rng = np.random.default_rng(seed=seed)
    out = np.zeros((batch, n_tok), dtype=bool)
    for b in range(batch):
        if len(mask_strides) == 1:
            s_in = int(mask_strides[0])
        else:
            s_in = int(mask_strides[stride_rng.integers(low=0, high=len(mask_strides), endpoint=False)])
        shift = 0
        if rand_shift:
            lo, hi, endpoint = rand_shift_range(s_in // s_out)
            shift = int(shift_rng.integers(low=lo, high=hi, endpoint=endpoint))
        out[b] = stride_mask(n_tok, s_out, s_in, shift_tokens=shift)
    return out


def window_source_frames(n_tok: int, s_out: int, T: int, center_frames: Sequence[int], pad_copy: bool = True) -> np.ndarray:
    """Source video frame of every token of the sliding windows centred on ``center_frames`` -> int32 (B, n_tok);
    -1 marks zero padding.  Mirrors uplifiting_dataset.py:341-375: frames outside [0, T) repeat the first / last
    in-range *strided* sample (PADDING_TYPE "copy" = np.pad mode "edge") or are zeros ("zeros")."""
    c = np.asarray(center_frames, dtype=np.int64)[:, None]
    k = np.arange(n_tok, dtype=np.int64)[None, :]
    f = (k - n_tok // 2) * s_out + c
    if not pad_copy:
        return np.where((f < 0) | (f >= T), -1, f).astype(np.int32)
    f0 = c - (n_tok // 2) * s_out
    k_min = np.where(f0 >= 0, 0, (-f0 + s_out - 1) // s_out)
    k_max = np.minimum((T - 1 - f0) // s_out, n_tok - 1)
    kc = np.clip(k, k_min, k_max)
    return (f0 + kc * s_out).astype(np.int32)
