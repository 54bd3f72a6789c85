"""TEST INFRASTRUCTURE — NumPy restatement of the evaluation glue around the forward pass.

flip_tta:      eval.py:152-180 (flip test-time augmentation of both outputs)
interpolate:   common/dataset/action_wise_eval.py:76-100 (key-frame interpolation)
Pinned against the reference's own code by tests/golden/tta_*.npz and interp_*.npz (scripts/make_golden.py)."""
import numpy as np

from . import forward_np


def flip(x, order, axis):
    """Negate the x coordinate (last axis, component 0) and gather joints by AUGM_FLIP_KEYPOINT_ORDER."""
    y = np.array(x, copy=True)
    y[..., 0] *= -1.0
    return np.take(y, order, axis=axis)


def flip_tta(spec, w, keypoints2d, stride_masks, order, dtype=np.float64):
    full, central = forward_np.test_step(spec, w, keypoints2d, stride_masks, dtype=dtype)            # :152-153
    ffull, fcentral = forward_np.test_step(spec, w, flip(np.asarray(keypoints2d, dtype=dtype), order, 2), stride_masks,
                                           dtype=dtype)                                               # :155-162
    central = (central + flip(fcentral, order, 1)) / 2.0                                             # :164-170
    if full is not None:
        full = (full + flip(ffull, order, 2)) / 2.0                                                  # :172-181
    return full, central


def interpolate_between_keyframes(pred, frame_indices, stride):
    """Frames on the key-frame grid keep their prediction, frames between two key frames of one video are linear in
    list position, frames after the last key frame copy it; a video ends where the index does not increase."""
    pred = np.asarray(pred, dtype=np.float64)
    out = pred.copy()
    n = len(frame_indices)
    last = None
    for i in range(n):
        f = int(frame_indices[i])
        if i > 0 and f <= int(frame_indices[i - 1]):
            last = None
        if f % stride == 0:
            if last is not None:
                for k in range(last + 1, i):
                    dl, dr = k - last, i - k
                    out[k] = pred[last] * (dr / (dl + dr)) + pred[i] * (dl / (dl + dr))
            last = i
        elif last is not None:
            out[i] = pred[last]
    return out
