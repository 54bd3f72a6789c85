"""TEST INFRASTRUCTURE — NumPy restatement of the evaluation glue around the forward pass.

flip_tta:      eval.py:152-180 (flip test-time augmentation of both outputs)
interpolate:   common/dataset/action_wise_eval.py:76-100 (key-frame interpolation)
mpjpe, nmpjpe: common/dataset/metrics.py:13-81, :120-133 (root alignment; N-MPJPE with optimal per-pose scale)
world_to_cam_and_2d: common/dataset/uplifiting_dataset.py:669-761 (quaternion world -> camera, H3.6M projection)
Pinned against the reference's own code by tests/golden/tta_*.npz, interp_*.npz, metrics_*.npz and projection_*.npz
(scripts/make_golden.py)."""
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


def _root_aligned(pred, gt, root):
    g = gt[:, :, :3] - gt[:, root, None, :3]
    p = pred - pred[:, root, None, :]
    return p, g, gt[:, :, 3] > 0


def mpjpe(pred, gt, root, normalize=True):
    """metrics.py:13-37: mean over valid joints of ||(p - p_root) - (g - g_root)||; per joint (-1 = invalid) otherwise."""
    p, g, valid = _root_aligned(np.asarray(pred, np.float64), np.asarray(gt, np.float64), root)
    d = np.sqrt(((p - g) ** 2).sum(-1))
    return np.where(valid, d, 0.0).sum() / valid.sum() if normalize else np.where(valid, d, -1.0)


def nmpjpe(pred, gt, root, normalize=True):
    """metrics.py:40-81 with alignment="root" and optimal_scaling (:120-133): s = <p, g> / <p, p> over valid joints."""
    p, g, valid = _root_aligned(np.asarray(pred, np.float64), np.asarray(gt, np.float64), root)
    v = valid[:, :, None]
    s = (p * g * v).sum((1, 2)) / (p * p * v).sum((1, 2))
    d = np.sqrt(((p * s[:, None, None] - g) ** 2).sum(-1))
    return np.where(valid, d, 0.0).sum() / valid.sum() if normalize else np.where(valid, d, -1.0)


def world_to_cam_and_2d(seq3d, cam):
    """uplifiting_dataset.py:669-761.  seq3d (..., 3) world coordinates of one sample, cam (18,) = unit quaternion (w, x, y, z),
    translation (3), intrinsics (res 2, focal 2, centre 2, radial 3, tangential 2).  Returns (camera-space 3-D, 2-D)."""
    x = np.asarray(seq3d, np.float64)
    cam = np.asarray(cam, np.float64)
    w, qv = cam[0], -cam[1:4]                                  # tf_qinverse: conjugate (:712-716)
    v = x - cam[4:7]                                           # :719-722
    uv = np.cross(np.broadcast_to(qv, v.shape), v)             # tf_qrot (:701-709)
    uuv = np.cross(np.broadcast_to(qv, v.shape), uv)
    xc = v + 2 * (w * uv + uuv)
    intr = cam[7:18]
    f, c, k, p = intr[2:4], intr[4:6], intr[6:9], intr[9:11]
    xx = np.clip(xc[..., :2] / xc[..., 2:], -1.0, 1.0)         # :746-747
    r2 = (xx ** 2).sum(-1, keepdims=True)
    radial = 1 + (k * np.concatenate([r2, r2 ** 2, r2 ** 3], axis=-1)).sum(-1, keepdims=True)
    tan = (p * xx).sum(-1, keepdims=True)
    return xc, f * (xx * (radial + tan) + p * r2) + c           # :752-757
