"""Host-side mirror of the reference's model interface for the hot path.

    model = build_uplift_upsample_transformer(config)            # constructor.py:14-50
    full, central = model([x2d, stride_mask], training=False)    # net:388-421
    full, central = test_step(model, keypoints2d, stride_masks)  # eval.py:63-71

PyTorch is only the tensor container (device memory + streams); all arithmetic
runs in libuu3d.so (hand-written sm_100a kernels) through the C ABI of
include/uu3d.h.  There is no CPU path: constructing a model without a CUDA
device raises.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_int, c_int64, c_void_p
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from . import weights as W
from .spec import ModelSpec, spec_from_config


def _torch():
    import torch
    return torch


class UpliftUpsampleTransformer:
    """Drop-in for the call contract of common/net/uplift_upsample_transformer.py:163-421."""

    def __init__(self, spec: ModelSpec, device: int = 0, precision: str = "fp32"):
        self.spec = spec
        self.device = int(device)
        self.has_strided_input = spec.has_strided_input      # attributes the callers read (eval.py:66, train.py:472)
        self.full_output = spec.full_output
        self._lib = _lib.load()
        self._h = c_void_p()
        cspec = _lib.make_spec(spec)
        _lib.check(self._lib.uu_create(byref(cspec), self.device, byref(self._h)))
        self.set_precision(precision)
        self._keys = W.flat_keys(spec)
        self._check_inventory()

    # ---- lifetime ------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.uu_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- precision -----------------------------------------------------------------------------
    def set_precision(self, precision: str) -> None:
        if precision not in _lib.PRECISION:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISION)}")
        _lib.check(self._lib.uu_set_precision(self._h, _lib.PRECISION[precision]))
        self.precision = precision

    # ---- weights -------------------------------------------------------------------------------
    def _check_inventory(self) -> None:
        """The C++ inventory and the Python one (weights.py) must agree tensor by tensor."""
        n = self._lib.uu_weight_count(self._h)
        if n != len(self._keys):
            raise _lib.UUError(f"weight inventory mismatch: library has {n}, host has {len(self._keys)}")
        buf = ctypes.create_string_buffer(128)
        idx, rank = c_int(), c_int()
        shape = (c_int64 * 4)()
        for i, ((g, k), shp) in enumerate(self._keys):
            _lib.check(self._lib.uu_weight_info(self._h, i, buf, 128, byref(idx), shape, byref(rank)))
            got = (buf.value.decode(), idx.value, tuple(shape[j] for j in range(rank.value)))
            if got != (g, k, tuple(shp)):
                raise _lib.UUError(f"weight inventory mismatch at {i}: library {got}, host {(g, k, tuple(shp))}")

    def set_weights(self, w: Dict[W.WeightKey, np.ndarray], partial: bool = False) -> None:
        """partial: tensors that ``w`` does not hold keep their current values (a by-name load of a file without some
        layers, weight_io.py:247-251); otherwise a missing tensor is an error."""
        for (g, k), shp in self._keys:
            if (g, k) not in w:
                if partial:
                    continue
                raise ValueError(f"missing weight {g}[{k}]")
            a = np.ascontiguousarray(w[(g, k)], dtype=np.float32)
            shape = (c_int64 * max(a.ndim, 1))(*a.shape)
            _lib.check(self._lib.uu_set_weight(self._h, g.encode(), k, a.ctypes.data_as(c_void_p), shape, a.ndim))

    def get_weights(self) -> Dict[W.WeightKey, np.ndarray]:
        out = {}
        for (g, k), shp in self._keys:
            a = np.empty(shp, dtype=np.float32)
            _lib.check(self._lib.uu_get_weight(self._h, g.encode(), k, a.ctypes.data_as(c_void_p), a.size))
            out[(g, k)] = a
        return out

    def load_weights(self, path: str, skip_mismatch: bool = False, verbose: bool = True) -> None:
        """``.h5`` (Keras layout, by-name loading with the rules of weight_io.py:76-263: layers the file does not hold keep
        their values and are reported, count / shape mismatches raise unless ``skip_mismatch``) or the ``.npz`` mirror."""
        if path.endswith(".npz"):
            self.set_weights(W.load_npz(path, self.spec))
        else:
            from . import h5lite
            self.set_weights(h5lite.load_keras_weights(path, self.spec, skip_mismatch=skip_mismatch, verbose=verbose),
                             partial=True)

    def save_weights(self, path: str) -> None:
        if path.endswith(".npz"):
            W.save_npz(path, self.spec, self.get_weights())
        else:
            from . import h5lite
            h5lite.save_keras_weights(path, self.spec, self.get_weights())

    @property
    def param_count(self) -> int:
        return int(self._lib.uu_param_count(self._h))

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.uu_last_launch_count(self._h))

    def set_profiling(self, on: bool) -> None:
        _lib.check(self._lib.uu_set_profiling(self._h, int(on)))

    def get_profile(self) -> Dict[str, Tuple[float, int]]:
        """{kernel kind: (device ms, launches)} of the most recent forward (profiling must be on)."""
        n = len(_lib.KINDS)
        ms = (ctypes.c_float * n)()
        cnt = (ctypes.c_int32 * n)()
        _lib.check(self._lib.uu_get_profile(self._h, ms, cnt, n))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_lib.KINDS)}

    # ---- forward -------------------------------------------------------------------------------
    def __call__(self, inputs, training: bool = False, want_full: bool = True):
        """inputs = [x2d (B,n_tok,J,2) float32 cuda, stride_mask (B,n_tok) bool/uint8 cuda] (or x2d alone when
        the model has no strided input).  Returns (full (B,n_tok,J,3) | None, central (B,J,3)) as fresh tensors."""
        torch = _torch()
        if training:
            raise NotImplementedError("training=True goes through uplift_upsample_3dhpe_b200.train.train_step")
        if self.has_strided_input:
            x, mask = inputs[0], inputs[1]
        else:
            x, mask = (inputs[0] if isinstance(inputs, (list, tuple)) else inputs), None
        s = self.spec
        if x.dim() != 4 or tuple(x.shape[1:]) != (s.n_tok, s.n_joints, 2):
            raise ValueError(f"x2d must be (B,{s.n_tok},{s.n_joints},2), got {tuple(x.shape)}")
        if not x.is_cuda or x.device.index != self.device:
            raise ValueError(f"x2d must live on cuda:{self.device}")
        x = x.contiguous().float()
        B = x.shape[0]
        mptr = None
        if mask is not None:
            if tuple(mask.shape) != (B, s.n_tok):
                raise ValueError(f"stride_mask must be (B,{s.n_tok}), got {tuple(mask.shape)}")
            mask = mask.to(device=x.device, dtype=torch.uint8).contiguous()
            mptr = mask.data_ptr()
        full = None
        if self.full_output and want_full:
            full = torch.empty((B, s.n_tok, s.n_joints, 3), dtype=torch.float32, device=x.device)
        central = torch.empty((B, s.n_joints, 3), dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self._lib.uu_forward(self._h, x.data_ptr(), mptr, B, full.data_ptr() if full is not None else None,
                                        central.data_ptr(), stream))
        return full, central

    def forward_raw(self, x_ptr: int, mask_ptr: Optional[int], B: int, full_ptr: Optional[int], central_ptr: int,
                    stream: int = 0) -> None:
        """Pointer-level call (what uu_forward binds); used by bench.py to time without tensor allocation."""
        _lib.check(self._lib.uu_forward(self._h, x_ptr, mask_ptr, B, full_ptr, central_ptr, stream))

    def forward_host(self, x2d: np.ndarray, stride_mask: Optional[np.ndarray], full_out: Optional[np.ndarray],
                     central_out: np.ndarray) -> None:
        """End-to-end call on HOST buffers: H2D, forward, D2H, sync (uu_forward_host)."""
        s = self.spec
        B = x2d.shape[0]
        assert x2d.dtype == np.float32 and x2d.flags.c_contiguous
        assert central_out.dtype == np.float32 and central_out.shape == (B, s.n_joints, 3)
        mptr = None
        if stride_mask is not None:
            assert stride_mask.dtype in (np.uint8, np.bool_) and stride_mask.flags.c_contiguous
            mptr = stride_mask.ctypes.data_as(c_void_p)
        fptr = full_out.ctypes.data_as(c_void_p) if full_out is not None else None
        _lib.check(self._lib.uu_forward_host(self._h, x2d.ctypes.data_as(c_void_p), mptr, B, fptr,
                                             central_out.ctypes.data_as(c_void_p)))


    # ---- evaluation glue on the device ------------------------------------
    def set_flip_order(self, order) -> None:
        """AUGM_FLIP_KEYPOINT_ORDER of the config (eval.py:159)."""
        a = np.ascontiguousarray(order, dtype=np.int32)
        _lib.check(self._lib.uu_set_flip_order(self._h, a.ctypes.data_as(c_void_p), int(a.size)))

    def forward_tta(self, inputs, want_full: bool = True):
        """Flip test-time augmentation (eval.py:152-180): (f(x) + unflip(f(flip(x)))) / 2 for both outputs."""
        torch = _torch()
        if self.has_strided_input:
            x, mask = inputs[0], inputs[1]
        else:
            x, mask = (inputs[0] if isinstance(inputs, (list, tuple)) else inputs), None
        s = self.spec
        x = x.contiguous().float()
        B = x.shape[0]
        mptr = None
        if mask is not None:
            mask = mask.to(device=x.device, dtype=torch.uint8).contiguous()
            mptr = mask.data_ptr()
        full = torch.empty((B, s.n_tok, s.n_joints, 3), dtype=torch.float32, device=x.device) \
            if (self.full_output and want_full) else None
        central = torch.empty((B, s.n_joints, 3), dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self._lib.uu_forward_tta(self._h, x.data_ptr(), mptr, B, full.data_ptr() if full is not None else None,
                                            central.data_ptr(), stream))
        return full, central

    # ---- sliding windows cut on the device from one video -------------------
    def forward_video(self, video2d, centers, s_out: int, s_in: int, pad_copy: bool = True, want_full: bool = True,
                      flip_tta: bool = False):
        """video2d (T,J,2) float32 cuda, centers (B,) int32 cuda frame indices.  Equivalent to building the
        reference generator's windows + globally aligned stride masks (uplifiting_dataset.py:341-394) and calling
        test_step on them, without materialising the (B, n_tok, J, 2) tensor."""
        torch = _torch()
        s = self.spec
        video2d = video2d.contiguous().float()
        centers = centers.to(dtype=torch.int32).contiguous()
        if video2d.dim() != 3 or tuple(video2d.shape[1:]) != (s.n_joints, 2):
            raise ValueError(f"video2d must be (T,{s.n_joints},2), got {tuple(video2d.shape)}")
        T, B = video2d.shape[0], centers.shape[0]
        full = None
        if self.full_output and want_full:
            full = torch.empty((B, s.n_tok, s.n_joints, 3), dtype=torch.float32, device=video2d.device)
        central = torch.empty((B, s.n_joints, 3), dtype=torch.float32, device=video2d.device)
        stream = torch.cuda.current_stream(video2d.device).cuda_stream
        fn = self._lib.uu_forward_video_tta if flip_tta else self._lib.uu_forward_video
        _lib.check(fn(self._h, video2d.data_ptr(), T, centers.data_ptr(), B, int(s_out), int(s_in),
                      1 if pad_copy else 0, full.data_ptr() if full is not None else None, central.data_ptr(), stream))
        return full, central

    def forward_video_host(self, video2d: np.ndarray, centers: np.ndarray, s_out: int, s_in: int,
                           central_out: np.ndarray, full_out: Optional[np.ndarray] = None, pad_copy: bool = True) -> None:
        """Host-buffer twin (uu_forward_video_host): H2D of the video and the centre list only."""
        s = self.spec
        assert video2d.dtype == np.float32 and video2d.flags.c_contiguous and video2d.shape[1:] == (s.n_joints, 2)
        assert centers.dtype == np.int32 and centers.flags.c_contiguous
        B = centers.shape[0]
        assert central_out.dtype == np.float32 and central_out.shape == (B, s.n_joints, 3)
        fptr = full_out.ctypes.data_as(c_void_p) if full_out is not None else None
        _lib.check(self._lib.uu_forward_video_host(self._h, video2d.ctypes.data_as(c_void_p), video2d.shape[0],
                                                   centers.ctypes.data_as(c_void_p), B, int(s_out), int(s_in),
                                                   1 if pad_copy else 0, fptr, central_out.ctypes.data_as(c_void_p)))


def build_uplift_upsample_transformer(config, device: int = 0, precision: str = "fp32",
                                      weights: Optional[Dict[W.WeightKey, np.ndarray]] = None,
                                      seed: int = 1) -> UpliftUpsampleTransformer:
    """constructor.py:14-50.  Weights are Keras-initialiser random draws unless given."""
    spec = spec_from_config(config)
    model = UpliftUpsampleTransformer(spec, device=device, precision=precision)
    model.set_weights(weights if weights is not None else W.init_weights(spec, seed))
    order = getattr(config, "AUGM_FLIP_KEYPOINT_ORDER", None)
    if order is not None and len(order) == spec.n_joints:
        model.set_flip_order(order)
    return model


def test_step(model: UpliftUpsampleTransformer, keypoints2d, stride_masks):
    """eval.py:63-71.  The reference zeroes frames without 2-D input before the call; here the
    kernels never read those frames, which is the same thing."""
    if model.has_strided_input:
        return model([keypoints2d, stride_masks], training=False)
    return model(keypoints2d, training=False)


def keyframe_interp(pred, frame_indices, keyframe_stride: int):
    """action_wise_eval.interpolate_between_keyframes on the device: pred (n, ...) float32 cuda, frame_indices (n,)."""
    torch = _torch()
    lib = _lib.load()
    pred = pred.contiguous().float()
    idx = frame_indices.to(device=pred.device, dtype=torch.int32).contiguous()
    out = torch.empty_like(pred)
    n = pred.shape[0]
    stream = torch.cuda.current_stream(pred.device).cuda_stream
    _lib.check(lib.uu_op_keyframe_interp(pred.data_ptr(), idx.data_ptr(), n, int(keyframe_stride), pred[0].numel() if n else 1,
                                         out.data_ptr(), stream))
    return out


def world_to_cam_and_2d(seq3d, cams):
    """uplifiting_dataset.tf_world_to_cam_and_2d on the device for a batch: seq3d (B, ..., 3) world coordinates and
    cams (B, 18) float32 cuda -> (camera-space 3-D, 2-D projection) with the leading shape of seq3d."""
    torch = _torch()
    lib = _lib.load()
    seq3d, cams = seq3d.contiguous().float(), cams.contiguous().float()
    B = seq3d.shape[0]
    assert seq3d.shape[-1] == 3 and tuple(cams.shape) == (B, 18) and cams.device == seq3d.device
    pps = seq3d[0].numel() // 3
    cam3d = torch.empty_like(seq3d)
    p2d = torch.empty(seq3d.shape[:-1] + (2,), device=seq3d.device)
    stream = torch.cuda.current_stream(seq3d.device).cuda_stream
    _lib.check(lib.uu_op_world_to_cam_and_2d(seq3d.data_ptr(), cams.data_ptr(), B, pps, cam3d.data_ptr(), p2d.data_ptr(), stream))
    return cam3d, p2d


def pose_metrics(pred, gt, root_index: int, per_joint: bool = False):
    """metrics.mpjpe / metrics.nmpjpe (root alignment) on the device: pred (n, J, 3), gt (n, J, 4 = x, y, z, valid) float32
    cuda.  Returns (mpjpe, nmpjpe) floats, plus the two (n, J) per-joint arrays (-1 = invalid joint) when per_joint."""
    import ctypes
    torch = _torch()
    lib = _lib.load()
    pred, gt = pred.contiguous().float(), gt.contiguous().float()
    n, J = pred.shape[0], pred.shape[1]
    assert tuple(pred.shape) == (n, J, 3) and tuple(gt.shape) == (n, J, 4) and gt.device == pred.device
    jpe = torch.empty((n, J), device=pred.device) if per_joint else None
    njpe = torch.empty((n, J), device=pred.device) if per_joint else None
    res = (ctypes.c_double * 3)()
    stream = torch.cuda.current_stream(pred.device).cuda_stream
    _lib.check(lib.uu_op_pose_metrics(pred.data_ptr(), gt.data_ptr(), n, J, int(root_index),
                                      jpe.data_ptr() if per_joint else None, njpe.data_ptr() if per_joint else None, res, stream))
    return (res[0], res[1], jpe, njpe) if per_joint else (res[0], res[1])


def host_stride_mask(n_tok: int, s_out: int, s_in: int, shift: int = 0) -> np.ndarray:
    """uu_stride_mask: the C-ABI twin of stride_mask.stride_mask (bit-exact check target)."""
    lib = _lib.load()
    out = np.zeros(n_tok, dtype=np.uint8)
    _lib.check(lib.uu_stride_mask(n_tok, s_out, s_in, int(shift), out.ctypes.data_as(c_void_p)))
    return out.astype(bool)
