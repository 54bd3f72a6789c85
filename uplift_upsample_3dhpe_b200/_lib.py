"""Builds (nvcc, sm_100a) and loads ``libuu3d.so`` through ctypes.

The shared library is the product; there is no Python/CPU fallback.  ``load()``
raises if the library is missing and cannot be built, and every compute entry
point of the library itself fails without a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libuu3d.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
SOURCES = ["kernels_f32.cu", "spatial_tc.cu", "attention_tc.cu", "gemm_tc.cu", "train_kernels.cu", "uu_train.cu",
           "uu_comm.cu", "wgrad_tc.cu", "attn_tc5.cu", "attn_mma.cu", "spatial_train.cu", "uu_api.cu"]
OBJ_DIR = os.path.join(_HERE, "build")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "550,177"]

UU_MAX_STRIDED = 8
PRECISION = {"fp32": 0, "bf16": 1, "tf32": 2}
KINDS = ["gather", "spatial", "token_fill", "layernorm", "attention", "gemm_tc", "gemm_f32", "cast"]


class UUSpec(ctypes.Structure):
    _fields_ = [
        ("n_tok", c_int32), ("n_joints", c_int32), ("d_spatial", c_int32), ("d_temporal", c_int32),
        ("spatial_depth", c_int32), ("temporal_depth", c_int32), ("num_heads", c_int32),
        ("h_spatial", c_int32), ("h_temporal", c_int32), ("n_strided", c_int32),
        ("strides", c_int32 * UU_MAX_STRIDED), ("pad_left", c_int32 * UU_MAX_STRIDED),
        ("pad_right", c_int32 * UU_MAX_STRIDED),
        ("has_strided_input", c_int32), ("first_strided_token_attention_layer", c_int32),
        ("full_output", c_int32),
    ]


def _headers() -> list:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join(INCLUDE, "uu3d.h")]


def _digest(paths) -> str:
    """Content hash (not mtimes: a repository snapshot copied to another machine does not keep their order)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _source_digest() -> str:
    return _digest([os.path.join(CSRC, f) for f in SOURCES] + _headers())


def _read(path) -> str:
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return ""


def _stale() -> bool:
    return not os.path.exists(LIB_PATH) or _read(LIB_PATH + ".srchash") != _source_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libuu3d.so (in-tree, so it travels with the repo).

    One object per translation unit (compiled in parallel, re-used while neither the source nor any header changed),
    linked into a temporary file that replaces libuu3d.so atomically; an exclusive file lock serialises concurrent
    builders (ranks of one torchrun launch), so nobody ever loads a half-written library."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libuu3d.so")
    import fcntl
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    with open(os.path.join(OBJ_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():          # another process built it while we waited
            return LIB_PATH
        hdrs = _headers()

        def compile_one(src):
            obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
            path = os.path.join(CSRC, src)
            dig = _digest([path] + hdrs)
            if not force and os.path.exists(obj) and _read(obj + ".srchash") == dig:
                return obj
            cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on " + src + ":\n" + r.stdout + r.stderr)
            with open(obj + ".srchash", "w") as f:
                f.write(dig)
            return obj

        with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
            objs = list(ex.map(compile_one, SOURCES))
        tmp = LIB_PATH + ".tmp.%d" % os.getpid()
        r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-ldl", "-o", tmp],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
        os.replace(tmp, LIB_PATH)
        with open(LIB_PATH + ".srchash", "w") as f:
            f.write(_source_digest())
    return LIB_PATH


_lib = None
EXPECTED_VERSION = 200          # uu_version() of the library this binding was written against


def load() -> ctypes.CDLL:
    """Load the library; (re)build it first when it is missing or older than its sources and nvcc is present."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("UU3D_LIB") or LIB_PATH      # UU3D_LIB: development override (A/B builds of the library)
    if path == LIB_PATH:
        have_nvcc = bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")
        if not os.path.exists(LIB_PATH) or (have_nvcc and _stale()):
            build()
    lib = ctypes.CDLL(path)
    _declare(lib)
    if lib.uu_version() != EXPECTED_VERSION:
        raise RuntimeError("libuu3d.so reports version %d, this binding expects %d: rebuild the library"
                           % (lib.uu_version(), EXPECTED_VERSION))
    _lib = lib
    return lib


def _declare(lib) -> None:
    P = POINTER
    lib.uu_last_error.restype = c_char_p
    lib.uu_last_error.argtypes = []
    lib.uu_version.restype = c_int
    lib.uu_create.argtypes = [P(UUSpec), c_int, P(c_void_p)]
    lib.uu_destroy.argtypes = [c_void_p]
    lib.uu_set_precision.argtypes = [c_void_p, c_int]
    lib.uu_get_precision.argtypes = [c_void_p]
    lib.uu_weight_count.argtypes = [c_void_p]
    lib.uu_param_count.argtypes = [c_void_p]
    lib.uu_param_count.restype = c_int64
    lib.uu_weight_info.argtypes = [c_void_p, c_int, c_char_p, c_int, P(c_int), P(c_int64), P(c_int)]
    lib.uu_set_weight.argtypes = [c_void_p, c_char_p, c_int, c_void_p, P(c_int64), c_int]
    lib.uu_get_weight.argtypes = [c_void_p, c_char_p, c_int, c_void_p, c_int64]
    lib.uu_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    lib.uu_forward_host.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    lib.uu_last_launch_count.argtypes = [c_void_p]
    lib.uu_set_profiling.argtypes = [c_void_p, c_int]
    lib.uu_get_profile.argtypes = [c_void_p, c_void_p, c_void_p, c_int]
    lib.uu_train_config.argtypes = [c_void_p, c_int, c_int, c_float, c_float, P(c_float), c_int, ctypes.c_uint64]
    lib.uu_train_forward_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]
    lib.uu_grad_buffer.argtypes = [c_void_p, P(c_void_p), P(c_int64)]
    lib.uu_optimizer_state.argtypes = [c_void_p, c_int, c_int, P(c_void_p), P(c_int64)]
    lib.uu_get_grad.argtypes = [c_void_p, c_char_p, c_int, c_void_p, c_int64]
    lib.uu_get_droppath_scale.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int64, P(c_float)]
    lib.uu_adamw_step.argtypes = [c_void_p, c_float, c_float, c_float, c_float, c_float, c_int64, c_float, c_void_p]
    lib.uu_get_ema_weight.argtypes = [c_void_p, c_char_p, c_int, c_void_p, c_int64]
    lib.uu_stride_mask.argtypes = [c_int, c_int, c_int, c_int64, c_void_p]
    lib.uu_op_build_gather.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.uu_op_token_fill.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.uu_op_layernorm.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p,
                                    c_int, c_void_p]
    lib.uu_op_attention.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]
    lib.uu_forward_video.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p]
    lib.uu_forward_video_host.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                          c_void_p]
    lib.uu_set_flip_order.argtypes = [c_void_p, c_void_p, c_int]
    lib.uu_forward_tta.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    lib.uu_forward_video_tta.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                         c_void_p]
    lib.uu_op_keyframe_interp.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.uu_op_window_gather.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p, c_void_p]
    lib.uu_op_spatial.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, P(c_int32), c_void_p]
    lib.uu_op_gemm_f32.argtypes = [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                   c_int64, c_void_p, c_int64, c_void_p]
    lib.uu_op_gemm_bf16.argtypes = [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int,
                                    c_void_p, c_int64, c_void_p, c_int, c_int64, c_void_p]
    lib.uu_op_world_to_cam_and_2d.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.uu_op_pose_metrics.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.uu_train_set_math.argtypes = [c_void_p, c_int]
    lib.uu_train_set_token_masking.argtypes = [c_void_p, c_float]
    lib.uu_get_token_mask.argtypes = [c_void_p, c_void_p, c_int64]
    lib.uu_op_gemm_tf32.argtypes = [c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p,
                                    c_int64, c_void_p, c_int64, c_void_p]
    lib.uu_op_ln_gemm_bf16.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int,
                                       c_int, c_void_p, c_void_p]
    lib.uu_op_resid_gemm_bf16.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                          c_int, c_void_p, c_void_p]
    lib.uu_op_attention_tc5.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]
    lib.uu_op_attention_train.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                          c_int, c_void_p]
    lib.uu_op_mlp_bf16.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]
    lib.uu_op_wgrad_tf32.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int, c_void_p]
    lib.uu_comm_unique_id.argtypes = [c_void_p, c_int]
    lib.uu_comm_init.argtypes = [c_void_p, c_void_p, c_int, c_int]
    lib.uu_comm_destroy.argtypes = [c_void_p]
    lib.uu_comm_world_size.argtypes = [c_void_p]
    lib.uu_allreduce_gradients.argtypes = [c_void_p, c_void_p]
    lib.uu_train_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_float, c_float, c_float, c_float,
                                  c_float, c_float, c_void_p, c_void_p]
    for name in EXPORTS:          # fail at load time, not at first use, if a symbol is missing
        getattr(lib, name)


# every symbol include/uu3d.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "uu_last_error", "uu_version", "uu_create", "uu_destroy", "uu_set_precision", "uu_get_precision",
    "uu_weight_count", "uu_param_count", "uu_weight_info", "uu_set_weight", "uu_get_weight",
    "uu_forward", "uu_forward_host", "uu_forward_video", "uu_forward_video_host", "uu_op_window_gather",
    "uu_set_flip_order", "uu_forward_tta", "uu_forward_video_tta", "uu_op_keyframe_interp", "uu_op_pose_metrics", "uu_op_world_to_cam_and_2d", "uu_last_launch_count", "uu_set_profiling", "uu_get_profile", "uu_stride_mask",
    "uu_train_config", "uu_train_forward_backward", "uu_grad_buffer", "uu_optimizer_state", "uu_get_grad", "uu_get_droppath_scale",
    "uu_adamw_step", "uu_get_ema_weight", "uu_train_set_math", "uu_train_set_token_masking", "uu_get_token_mask",
    "uu_op_build_gather", "uu_op_token_fill", "uu_op_layernorm", "uu_op_attention", "uu_op_spatial", "uu_op_gemm_f32",
    "uu_op_gemm_bf16", "uu_op_gemm_tf32", "uu_op_ln_gemm_bf16", "uu_op_resid_gemm_bf16",
    "uu_op_wgrad_tf32", "uu_op_mlp_bf16", "uu_op_attention_tc5", "uu_op_attention_train", "uu_comm_unique_id", "uu_comm_init", "uu_comm_destroy", "uu_comm_world_size", "uu_allreduce_gradients", "uu_train_step",
]


class UUError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise UUError(load().uu_last_error().decode("utf8", "replace"))


def make_spec(spec) -> UUSpec:
    s = UUSpec()
    s.n_tok, s.n_joints = spec.n_tok, spec.n_joints
    s.d_spatial, s.d_temporal = spec.d_spatial, spec.d_temporal
    s.spatial_depth, s.temporal_depth = spec.spatial_depth, spec.temporal_depth
    s.num_heads = spec.num_heads
    s.h_spatial, s.h_temporal = spec.h_spatial, spec.h_temporal
    if len(spec.strides) > UU_MAX_STRIDED:
        raise ValueError("too many strided blocks")
    s.n_strided = len(spec.strides)
    for i, (st, p) in enumerate(zip(spec.strides, spec.paddings)):
        s.strides[i], s.pad_left[i], s.pad_right[i] = st, p[0], p[1]
    s.has_strided_input = int(spec.has_strided_input)
    s.first_strided_token_attention_layer = spec.first_strided_token_attention_layer
    s.full_output = int(spec.full_output)
    return s
