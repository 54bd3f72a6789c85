"""Weight contract: the ordered tensor inventory of the Keras ``.h5`` file.

Matching in the reference loader is by top-level layer (group) name, then by
position inside the group (common/utils/weight_io.py:155-198) — inner variable
names are never compared.  The order inside a block follows attribute creation
order: norm1, wq, wk, wv, projection, norm2, mlp.fc1, mlp.fc2 / mlp.strided_conv
(common/net/vision_transformer.py:82-86, :168-174;
common/net/uplift_upsample_transformer.py:67-77, :109-115).
Dense kernels are (in, out); Conv1D kernels are (k, in, out).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np

from .spec import ModelSpec

WeightKey = Tuple[str, int]          # (top-level group, position in group)


def _block(prefix: str, d: int, h: int, strided: bool) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(inner name, shape, init kind) of one transformer block, in file order."""
    t = [("layer_normalization/gamma:0", (d,), "ones"), ("layer_normalization/beta:0", (d,), "zeros")]
    for n in ("dense", "dense_1", "dense_2", "dense_3"):          # wq, wk, wv, projection
        t += [(f"mha/{n}/kernel:0", (d, d), "glorot"), (f"mha/{n}/bias:0", (d,), "zeros")]
    t += [("layer_normalization_1/gamma:0", (d,), "ones"), ("layer_normalization_1/beta:0", (d,), "zeros")]
    if strided:
        t += [("strided_mlp/conv1d/kernel:0", (1, d, h), "glorot"), ("strided_mlp/conv1d/bias:0", (h,), "zeros"),
              ("strided_mlp/conv1d_1/kernel:0", (3, h, d), "glorot"), ("strided_mlp/conv1d_1/bias:0", (d,), "zeros")]
    else:
        t += [("mlp/dense_4/kernel:0", (d, h), "glorot"), ("mlp/dense_4/bias:0", (h,), "zeros"),
              ("mlp/dense_5/kernel:0", (h, d), "glorot"), ("mlp/dense_5/bias:0", (d,), "zeros")]
    return [(f"{prefix}/{n}", s, k) for n, s, k in t]


def inventory(spec: ModelSpec) -> "OrderedDict[str, List[Tuple[str, Tuple[int, ...], str]]]":
    """group -> ordered [(weight_name, shape, init kind)], groups in Keras ``layer_names`` order."""
    J, ds, dt = spec.n_joints, spec.d_spatial, spec.d_temporal
    inv: "OrderedDict[str, list]" = OrderedDict()
    inv["keypoint_embedding"] = [("keypoint_embedding/kernel:0", (2, ds), "glorot"),
                                 ("keypoint_embedding/bias:0", (ds,), "zeros")]
    inv["token_dropout"] = []
    inv["spatial_pe"] = [("spatial_pe/positional_encoding_weights:0", (J, ds), "trunc")]
    inv["temporal_pe"] = [("temporal_pe/positional_encoding_weights:0", (spec.n_tok, dt), "trunc")]
    for i in range(len(spec.strides)):
        inv[f"strided_temporal_pe_{i + 1}"] = [
            (f"strided_temporal_pe_{i + 1}/positional_encoding_weights:0", (spec.seq_lens[i], dt), "trunc")]
    if spec.has_strided_input:
        inv["strided_input_token_layer"] = [("strided_input_token_layer/learnable_masked_token:0", (dt,), "trunc")]
    for i in range(spec.spatial_depth):
        inv[f"spatial_block_{i + 1}"] = _block(f"spatial_block_{i + 1}", ds, spec.h_spatial, False)
    inv["spatial_norm"] = [("spatial_norm/gamma:0", (ds,), "ones"), ("spatial_norm/beta:0", (ds,), "zeros")]
    inv["spatial_to_temporal_fc"] = [("spatial_to_temporal_fc/kernel:0", (J * ds, dt), "glorot"),
                                     ("spatial_to_temporal_fc/bias:0", (dt,), "zeros")]
    for i in range(spec.temporal_depth):
        inv[f"temporal_block_{i + 1}"] = _block(f"temporal_block_{i + 1}", dt, spec.h_temporal, False)
    for i in range(len(spec.strides)):
        inv[f"strided_temporal_block_{i + 1}"] = _block(f"strided_temporal_block_{i + 1}", dt, spec.h_temporal, True)
    if spec.full_output:
        inv["temporal_fc"] = [("temporal_fc/kernel:0", (dt, spec.out_dim), "glorot"),
                              ("temporal_fc/bias:0", (spec.out_dim,), "zeros")]
    inv["strided_temporal_fc"] = [("strided_temporal_fc/kernel:0", (dt, spec.out_dim), "glorot"),
                                  ("strided_temporal_fc/bias:0", (spec.out_dim,), "zeros")]
    return inv


def flat_keys(spec: ModelSpec) -> List[Tuple[WeightKey, Tuple[int, ...]]]:
    """[((group, index), shape)] in canonical flat order == model.trainable_variables order."""
    out = []
    for g, tensors in inventory(spec).items():
        for i, (_, shape, _) in enumerate(tensors):
            out.append(((g, i), shape))
    return out


def param_count(spec: ModelSpec) -> int:
    return sum(int(np.prod(s)) for _, s in flat_keys(spec))


def _glorot(rng, shape):
    # Keras GlorotUniform: limit = sqrt(6 / (fan_in + fan_out)); conv fans are multiplied by the receptive field.
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=shape)


def _trunc_normal(rng, shape, std=0.02):
    # Keras TruncatedNormal(stddev=0.02): resample outside +-2 sigma (net:30, :45).
    x = rng.normal(0.0, std, size=shape)
    bad = np.abs(x) > 2 * std
    while bad.any():
        x[bad] = rng.normal(0.0, std, size=int(bad.sum()))
        bad = np.abs(x) > 2 * std
    return x


def init_weights(spec: ModelSpec, seed: int = 1, perturb: bool = False) -> Dict[WeightKey, np.ndarray]:
    """Random-init weights per the Keras initialisers.  ``perturb`` additionally draws
    non-trivial biases / LN affine parameters so bias and affine bugs are visible
    (SURVEY.md §8d config 1): biases N(0,0.02), gamma 1+N(0,0.1), beta N(0,0.1)."""
    rng = np.random.default_rng(seed)
    w: Dict[WeightKey, np.ndarray] = {}
    for g, tensors in inventory(spec).items():
        for i, (name, shape, kind) in enumerate(tensors):
            if kind == "glorot":
                a = _glorot(rng, shape)
            elif kind == "trunc":
                a = _trunc_normal(rng, shape)
            elif kind == "ones":
                a = np.ones(shape) + (rng.normal(0, 0.1, size=shape) if perturb else 0.0)
            else:
                is_beta = name.endswith("beta:0")
                a = rng.normal(0, 0.1 if is_beta else 0.02, size=shape) if perturb else np.zeros(shape)
            w[(g, i)] = np.ascontiguousarray(a, dtype=np.float32)
    return w


def to_flat(spec: ModelSpec, w: Dict[WeightKey, np.ndarray]) -> np.ndarray:
    return np.concatenate([np.asarray(w[k], dtype=np.float32).reshape(-1) for k, _ in flat_keys(spec)])


def from_flat(spec: ModelSpec, flat: np.ndarray) -> Dict[WeightKey, np.ndarray]:
    out, off = {}, 0
    for k, shape in flat_keys(spec):
        n = int(np.prod(shape))
        out[k] = np.asarray(flat[off:off + n], dtype=np.float32).reshape(shape).copy()
        off += n
    if off != flat.size:
        raise ValueError(f"flat weight vector has {flat.size} values, model needs {off}")
    return out


def save_npz(path: str, spec: ModelSpec, w: Dict[WeightKey, np.ndarray]) -> None:
    """``.npz`` mirror of the .h5 layout: arrays named ``<group>/<index>``."""
    np.savez(path, **{f"{g}/{i}": w[(g, i)] for (g, i), _ in flat_keys(spec)})


def load_npz(path: str, spec: ModelSpec) -> Dict[WeightKey, np.ndarray]:
    z = np.load(path)
    out = {}
    for (g, i), shape in flat_keys(spec):
        a = z[f"{g}/{i}"]
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{g}[{i}]: file has shape {a.shape}, model expects {shape}")   # weight_io.py:219-232
        out[(g, i)] = a.astype(np.float32)
    return out
