"""ORACLE (test infrastructure, never shipped, never the thing measured).

NumPy restatement of the reference forward pass, layer by layer.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.

PARITY UNPINNED: the reference ships no tests, fixtures, golden vectors or
weights, and TensorFlow 2.4.3 / tensorflow-addons 0.13.0 / h5py (where the
arithmetic actually lives, requirements.txt:1-4) are not installable here, so
this restatement cannot be checked against the real reference.  It is pinned
only by known answers derivable from the source (tests/test_oracle.py).

Every function cites the reference lines it follows.  Shorthands:
  net = common/net/uplift_upsample_transformer.py, vit = common/net/vision_transformer.py.
``dtype`` selects float64 ("truth") or float32 (stand-in for TF-CPU fp32).
"""
from __future__ import annotations

import math

import numpy as np

try:                                    # scipy is in the image; keep a slow exact fallback
    from scipy.special import erf as _erf
except Exception:                       # pragma: no cover
    _erf = np.vectorize(math.erf)


def dense(x, W, b):
    """Keras Dense on the last axis: x @ W(in,out) + b."""
    return x @ W + b


def layer_norm(x, gamma, beta, eps):
    """Keras LayerNormalization, non-fused path (eps < 1.001e-5): biased moments, then
    x*(gamma*rsqrt(var+eps)) + (beta - mean*gamma*rsqrt(var+eps)).  vit:168,171 (1e-5), net:238 (1e-6)."""
    mean = x.mean(axis=-1, keepdims=True)
    var = np.square(x - mean).mean(axis=-1, keepdims=True)
    inv = gamma / np.sqrt(var + x.dtype.type(eps))
    return x * inv + (beta - mean * inv)


def gelu_erf(x):
    """keras.activations.gelu(approximate=False), the default activation of vit.TransformerBlock (vit:164)."""
    return x.dtype.type(0.5) * x * (x.dtype.type(1.0) + _erf(x / x.dtype.type(math.sqrt(2.0))).astype(x.dtype))


def relu(x):
    return np.maximum(x, 0)


def softmax(x):
    """tf.nn.softmax over the last axis."""
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def mha(y, p, num_heads, key_mask=None):
    """vit.MHA.call (vit:132-156) with scaled_dot_product_attention (vit:99-130).
    p = [wq_k, wq_b, wk_k, wk_b, wv_k, wv_b, proj_k, proj_b].
    key_mask: float (B, S), 1 on keys that must NOT be attended (the *inverted* stride mask, net:358-363)."""
    B, S, D = y.shape
    dh = D // num_heads

    def split(t):                        # vit:92-97: channel c -> head c // dh, dim c % dh
        return t.reshape(B, S, num_heads, dh).transpose(0, 2, 1, 3)

    q, k, v = split(dense(y, p[0], p[1])), split(dense(y, p[2], p[3])), split(dense(y, p[4], p[5]))
    logits = (q @ k.transpose(0, 1, 3, 2)) / y.dtype.type(math.sqrt(dh))       # vit:117-120
    if key_mask is not None:
        logits = logits + key_mask[:, None, None, :].astype(y.dtype) * y.dtype.type(-1e9)   # vit:122-123
    a = softmax(logits)
    o = (a @ v).transpose(0, 2, 1, 3).reshape(B, S, D)
    return dense(o, p[6], p[7])


def transformer_block(x, p, num_heads, act, key_mask=None):
    """vit.TransformerBlock.call (vit:176-195), inference (no DropPath).
    p = 16 tensors: norm1 g,b; wq,wk,wv,proj (k,b each); norm2 g,b; fc1 k,b; fc2 k,b."""
    y = layer_norm(x, p[0], p[1], 1e-5)
    x = x + mha(y, p[2:10], num_heads, key_mask)
    z = layer_norm(x, p[10], p[11], 1e-5)
    z = dense(act(dense(z, p[12], p[13])), p[14], p[15])
    return x + z


def strided_conv1d(a, W, b, stride, pad):
    """ZeroPadding1D(pad) + Conv1D(k=3, strides=s, 'valid') (net:72-77, :85-86).
    a (B, L, Cin); W (3, Cin, Cout): z[t] = b + sum_k a_pad[t*s + k] @ W[k] (cross-correlation)."""
    B, L, Cin = a.shape
    ap = np.zeros((B, L + pad[0] + pad[1], Cin), dtype=a.dtype)
    ap[:, pad[0]:pad[0] + L] = a
    Lo = (ap.shape[1] - 3) // stride + 1
    z = np.zeros((B, Lo, W.shape[2]), dtype=a.dtype)
    for k in range(3):
        z = z + ap[:, k:k + (Lo - 1) * stride + 1:stride] @ W[k]
    return z + b


def strided_block(x, pe, p, num_heads, stride, pad):
    """StridedTransformerBlock.call (net:122-160) with StridedMLP.call (net:81-90).
    p = 16 tensors: norm1; wq,wk,wv,proj; norm2; fc1 Conv1D k=1 (1,d,h); strided_conv (3,h,d)."""
    x = x + pe                                                     # net:126-128
    y = layer_norm(x, p[0], p[1], 1e-5)
    x = x + mha(y, p[2:10], num_heads)                             # net:129-133
    z = layer_norm(x, p[10], p[11], 1e-5)
    z = relu(dense(z, p[12][0], p[13]))                            # Conv1D k=1 == Dense
    z = strided_conv1d(z, p[14], p[15], stride, pad)
    ident = x
    if stride > 1:                                                 # net:137-152
        if pad[0] == 0:
            ident = ident[:, 1:]
        if pad[1] == 0:
            ident = ident[:, :-1]
        ident = ident[:, ::stride]                                 # MaxPool1D(pool_size=1, strides=s)
    return ident + z                                               # net:156


def group(w, name):
    out, i = [], 0
    while (name, i) in w:
        out.append(w[(name, i)])
        i += 1
    return out


def forward(spec, w, x2d, stride_mask, dtype=np.float64, return_intermediates=False):
    """UpliftUpsampleTransformer.call (net:388-421), training=False.

    x2d (B, n_tok, J, 2) — the *caller-masked* key-points (eval.py:67); stride_mask (B, n_tok) bool,
    True on tokens with a 2-D pose.  Returns (full (B,n_tok,J,3) | None, central (B,J,3))."""
    w = {k: np.asarray(v, dtype=dtype) for k, v in w.items()}
    x = np.asarray(x2d, dtype=dtype)
    B, N, J, _ = x.shape
    H = spec.num_heads
    inter = {}
    # --- spatial_transformation (net:313-333)
    x = x.reshape(B * N, J, 2)
    ke = group(w, "keypoint_embedding")
    x = dense(x, ke[0], ke[1]) + w[("spatial_pe", 0)]              # net:321-323
    for i in range(spec.spatial_depth):
        x = transformer_block(x, group(w, f"spatial_block_{i + 1}"), H, gelu_erf)
    sn = group(w, "spatial_norm")
    x = layer_norm(x, sn[0], sn[1], 1e-6)                          # net:238, :329
    x = x.reshape(B, N, J * spec.d_spatial)                        # joint-major flatten, net:330
    inter["spatial"] = x
    fc = group(w, "spatial_to_temporal_fc")
    x = dense(x, fc[0], fc[1])                                     # net:332
    # --- temporal_transformation (net:335-367)
    inv = None
    if spec.has_strided_input:
        m = np.asarray(stride_mask, dtype=dtype)
        inv = dtype(1.0) - m
        tok = w[("strided_input_token_layer", 0)]
        x = m[..., None] * x + inv[..., None] * tok                # net:350
    x = x + w[("temporal_pe", 0)]                                  # net:352
    inter["temporal_in"] = x
    for i in range(spec.temporal_depth):
        km = inv if (spec.has_strided_input and i < spec.first_strided_token_attention_layer) else None
        x = transformer_block(x, group(w, f"temporal_block_{i + 1}"), H, relu, km)
    inter["temporal_out"] = x
    full = None
    if spec.full_output:                                           # net:399-404
        h1 = group(w, "temporal_fc")
        full = dense(x, h1[0], h1[1]).reshape(B, N, J, 3)
    # --- strided_temporal_transformation (net:369-386)
    for i, s in enumerate(spec.strides):
        x = strided_block(x, w[(f"strided_temporal_pe_{i + 1}", 0)], group(w, f"strided_temporal_block_{i + 1}"),
                          H, s, spec.paddings[i])
        inter[f"strided_{i + 1}"] = x
    assert x.shape[1] == 1
    h2 = group(w, "strided_temporal_fc")
    central = dense(x, h2[0], h2[1]).reshape(B, J, 3)              # net:414-416
    if return_intermediates:
        return full, central, inter
    return full, central


def test_step(spec, w, keypoints2d, stride_masks, dtype=np.float64):
    """eval.py:63-71: zero the frames without 2-D input, then forward."""
    x = np.asarray(keypoints2d, dtype=dtype)
    if spec.has_strided_input:
        x = x * np.asarray(stride_masks, dtype=dtype)[:, :, None, None]
    return forward(spec, w, x, stride_masks, dtype=dtype)
