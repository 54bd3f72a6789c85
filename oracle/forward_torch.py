"""ORACLE (test infrastructure, never shipped): the same restatement as forward_np.py written with
PyTorch *CPU* ops, so that (a) the training-step oracle can use CPU autograd and (b) bench.py's
``cpu_baseline`` / ``--impl reference`` legs can run the reference algorithm on all host cores
(TensorFlow, where the reference actually runs, is not installable here — BASELINE.md §3).

PARITY UNPINNED, like forward_np.py: no reference vectors exist.  tests/test_oracle.py pins this
file to forward_np.py (two independent restatements must agree to fp64 round-off).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
net = common/net/uplift_upsample_transformer.py, vit = common/net/vision_transformer.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def layer_norm(x, g, b, eps):
    """Keras LayerNormalization (non-fused form), vit:168,171 / net:238."""
    mean = x.mean(dim=-1, keepdim=True)
    var = (x - mean).pow(2).mean(dim=-1, keepdim=True)
    inv = g * torch.rsqrt(var + eps)
    return x * inv + (b - mean * inv)


def mha(y, p, H, key_mask=None, keep=None):
    """vit:99-156.  key_mask (B,S) float, 1 = do not attend."""
    B, S, D = y.shape
    dh = D // H
    q = (y @ p[0] + p[1]).view(B, S, H, dh).transpose(1, 2)
    k = (y @ p[2] + p[3]).view(B, S, H, dh).transpose(1, 2)
    v = (y @ p[4] + p[5]).view(B, S, H, dh).transpose(1, 2)
    logits = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if key_mask is not None:
        logits = logits + key_mask[:, None, None, :] * (-1e9)
    a = torch.softmax(logits, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, S, D)
    return o @ p[6] + p[7]


def _drop(z, keep):
    """vit:16-28 with an injected per-sample keep mask: (z / keep_prob) * mask."""
    if keep is None:
        return z
    keep_prob, mask = keep
    return (z / keep_prob) * mask.view(-1, *([1] * (z.dim() - 1)))


def block(x, p, H, act, key_mask=None, keep=None, keep2=None):
    """vit:176-195.  keep / keep2: the DropPath draws of the attention branch (vit:185-186) and of the MLP branch
    (vit:189-190) — the layer is called twice, each call draws its own tf.random.uniform."""
    x = x + _drop(mha(layer_norm(x, p[0], p[1], 1e-5), p[2:10], H, key_mask), keep)
    z = layer_norm(x, p[10], p[11], 1e-5)
    z = act(z @ p[12] + p[13]) @ p[14] + p[15]
    return x + _drop(z, keep2)


def strided_block(x, pe, p, H, stride, pad, keep=None, keep2=None):
    """net:122-160 with StridedMLP net:81-90 (DropPath draws: net:131-132 and net:134-135)."""
    x = x + pe
    x = x + _drop(mha(layer_norm(x, p[0], p[1], 1e-5), p[2:10], H), keep)
    z = layer_norm(x, p[10], p[11], 1e-5)
    z = torch.relu(z @ p[12][0] + p[13])
    z = F.pad(z, (0, 0, pad[0], pad[1]))
    # Conv1D channels-last, kernel (k, in, out) -> torch conv1d (out, in, k) on (B, C, L)
    z = F.conv1d(z.transpose(1, 2), p[14].permute(2, 1, 0), p[15], stride=stride).transpose(1, 2)
    z = _drop(z, keep2)
    ident = x
    if stride > 1:
        if pad[0] == 0:
            ident = ident[:, 1:]
        if pad[1] == 0:
            ident = ident[:, :-1]
        ident = ident[:, ::stride]
    return ident + z


def group(w, name):
    out, i = [], 0
    while (name, i) in w:
        out.append(w[(name, i)])
        i += 1
    return out


def forward(spec, w, x2d, stride_mask, keeps=None, token_keep=None):
    """net:388-421.  w: {(group, index): tensor}; x2d (B,N,J,2) already masked by the caller;
    stride_mask (B,N) bool or None.  keeps: optional {(stage, i, branch): (keep_prob, mask)} DropPath masks
    for training-mode parity (stage in 'spatial' | 'temporal' | 'strided'; branch 0 = attention, 1 = MLP).  token_keep: optional (B,N) 0/1 factors
    = 1 - token_mask of random_token_masking (net:287-311, masked value 0), training mode only."""
    keeps = keeps or {}
    B, N, J, _ = x2d.shape
    H = spec.num_heads
    x = x2d.reshape(B * N, J, 2)
    ke = group(w, "keypoint_embedding")
    x = x @ ke[0] + ke[1] + w[("spatial_pe", 0)]
    for i in range(spec.spatial_depth):
        x = block(x, group(w, f"spatial_block_{i + 1}"), H, lambda t: F.gelu(t), keep=keeps.get(("spatial", i, 0)),
                  keep2=keeps.get(("spatial", i, 1)))
    sn = group(w, "spatial_norm")
    x = layer_norm(x, sn[0], sn[1], 1e-6).reshape(B, N, J * spec.d_spatial)
    fc = group(w, "spatial_to_temporal_fc")
    x = x @ fc[0] + fc[1]
    if token_keep is not None:                       # net:336-338, before the strided-input token and the PE
        x = x * token_keep.to(x.dtype)[..., None]
    inv = None
    if spec.has_strided_input:
        m = stride_mask.to(x.dtype)
        inv = 1.0 - m
        x = m[..., None] * x + inv[..., None] * w[("strided_input_token_layer", 0)]
    x = x + w[("temporal_pe", 0)]
    for i in range(spec.temporal_depth):
        km = inv if (spec.has_strided_input and i < spec.first_strided_token_attention_layer) else None
        x = block(x, group(w, f"temporal_block_{i + 1}"), H, torch.relu, km, keep=keeps.get(("temporal", i, 0)),
                  keep2=keeps.get(("temporal", i, 1)))
    full = None
    if spec.full_output:
        h1 = group(w, "temporal_fc")
        full = (x @ h1[0] + h1[1]).view(B, N, J, 3)
    for i, s in enumerate(spec.strides):
        x = strided_block(x, w[(f"strided_temporal_pe_{i + 1}", 0)], group(w, f"strided_temporal_block_{i + 1}"),
                          H, s, spec.paddings[i], keep=keeps.get(("strided", i, 0)), keep2=keeps.get(("strided", i, 1)))
    h2 = group(w, "strided_temporal_fc")
    central = (x @ h2[0] + h2[1]).view(B, J, 3)
    return full, central


def to_torch(w, dtype=torch.float32, requires_grad=False):
    return {k: torch.tensor(v, dtype=dtype, requires_grad=requires_grad) for k, v in w.items()}


def test_step(spec, w, keypoints2d, stride_masks):
    """eval.py:63-71."""
    x = keypoints2d
    if spec.has_strided_input:
        x = x * stride_masks.to(x.dtype)[:, :, None, None]
    with torch.no_grad():
        return forward(spec, w, x, stride_masks)
