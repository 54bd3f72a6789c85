"""The oracle against hand-derivable known answers (the reference ships no golden vectors — SURVEY.md §8c)."""
import math

import numpy as np
import pytest

from oracle import forward_np as O
from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, spec_from_config, stride_mask, weights


def _setup(name, B=3, seed=0, perturb=True, **over):
    cfg = UpliftUpsampleConfig.preset(name, **over)
    spec = spec_from_config(cfg)
    w = weights.init_weights(spec, 1, perturb=perturb)
    x = np.random.default_rng(seed).uniform(-1, 1, (B, spec.n_tok, 17, 2)).astype(np.float32)
    return cfg, spec, w, x


def test_layer_norm_and_gelu_known_values():
    x = np.array([[1.0, 2.0, 3.0, 4.0]])
    y = O.layer_norm(x, np.ones(4), np.zeros(4), 0.0)
    assert np.allclose(y, (x - 2.5) / math.sqrt(1.25))
    assert np.allclose(O.gelu_erf(np.array([0.0, 1.0, -1.0])), [0.0, 0.8413447460685429, -0.15865525393145707])
    assert np.allclose(O.softmax(np.array([[0.0, math.log(3.0)]])), [[0.25, 0.75]])


def test_strided_conv_matches_direct_loops():
    rng = np.random.default_rng(3)
    a = rng.normal(size=(2, 11, 5)); Wc = rng.normal(size=(3, 5, 4)); b = rng.normal(size=4)
    for stride, pad in ((4, (0, 0)), (4, (1, 1)), (3, (0, 0))):
        z = O.strided_conv1d(a, Wc, b, stride, pad)
        ap = np.pad(a, ((0, 0), pad, (0, 0)))
        Lo = (ap.shape[1] - 3) // stride + 1
        want = np.zeros((2, Lo, 4))
        for n in range(2):
            for t in range(Lo):
                want[n, t] = b + sum(ap[n, t * stride + k] @ Wc[k] for k in range(3))
        assert z.shape == want.shape and np.allclose(z, want)


def test_identity_gather_indices():
    # 351 cfg -> {1,4,..,67}, {1,11,21}, {1}; 81 cfg -> {0,4,..,40}, {1,5,9}, {1}
    for name, expect in (("h36m_351", [list(range(1, 68, 3)), [1, 11, 21], [1]]),
                         ("h36m_81", [list(range(0, 41, 4)), [1, 5, 9], [1]])):
        spec = spec_from_config(UpliftUpsampleConfig.preset(name))
        for i, s in enumerate(spec.strides):
            L = spec.seq_lens[i]
            x = np.arange(L, dtype=np.float64)[None, :, None] * np.ones((1, 1, 384))
            zero = [np.zeros_like(t) for t in weights.init_weights(spec, 1).values() if False]
            p = O.group(weights.init_weights(spec, 1), f"strided_temporal_block_{i + 1}")
            p = [np.zeros_like(t, dtype=np.float64) for t in p]          # all-zero block: out == identity path
            out = O.strided_block(x, np.zeros((L, 384)), p, 8, s, spec.paddings[i])
            assert out[0, :, 0].astype(int).tolist() == expect[i]


def test_forward_shapes_and_fp32_floor():
    for name in ("h36m_351", "h36m_81"):
        cfg, spec, w, x = _setup(name)
        m = np.stack([stride_mask.stride_mask(spec.n_tok, cfg.SEQUENCE_STRIDE, cfg.MASK_STRIDE[1])] * 3)
        f64, c64 = O.test_step(spec, w, x, m, np.float64)
        f32, c32 = O.test_step(spec, w, x, m, np.float32)
        assert f64.shape == (3, spec.n_tok, 17, 3) and c64.shape == (3, 17, 3)
        assert np.abs(f64 - f32).max() < 1e-4 and np.abs(c64 - c32).max() < 1e-4


def test_masked_frames_never_influence_outputs():
    cfg, spec, w, x = _setup("h36m_351")
    m = np.stack([stride_mask.stride_mask(71, 5, 20, shift_tokens=s) for s in (0, 1, -2)])
    f1, c1 = O.forward(spec, w, x * m[:, :, None, None], m)
    x2 = x.copy()
    x2[~m] = 1e3 * np.random.default_rng(9).normal(size=x2[~m].shape)
    f2, c2 = O.forward(spec, w, x2 * m[:, :, None, None], m)
    assert np.array_equal(f1, f2) and np.array_equal(c1, c2)
    # token fill is exact selection: valid rows keep the spatial result, masked rows get token + PE
    _, _, inter = O.forward(spec, w, x * m[:, :, None, None], m, return_intermediates=True)
    tok, pe = w[("strided_input_token_layer", 0)], w[("temporal_pe", 0)]
    assert np.allclose(inter["temporal_in"][~m], (tok + pe)[np.nonzero(~m)[1]])


def test_sin_equals_sout_is_noop_mask():
    cfg, spec, w, x = _setup("h36m_351")
    m = np.ones((3, 71), dtype=bool)
    f1, c1 = O.forward(spec, w, x, m)
    spec2 = spec_from_config(UpliftUpsampleConfig.preset("h36m_351", MASK_STRIDE=None))
    w2 = {k: v for k, v in w.items() if k[0] != "strided_input_token_layer"}
    f2, c2 = O.forward(spec2, w2, x, None)
    assert np.allclose(f1, f2, atol=1e-12) and np.allclose(c1, c2, atol=1e-12)


def test_batch_permutation_equivariance():
    cfg, spec, w, x = _setup("h36m_81", B=4)
    m = stride_mask.batch_stride_masks_train(41, 2, [4, 10, 20], 4, seed=0)
    f, c = O.test_step(spec, w, x, m)
    perm = [2, 0, 3, 1]
    fp, cp = O.test_step(spec, w, x[perm], m[perm])
    assert np.allclose(fp, f[perm], atol=1e-12) and np.allclose(cp, c[perm], atol=1e-12)


def test_all_masked_window_gives_uniform_attention_in_fp32():
    # vit:122-123 in fp32: x - 1e9 rounds to -1e9 for |x| < 32 -> uniform attention over all keys
    rng = np.random.default_rng(0)
    y = rng.normal(size=(1, 71, 384)).astype(np.float32)
    cfg, spec, w, _ = _setup("h36m_351", perturb=False)
    p = [t.astype(np.float32) for t in O.group(w, "temporal_block_1")]
    km = np.ones((1, 71), dtype=np.float32)
    out = O.mha(y, p[2:10], 8, km)
    v = O.dense(y, p[6], p[7])
    want = O.dense(np.broadcast_to(v.mean(axis=1, keepdims=True), v.shape), p[8], p[9])
    assert np.allclose(out, want, atol=1e-5)


def test_torch_restatement_agrees_with_numpy_restatement():
    """Two independent restatements (numpy / torch-CPU incl. F.conv1d and F.gelu) must agree in fp64."""
    import torch
    from oracle import forward_torch as OT
    for name, s_in in (("h36m_351", 20), ("h36m_81", 4)):
        cfg, spec, w, x = _setup(name, B=3)
        m = stride_mask.batch_stride_masks_train(spec.n_tok, cfg.SEQUENCE_STRIDE, [s_in], 3, seed=1)
        f, c = O.test_step(spec, w, x, m, np.float64)
        wt = OT.to_torch(w, torch.float64)
        ft, ct = OT.test_step(spec, wt, torch.tensor(x, dtype=torch.float64), torch.tensor(m))
        assert np.abs(ft.numpy() - f).max() < 1e-10 and np.abs(ct.numpy() - c).max() < 1e-10
