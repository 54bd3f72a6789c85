"""End-to-end parity on the B200: the CUDA forward (through the C ABI) against the oracle on the
same seeded inputs, both precisions, BASELINE configs at oracle-sized batches."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import forward_np as O
from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, spec_from_config, stride_mask, weights
from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer
from uplift_upsample_3dhpe_b200.model import test_step as run_test_step

# Tolerances (absolute, outputs are O(1..10) with the perturbed random-init weights):
#   fp32 path: <= 1e-4 against the fp64 oracle (north_star's fp32 bound; the fp32 oracle itself sits ~1e-5 away)
#   bf16 path: bf16/fp16 operands, fp32 accumulation, LayerNorm/softmax/residual adds in fp32 registers, bf16-resident
#              activations — measured 0.09-0.15 on outputs of magnitude ~10 (rms ~3); bound 0.2 = 1.33 x the largest measurement
TOL = {"fp32": 1e-4, "bf16": 0.2}


def _case(name, s_in, B, mode, seed=0):
    cfg = UpliftUpsampleConfig.preset(name)
    spec = spec_from_config(cfg)
    w = weights.init_weights(spec, 1, perturb=True)
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, (B, spec.n_tok, 17, 2)).astype(np.float32)
    so = cfg.SEQUENCE_STRIDE
    if mode == "centred":
        m = np.stack([stride_mask.stride_mask(spec.n_tok, so, s_in)] * B)
    elif mode == "shifted":
        lo, hi, ep = stride_mask.rand_shift_range(s_in // so)
        sh = rng.integers(lo, hi, size=B, endpoint=ep)
        m = np.stack([stride_mask.stride_mask(spec.n_tok, so, s_in, shift_tokens=int(s)) for s in sh])
    else:  # "global": eval alignment, includes all-masked windows when i % s_out != 0
        m = stride_mask.batch_stride_masks_eval(spec.n_tok, so, s_in, center_frames=np.arange(B) * 3)
    return cfg, spec, w, x, m


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,s_in,B,mode", [
    ("h36m_81", 4, 16, "centred"),
    ("h36m_81", 10, 9, "shifted"),
    ("h36m_81", 20, 5, "global"),
    ("h36m_351", 5, 6, "centred"),
    ("h36m_351", 10, 7, "shifted"),
    ("h36m_351", 20, 8, "shifted"),
    ("h36m_351", 20, 11, "global"),
])
def test_forward_matches_oracle(name, s_in, B, mode, precision):
    cfg, spec, w, x, m = _case(name, s_in, B, mode)
    # Windows without any valid token (global alignment, i % s_out != 0) are defined by the reference's
    # fp32 arithmetic (x - 1e9 rounds to -1e9 => uniform attention, vit:122-123); fp64 would not round.
    all_masked = bool((m.sum(axis=1) == 0).any())
    want_full, want_central = O.test_step(spec, w, x, m, dtype=np.float32 if all_masked else np.float64)
    model = build_uplift_upsample_transformer(cfg, precision=precision, weights=w)
    full, central = run_test_step(model, torch.from_numpy(x).cuda(), torch.from_numpy(m).cuda())
    torch.cuda.synchronize()
    e_f = np.abs(full.cpu().numpy() - want_full).max()
    e_c = np.abs(central.cpu().numpy() - want_central).max()
    print(f"{name} s_in={s_in} {mode} {precision}: max|err| full {e_f:.3e} central {e_c:.3e}")
    assert e_f <= TOL[precision] and e_c <= TOL[precision]
    assert model.last_launch_count > 0
    model.close()


def test_masked_frame_values_are_never_read():
    cfg, spec, w, x, m = _case("h36m_351", 20, 6, "shifted")
    model = build_uplift_upsample_transformer(cfg, precision="fp32", weights=w)
    md = torch.from_numpy(m).cuda()
    f1, c1 = run_test_step(model, torch.from_numpy(x).cuda(), md)
    x2 = x.copy()
    x2[~m] = np.nan                      # garbage in frames without 2-D input
    f2, c2 = run_test_step(model, torch.from_numpy(x2).cuda(), md)
    torch.cuda.synchronize()
    assert torch.equal(f1, f2) and torch.equal(c1, c2)
    model.close()


def test_no_strided_input_model_and_weight_roundtrip(tmp_path):
    cfg = UpliftUpsampleConfig.preset("h36m_81", MASK_STRIDE=None)
    spec = spec_from_config(cfg)
    w = weights.init_weights(spec, 3, perturb=True)
    x = np.random.default_rng(5).uniform(-1, 1, (4, 41, 17, 2)).astype(np.float32)
    want_full, want_central = O.forward(spec, w, x, None)
    model = build_uplift_upsample_transformer(cfg, precision="fp32", weights=w)
    full, central = model(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    assert np.abs(full.cpu().numpy() - want_full).max() < 1e-4
    assert np.abs(central.cpu().numpy() - want_central).max() < 1e-4
    got = model.get_weights()
    assert all(np.array_equal(got[k], w[k]) for k in w)
    p = str(tmp_path / "w.npz")
    model.save_weights(p)
    m2 = build_uplift_upsample_transformer(cfg, precision="fp32")
    m2.load_weights(p)
    f2, c2 = m2(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    assert torch.equal(c2, central) and torch.equal(f2, full)
    with pytest.raises(Exception):
        bad = dict(w); bad[("temporal_fc", 0)] = np.zeros((384, 50), np.float32)
        model.set_weights(bad)
    model.close(); m2.close()


def test_forward_host_and_batch_growth():
    cfg, spec, w, x, m = _case("h36m_351", 10, 12, "shifted")
    want_full, want_central = O.test_step(spec, w, x, m, dtype=np.float64)
    model = build_uplift_upsample_transformer(cfg, precision="fp32", weights=w)
    for B in (3, 12, 5):                 # workspace grows, then a smaller batch reuses it
        full = np.empty((B, 71, 17, 3), np.float32)
        central = np.empty((B, 17, 3), np.float32)
        model.forward_host(x[:B].copy(), m[:B].astype(np.uint8).copy(), full, central)
        assert np.abs(full - want_full[:B]).max() < 1e-4 and np.abs(central - want_central[:B]).max() < 1e-4
    model.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_large_batch_properties(precision):
    """BASELINE-size batch (512 windows): batch-permutation equivariance and agreement with a
    window-by-window evaluation of a subset — properties that do not need the oracle at full size."""
    cfg, spec, w, x, m = _case("h36m_351", 20, 512, "shifted", seed=3)
    model = build_uplift_upsample_transformer(cfg, precision=precision, weights=w)
    xd, md = torch.from_numpy(x).cuda(), torch.from_numpy(m).cuda()
    full, central = run_test_step(model, xd, md)
    perm = torch.randperm(512, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    fp, cp = run_test_step(model, xd[perm].contiguous(), md[perm].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(fp, full[perm]) and torch.equal(cp, central[perm])
    fs, cs = run_test_step(model, xd[100:116].contiguous(), md[100:116].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(cs, central[100:116]) and torch.equal(fs, full[100:116])
    want_full, want_central = O.test_step(spec, w, x[:4], m[:4], dtype=np.float64)
    assert np.abs(central[:4].cpu().numpy() - want_central).max() <= TOL[precision]
    model.close()


def test_benched_batch_against_the_oracle():
    """The benched configuration at its full size (h36m_351, s_in = 5, 4096 windows, bf16 schedule — bench.py's step) against
    the torch-CPU fp32 restatement of the reference forward over EVERY window (the fp32 restatement sits ~1e-5 from the float64
    one; a few seconds on the host cores), plus the error statistics the bench line quotes."""
    from oracle import forward_torch as OT
    cfg, spec, w, x, m = _case("h36m_351", 5, 4096, "centred", seed=11)
    model = build_uplift_upsample_transformer(cfg, precision="bf16", weights=w)
    full, central = run_test_step(model, torch.from_numpy(x).cuda(), torch.from_numpy(m).cuda())
    torch.cuda.synchronize()
    wt = OT.to_torch(w, torch.float32)
    ref_f, ref_c = [], []
    with torch.no_grad():
        for i in range(0, 4096, 256):
            f, c = OT.test_step(spec, wt, torch.from_numpy(x[i:i + 256]), torch.from_numpy(m[i:i + 256].astype(bool)))
            ref_f.append(f.numpy()); ref_c.append(c.numpy())
    ref_f, ref_c = np.concatenate(ref_f), np.concatenate(ref_c)
    e_f, e_c = np.abs(full.cpu().numpy() - ref_f), np.abs(central.cpu().numpy() - ref_c)
    print(f"B=4096 bf16: max|err| full {e_f.max():.3e} central {e_c.max():.3e}, rms {np.sqrt((e_f ** 2).mean()):.3e} "
          f"(output rms {np.sqrt((ref_f ** 2).mean()):.3f})")
    assert np.isfinite(full.cpu().numpy()).all()
    assert e_f.max() <= TOL["bf16"] and e_c.max() <= TOL["bf16"]
    assert np.sqrt((e_f ** 2).mean()) <= 0.04
    model.close()


@pytest.mark.parametrize("s_in", [5, 20])
def test_chunked_host_forward_is_identical_to_device_forward(s_in):
    """uu_forward_host splits the input copy of large bf16 batches into chunks overlapped with the spatial
    kernel; the result must be bit-identical to uu_forward on device-resident inputs (ragged last chunk, masks
    with different valid counts per window so the chunk boundaries fall at arbitrary gather-list positions)."""
    cfg, spec, w, x, m = _case("h36m_351", s_in, 2050, "shifted" if s_in > 5 else "centred", seed=9)
    m[7] = False                                  # an all-masked window inside a chunk
    model = build_uplift_upsample_transformer(cfg, precision="bf16", weights=w)
    full, central = run_test_step(model, torch.from_numpy(x).cuda(), torch.from_numpy(m).cuda())
    torch.cuda.synchronize()
    hf = np.empty((2050, 71, 17, 3), np.float32)
    hc = np.empty((2050, 17, 3), np.float32)
    model.forward_host(x, m.astype(np.uint8), hf, hc)
    assert np.array_equal(hc, central.cpu().numpy()) and np.array_equal(hf, full.cpu().numpy())
    model.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B,kind", [(1, "valid"), (3, "all_masked"), (130, "ragged")])
def test_edge_batches(precision, B, kind):
    """B = 1, a batch in which NO window has a valid token (empty gather list: zero-row GEMM, token fill everywhere,
    uniform attention in block 1), and a batch that is not a multiple of any tile size with ragged masks."""
    cfg = UpliftUpsampleConfig.preset("h36m_351", MASK_STRIDE=20)
    spec = spec_from_config(cfg)
    w = weights.init_weights(spec, 2, perturb=True)
    rng = np.random.default_rng(B)
    x = rng.uniform(-1, 1, (B, spec.n_tok, 17, 2)).astype(np.float32)
    if kind == "all_masked":
        m = np.zeros((B, spec.n_tok), dtype=bool)
    elif kind == "ragged":
        m = rng.random((B, spec.n_tok)) < 0.3
        m[5] = False
    else:
        m = np.stack([stride_mask.stride_mask(spec.n_tok, 5, 20)] * B)
    want_full, want_central = O.test_step(spec, w, x, m, dtype=np.float32)     # fp32 defines all-masked windows
    model = build_uplift_upsample_transformer(cfg, precision=precision, weights=w)
    full, central = run_test_step(model, torch.from_numpy(x).cuda(), torch.from_numpy(m).cuda())
    torch.cuda.synchronize()
    e = max(np.abs(full.cpu().numpy() - want_full).max(), np.abs(central.cpu().numpy() - want_central).max())
    print(f"edge {kind} B={B} {precision}: max|err| {e:.3e}")
    assert np.isfinite(full.cpu().numpy()).all() and e <= (2e-4 if precision == "fp32" else TOL["bf16"])
    model.close()


def test_graph_replay_matches_eager_and_follows_buffer_contents():
    """uu_forward on a non-default stream replays a captured CUDA graph from the third call with the same buffers:
    results must be bit-identical to the eager first call and must track new contents of the same buffers."""
    cfg, spec, w, x, m = _case("h36m_351", 10, 64, "shifted", seed=4)
    model = build_uplift_upsample_transformer(cfg, precision="bf16", weights=w)
    s = torch.cuda.Stream()
    xd, md = torch.from_numpy(x).cuda(), torch.from_numpy(m.astype(np.uint8)).cuda()
    full = torch.empty((64, 71, 17, 3), device="cuda")
    central = torch.empty((64, 17, 3), device="cuda")
    torch.cuda.synchronize()
    outs = []
    with torch.cuda.stream(s):
        for it in range(4):                                  # eager, capture, replay, replay
            model.forward_raw(xd.data_ptr(), md.data_ptr(), 64, full.data_ptr(), central.data_ptr(), s.cuda_stream)
            s.synchronize()
            outs.append((full.clone(), central.clone()))
        for f, c in outs[1:]:
            assert torch.equal(f, outs[0][0]) and torch.equal(c, outs[0][1])
        # new inputs and a different mask in the SAME buffers
        x2 = torch.from_numpy(np.ascontiguousarray(x[::-1])).cuda()
        m2 = torch.from_numpy(np.ascontiguousarray(m[::-1]).astype(np.uint8)).cuda()
        xd.copy_(x2); md.copy_(m2)
        model.forward_raw(xd.data_ptr(), md.data_ptr(), 64, full.data_ptr(), central.data_ptr(), s.cuda_stream)
        s.synchronize()
    assert torch.equal(central, outs[0][1].flip(0)) and torch.equal(full, outs[0][0].flip(0))
    # weights change under a cached graph: derived packs are refreshed in place, the replay must see them
    w2 = {k: (v * 1.01).astype(np.float32) for k, v in w.items()}
    model.set_weights(w2)
    with torch.cuda.stream(s):
        model.forward_raw(xd.data_ptr(), md.data_ptr(), 64, full.data_ptr(), central.data_ptr(), s.cuda_stream)
        s.synchronize()
    ref = build_uplift_upsample_transformer(cfg, precision="bf16", weights=w2)
    f2, c2 = run_test_step(ref, xd, md.bool())
    torch.cuda.synchronize()
    assert torch.equal(c2, central) and torch.equal(f2, full)
    model.close(); ref.close()
