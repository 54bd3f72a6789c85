"""Single-kernel parity on the B200: each CUDA kernel against the oracle, through the C ABI."""
import ctypes
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import forward_np as O
from uplift_upsample_3dhpe_b200 import _lib, stride_mask


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


_KEEP = []          # ctypes pointers do not own the tensors: keep them alive until the test ends


@pytest.fixture(autouse=True)
def _release():
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    t = t.to(dtype) if dtype is not None else t
    _KEEP.append(t)
    return t


@pytest.mark.parametrize("B,n_tok", [(1, 71), (7, 41), (513, 71), (3000, 71), (5, 128), (4, 1)])
def test_gather_list_is_bit_exact(lib, B, n_tok):
    rng = np.random.default_rng(B)
    mask = rng.random((B, n_tok)) < 0.4
    if B > 2:
        mask[1] = False          # an all-masked window
        mask[2] = True
    m = dev(mask.astype(np.uint8))
    scratch = torch.zeros(B + 1, dtype=torch.int32, device="cuda")
    lst = torch.full((B * n_tok,), -1, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
    _lib.check(lib.uu_op_build_gather(P(m), B, n_tok, P(scratch), P(lst), P(cnt), None))
    torch.cuda.synchronize()
    want = np.nonzero(mask.reshape(-1))[0]
    assert int(cnt[0]) == want.size
    assert np.array_equal(lst.cpu().numpy()[:want.size], want.astype(np.int32))
    assert np.array_equal(scratch.cpu().numpy()[:B], np.concatenate([[0], np.cumsum(mask.sum(1))[:-1]]))


def test_gather_empty_mask(lib):
    m = torch.zeros((9, 71), dtype=torch.uint8, device="cuda")
    scratch = torch.zeros(10, dtype=torch.int32, device="cuda")
    lst = torch.full((9 * 71,), -1, dtype=torch.int32, device="cuda")
    cnt = torch.ones(4, dtype=torch.int32, device="cuda")
    _lib.check(lib.uu_op_build_gather(P(m), 9, 71, P(scratch), P(lst), P(cnt), None))
    torch.cuda.synchronize()
    assert int(cnt[0]) == 0 and int(lst.max()) == -1


def test_token_fill(lib):
    rng = np.random.default_rng(0)
    B, N, d = 37, 71, 384
    mask = np.stack([stride_mask.stride_mask(N, 5, 20, shift_tokens=b % 4 - 2) for b in range(B)])
    x0 = rng.normal(size=(B * N, d)).astype(np.float32)
    tok = rng.normal(size=d).astype(np.float32)
    pe = rng.normal(size=(N, d)).astype(np.float32)
    x = dev(x0)
    _lib.check(lib.uu_op_token_fill(P(dev(mask.astype(np.uint8))), B * N, N, d, P(dev(tok)), P(dev(pe)), P(x), None))
    torch.cuda.synchronize()
    want = x0.copy().reshape(B, N, d)
    want[~mask] = (tok + pe)[np.nonzero(~mask)[1]]
    assert np.array_equal(x.cpu().numpy().reshape(B, N, d), want)      # selection + one fp32 add: exact


@pytest.mark.parametrize("with_table", [False, True])
def test_layernorm(lib, with_table):
    rng = np.random.default_rng(1)
    rows, d, period = 1001, 384, 23
    x0 = (rng.normal(size=(rows, d)) * 3 + 1.5).astype(np.float32)
    g = (1 + 0.1 * rng.normal(size=d)).astype(np.float32)
    b = (0.1 * rng.normal(size=d)).astype(np.float32)
    table = rng.normal(size=(period, d)).astype(np.float32)
    xin = x0 + table[np.arange(rows) % period] if with_table else x0
    want = O.layer_norm(xin.astype(np.float64), g.astype(np.float64), b.astype(np.float64), 1e-5)
    for y_bf16 in (0, 1):
        x = dev(x0)
        y = torch.empty((rows, d), dtype=torch.bfloat16 if y_bf16 else torch.float32, device="cuda")
        _lib.check(lib.uu_op_layernorm(P(x), rows, d, P(dev(g)), P(dev(b)), 1e-5, P(dev(table)) if with_table else None,
                                       period, P(y), y_bf16, None))
        torch.cuda.synchronize()
        err = np.abs(y.float().cpu().numpy() - want).max()
        assert err < (3e-2 if y_bf16 else 2e-5), err
        if with_table:
            assert np.array_equal(x.cpu().numpy(), xin)                 # PE add is written back


@pytest.mark.parametrize("S,dh,masked", [(71, 48, True), (71, 48, False), (23, 48, False), (3, 48, False),
                                          (41, 48, True), (128, 64, False), (11, 16, False)])
def test_attention(lib, S, dh, masked):
    rng = np.random.default_rng(S)
    B, H = 5, 8
    d = H * dh
    qkv = rng.normal(size=(B, S, 3 * d)).astype(np.float32)
    keep = np.ones((B, S), dtype=bool)
    if masked:
        keep = rng.random((B, S)) < 0.3
        keep[0] = False                      # all-masked window -> uniform attention (fp32 rounding of -1e9)
        keep[1, :] = False; keep[1, S // 2] = True

    def ref(dtype):
        q, k, v = [qkv[..., i * d:(i + 1) * d].astype(dtype).reshape(B, S, H, dh).transpose(0, 2, 1, 3) for i in range(3)]
        logits = (q @ k.transpose(0, 1, 3, 2)) / dtype(math.sqrt(dh))
        if masked:
            logits = logits + (1 - keep[:, None, None, :].astype(dtype)) * dtype(-1e9)
        return (O.softmax(logits) @ v).transpose(0, 2, 1, 3).reshape(B, S, d)

    want = ref(np.float32)                   # fp32 semantics define the all-masked case
    out = torch.empty((B, S, d), dtype=torch.float32, device="cuda")
    _lib.check(lib.uu_op_attention(P(dev(qkv)), 0, B, S, H, dh, P(dev(keep.astype(np.uint8))) if masked else None, S,
                                   P(out), None))
    torch.cuda.synchronize()
    assert np.abs(out.cpu().numpy() - want).max() < 2e-5
    q16 = dev(qkv, torch.bfloat16)
    out16 = torch.empty((B, S, d), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.uu_op_attention(P(q16), 1, B, S, H, dh, P(dev(keep.astype(np.uint8))) if masked else None, S,
                                   P(out16), None))
    torch.cuda.synchronize()
    assert np.abs(out16.float().cpu().numpy() - want).max() < 6e-2


@pytest.mark.parametrize("S,B,masked", [(71, 5, False), (71, 300, True), (41, 7, True), (24, 9, False), (8, 33, False),
                                        (3, 4, False), (80, 3, True), (1, 5, False), (16, 2, True)])
def test_attention_tcgen05(lib, S, B, masked):
    """attn_tc5.cu: q k^T and P v on tcgen05 (scores / probabilities in tensor memory, P read back as the TMEM A operand,
    V read MN-major from the TMA box), softmax with the reference's literal -1e9 key term (vit:117-123): all-masked
    windows give uniform attention.  Reference: fp32 numpy on the bf16-rounded inputs."""
    rng = np.random.default_rng(S + B)
    H, dh = 8, 48
    d = H * dh
    qkv = rng.normal(size=(B, S, 3 * d)).astype(np.float32)
    q16 = dev(qkv, torch.bfloat16)
    qr = q16.float().cpu().numpy()
    keep = np.ones((B, S), dtype=bool)
    if masked:
        keep = rng.random((B, S)) < 0.3
        keep[0] = False
        keep[1, :] = False; keep[1, S // 2] = True
    q, k, v = [qr[..., i * d:(i + 1) * d].reshape(B, S, H, dh).transpose(0, 2, 1, 3) for i in range(3)]
    logits = (q @ k.transpose(0, 1, 3, 2)) / np.float32(math.sqrt(dh))
    if masked:
        logits = logits + (1 - keep[:, None, None, :].astype(np.float32)) * np.float32(-1e9)
    want = (O.softmax(logits) @ v).transpose(0, 2, 1, 3).reshape(B, S, d)
    out16 = torch.full((B, S, d), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.uu_op_attention_tc5(P(q16), B, S, P(dev(keep.astype(np.uint8))) if masked else None, S, P(out16), None))
    torch.cuda.synchronize()
    got = out16.float().cpu().numpy()
    assert np.isfinite(got).all()
    err = np.abs(got - want).max()
    print(f"tcgen05 attention S={S} B={B}: max err {err:.3e}")
    assert err < 6e-2, err


@pytest.mark.parametrize("S,B,H,dh,masked", [(71, 5, 8, 48, True), (71, 3, 8, 48, False), (41, 4, 8, 48, True),
                                             (24, 9, 8, 48, False), (8, 17, 8, 48, False), (3, 4, 8, 48, False),
                                             (80, 2, 8, 48, True), (1, 5, 8, 48, False), (30, 3, 4, 32, True),
                                             (71, 2, 8, 64, True)])
@pytest.mark.parametrize("nsplit", [2, 3, 1])
def test_attention_train_mma(lib, S, B, H, dh, masked, nsplit):
    """attn_mma.cu (training step, vit:99-130 and its gradient): forward and backward on mma.sync TF32 against a float64
    autograd evaluation of the same lines.  nsplit = 3 (TF32 hi / lo operand split) must be fp32-grade: 2e-5 of the tensor
    scale; nsplit = 2 (bf16 hi + lo planes, 16 mantissa bits per operand: the training default) 2e-4; nsplit = 1 (plain TF32,
    10-bit operands) 1e-2.  All-masked windows give uniform attention (-1e9 term)."""
    rng = np.random.default_rng(S * 7 + B)
    d = H * dh
    qkv = rng.normal(size=(B, S, 3 * d)).astype(np.float32)
    dO = rng.normal(size=(B, S, d)).astype(np.float32)
    keep = np.ones((B, S), dtype=bool)
    if masked:
        keep = rng.random((B, S)) < 0.4
        keep[0] = False
        keep[1, :] = False; keep[1, S // 2] = True
    t = torch.tensor(qkv, dtype=torch.float64, requires_grad=True)
    q, k, v = [t[..., i * d:(i + 1) * d].reshape(B, S, H, dh).permute(0, 2, 1, 3) for i in range(3)]
    logits = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if masked:
        # fp32 semantics of `logits + (1 - mask) * -1e9` (vit:117-119): the score is absorbed (|score| < 32 = half an ulp
        # of 1e9), the value is exactly -1e9 and the gradient still passes through the addition
        drop = torch.tensor(~keep[:, None, None, :]).expand_as(logits)
        logits = torch.where(drop, logits + (-1e9 - logits).detach(), logits)
    want = (torch.softmax(logits, -1) @ v).permute(0, 2, 1, 3).reshape(B, S, d)
    want.backward(torch.tensor(dO, dtype=torch.float64))
    want_g = t.grad.numpy()
    want = want.detach().numpy()
    qd, gd = dev(qkv), dev(dO)
    km = P(dev(keep.astype(np.uint8))) if masked else None
    out = torch.full((B, S, d), float("nan"), dtype=torch.float32, device="cuda")
    dq = torch.full((B, S, 3 * d), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(lib.uu_op_attention_train(P(qd), None, B, S, H, dh, km, S, P(out), None, nsplit, None))
    _lib.check(lib.uu_op_attention_train(P(qd), P(gd), B, S, H, dh, km, S, None, P(dq), nsplit, None))
    torch.cuda.synchronize()
    got, got_g = out.cpu().numpy(), dq.cpu().numpy()
    assert np.isfinite(got).all() and np.isfinite(got_g).all()
    e_f = np.abs(got - want).max() / np.abs(want).max()
    e_b = np.abs(got_g - want_g).max() / np.abs(want_g).max()
    print(f"mma attention S={S} B={B} dh={dh} nsplit={nsplit}: fwd {e_f:.2e} bwd {e_b:.2e} (relative to the tensor max)")
    tol = {3: 2e-5, 2: 2e-4, 1: 1e-2}[nsplit]
    assert e_f < tol and e_b < tol, (e_f, e_b)


@pytest.mark.parametrize("M,N,K,flags", [(300, 384, 384, 0), (129, 51, 384, 0), (1000, 768, 384, 1),
                                         (77, 384, 2304, 2), (64, 1152, 544, 3)])
def test_gemm_f32(lib, M, N, K, flags):
    rng = np.random.default_rng(M)
    A = rng.normal(size=(M, K)).astype(np.float32)
    Wm = (rng.normal(size=(K, N)) / math.sqrt(K)).astype(np.float32)
    bias = rng.normal(size=N).astype(np.float32)
    res = rng.normal(size=(M, N)).astype(np.float32)
    want = A.astype(np.float64) @ Wm + bias
    if flags & 1:
        want = np.maximum(want, 0)
    if flags & 2:
        want = want + res
    C = dev(res.copy())                       # residual aliases the output, as in the forward pass
    _lib.check(lib.uu_op_gemm_f32(P(dev(A)), K, P(dev(Wm)), M, N, K, P(dev(bias)), flags, P(C), N, P(C), N, None))
    torch.cuda.synchronize()
    assert np.abs(C.cpu().numpy() - want).max() < 2e-5 * math.sqrt(K)


@pytest.mark.parametrize("M,N,K,flags,c_bf16", [(128, 128, 64, 0, 0), (300, 384, 384, 0, 0), (129, 51, 384, 0, 0),
                                                (1000, 768, 384, 1, 1), (77, 384, 2304, 2, 0), (640, 1152, 544, 1, 1),
                                                (36352, 1152, 384, 0, 1)])
def test_gemm_bf16_tcgen05(lib, M, N, K, flags, c_bf16):
    rng = np.random.default_rng(M + N)
    A = dev(rng.normal(size=(M, K)).astype(np.float32), torch.bfloat16)
    Wm = dev((rng.normal(size=(K, N)) / math.sqrt(K)).astype(np.float32), torch.bfloat16)
    n_pad = (N + 63) // 64 * 64
    Wt = torch.zeros((n_pad, K), dtype=torch.bfloat16, device="cuda")
    Wt[:N] = Wm.t()
    bias = rng.normal(size=N).astype(np.float32)
    res = rng.normal(size=(M, N)).astype(np.float32)
    want = A.float().cpu().numpy().astype(np.float64) @ Wm.float().cpu().numpy().astype(np.float64) + bias
    if flags & 1:
        want = np.maximum(want, 0)
    if flags & 2:
        want = want + res
    if c_bf16:
        C = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
        r = dev(res)
    else:
        C = dev(res.copy())
        r = C
    _lib.check(lib.uu_op_gemm_bf16(P(A), K, M, K, P(Wt), n_pad, N, P(dev(bias)), flags, P(r), N, P(C), c_bf16, N, None))
    torch.cuda.synchronize()
    got = C.float().cpu().numpy()
    assert np.isfinite(got).all()
    err = np.abs(got - want).max()
    assert err < (5e-2 if c_bf16 else 1e-3), err


@pytest.mark.parametrize("M,N,K,flags", [(300, 384, 384, 0), (1000, 768, 384, 1), (355, 384, 768, 2), (4096, 1152, 384, 0),
                                         (512, 64, 2304, 2)])
def test_gemm_tf32_tcgen05(lib, M, N, K, flags):
    """kind::tf32: fp32 operands straight from TMA, TF32 products (the tensor core drops the low 13 mantissa bits of
    each operand: truncation, up to 2^-10 relative per operand, biased towards zero), fp32 accumulation.  Against
    float64 on the same fp32 data the measured max error is 3.4e-3 .. 4.2e-3 on unit-variance outputs (|max| ~ 5):
    ~1e-3 of the output scale."""
    rng = np.random.default_rng(M + N + K)
    A = rng.normal(size=(M, K)).astype(np.float32)
    Bt = (rng.normal(size=(N, K)) / math.sqrt(K)).astype(np.float32)
    bias = rng.normal(size=N).astype(np.float32)
    res = rng.normal(size=(M, N)).astype(np.float32)
    want = A.astype(np.float64) @ Bt.astype(np.float64).T + bias
    if flags & 1:
        want = np.maximum(want, 0)
    if flags & 2:
        want = want + res
    C = dev(res.copy())
    _lib.check(lib.uu_op_gemm_tf32(P(dev(A)), K, M, K, P(dev(Bt)), K, N, P(dev(bias)), flags, P(C), N, P(C), N, None))
    torch.cuda.synchronize()
    err = np.abs(C.cpu().numpy() - want).max()
    print(f"tf32 gemm {M}x{N}x{K}: max err {err:.2e}")
    assert err < 8e-3, err


@pytest.mark.parametrize("R,Kd,Nd,ldx_extra,acc", [(300, 384, 384, 0, 0), (5000, 384, 768, 0, 1), (36352, 768, 384, 0, 1),
                                                    (1000, 544, 384, 0, 0), (777, 2304, 384, 768, 1), (4100, 384, 64, 0, 0),
                                                    (2900, 384, 384, 768, 1)])
def test_wgrad_tf32_tcgen05(lib, R, Kd, Nd, ldx_extra, acc):
    """dW (+)= X^T dY on tcgen05 kind::tf32 with MN-major operands (contraction over the row index, TMA straight from the
    row-major tape), split-K partials summed in a fixed order.  Reference: float64 on the same fp32 data; TF32 truncation
    gives ~1e-3 of the output scale (outputs here have unit variance).  ldx_extra > 0: strided rows (the zero-padded
    conv input / a column slice of the q|k|v gradient).  Two runs must agree bit for bit (no atomics)."""
    rng = np.random.default_rng(R + Kd + Nd)
    ldx, ldy = Kd + ldx_extra, Nd + (128 if ldx_extra else 0)
    Xf = rng.normal(size=(R, ldx)).astype(np.float32)
    Yf = (rng.normal(size=(R, ldy)) / math.sqrt(R)).astype(np.float32)
    W0 = rng.normal(size=(Kd, Nd)).astype(np.float32)
    want = Xf[:, :Kd].astype(np.float64).T @ Yf[:, :Nd].astype(np.float64) + (W0 if acc else 0)
    Xd, Yd = dev(Xf), dev(Yf)
    outs = []
    for _ in range(2):
        Wd = dev(W0.copy())
        _lib.check(lib.uu_op_wgrad_tf32(P(Xd), ldx, P(Yd), ldy, R, Kd, Nd, P(Wd), acc, None))
        torch.cuda.synchronize()
        outs.append(Wd.cpu().numpy())
    err = np.abs(outs[0] - want).max()
    print(f"tf32 wgrad R={R} {Kd}x{Nd}: max err {err:.2e}")
    assert err < 8e-3, err
    assert np.array_equal(outs[0], outs[1]), "split-K reduction must be deterministic"


@pytest.mark.parametrize("rows,N,relu", [(71, 1152, 0), (300, 768, 1), (36352, 1152, 0), (5000, 128, 1)])
def test_layernorm_folded_into_gemm(lib, rows, N, relu):
    """EPI_LNFOLD: the GEMM reads the raw bf16 stream with gamma folded into W and finishes LayerNorm (vit:168-171) in
    the epilogue from per-row statistics.  Reference: float64 LN of the same bf16 rows, fp32 weights.  Rows get a large
    common offset (|mean| >> std) so that a wrong mean / column-sum term cannot hide."""
    d = 384
    rng = np.random.default_rng(rows + N)
    x = (rng.normal(size=(rows, d)) * rng.uniform(0.5, 3.0, (rows, 1)) + rng.normal(size=(rows, 1)) * 4).astype(np.float32)
    xb = dev(x, torch.bfloat16)
    gamma = (1 + 0.3 * rng.normal(size=d)).astype(np.float32)
    beta = (0.3 * rng.normal(size=d)).astype(np.float32)
    Wm = (rng.normal(size=(d, N)) / math.sqrt(d)).astype(np.float32)
    bias = rng.normal(size=N).astype(np.float32)
    x64 = xb.float().cpu().numpy().astype(np.float64)
    mu, var = x64.mean(1, keepdims=True), x64.var(1, keepdims=True)
    want = ((x64 - mu) / np.sqrt(var + 1e-5) * gamma + beta) @ Wm.astype(np.float64) + bias
    if relu:
        want = np.maximum(want, 0)
    out = torch.full((rows, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.uu_op_ln_gemm_bf16(P(xb), rows, d, P(dev(gamma)), P(dev(beta)), 1e-5, P(dev(Wm)), P(dev(bias)), N, relu,
                                      P(out), None))
    got = out.float().cpu().numpy()
    assert np.isfinite(got).all()
    err = np.abs(got - want)
    # bf16 weights (2^-9 relative, sqrt(d) terms) + bf16 output rounding of O(1..4) values
    assert err.max() < 6e-2 and np.sqrt((err ** 2).mean()) < 1e-2, (err.max(), np.sqrt((err ** 2).mean()))


@pytest.mark.parametrize("rows,h", [(256, 768), (300, 768), (71, 768), (36352, 768), (5000, 128), (1000, 512)])
def test_fused_mlp_tcgen05(lib, rows, h):
    """mlp_tc.cuh: x += fc2(ReLU(fc1(LN2(x)))) (vit:190-195) as ONE cta_group::2 tcgen05 kernel, hidden chunks handed from
    the fc1 accumulator to the fc2 operand through shared memory.  Reference: float64 on the same bf16 rows with the fp32
    weights; the kernel rounds weights and the hidden activation to bf16 (same error budget as the two-GEMM path)."""
    d = 384
    rng = np.random.default_rng(rows + h)
    x = (rng.normal(size=(rows, d)) * rng.uniform(0.5, 3.0, (rows, 1)) + rng.normal(size=(rows, 1)) * 4).astype(np.float32)
    xb = dev(x, torch.bfloat16)
    gamma = (1 + 0.3 * rng.normal(size=d)).astype(np.float32)
    beta = (0.3 * rng.normal(size=d)).astype(np.float32)
    W1 = (rng.normal(size=(d, h)) / math.sqrt(d)).astype(np.float32)
    b1 = rng.normal(size=h).astype(np.float32)
    W2 = (rng.normal(size=(h, d)) / math.sqrt(h)).astype(np.float32)
    b2 = rng.normal(size=d).astype(np.float32)
    x64 = xb.float().cpu().numpy().astype(np.float64)
    mu, var = x64.mean(1, keepdims=True), x64.var(1, keepdims=True)
    hid = np.maximum(((x64 - mu) / np.sqrt(var + 1e-5) * gamma + beta) @ W1.astype(np.float64) + b1, 0)
    want = x64 + hid @ W2.astype(np.float64) + b2
    stats = torch.full((rows, d // 64, 2), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(lib.uu_op_mlp_bf16(P(xb), rows, d, h, P(dev(gamma)), P(dev(beta)), 1e-5, P(dev(W1)), P(dev(b1)), P(dev(W2)),
                                  P(dev(b2)), P(stats), None))
    got = xb.float().cpu().numpy()
    assert np.isfinite(got).all()
    err = np.abs(got - want)
    print(f"fused mlp rows={rows} h={h}: max err {err.max():.3e} rms {np.sqrt((err ** 2).mean()):.3e}")
    assert err.max() < 0.12 and np.sqrt((err ** 2).mean()) < 2e-2, (err.max(), np.sqrt((err ** 2).mean()))
    st = stats.cpu().numpy().astype(np.float64)
    g64 = got.astype(np.float64).reshape(rows, d // 64, 64)
    assert np.abs(st[..., 0] - g64.sum(-1)).max() < 1.5 and np.abs(st[..., 1] - (g64 ** 2).sum(-1)).max() < 0.02 * (g64 ** 2).sum(-1).max()


@pytest.mark.parametrize("rows,K,resid", [(71, 384, 1), (300, 768, 1), (36352, 384, 1), (2048, 768, 1), (5000, 544, 0), (213, 768, 0)])
def test_residual_and_table_epilogues(lib, rows, K, resid):
    """EPI_RESID_BF16 (x += A W + b in place) and the positional-table epilogue (x = A W + b + pe[row % period]), both
    with the (sum, sum of squares) partials per 64-column slot that the next folded LayerNorm consumes."""
    d, period = 384, 71
    rng = np.random.default_rng(rows + K)
    A = dev(rng.normal(size=(rows, K)).astype(np.float32), torch.bfloat16)
    Wm = (rng.normal(size=(K, d)) / math.sqrt(K)).astype(np.float32)
    bias = rng.normal(size=d).astype(np.float32)
    x0 = dev((rng.normal(size=(rows, d)) * 2).astype(np.float32), torch.bfloat16)
    table = rng.normal(size=(period, d)).astype(np.float32)
    # the kernel multiplies bf16(A) by bf16(W): same operands for the reference
    Wb = dev(Wm, torch.bfloat16).float().cpu().numpy().astype(np.float64)
    want = A.float().cpu().numpy().astype(np.float64) @ Wb + bias
    if resid:
        want = want + x0.float().cpu().numpy()
    else:
        want = want + table[np.arange(rows) % period]
    x = x0.clone()
    stats = torch.full((rows, d // 64, 2), float("nan"), device="cuda")
    _lib.check(lib.uu_op_resid_gemm_bf16(P(A), rows, K, P(dev(Wm)), P(dev(bias)), d, P(x), resid, P(dev(table)), period,
                                         P(stats), None))
    got = x.float().cpu().numpy()
    assert np.abs(got - want).max() < 4e-2          # bf16 rounding of O(1..8) outputs
    w = want.reshape(rows, d // 64, 64)
    st = stats.cpu().numpy()
    assert np.isfinite(st).all()
    assert np.abs(st[..., 0] - w.sum(-1)).max() < 2e-3 and np.abs(st[..., 1] - (w ** 2).sum(-1)).max() < 2e-2


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B,masked", [(3, False), (5, True), (40, True)])
def test_spatial_transformer(lib, precision, B, masked):
    """K2 alone (S1-S3 + spatial_norm) against the oracle's LayerNorm'd (frames, 544) features, valid frames only."""
    from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, spec_from_config, weights
    from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer
    cfg = UpliftUpsampleConfig.preset("h36m_351", MASK_STRIDE=20 if masked else 5)
    spec = spec_from_config(cfg)
    w = weights.init_weights(spec, 3, perturb=True)
    rng = np.random.default_rng(B)
    x = rng.uniform(-1, 1, (B, spec.n_tok, 17, 2)).astype(np.float32)
    if masked:
        m = np.stack([stride_mask.stride_mask(spec.n_tok, 5, 20, shift_tokens=b % 4 - 2) for b in range(B)])
        m[1] = False
    else:
        m = np.ones((B, spec.n_tok), dtype=bool)
    _, _, inter = O.forward(spec, w, x * m[:, :, None, None], m, dtype=np.float64, return_intermediates=True)
    want = inter["spatial"].reshape(B * spec.n_tok, 544)[m.reshape(-1)]
    model = build_uplift_upsample_transformer(cfg, precision=precision, weights=w)
    out = torch.full((B * spec.n_tok, 544), float("nan"), dtype=torch.bfloat16 if precision == "bf16" else torch.float32,
                     device="cuda")
    n = ctypes.c_int32(-1)
    _lib.check(lib.uu_op_spatial(model._h, P(dev(x)), P(dev(m.astype(np.uint8))) if masked else None, B, P(out),
                                 ctypes.byref(n), None))
    assert n.value == want.shape[0]
    got = out.float().cpu().numpy()[:n.value]
    assert np.isfinite(got).all()
    err = np.abs(got - want).max()
    rms = float(np.sqrt(((got - want) ** 2).mean()))
    print(f"spatial {precision} B={B} masked={masked}: max|err| {err:.3e} rms {rms:.3e} (|want|max {np.abs(want).max():.2f})")
    # bf16 path: four blocks of bf16/fp16-operand MMAs on O(1) LayerNorm'd features; measured rms 7e-3 (4x the bf16
    # rounding of the output itself), max 0.05-0.11 over 1e5..1e6 values
    assert err < (1e-4 if precision == "fp32" else 0.2), err
    assert rms < (1e-5 if precision == "fp32" else 1.2e-2), rms
    model.close()
