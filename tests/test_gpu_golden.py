"""CUDA path (through the C ABI) against the golden vectors the reference's own code produced
(tests/golden, scripts/make_golden.py).  fp32 path: <= 1e-4 of the reference's float32 run on every
window (incl. all-masked ones) and of its float64 run where a valid token exists; bf16 path: <= 0.2
absolute on O(10) outputs (bf16 operands, fp32 accumulation)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from test_golden import FORWARD, load_forward_case  # noqa: E402
from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer  # noqa: E402
from uplift_upsample_3dhpe_b200.model import test_step as run_test_step  # noqa: E402


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("path", FORWARD, ids=[os.path.basename(p)[8:-4] for p in FORWARD])
def test_cuda_forward_matches_reference_golden(path, precision):
    cfg, spec, w, z = load_forward_case(path)
    x, m = z["x"], z["mask"]
    model = build_uplift_upsample_transformer(cfg, precision=precision, weights=w)
    full, central = run_test_step(model, torch.from_numpy(x).cuda(), torch.from_numpy(m).cuda())
    torch.cuda.synchronize()
    full, central = full.cpu().numpy(), central.cpu().numpy()
    model.close()
    # tf32: the fp32 schedule with TF32 products in the large GEMMs (10-bit mantissa operands): measured up to 1.9e-2 (truncation, not rounding, of the operands)
    tol = {"fp32": 1e-4, "tf32": 3e-2, "bf16": 0.2}[precision]
    e32 = max(np.abs(central - z["central_f32"]).max(), np.abs(full - z["full_f32"]).max())
    valid = m.sum(axis=1) > 0
    e64 = max(np.abs(central[valid] - z["central"][valid]).max(), np.abs(full[valid] - z["full"][valid]).max())
    print(f"{os.path.basename(path)} {precision}: max|err| vs reference f32 {e32:.3e}, vs f64 (valid windows) {e64:.3e}")
    assert e32 <= tol and e64 <= tol


def test_h5_written_by_us_loads_through_c_abi(tmp_path):
    """Same file path the reference's loader was fed in make_golden.py: write .h5, load, compare outputs."""
    cfg, spec, w, z = load_forward_case(FORWARD[0])
    from uplift_upsample_3dhpe_b200 import h5lite
    p = str(tmp_path / "w.h5")
    h5lite.save_keras_weights(p, spec, w)
    model = build_uplift_upsample_transformer(cfg, precision="fp32")
    model.load_weights(p)
    full, central = run_test_step(model, torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["mask"]).cuda())
    torch.cuda.synchronize()
    assert np.abs(central.cpu().numpy() - z["central_f32"]).max() <= 1e-4
    model.close()


from test_golden import WINDOWS  # noqa: E402


@pytest.mark.parametrize("path", [p for p in WINDOWS if "eval" in p], ids=lambda p: os.path.basename(p)[8:-4])
def test_device_window_gather_is_bit_exact(path):
    """uu_op_window_gather against the reference generator's windows and stride masks (tests/golden)."""
    import ctypes
    from uplift_upsample_3dhpe_b200 import _lib
    lib = _lib.load()
    z = np.load(path, allow_pickle=False)
    n_tok, stride, s_in = int(z["n_tok"]), int(z["stride"]), int(z["mask_stride"])
    video = torch.from_numpy(z["video_2d"]).cuda()
    centers = torch.from_numpy(z["centers"].astype(np.int32)).cuda()
    B, T = centers.shape[0], video.shape[0]
    src = torch.empty((B, n_tok), dtype=torch.int32, device="cuda")
    mask = torch.empty((B, n_tok), dtype=torch.uint8, device="cuda")
    x = torch.empty((B, n_tok, 17, 2), dtype=torch.float32, device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(lib.uu_op_window_gather(P(video), T, P(centers), B, n_tok, 17, stride, s_in, 1, P(src), P(mask), P(x), None))
    torch.cuda.synchronize()
    assert np.array_equal(mask.cpu().numpy().astype(bool), z["stride_masks"])
    assert np.array_equal(x.cpu().numpy(), z["seq_2d"])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_video_equals_forward_on_materialised_windows(precision):
    """The fused gather changes where the key-points are read from, not a single arithmetic operation."""
    from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, spec_from_config, stride_mask, weights
    cfg = UpliftUpsampleConfig.preset("h36m_351", MASK_STRIDE=20)
    spec = spec_from_config(cfg)
    w = weights.init_weights(spec, 5, perturb=True)
    rng = np.random.default_rng(11)
    T = 260
    video = rng.uniform(-1, 1, (T, 17, 2)).astype(np.float32)
    centers = np.concatenate([np.arange(0, 40, 5), np.arange(100, 140, 5), np.arange(T - 40, T, 5), [3, 7]]).astype(np.int32)
    src = stride_mask.window_source_frames(spec.n_tok, 5, T, centers)
    m = stride_mask.batch_stride_masks_eval(spec.n_tok, 5, 20, centers)
    model = build_uplift_upsample_transformer(cfg, precision=precision, weights=w)
    f1, c1 = run_test_step(model, torch.from_numpy(video[src]).cuda(), torch.from_numpy(m).cuda())
    f2, c2 = model.forward_video(torch.from_numpy(video).cuda(), torch.from_numpy(centers).cuda(), 5, 20)
    torch.cuda.synchronize()
    assert torch.equal(c1, c2) and torch.equal(f1, f2)
    hc = np.empty((len(centers), 17, 3), dtype=np.float32)
    model.forward_video_host(video, centers, 5, 20, hc)
    assert np.array_equal(hc, c1.cpu().numpy())
    model.close()


import glob  # noqa: E402

from test_golden import GOLDEN  # noqa: E402


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_flip_tta_matches_reference_golden(precision):
    """uu_forward_tta against eval.py:152-180 executed on the reference model (tests/golden/tta_*.npz)."""
    for path in sorted(glob.glob(os.path.join(GOLDEN, "tta_*.npz"))):
        cfg, spec, w, z = load_forward_case(path)
        model = build_uplift_upsample_transformer(cfg, precision=precision, weights=w)
        full, central = model.forward_tta([torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["mask"]).cuda()])
        torch.cuda.synchronize()
        tol = 1e-4 if precision == "fp32" else 0.2
        e = max(np.abs(central.cpu().numpy() - z["central"]).max(), np.abs(full.cpu().numpy() - z["full"]).max())
        print(f"tta {os.path.basename(path)} {precision}: max|err| {e:.3e}")
        assert e <= tol
        model.close()


def test_keyframe_interpolation_matches_reference_golden():
    from uplift_upsample_3dhpe_b200.model import keyframe_interp
    for path in sorted(glob.glob(os.path.join(GOLDEN, "interp_*.npz"))):
        z = np.load(path)
        out = keyframe_interp(torch.from_numpy(z["pred"]).cuda(), torch.from_numpy(z["frame_indices"]).cuda(), int(z["stride"]))
        torch.cuda.synchronize()
        # the reference interpolates in float64; fp32 weights and one rounding per product
        assert np.abs(out.cpu().numpy() - z["out"]).max() < 1e-6
        key = z["keyframes"]
        assert np.array_equal(out.cpu().numpy()[key], z["pred"][key])


def test_video_tta_equals_tta_on_materialised_windows():
    from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, spec_from_config, stride_mask, weights
    cfg = UpliftUpsampleConfig.preset("h36m_81", MASK_STRIDE=4)
    spec = spec_from_config(cfg)
    w = weights.init_weights(spec, 5, perturb=True)
    rng = np.random.default_rng(3)
    T = 90
    video = rng.uniform(-1, 1, (T, 17, 2)).astype(np.float32)
    centers = np.arange(0, T, 2, dtype=np.int32)
    src = stride_mask.window_source_frames(spec.n_tok, 2, T, centers)
    m = stride_mask.batch_stride_masks_eval(spec.n_tok, 2, 4, centers)
    model = build_uplift_upsample_transformer(cfg, precision="fp32", weights=w)
    f1, c1 = model.forward_tta([torch.from_numpy(video[src]).cuda(), torch.from_numpy(m).cuda()])
    f2, c2 = model.forward_video(torch.from_numpy(video).cuda(), torch.from_numpy(centers).cuda(), 2, 4, flip_tta=True)
    torch.cuda.synchronize()
    assert torch.equal(c1, c2) and torch.equal(f1, f2)
    model.close()


def test_pose_metrics_on_device_match_reference_golden():
    """uu_op_pose_metrics against common/dataset/metrics.py (tests/golden/metrics_*.npz): fp32 on the device, float64 in
    the reference."""
    import ctypes
    from uplift_upsample_3dhpe_b200 import _lib
    lib = _lib.load()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    for path in sorted(glob.glob(os.path.join(GOLDEN, "metrics_*.npz"))):
        z = np.load(path, allow_pickle=False)
        pred, gt = torch.from_numpy(z["pred"]).cuda(), torch.from_numpy(z["gt"]).cuda()
        n, J = pred.shape[0], pred.shape[1]
        jpe = torch.empty((n, J), device="cuda")
        njpe = torch.empty((n, J), device="cuda")
        res = (ctypes.c_double * 3)()
        _lib.check(lib.uu_op_pose_metrics(P(pred), P(gt), n, J, int(z["root"]), P(jpe), P(njpe), res, None))
        assert abs(res[0] - float(z["mpjpe"])) < 1e-6 and abs(res[1] - float(z["nmpjpe"])) < 1e-6
        assert res[2] == float((z["gt"][:, :, 3] > 0).sum())
        assert np.abs(jpe.cpu().numpy() - z["jpe"]).max() < 2e-6
        assert np.abs(njpe.cpu().numpy() - z["njpe"]).max() < 2e-6


def test_camera_projection_on_device_matches_reference_golden():
    """uu_op_world_to_cam_and_2d against tf_world_to_cam_and_2d (tests/golden/projection_*.npz); fp32 on the device:
    camera-space poses to 1e-5, pixel coordinates (|value| up to ~1.6e3) to 2e-3."""
    from uplift_upsample_3dhpe_b200.model import world_to_cam_and_2d
    for path in sorted(glob.glob(os.path.join(GOLDEN, "projection_*.npz"))):
        z = np.load(path, allow_pickle=False)
        c3, p2 = world_to_cam_and_2d(torch.from_numpy(z["seq3d"]).cuda(), torch.from_numpy(z["cams"]).cuda())
        assert np.abs(c3.cpu().numpy() - z["cam3d"]).max() < 1e-5
        assert np.abs(p2.cpu().numpy() - z["p2d"]).max() < 2e-3
