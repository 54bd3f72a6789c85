"""The oracle and the host logic against golden vectors produced by the REFERENCE's own code.

tests/golden/*.npz were written by scripts/make_golden.py, which imports the reference's unmodified
modules (constructor, network, weight_io loader, sequence generator) over the NumPy TensorFlow
stand-in in oracle/tfshim and runs them on seeded inputs.  These tests need neither the reference
nor a GPU: they pin oracle/forward_np.py, the weight inventory order, the .h5 writer and the
stride-mask rule to what the reference computes."""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import forward_np as O
from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, spec_from_config, stride_mask, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FORWARD = sorted(glob.glob(os.path.join(GOLDEN, "forward_*.npz")))
WINDOWS = sorted(glob.glob(os.path.join(GOLDEN, "windows_*.npz")))


def load_forward_case(path):
    z = np.load(path, allow_pickle=False)
    ms = z["mask_stride"]
    ms = int(ms) if ms.ndim == 0 else [int(v) for v in ms]
    cfg = UpliftUpsampleConfig.preset(str(z["config"]), MASK_STRIDE=ms)
    spec = spec_from_config(cfg)
    w = weights.init_weights(spec, seed=int(z["seed"]), perturb=True)
    sha = hashlib.sha256(np.ascontiguousarray(weights.to_flat(spec, w)).tobytes()).hexdigest()[:16]
    assert sha == str(z["weights_sha"]), "weights.init_weights changed: regenerate tests/golden (scripts/make_golden.py)"
    return cfg, spec, w, z


def test_fixtures_present():
    assert len(FORWARD) >= 5 and len(WINDOWS) >= 4


@pytest.mark.parametrize("path", FORWARD, ids=[os.path.basename(p)[8:-4] for p in FORWARD])
def test_oracle_matches_reference_forward(path):
    cfg, spec, w, z = load_forward_case(path)
    x, m = z["x"], z["mask"]
    valid = m.sum(axis=1) > 0
    # float64: exact restatement => agreement to rounding noise on every window that has a valid token
    full, central = O.test_step(spec, w, x, m, dtype=np.float64)
    assert np.abs(central[valid] - z["central"][valid]).max() < 1e-9
    assert np.abs(full[valid] - z["full"][valid]).max() < 1e-9
    # float32: the precision TensorFlow computes in; it defines the all-masked windows (x - 1e9 rounds to -1e9)
    full32, central32 = O.test_step(spec, w, x, m, dtype=np.float32)
    assert full32.dtype == np.float32
    assert np.abs(central32 - z["central_f32"]).max() < 2e-4
    assert np.abs(full32 - z["full_f32"]).max() < 2e-4
    if (~valid).any():     # the two precisions really disagree there, so the fp32 definition is load-bearing
        assert np.abs(z["central_f32"][~valid] - z["central"][~valid]).max() > 1e-2


@pytest.mark.parametrize("path", WINDOWS, ids=[os.path.basename(p)[8:-4] for p in WINDOWS])
def test_stride_masks_are_bit_exact(path):
    z = np.load(path, allow_pickle=False)
    n_tok, stride, mode = int(z["n_tok"]), int(z["stride"]), str(z["mode"])
    want = z["stride_masks"]
    if mode == "eval":
        got = stride_mask.batch_stride_masks_eval(n_tok, stride, int(z["mask_stride"]), z["centers"])
    else:
        got = stride_mask.batch_stride_masks_train(n_tok, stride, [int(v) for v in z["mask_stride"]], want.shape[0], seed=0)
    assert got.dtype == np.bool_ and np.array_equal(got, want)
    from uplift_upsample_3dhpe_b200 import _lib     # the C-ABI rule must agree too (host-only entry point)
    import ctypes
    lib = _lib.load()
    if mode == "eval":
        buf = (ctypes.c_uint8 * n_tok)()
        for c, row in zip(z["centers"], want):
            _lib.check(lib.uu_stride_mask(n_tok, stride, int(z["mask_stride"]), int(c), buf))
            assert np.array_equal(np.frombuffer(buf, dtype=np.uint8).astype(bool), row)


@pytest.mark.parametrize("path", [p for p in WINDOWS if "eval" in p], ids=lambda p: os.path.basename(p)[8:-4])
def test_window_source_frames_match_reference_generator(path):
    """Host twin of the device window gather: the reference generator's windows (edge-padded, strided) bit for bit."""
    z = np.load(path, allow_pickle=False)
    n_tok, stride = int(z["n_tok"]), int(z["stride"])
    video = z["video_2d"]
    src = stride_mask.window_source_frames(n_tok, stride, video.shape[0], z["centers"], pad_copy=True)
    assert src.min() >= 0 and src.max() < video.shape[0]
    assert np.array_equal(video[src], z["seq_2d"])
    # padded positions are exactly those the generator's pad mask marks with 0
    f = (np.arange(n_tok)[None, :] - n_tok // 2) * stride + z["centers"][:, None]
    assert np.array_equal(((f >= 0) & (f < video.shape[0])).astype(np.float32), z["pad_masks"])


def test_oracle_tta_and_interpolation_match_reference():
    from oracle import eval_np
    tta = sorted(glob.glob(os.path.join(GOLDEN, "tta_*.npz")))
    interp = sorted(glob.glob(os.path.join(GOLDEN, "interp_*.npz")))
    assert tta and interp
    for path in tta:
        cfg, spec, w, z = load_forward_case(path)
        full, central = eval_np.flip_tta(spec, w, z["x"], z["mask"], cfg.AUGM_FLIP_KEYPOINT_ORDER)
        assert np.abs(central - z["central"]).max() < 1e-9 and np.abs(full - z["full"]).max() < 1e-9
    for path in interp:
        z = np.load(path)
        out = eval_np.interpolate_between_keyframes(z["pred"], z["frame_indices"], int(z["stride"]))
        assert np.array_equal(out, z["out"])           # same float64 operations in the same order


def test_random_token_masking_matches_reference_training_forward():
    """D2 (net:287-311, :336-338): the reference model run with training=True, TOKEN_MASK_RATE = 0.3 and the recorded
    uniform draw; the oracle's token_keep path must reproduce it (this is what the CUDA training step is tested against
    in tests/test_gpu_train.py)."""
    torch = pytest.importorskip("torch")
    from oracle import forward_torch as OT
    paths = sorted(glob.glob(os.path.join(GOLDEN, "tokenmask_*.npz")))
    assert paths
    for path in paths:
        z = np.load(path, allow_pickle=False)
        rate = float(z["rate"])
        cfg = UpliftUpsampleConfig.preset(str(z["config"]), MASK_STRIDE=int(z["mask_stride"]), TOKEN_MASK_RATE=rate)
        spec = spec_from_config(cfg)
        w = weights.init_weights(spec, seed=int(z["seed"]), perturb=True)
        u, m = z["uniform"], z["mask"]
        token_mask = (u < rate) & (np.arange(spec.n_tok) != spec.n_tok // 2)[None, :]      # net:291-303
        assert token_mask.any() and not token_mask[:, spec.n_tok // 2].any()
        keep = 1.0 - token_mask.astype(np.float64)
        x = torch.tensor(z["x"], dtype=torch.float64) * torch.tensor(m, dtype=torch.float64)[:, :, None, None]
        full, central = OT.forward(spec, OT.to_torch(w, torch.float64), x, torch.tensor(m), token_keep=torch.tensor(keep))
        valid = m.sum(axis=1) > 0
        assert np.abs(central.numpy()[valid] - z["central"][valid]).max() < 1e-9
        assert np.abs(full.numpy()[valid] - z["full"][valid]).max() < 1e-9
        # and the mask matters: without it the outputs differ
        _, c0 = OT.forward(spec, OT.to_torch(w, torch.float64), x, torch.tensor(m))
        assert np.abs(c0.numpy()[valid] - z["central"][valid]).max() > 1e-4


def test_pose_metrics_match_reference():
    """MPJPE / N-MPJPE restatement (oracle/eval_np.py) against common/dataset/metrics.py run on the same poses."""
    from oracle import eval_np as E
    paths = sorted(glob.glob(os.path.join(GOLDEN, "metrics_*.npz")))
    assert len(paths) >= 2
    for path in paths:
        z = np.load(path, allow_pickle=False)
        root = int(z["root"])
        assert abs(E.mpjpe(z["pred"], z["gt"], root) - float(z["mpjpe"])) < 1e-12
        assert abs(E.nmpjpe(z["pred"], z["gt"], root) - float(z["nmpjpe"])) < 1e-12
        assert np.abs(E.mpjpe(z["pred"], z["gt"], root, normalize=False) - z["jpe"]).max() < 1e-12
        assert np.abs(E.nmpjpe(z["pred"], z["gt"], root, normalize=False) - z["njpe"]).max() < 1e-12


def test_camera_projection_matches_reference():
    """world -> camera -> 2-D restatement against uplifiting_dataset.tf_world_to_cam_and_2d (run over the shim)."""
    from oracle import eval_np as E
    paths = sorted(glob.glob(os.path.join(GOLDEN, "projection_*.npz")))
    assert paths
    for path in paths:
        z = np.load(path, allow_pickle=False)
        for b in range(z["seq3d"].shape[0]):
            c3, p2 = E.world_to_cam_and_2d(z["seq3d"][b], z["cams"][b])
            assert np.abs(c3 - z["cam3d"][b]).max() < 1e-12 and np.abs(p2 - z["p2d"][b]).max() < 1e-9
