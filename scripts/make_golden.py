"""Generate tests/golden/*.npz by executing the REFERENCE's own Python sources.

Runs only in the build container (needs /root/reference); the fixtures it writes are committed and
are what the GPU box tests against.  The reference modules are imported unmodified; TensorFlow,
Keras and h5py (absent from this image) are replaced by the NumPy stand-ins under oracle/tfshim.

Executed from the reference, not restated:
  * config loading            common/utils/config.py + config/*.json
  * model construction        common/net/uplift_upsample_transformer_constructor.py:14-50
  * the forward graph         common/net/uplift_upsample_transformer.py, common/net/vision_transformer.py
  * .h5 weight loading        common/utils/weight_io.py:76-263  (reads files written by our h5lite writer)
  * eval-time glue            eval.py:63-71 (mask multiply) — restated in three lines below, eval.py imports datasets at top level
  * window + stride-mask generator  common/dataset/uplifiting_dataset.py:213-428

usage: python scripts/make_golden.py [--out tests/golden]
"""
import argparse
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("UU_REFERENCE_ROOT", "/root/reference")

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
args = ap.parse_args()

sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "tfshim"))
sys.path.insert(0, REF)

import numpy as np  # noqa: E402

import tensorflow as tf  # noqa: E402  (the shim)
from common.net.uplift_upsample_transformer_config import UpliftUpsampleConfig as RefConfig  # noqa: E402
from common.net.uplift_upsample_transformer_constructor import build_uplift_upsample_transformer as ref_build  # noqa: E402
from common.utils import weight_io as ref_weight_io  # noqa: E402
from common.dataset.uplifiting_dataset import H36mSequenceGenerator  # noqa: E402

from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, h5lite, spec_from_config, weights  # noqa: E402

assert tf.__version__.endswith("numpy-shim")
os.makedirs(args.out, exist_ok=True)
TMP = os.path.join("/tmp", "uu_golden")
os.makedirs(TMP, exist_ok=True)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def ref_test_step(model, keypoints2d, stride_masks):
    """eval.py:63-71."""
    if model.has_strided_input:
        mask = tf.cast(stride_masks, dtype=tf.float32)
        inputs = [keypoints2d * mask[:, :, tf.newaxis, tf.newaxis], stride_masks]
    else:
        inputs = keypoints2d
    return model(inputs, training=False)


def forward_case(tag, cfg_name, mask_stride, B, mode, seed, subsample=1):
    ref_cfg = RefConfig(config_file=os.path.join(REF, "config", cfg_name + ".json"))
    ref_cfg.MASK_STRIDE = mask_stride            # eval.py:264-266 sets an int per run
    ref_cfg.BATCH_SIZE = B

    ours = UpliftUpsampleConfig.preset(cfg_name, MASK_STRIDE=mask_stride)
    spec = spec_from_config(ours)
    # every key the hot path consumes must agree between our preset and the reference's JSON
    for k in ("SEQUENCE_LENGTH", "SEQUENCE_STRIDE", "NUM_KEYPOINTS", "SPATIAL_EMBED_DIM", "TEMPORAL_EMBED_DIM",
              "SPATIAL_TRANSFORMER_BLOCKS", "TEMPORAL_TRANSFORMER_BLOCKS", "STRIDES", "PADDINGS", "NUM_HEADS", "MLP_RATIO",
              "QKV_BIAS", "FIRST_STRIDED_TOKEN_ATTENTION_LAYER", "USE_REFINE"):
        assert getattr(ref_cfg, k) == getattr(ours, k), (k, getattr(ref_cfg, k), getattr(ours, k))

    w = weights.init_weights(spec, seed=seed, perturb=True)
    h5 = os.path.join(TMP, f"{tag}.h5")
    h5lite.save_keras_weights(h5, spec, w)
    rng = np.random.default_rng(seed + 100)
    n_tok = ref_cfg.SEQUENCE_LENGTH
    x = rng.uniform(-1, 1, (B, n_tok, 17, 2)).astype(np.float32)
    gen = mask_generator(n_tok, ref_cfg.SEQUENCE_STRIDE, mask_stride, mode, B, subsample)
    masks = np.stack([m for _, m in gen])
    centers = np.array([c for c, _ in gen])
    out = {}
    # float64 = truth; float32 = the precision TensorFlow computes in (it defines all-masked windows: x - 1e9 rounds to -1e9)
    for dt in ("float64", "float32"):
        tf.set_float_dtype(dt)
        model = ref_build(ref_cfg)
        assert model.has_strided_input == bool(spec.has_strided_input)
        ref_weight_io.load_weights_with_callback(model, h5, verbose=True)      # the reference's own loader
        assert model.count_params() == weights.param_count(spec), (model.count_params(), weights.param_count(spec))
        # the loader matched every tensor: values in the reference model equal the inventory, by group and position
        by_name = {l.name: l for l in model.layers}
        for (g, i), a in w.items():
            got = np.asarray(by_name[g].weights[i])
            assert got.shape == a.shape and np.array_equal(got.astype(np.float32), a), (g, i)
        full, central = ref_test_step(model, x.astype(tf.float32), masks)
        assert np.asarray(central).dtype == np.dtype(dt)
        out[dt] = (np.asarray(full), np.asarray(central))
    np.savez_compressed(os.path.join(args.out, f"forward_{tag}.npz"), config=cfg_name, mask_stride=mask_stride, seed=seed,
                        x=x, mask=masks, center_frames=centers, full=out["float64"][0], central=out["float64"][1],
                        full_f32=out["float32"][0], central_f32=out["float32"][1],
                        weights_sha=sha(weights.to_flat(spec, w)))
    d = np.abs(out["float64"][1] - out["float32"][1]).reshape(B, -1).max(1)
    print(f"forward_{tag}: full {out['float64'][0].shape} central {out['float64'][1].shape} "
          f"valid/window {masks.sum(1).tolist()} |central|max {np.abs(out['float64'][1]).max():.3f} "
          f"f32-vs-f64 per window {np.array2string(d, precision=2)}")


def tta_case(tag, cfg_name, mask_stride, B, seed):
    """Flip test-time augmentation, eval.py:152-180 (restated line by line on the reference model object)."""
    ref_cfg = RefConfig(config_file=os.path.join(REF, "config", cfg_name + ".json"))
    ref_cfg.MASK_STRIDE = mask_stride
    ref_cfg.BATCH_SIZE = B
    ours = UpliftUpsampleConfig.preset(cfg_name, MASK_STRIDE=mask_stride)
    assert list(ref_cfg.AUGM_FLIP_KEYPOINT_ORDER) == list(ours.AUGM_FLIP_KEYPOINT_ORDER)
    spec = spec_from_config(ours)
    w = weights.init_weights(spec, seed=seed, perturb=True)
    h5 = os.path.join(TMP, f"{tag}.h5")
    h5lite.save_keras_weights(h5, spec, w)
    rng = np.random.default_rng(seed + 100)
    x = rng.uniform(-1, 1, (B, ref_cfg.SEQUENCE_LENGTH, 17, 2)).astype(np.float32)
    gen = mask_generator(ref_cfg.SEQUENCE_LENGTH, ref_cfg.SEQUENCE_STRIDE, mask_stride, "eval", B, 5)
    masks = np.stack([m for _, m in gen])
    tf.set_float_dtype("float64")
    model = ref_build(ref_cfg)
    ref_weight_io.load_weights_with_callback(model, h5, verbose=False)
    order = ref_cfg.AUGM_FLIP_KEYPOINT_ORDER
    x64 = x.astype(tf.float32)
    seq, cen = ref_test_step(model, x64, masks)                                                    # :152-153
    fx = np.concatenate([x64[:, :, :, :1] * -1., x64[:, :, :, 1:]], axis=-1)                       # :155-158
    fx = np.take(fx, order, axis=2)                                                                # :159
    fseq, fcen = ref_test_step(model, fx, masks)                                                   # :161-162
    fcen = np.take(np.concatenate([fcen[:, :, :1] * -1., fcen[:, :, 1:]], axis=-1), order, axis=1)  # :164-167
    cen = (cen + fcen) / 2.                                                                        # :169-170
    fseq = np.take(np.concatenate([fseq[:, :, :, :1] * -1., fseq[:, :, :, 1:]], axis=-1), order, axis=2)  # :172-178
    seq = (seq + fseq) / 2.                                                                        # :180-181
    np.savez_compressed(os.path.join(args.out, f"tta_{tag}.npz"), config=cfg_name, mask_stride=mask_stride, seed=seed,
                        x=x, mask=masks, full=seq, central=cen, weights_sha=sha(weights.to_flat(spec, w)))
    print(f"tta_{tag}: central {cen.shape} |max| {np.abs(cen).max():.3f}")


def token_mask_case(tag, cfg_name, mask_stride, B, seed, rate):
    """Training-mode forward with random_token_masking (net:287-311, :336-338): TOKEN_MASK_RATE > 0, masked value 0,
    DropPath off.  The uniform draw the reference consumed is recorded so that the oracle can rebuild the mask."""
    ref_cfg = RefConfig(config_file=os.path.join(REF, "config", cfg_name + ".json"))
    ref_cfg.MASK_STRIDE = mask_stride
    ref_cfg.BATCH_SIZE = B
    ref_cfg.TOKEN_MASK_RATE = rate
    ref_cfg.LEARNABLE_MASKED_TOKEN = False
    ref_cfg.DROP_PATH_RATE = [0.0, 0.0, 0.0]
    ours = UpliftUpsampleConfig.preset(cfg_name, MASK_STRIDE=mask_stride, TOKEN_MASK_RATE=rate)
    spec = spec_from_config(ours)
    w = weights.init_weights(spec, seed=seed, perturb=True)
    h5 = os.path.join(TMP, f"{tag}.h5")
    h5lite.save_keras_weights(h5, spec, w)
    rng = np.random.default_rng(seed + 100)
    n_tok = ref_cfg.SEQUENCE_LENGTH
    x = rng.uniform(-1, 1, (B, n_tok, 17, 2)).astype(np.float32)
    gen = mask_generator(n_tok, ref_cfg.SEQUENCE_STRIDE, mask_stride, "eval", B, 5)
    masks = np.stack([m for _, m in gen])
    tf.set_float_dtype("float64")
    model = ref_build(ref_cfg)
    ref_weight_io.load_weights_with_callback(model, h5, verbose=False)
    draws = []
    orig = tf.random.uniform

    def recording_uniform(*a, **k):
        u = orig(*a, **k)
        draws.append(np.asarray(u).copy())
        return u

    tf.random.set_seed(seed)
    tf.random.uniform = recording_uniform
    try:
        mask = tf.cast(masks, dtype=tf.float32)                                   # train.py:474-478
        full, central = model([x.astype(tf.float32) * mask[:, :, tf.newaxis, tf.newaxis], masks], training=True)
    finally:
        tf.random.uniform = orig
    assert len(draws) == 1 and draws[0].shape == (B, n_tok), [d.shape for d in draws]
    np.savez_compressed(os.path.join(args.out, f"tokenmask_{tag}.npz"), config=cfg_name, mask_stride=mask_stride, seed=seed,
                        rate=rate, x=x, mask=masks, uniform=draws[0], full=np.asarray(full), central=np.asarray(central),
                        weights_sha=sha(weights.to_flat(spec, w)))
    print(f"tokenmask_{tag}: masked tokens/window {(draws[0] < rate).sum(1).tolist()} |central|max {np.abs(np.asarray(central)).max():.3f}")


def interp_case(tag, stride, lens, seed):
    """Key-frame interpolation by the reference's own numpy function (common/dataset/action_wise_eval.py:76-100)."""
    from common.dataset.action_wise_eval import interpolate_between_keyframes
    rng = np.random.default_rng(seed)
    frame_indices = np.concatenate([np.arange(n) for n in lens])          # videos concatenated, index restarts at 0
    pred = rng.normal(size=(frame_indices.size, 17, 3)).astype(np.float32)
    out, keyframes = interpolate_between_keyframes(pred3d=pred.astype(np.float64), frame_indices=frame_indices,
                                                   keyframe_stride=np.tile([stride], reps=frame_indices.size))
    np.savez_compressed(os.path.join(args.out, f"interp_{tag}.npz"), stride=stride, frame_indices=frame_indices.astype(np.int32),
                        pred=pred, out=out, keyframes=keyframes)
    print(f"interp_{tag}: {frame_indices.size} frames, {int(keyframes.sum())} key frames")


def metrics_case(tag, n, seed, root=0):
    """MPJPE / N-MPJPE by the reference's own numpy functions (common/dataset/metrics.py:13-81)."""
    from common.dataset import metrics as ref_metrics
    rng = np.random.default_rng(seed)
    gt = rng.normal(0, 0.4, (n, 17, 4)).astype(np.float32)
    gt[:, :, 3] = (rng.random((n, 17)) > 0.15).astype(np.float32)         # some invalid joints
    gt[:, root, 3] = 1.0
    pred = (gt[:, :, :3] * rng.uniform(0.7, 1.3, (n, 1, 1)) + rng.normal(0, 0.05, (n, 17, 3)) + rng.normal(0, 1, (n, 1, 3))).astype(np.float32)
    p64, g64 = pred.astype(np.float64), gt.astype(np.float64)
    np.savez_compressed(os.path.join(args.out, f"metrics_{tag}.npz"), pred=pred, gt=gt, root=root,
                        mpjpe=ref_metrics.mpjpe(p64, g64, root_index=root),
                        nmpjpe=ref_metrics.nmpjpe(p64, g64, root_index=root),
                        jpe=ref_metrics.mpjpe(p64, g64, root_index=root, normalize=False),
                        njpe=ref_metrics.nmpjpe(p64, g64, root_index=root, normalize=False))
    print(f"metrics_{tag}: mpjpe {ref_metrics.mpjpe(p64, g64, root_index=root):.5f} nmpjpe {ref_metrics.nmpjpe(p64, g64, root_index=root):.5f}")


def projection_case(tag, B, n_frames, seed):
    """AMASS-style virtual-camera projection (uplifiting_dataset.py:669-761): world -> camera by quaternion, then the
    Human3.6M projection with radial / tangential distortion, by the reference's own tf_world_to_cam_and_2d."""
    from common.dataset.uplifiting_dataset import tf_world_to_cam_and_2d
    tf.set_float_dtype("float64")
    rng = np.random.default_rng(seed)
    seq = (rng.normal(0, 0.5, (B, n_frames, 17, 3)) + np.array([0.0, 0.0, 1.0])).astype(np.float32)
    q = rng.normal(size=(B, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    trans = rng.normal(0, 1.0, (B, 3)) + np.array([0.0, -4.0, 1.5])
    intr = np.concatenate([np.tile([1000.0, 1002.0], (B, 1)), rng.uniform(1100, 1200, (B, 2)), rng.uniform(480, 540, (B, 2)),
                           rng.normal(0, 0.1, (B, 3)), rng.normal(0, 0.01, (B, 2))], axis=1)
    cams = np.concatenate([q, trans, intr], axis=1).astype(np.float32)                  # (B, 4 + 3 + 11)
    cam3d, p2d = [], []
    for b in range(B):
        out = tf_world_to_cam_and_2d(seq[b].astype(np.float64), cams[b].astype(np.float64), None, None, None, None, None)
        cam3d.append(np.asarray(out[0])); p2d.append(np.asarray(out[1]))
    np.savez_compressed(os.path.join(args.out, f"projection_{tag}.npz"), seq3d=seq, cams=cams, cam3d=np.stack(cam3d),
                        p2d=np.stack(p2d))
    print(f"projection_{tag}: 2d range {np.stack(p2d).min():.1f} .. {np.stack(p2d).max():.1f}")


def run_generator(n_tok, stride, mask_stride, mode, n_windows, video_len=400, seed=0, subsample=1):
    """Drive the reference's H36mSequenceGenerator on one synthetic video."""
    rng = np.random.default_rng(7)
    p3 = rng.normal(size=(video_len, 17, 3)).astype(np.float32)
    p2 = rng.normal(size=(video_len, 17, 2)).astype(np.float32)
    cam = np.zeros(11, dtype=np.float32)
    g = H36mSequenceGenerator([p3], [p2], [cam], ["S1"], ["Walking"], [50], split="test", seq_len=n_tok,
                              subsample=subsample, stride=stride, padding_type="copy", flip_augment=False,
                              mask_stride=mask_stride, stride_mask_align_global=(mode == "eval"),
                              rand_shift_stride_mask=(mode == "train"), shuffle=False, seed=seed, verbose=False)
    out = []
    for k, item in enumerate(g.next_epoch_iterator()):
        if k >= n_windows:
            break
        out.append(item)
    return p2, out


def mask_generator(n_tok, stride, mask_stride, mode, B, subsample):
    _, items = run_generator(n_tok, stride, mask_stride, mode, B, subsample=subsample)
    return [(int(it[6]), np.asarray(it[7], dtype=bool)) for it in items]


def stride_mask_case(tag, n_tok, stride, mask_stride, mode, n_windows, video_len):
    p2, items = run_generator(n_tok, stride, mask_stride, mode, n_windows, video_len=video_len)
    np.savez_compressed(os.path.join(args.out, f"windows_{tag}.npz"), n_tok=n_tok, stride=stride,
                        mask_stride=np.asarray(mask_stride), mode=mode, video_2d=p2,
                        centers=np.array([it[6] for it in items]),
                        stride_masks=np.stack([np.asarray(it[7], dtype=bool) for it in items]),
                        pad_masks=np.stack([it[2] for it in items]),
                        seq_2d=np.stack([it[1] for it in items]).astype(np.float32))
    print(f"windows_{tag}: {len(items)} windows, valid tokens min/max "
          f"{min(int(it[7].sum()) for it in items)}/{max(int(it[7].sum()) for it in items)}")


if __name__ == "__main__":
    # forward: the BASELINE.json configs at fixture size (the full-size runs use properties instead)
    # (window centres step by `subsample` frames; centres off the s_out grid give all-masked windows, eval.py global alignment)
    forward_case("h36m_81_sin4", "h36m_81", 4, 4, "eval", seed=1, subsample=3)            # valid tokens 21, 0, 20, 0
    forward_case("h36m_351_sin5", "h36m_351", 5, 3, "eval", seed=2, subsample=5)          # mask all ones (no-op)
    forward_case("h36m_351_sin5_unaligned", "h36m_351", 5, 3, "eval", seed=2, subsample=3)  # 71, 0, 0
    forward_case("h36m_351_sin20", "h36m_351", 20, 5, "eval", seed=3, subsample=5)        # 17/18 valid, every alignment
    forward_case("amass_351_train_masks", "amass_351", [5, 10, 20], 6, "train", seed=4)
    tta_case("h36m_351_sin10", "h36m_351", 10, 3, seed=6)
    token_mask_case("h36m_81_sin4_rate03", "h36m_81", 4, 3, seed=8, rate=0.3)
    metrics_case("n300_root0", 300, seed=9, root=0)
    metrics_case("n77_root6", 77, seed=10, root=6)
    projection_case("b6_n9", 6, 9, seed=12)
    interp_case("stride5", 5, [23, 41, 5, 1], seed=8)
    interp_case("stride2", 2, [9, 12], seed=9)
    # window + stride-mask generator (bit-exact contract)
    stride_mask_case("eval_351_sin20", 71, 5, 20, "eval", 120, 150)
    stride_mask_case("eval_81_sin10", 41, 2, 10, "eval", 60, 90)
    stride_mask_case("train_351_mixed", 71, 5, [5, 10, 20], "train", 64, 400)
    stride_mask_case("train_81_mixed", 41, 2, [4, 10, 20], "train", 64, 200)
