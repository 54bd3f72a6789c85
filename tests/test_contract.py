"""CPU tests of the host-side contract: config keys, spec derivation, stride-mask rule, weight
inventory, and the known answers of SURVEY.md §8c."""
import json
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, forward_macs, spec_from_config, stride_mask, weights
from uplift_upsample_3dhpe_b200.spec import has_strided_input, strided_seq_lens


def test_presets_match_survey():
    s351 = spec_from_config(UpliftUpsampleConfig.preset("h36m_351"))
    s81 = spec_from_config(UpliftUpsampleConfig.preset("h36m_81"))
    assert s351.seq_lens == (71, 23, 3, 1) and s81.seq_lens == (41, 11, 3, 1)
    assert s351.receptive_field == 351 and s81.receptive_field == 81
    assert weights.param_count(s351) == 10_404_902 and weights.param_count(s81) == 10_377_254
    assert forward_macs(s351) == 525_620_672 and forward_macs(s81) == 297_253_184
    # F_alg by s_in (valid frames 71/35/17 and 21/9/5)
    assert forward_macs(s351, 71) - forward_macs(s351, 35) == 36 * 841_024
    assert round(2 * forward_macs(s351, 17) / 1e9, 4) == 0.9604
    assert abs(2 * forward_macs(s81, 21) / 1e9 - 0.5608) < 1e-4


def test_config_json_roundtrip(tmp_path):
    cfg = UpliftUpsampleConfig.preset("h36m_351")
    p = tmp_path / "c.json"
    cfg.dump(str(p))
    cfg2 = UpliftUpsampleConfig(str(p))
    assert cfg2.as_dict() == cfg.as_dict()
    d = json.loads(p.read_text())
    assert d["ROOT_KEYTPOINT"] == 6 and d["SEQUENCE_LENGTH"] == 71 and d["STRIDES"] == [3, 10, 3]
    # txt format: KEY <json>
    t = tmp_path / "c.txt"
    t.write_text("# comment\nSEQUENCE_LENGTH 41\nSTRIDES [4, 4, 3]\nMASK_STRIDE null\n")
    c3 = UpliftUpsampleConfig(str(t))
    assert c3.SEQUENCE_LENGTH == 41 and c3.STRIDES == [4, 4, 3] and c3.MASK_STRIDE is None


def test_unknown_keys_are_kept_and_unsupported_rejected():
    cfg = UpliftUpsampleConfig.preset("h36m_351", SOME_FUTURE_KEY=3)
    assert cfg.as_dict()["SOME_FUTURE_KEY"] == 3
    spec_from_config(UpliftUpsampleConfig.preset("h36m_351", TOKEN_MASK_RATE=0.2))     # masked value 0: supported
    for bad in (dict(OUTPUT_BN=True), dict(DROP_RATE=0.1), dict(TOKEN_MASK_RATE=0.2, LEARNABLE_MASKED_TOKEN=True),
                dict(ATTENTION_DROP_RATE=0.1)):
        with pytest.raises(NotImplementedError):
            spec_from_config(UpliftUpsampleConfig.preset("h36m_351", **bad))


def test_has_strided_input_rule():
    # constructor.py:16-21 — including s_in == s_out
    assert not has_strided_input(None) and not has_strided_input(1) and not has_strided_input([1, 5])
    assert has_strided_input(5) and has_strided_input([5, 10, 20]) and has_strided_input([4])


def test_seq_len_recurrence_default_padding():
    assert strided_seq_lens(27, [3, 3, 3], [(1, 1)] * 3) == [27, 9, 3, 1]       # reference defaults


def test_centred_masks_known_answers():
    m = lambda n, so, si: np.nonzero(stride_mask.stride_mask(n, so, si))[0].tolist()
    assert m(71, 5, 5) == list(range(71))
    assert m(71, 5, 10) == list(range(1, 70, 2))
    assert m(71, 5, 20) == list(range(3, 68, 4))
    assert m(41, 2, 4) == list(range(0, 41, 2))
    assert m(41, 2, 10) == list(range(0, 41, 5))
    assert m(41, 2, 20) == list(range(0, 41, 10))


def test_global_alignment_known_answers():
    cnt = lambda i, si: int(stride_mask.stride_mask(71, 5, si, center_frame=i).sum())
    assert (cnt(0, 5), cnt(0, 20)) == (71, 17)
    assert (cnt(5, 5), cnt(5, 20)) == (71, 18)
    assert (cnt(3, 5), cnt(3, 20), cnt(23, 5), cnt(23, 20)) == (0, 0, 0, 0)


def test_rand_shift_ranges_and_rng_stream():
    rs = stride_mask.rand_shift_range
    rng = lambda ms: (lambda lo, hi, ep: list(range(lo, hi + (1 if ep else 0))))(*rs(ms))
    assert rng(2) == [-1, 0] and rng(4) == [-2, -1, 0, 1] and rng(5) == [-2, -1, 0, 1, 2]
    assert rng(10) == list(range(-5, 5)) and rng(1) == [0]
    assert np.random.default_rng(0).integers(0, 3, size=8).tolist() == [2, 1, 1, 0, 0, 0, 0, 0]
    g = np.random.default_rng(0)
    assert [int(g.integers(-2, 2)) for _ in range(8)] == [1, 0, 0, -1, -1, -2, -2, -2]


@settings(max_examples=200, deadline=None)
@given(n_tok=st.integers(1, 128), s_out=st.integers(1, 10), k=st.integers(1, 8), shift=st.integers(-5000, 5000))
def test_mask_is_python_floor_mod(n_tok, s_out, k, shift):
    s_in = s_out * k
    got = stride_mask.stride_mask(n_tok, s_out, s_in, center_frame=shift)
    want = [((n - n_tok // 2) * s_out + shift) % s_in == 0 for n in range(n_tok)]   # python % is floor-mod
    assert got.tolist() == want


def test_train_masks_mix_strides():
    m = stride_mask.batch_stride_masks_train(71, 5, [5, 10, 20], 64, seed=0)
    counts = set(m.sum(axis=1).tolist())
    assert counts <= {71, 35, 36, 17, 18} and len(counts) >= 3
    assert m.dtype == bool and m.shape == (64, 71)


def test_mask_stride_must_divide():
    with pytest.raises(ValueError):
        stride_mask.stride_mask(71, 5, 4)


def test_inventory_order_and_shapes():
    spec = spec_from_config(UpliftUpsampleConfig.preset("h36m_351"))
    inv = weights.inventory(spec)
    names = list(inv)
    assert names[:8] == ["keypoint_embedding", "token_dropout", "spatial_pe", "temporal_pe", "strided_temporal_pe_1",
                         "strided_temporal_pe_2", "strided_temporal_pe_3", "strided_input_token_layer"]
    assert names[-2:] == ["temporal_fc", "strided_temporal_fc"]
    assert [s for _, s, _ in inv["strided_temporal_block_1"]][12:] == [(1, 384, 768), (768,), (3, 768, 384), (384,)]
    assert [s for _, s, _ in inv["strided_temporal_pe_2"]] == [(23, 384)]
    assert len(inv["spatial_block_1"]) == 16 and inv["token_dropout"] == []
    w = weights.init_weights(spec, 1)
    flat = weights.to_flat(spec, w)
    assert flat.size == 10_404_902
    back = weights.from_flat(spec, flat)
    assert all(np.array_equal(back[k], w[k]) for k in w)
    tok = w[("strided_input_token_layer", 0)]
    assert np.abs(tok).max() <= 0.04 + 1e-9                         # truncated normal, 2 sigma
    k = w[("spatial_to_temporal_fc", 0)]
    assert np.abs(k).max() <= math.sqrt(6 / (544 + 384)) + 1e-7     # glorot limit


def test_no_mask_stride_drops_token_tensor():
    spec = spec_from_config(UpliftUpsampleConfig.preset("h36m_351", MASK_STRIDE=None))
    assert not spec.has_strided_input and "strided_input_token_layer" not in weights.inventory(spec)
    assert weights.param_count(spec) == 10_404_902 - 384


def test_host_side_schedules():
    """common/utils/schedules.py: staircase ExponentialDecay (every shipped config) and ExponentialDecayWithSteps
    (:36-99: the small staircase skips the steps where the large one fires)."""
    torch = pytest.importorskip("torch")  # noqa: F841  (train.py imports lazily, the schedules are pure Python)
    from uplift_upsample_3dhpe_b200.train import scheduler_by_name
    s = scheduler_by_name("ExponentialDecay")(initial_learning_rate=4e-5, decay_steps=6000, decay_rate=0.99, staircase=True)
    assert s(0) == 4e-5 and s(5999) == 4e-5 and abs(s(6000) - 4e-5 * 0.99) < 1e-18 and abs(s(12001) - 4e-5 * 0.99 ** 2) < 1e-18
    w = scheduler_by_name("ExponentialDecayWithSteps")(initial_learning_rate=1e-3, decay_steps=100, decay_rate=0.9,
                                                       large_decay_steps=1000, large_decay_rate=0.5)
    assert w(0) == 1e-3 and w(99) == 1e-3
    assert abs(w(100) - 1e-3 * 0.9) < 1e-15 and abs(w(950) - 1e-3 * 0.9 ** 9) < 1e-15
    # step 1000: floor(1000/100) - floor(1000/1000) = 9 small decays and one large one
    assert abs(w(1000) - 1e-3 * 0.9 ** 9 * 0.5) < 1e-15
    assert abs(w(2345) - 1e-3 * 0.9 ** (23 - 2) * 0.5 ** 2) < 1e-15
    with pytest.raises(NotImplementedError):
        scheduler_by_name("CosineDecayRestarts")
