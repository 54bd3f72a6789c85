"""ModelSpec: everything the kernels need to know, derived from a config.

Follows the constructor of the reference
(common/net/uplift_upsample_transformer_constructor.py:14-50) and the shape
logic of UpliftUpsampleTransformer.__init__
(common/net/uplift_upsample_transformer.py:165-285).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Tuple


def has_strided_input(mask_stride) -> bool:
    """constructor.py:16-21 — True whenever MASK_STRIDE is set and is not 1 / [1, ...]."""
    if mask_stride is None:
        return False
    if type(mask_stride) is int and mask_stride == 1:
        return False
    if type(mask_stride) is list and mask_stride[0] == 1:
        return False
    return True


@dataclass(frozen=True)
class ModelSpec:
    n_tok: int                 # SEQUENCE_LENGTH: frame tokens the network sees (71 for "N=351")
    n_joints: int              # NUM_KEYPOINTS
    d_spatial: int             # SPATIAL_EMBED_DIM
    d_temporal: int            # TEMPORAL_EMBED_DIM
    spatial_depth: int
    temporal_depth: int
    strides: Tuple[int, ...]
    paddings: Tuple[Tuple[int, int], ...]
    num_heads: int
    mlp_ratio: float
    has_strided_input: bool
    first_strided_token_attention_layer: int
    full_output: bool
    seq_lens: Tuple[int, ...] = field(default=())   # n_tok, then after every strided block
    # training-side constants (train.py:464-506)
    batch_size: int = 256
    sequence_stride: int = 1
    root_keypoint: int = 6
    loss_weight_center: float = 1.0
    loss_weight_sequence: float = 1.0
    drop_path_rate: Tuple[float, float, float] = (0.0, 0.0, 0.0)

    @property
    def h_spatial(self) -> int:
        return int(self.d_spatial * self.mlp_ratio)

    @property
    def h_temporal(self) -> int:
        return int(self.d_temporal * self.mlp_ratio)

    @property
    def out_dim(self) -> int:
        return 3 * self.n_joints

    @property
    def receptive_field(self) -> int:
        return (self.n_tok - 1) * self.sequence_stride + 1


def strided_seq_lens(n_tok: int, strides, paddings) -> List[int]:
    """Sequence-length recurrence of the strided blocks (net:209-216):
    L' = ceil((L + p0 + p1 - 2) / s)  (kernel 3, 'valid' conv after explicit zero padding)."""
    out = [n_tok]
    L = n_tok
    for s, p in zip(strides, paddings):
        L = math.ceil((L + p[0] + p[1] - 2) / s)
        out.append(L)
    return out


def spec_from_config(cfg) -> ModelSpec:
    """Validate a config for the hot path and derive the ModelSpec.

    Features the BASELINE configs never enable are rejected loudly rather than
    silently ignored."""
    if cfg.OUTPUT_BN:
        raise NotImplementedError("OUTPUT_BN=true is not on the hot path (no shipped config enables it)")
    if cfg.DROP_RATE != 0.0 or cfg.ATTENTION_DROP_RATE != 0.0:
        raise NotImplementedError("DROP_RATE / ATTENTION_DROP_RATE > 0 are not supported (0.0 in every shipped config)")
    if not 0.0 <= float(cfg.TOKEN_MASK_RATE) < 1.0:
        raise ValueError("TOKEN_MASK_RATE must be in [0, 1)")
    if cfg.TOKEN_MASK_RATE != 0.0 and cfg.LEARNABLE_MASKED_TOKEN:
        # the learnable variant adds a weight (net:219-220) that no shipped checkpoint contains
        raise NotImplementedError("TOKEN_MASK_RATE > 0 with LEARNABLE_MASKED_TOKEN=true is not supported "
                                  "(masked value 0 is; 0.0 in every shipped config)")
    if cfg.SPATIAL_TRANSFORMER_BLOCKS <= 0 or cfg.TEMPORAL_TRANSFORMER_BLOCKS <= 0:
        raise NotImplementedError("spatial/temporal depth 0 variants are not supported")
    if not cfg.QKV_BIAS:
        raise NotImplementedError("QKV_BIAS=false is not supported (true in every shipped config)")
    strides = tuple(int(s) for s in cfg.STRIDES)
    if len(strides) == 0:
        raise NotImplementedError("a model without strided blocks is not supported")
    if cfg.PADDINGS is None:
        paddings = tuple((1, 1) for _ in strides)          # net:212
    else:
        paddings = tuple((int(p[0]), int(p[1])) for p in cfg.PADDINGS)
    if len(paddings) != len(strides):
        raise ValueError("PADDINGS and STRIDES must have the same length")
    if cfg.SPATIAL_EMBED_DIM % cfg.NUM_HEADS or cfg.TEMPORAL_EMBED_DIM % cfg.NUM_HEADS:
        raise ValueError("embedding widths must be divisible by NUM_HEADS (vit:79)")
    lens = strided_seq_lens(int(cfg.SEQUENCE_LENGTH), strides, paddings)
    if lens[-1] != 1:
        raise ValueError(f"strided blocks must reduce the sequence to one token, got {lens}")
    dpr = cfg.DROP_PATH_RATE
    if not isinstance(dpr, (list, tuple)):
        dpr = [dpr, dpr, dpr]
    return ModelSpec(
        n_tok=int(cfg.SEQUENCE_LENGTH), n_joints=int(cfg.NUM_KEYPOINTS),
        d_spatial=int(cfg.SPATIAL_EMBED_DIM), d_temporal=int(cfg.TEMPORAL_EMBED_DIM),
        spatial_depth=int(cfg.SPATIAL_TRANSFORMER_BLOCKS), temporal_depth=int(cfg.TEMPORAL_TRANSFORMER_BLOCKS),
        strides=strides, paddings=paddings, num_heads=int(cfg.NUM_HEADS), mlp_ratio=cfg.MLP_RATIO,
        has_strided_input=has_strided_input(cfg.MASK_STRIDE),
        first_strided_token_attention_layer=int(cfg.FIRST_STRIDED_TOKEN_ATTENTION_LAYER),
        full_output=not cfg.USE_REFINE, seq_lens=tuple(lens),
        batch_size=int(cfg.BATCH_SIZE), sequence_stride=int(cfg.SEQUENCE_STRIDE),
        root_keypoint=int(cfg.ROOT_KEYTPOINT),
        loss_weight_center=float(cfg.LOSS_WEIGHT_CENTER), loss_weight_sequence=float(cfg.LOSS_WEIGHT_SEQUENCE),
        drop_path_rate=tuple(float(x) for x in dpr),
    )


def forward_macs(spec: ModelSpec, valid_frames: int | None = None) -> int:
    """Algorithmic multiply-accumulates of one window: every
    GEMM/attention/conv of the reference at full density, spatial stages only
    for frames that carry 2-D input."""
    N, J, ds, dt, H = spec.n_tok, spec.n_joints, spec.d_spatial, spec.d_temporal, spec.num_heads
    hs, ht = spec.h_spatial, spec.h_temporal
    v = N if valid_frames is None else valid_frames
    per_frame = J * 2 * ds                                                   # S1
    per_frame += spec.spatial_depth * J * (4 * ds * ds + 2 * ds * hs)        # S2
    per_frame += spec.spatial_depth * 2 * J * J * ds                         # S3
    per_frame += J * ds * dt                                                 # S4
    macs = v * per_frame
    macs += spec.temporal_depth * N * (4 * dt * dt + 2 * dt * ht)            # T2
    macs += spec.temporal_depth * 2 * N * N * dt                             # T3
    if spec.full_output:
        macs += N * dt * spec.out_dim                                        # T4
    for i, s in enumerate(spec.strides):
        L, Lo = spec.seq_lens[i], spec.seq_lens[i + 1]
        macs += L * (4 * dt * dt + dt * ht)                                  # Q1 linears (qkv, proj, fc1)
        macs += 2 * L * L * dt                                               # Q1 attention
        macs += Lo * 3 * ht * dt                                             # Q2 strided conv
    macs += dt * spec.out_dim                                                # Q3
    return macs
