"""ORACLE (test infrastructure, never shipped): training step restated with PyTorch CPU autograd.

PARITY UNPINNED: TensorFlow 2.4.3 autodiff, Keras Adam and tensorflow-addons 0.13.0 AdamW are not
installable here; their arithmetic is restated from the pinned versions' documented behaviour
(SURVEY.md §8c, Appendix C).  Only tests/, smoke() and bench.py's CPU legs may import this.

reference: train.py:464-506 (train_step), :403-415 (optimizer), common/utils/losses_3d.py:13-14,
common/utils/schedules.py:17-32 (Keras ExponentialDecay, staircase).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import forward_torch as OT


def exponential_decay(initial, decay_steps, decay_rate, staircase, step):
    """keras.optimizers.schedules.ExponentialDecay.__call__ (config/*.json SCHEDULE_PARAMS)."""
    p = step / decay_steps
    if staircase:
        p = math.floor(p)
    return initial * decay_rate ** p


def loss_fn(spec, full, central, keypoints3d, batch_size):
    """train.py:467-491: root-centre the GT, per-joint L2 norm (tf_mpjpe), two weighted means over the
    GLOBAL config BATCH_SIZE."""
    gt = keypoints3d - keypoints3d[:, :, spec.root_keypoint:spec.root_keypoint + 1, :]
    gt_c = gt[:, spec.n_tok // 2]
    central_loss = torch.linalg.norm(gt_c - central, dim=-1).sum() / (batch_size * spec.n_joints)
    seq_loss = torch.linalg.norm(gt - full, dim=-1).sum() / (batch_size * spec.n_tok * spec.n_joints)
    return spec.loss_weight_center * central_loss + spec.loss_weight_sequence * seq_loss


def loss_and_grads(spec, w_np, x2d, keypoints3d, stride_mask, batch_size, keeps=None, dtype=torch.float64,
                   token_keep=None):
    """Returns (loss, {key: grad ndarray}).  x2d is masked here like train.py:474-475."""
    w = OT.to_torch(w_np, dtype, requires_grad=True)
    x = torch.tensor(x2d, dtype=dtype)
    m = torch.tensor(stride_mask) if stride_mask is not None else None
    if spec.has_strided_input:
        x = x * m.to(dtype)[:, :, None, None]
    tk = torch.tensor(token_keep, dtype=dtype) if token_keep is not None else None
    full, central = OT.forward(spec, w, x, m, keeps=keeps, token_keep=tk)
    loss = loss_fn(spec, full, central, torch.tensor(keypoints3d, dtype=dtype), batch_size)
    loss.backward()
    return float(loss.detach()), {k: v.grad.numpy().copy() for k, v in w.items()}


def adamw_step(w, g, m, v, lr_t, wd_t, t, beta1=0.9, beta2=0.999, eps=1e-8):
    """tfa.optimizers.AdamW = DecoupledWeightDecayExtension + Keras Adam (fused kernel form):
    var -= wd_t * var (no lr factor, every variable); m, v update;
    var -= lr_t * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps).  In place on numpy dicts."""
    alpha = lr_t * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    for k in w:
        w[k] = w[k] - wd_t * w[k]
        m[k] = beta1 * m[k] + (1 - beta1) * g[k]
        v[k] = beta2 * v[k] + (1 - beta2) * g[k] * g[k]
        w[k] = w[k] - alpha * m[k] / (np.sqrt(v[k]) + eps)


def ema_update(ema, w, decay):
    """train.py:502-504: ema_w -= (1 - d) * (ema_w - w)."""
    for k in w:
        ema[k] = ema[k] - (1 - decay) * (ema[k] - w[k])
