"""Host-side mirror of the reference's training step (train.py:464-506) and optimizer setup (:403-415).

    trainer = Trainer(model, config)                       # AdamW + staircase ExponentialDecay for lr and wd
    loss = trainer.train_step(keypoints2d, keypoints3d, stride_masks)

The arithmetic (forward with stochastic depth, MPJPE loss, backward, fused AdamW/EMA) runs in libuu3d.so.
Data-parallel training: one process per GPU, every rank passes its local windows, the flat gradient buffer
is summed with one NCCL all-reduce (losses are normalised by the GLOBAL config BATCH_SIZE, so the combine
is a sum, not a mean — train.py:482, :488-489).
"""
from __future__ import annotations

import copy
import ctypes
import math
from ctypes import byref, c_float, c_int64, c_void_p
from typing import Optional

import numpy as np

from . import _lib
from .sharding import allreduce_gradients


def scheduler_by_name(name: str):
    """common/utils/schedules.py:16-32: schedule name -> factory of a host-side callable step -> value.
    ExponentialDecay (every shipped config) and the repository's own ExponentialDecayWithSteps (:36-99) are implemented;
    the Keras-only PiecewiseConstantDecay / CosineDecayRestarts are not."""
    if name == "ExponentialDecay":
        def make(initial_learning_rate, decay_steps, decay_rate, staircase=False):
            def schedule(step: int) -> float:
                p = step / decay_steps
                if staircase:
                    p = math.floor(p)
                return initial_learning_rate * decay_rate ** p
            return schedule
        return make
    if name == "ExponentialDecayWithSteps":
        # two staircases: every `decay_steps` steps one factor `decay_rate`, except that every `large_decay_steps`-th
        # step applies `large_decay_rate` instead (schedules.py:84-97: p = floor(t / small) - floor(t / large))
        def make(initial_learning_rate, decay_steps, decay_rate, large_decay_steps, large_decay_rate, name=None):
            def schedule(step: int) -> float:
                large_p = math.floor(step / large_decay_steps)
                small_p = math.floor(step / decay_steps) - large_p
                return initial_learning_rate * decay_rate ** small_p * large_decay_rate ** large_p
            return schedule
        return make
    raise NotImplementedError(f"schedule {name!r} (implemented: ExponentialDecay, ExponentialDecayWithSteps)")


class _DevView:
    """Zero-copy torch view of a library-owned device buffer (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class Trainer:
    def __init__(self, model, config, droppath: bool = True, seed: int = 0, math: str = "fp32"):
        """math: "fp32" (CUDA-core GEMMs, fp32-grade tensor-core attention, gradients within 2e-3 of fp32 autograd) or
        "tf32" (forward, dgrad and wgrad GEMMs of the temporal / strided blocks on the tcgen05 tensor cores with TF32
        products — TensorFlow's own default on Ampere-or-newer GPUs — and their attention on bf16 hi + lo operand planes)."""
        import torch
        if math not in ("fp32", "tf32"):
            raise ValueError("math must be 'fp32' or 'tf32'")
        self.math = math
        self.torch = torch
        self.model = model
        self.config = config
        self.lib = model._lib
        if config.OPTIMIZER != "AdamW":
            raise NotImplementedError("only OPTIMIZER == 'AdamW' (all shipped configs) is implemented")
        mk = scheduler_by_name(config.SCHEDULE)
        self.lr_schedule = mk(**config.SCHEDULE_PARAMS)
        wd_params = copy.deepcopy(config.SCHEDULE_PARAMS)            # train.py:408-411
        wd_params["initial_learning_rate"] = config.WEIGHT_DECAY
        self.wd_schedule = mk(**wd_params)
        # train.py:412-415 forwards **OPTIMIZER_PARAMS to tfa.optimizers.AdamW: honour what the fused update implements,
        # reject the rest loudly
        op = dict(config.OPTIMIZER_PARAMS or {})
        self.beta1 = float(op.pop("beta_1", 0.9))
        self.beta2 = float(op.pop("beta_2", 0.999))
        self.epsilon = float(op.pop("epsilon", 1e-8))                           # train.py:414 passes epsilon=1e-8
        if op.pop("amsgrad", False):
            raise NotImplementedError("OPTIMIZER_PARAMS amsgrad=True is not implemented by the fused AdamW update")
        op.pop("name", None)
        if op:
            raise NotImplementedError(f"unsupported OPTIMIZER_PARAMS keys: {sorted(op)}")
        self.iterations = 0
        self.ema_enabled = bool(config.EMA_ENABLED)
        self.ema_decay = config.EMA_DECAY
        dpr = config.DROP_PATH_RATE if isinstance(config.DROP_PATH_RATE, (list, tuple)) else [config.DROP_PATH_RATE] * 3
        arr = (c_float * 3)(*[float(x) for x in dpr])
        _lib.check(self.lib.uu_train_config(model._h, int(config.BATCH_SIZE), int(config.ROOT_KEYTPOINT),
                                            float(config.LOSS_WEIGHT_CENTER), float(config.LOSS_WEIGHT_SEQUENCE),
                                            arr, 1 if droppath else 0, seed))
        _lib.check(self.lib.uu_train_set_math(model._h, 1 if math == "tf32" else 0))
        # random token masking of the temporal input (net:287-311), masked value 0; 0.0 in every shipped config
        _lib.check(self.lib.uu_train_set_token_masking(model._h, float(getattr(config, "TOKEN_MASK_RATE", 0.0) or 0.0)))
        self._loss = torch.zeros(1, dtype=torch.float32, device=f"cuda:{model.device}")
        self._grad_view = None

    def grad_view(self):
        """Flat fp32 gradient buffer as a torch tensor (no copy) — the all-reduce operand."""
        if self._grad_view is None:
            ptr, n = c_void_p(), c_int64()
            _lib.check(self.lib.uu_grad_buffer(self.model._h, byref(ptr), byref(n)))
            self._grad_view = self.torch.as_tensor(_DevView(ptr.value, n.value), device=f"cuda:{self.model.device}")
        return self._grad_view

    def forward_backward(self, keypoints2d, keypoints3d, stride_masks):
        """Loss and gradients of the local windows; returns the loss as a device tensor (1,)."""
        torch = self.torch
        s = self.model.spec
        x = keypoints2d.contiguous().float()
        g = keypoints3d.contiguous().float()
        B = x.shape[0]
        assert tuple(x.shape[1:]) == (s.n_tok, s.n_joints, 2) and tuple(g.shape) == (B, s.n_tok, s.n_joints, 3)
        mptr = None
        if self.model.has_strided_input:
            mk = stride_masks.to(device=x.device, dtype=torch.uint8).contiguous()
            mptr = mk.data_ptr()
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self.lib.uu_train_forward_backward(self.model._h, x.data_ptr(), mptr, g.data_ptr(), B,
                                                      self.iterations, self._loss.data_ptr(), stream))
        return self._loss

    def apply_gradients(self):
        """optimizer.apply_gradients (train.py:499) + optional EMA (train.py:502-504, :554-556)."""
        it = self.iterations
        lr_t, wd_t = self.lr_schedule(it), self.wd_schedule(it)
        ema = -1.0
        if self.ema_enabled:
            ema = min(self.ema_decay, (1 + it) / (10 + it))
        stream = self.torch.cuda.current_stream(self.model.device).cuda_stream
        _lib.check(self.lib.uu_adamw_step(self.model._h, lr_t, wd_t, self.beta1, self.beta2, self.epsilon, it + 1,
                                          ema, stream))
        self.iterations += 1

    def init_comm(self, dist) -> None:
        """Create the library's own NCCL communicator (uu_comm_init): rank 0 draws the 128-byte id, torch.distributed
        only carries it to the other ranks.  Afterwards train_step runs the whole step inside the library, with the
        gradient all-reduce bucketed and overlapped with the backward pass."""
        torch = self.torch
        world, rank = dist.get_world_size(), dist.get_rank()
        if world <= 1:
            return
        buf = (ctypes.c_uint8 * 128)()
        if rank == 0:
            _lib.check(self.lib.uu_comm_unique_id(buf, 128))
        idt = torch.tensor(list(buf), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            idt = idt.to(f"cuda:{self.model.device}")
        dist.broadcast(idt, src=0)
        raw = (ctypes.c_uint8 * 128)(*idt.cpu().tolist())
        _lib.check(self.lib.uu_comm_init(self.model._h, raw, rank, world))
        self._comm = True

    def train_step(self, keypoints2d, keypoints3d, stride_masks, dist=None):
        """One optimisation step (train.py:464-506).  With a library communicator (init_comm) the step is ONE C-ABI call:
        forward, backward, bucketed all-reduce overlapped with the backward pass, AdamW.  Otherwise the gradients are
        summed with torch.distributed between forward_backward and apply_gradients (dist given), or not at all."""
        if getattr(self, "_comm", False):
            return self._fused_step(keypoints2d, keypoints3d, stride_masks)
        loss = self.forward_backward(keypoints2d, keypoints3d, stride_masks)
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            allreduce_gradients(self.grad_view(), dist)                   # NCCL sum over NVLink
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
        self.apply_gradients()
        return loss

    def _fused_step(self, keypoints2d, keypoints3d, stride_masks):
        torch = self.torch
        s = self.model.spec
        x = keypoints2d.contiguous().float()
        g = keypoints3d.contiguous().float()
        B = x.shape[0]
        assert tuple(x.shape[1:]) == (s.n_tok, s.n_joints, 2) and tuple(g.shape) == (B, s.n_tok, s.n_joints, 3)
        mptr = None
        if self.model.has_strided_input:
            mk = stride_masks.to(device=x.device, dtype=torch.uint8).contiguous()
            mptr = mk.data_ptr()
        it = self.iterations
        ema = min(self.ema_decay, (1 + it) / (10 + it)) if self.ema_enabled else -1.0
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self.lib.uu_train_step(self.model._h, x.data_ptr(), mptr, g.data_ptr(), B, it, self.lr_schedule(it),
                                          self.wd_schedule(it), self.beta1, self.beta2, self.epsilon, ema,
                                          self._loss.data_ptr(), stream))
        self.iterations += 1
        return self._loss

    # ---- introspection for the parity tests ----------------------------------------------------------
    def _state_view(self, which: int, allocate: bool):
        ptr, n = c_void_p(), c_int64()
        _lib.check(self.lib.uu_optimizer_state(self.model._h, which, 1 if allocate else 0, byref(ptr), byref(n)))
        if not ptr.value:
            return None
        return self.torch.as_tensor(_DevView(ptr.value, n.value), device=f"cuda:{self.model.device}")

    def state_dict(self) -> dict:
        """Optimizer state for checkpoint / resume (what the reference's tf.train.Checkpoint holds next to the weights,
        train.py:417-430): the step counter, Adam's flat first / second moments and, when EMA is on, the EMA weights."""
        self.torch.cuda.synchronize(self.model.device)
        out = {"iterations": np.int64(self.iterations)}
        for name, which in (("adam_m", 0), ("adam_v", 1), ("ema", 2)):
            v = self._state_view(which, False)
            if v is not None:
                out[name] = v.cpu().numpy().copy()
        return out

    def load_state_dict(self, state: dict) -> None:
        self.iterations = int(state["iterations"])
        for name, which in (("adam_m", 0), ("adam_v", 1), ("ema", 2)):
            if name in state:
                v = self._state_view(which, True)
                a = np.ascontiguousarray(state[name], dtype=np.float32)
                if a.size != v.numel():
                    raise ValueError(f"{name}: {a.size} values for a model with {v.numel()} parameter slots")
                v.copy_(self.torch.from_numpy(a))
        self.torch.cuda.synchronize(self.model.device)

    def get_grads(self):
        out = {}
        for (g, k), shp in self.model._keys:
            a = np.empty(shp, dtype=np.float32)
            _lib.check(self.lib.uu_get_grad(self.model._h, g.encode(), k, a.ctypes.data_as(c_void_p), a.size))
            out[(g, k)] = a
        return out

    def get_ema_weights(self):
        out = {}
        for (g, k), shp in self.model._keys:
            a = np.empty(shp, dtype=np.float32)
            _lib.check(self.lib.uu_get_ema_weight(self.model._h, g.encode(), k, a.ctypes.data_as(c_void_p), a.size))
            out[(g, k)] = a
        return out

    def token_keep(self, B: int):
        """(B, n_tok) 0/1 factors (1 - token mask) drawn by the last step's random token masking."""
        n = B * self.model.spec.n_tok
        buf = np.empty(n, dtype=np.float32)
        _lib.check(self.lib.uu_get_token_mask(self.model._h, buf.ctypes.data_as(c_void_p), n))
        return buf.reshape(B, self.model.spec.n_tok)

    def droppath_keeps(self, B: int):
        """{(stage, block, branch): (keep_prob, mask ndarray)} actually used by the last step, in the oracle's format
        (branch 0 = attention residual, 1 = MLP residual: independent draws, vision_transformer.py:185-190)."""
        s = self.model.spec
        out = {}
        for si, (stage, depth, n) in enumerate((("spatial", s.spatial_depth, B * s.n_tok), ("temporal", s.temporal_depth, B),
                                                 ("strided", len(s.strides), B))):
            for i in range(depth):
                for branch in (0, 1):
                    buf = np.empty(n, dtype=np.float32)
                    kp = c_float()
                    _lib.check(self.lib.uu_get_droppath_scale(self.model._h, si, i, branch, buf.ctypes.data_as(c_void_p), n,
                                                              byref(kp)))
                    if kp.value < 1.0:
                        out[(stage, i, branch)] = (kp.value, buf * kp.value)
        return out
