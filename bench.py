#!/usr/bin/env python
"""Headline benchmark: 3-D poses/sec of batched sliding-window inference (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm (oracle port) on the host CPU cores

A step = one forward pass over one batch of synthetic windows per GPU (weak scaling: the per-GPU
batch is fixed, windows are independent, no data-path collective).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", os.devnull)     # keep NCCL's banner off stdout: rank 0 prints ONE JSON line

import numpy as np  # noqa: E402
import torch  # noqa: E402

from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, spec_from_config, stride_mask  # noqa: E402
from uplift_upsample_3dhpe_b200.spec import ModelSpec  # noqa: E402

METRIC = "3D poses/sec @N=351 s_in=5"
UNIT = "poses/s"


# ---- algorithmic work per window, split by kernel kind --------------------------------
def macs_by_kind(spec: ModelSpec, valid: int) -> dict:
    N, J, ds, dt, hs, ht = spec.n_tok, spec.n_joints, spec.d_spatial, spec.d_temporal, spec.h_spatial, spec.h_temporal
    spatial = valid * (J * 2 * ds + spec.spatial_depth * (J * (4 * ds * ds + 2 * ds * hs) + 2 * J * J * ds))
    attention = spec.temporal_depth * 2 * N * N * dt
    gemm = valid * J * ds * dt + spec.temporal_depth * N * (4 * dt * dt + 2 * dt * ht) + N * dt * spec.out_dim
    for i in range(len(spec.strides)):
        L, Lo = spec.seq_lens[i], spec.seq_lens[i + 1]
        attention += 2 * L * L * dt
        gemm += L * (4 * dt * dt + dt * ht) + Lo * 3 * ht * dt
    gemm += dt * spec.out_dim
    return {"spatial": spatial, "attention": attention, "gemm_tc": gemm, "gemm_f32": gemm}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tensor_burst": d["bf16_tflops"], "tensor_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm": d["hbm_gbs"], "source": "measured"}
    # /opt/skills/guides/B200_PROFILING.md fallback
    return {"tensor_burst": 1590.0, "tensor_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every ~5 ms from a thread
    (nvidia-smi -lms is the fallback; its first sample can arrive after a short timed region has ended)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index, self.samples, self.mask, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _visible_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            h = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None

    def stop(self) -> dict:
        if self._nvml is None:
            return self._smi_once()
        self._stop.set()
        self._thread.join(timeout=1.0)
        if not self.samples:
            return self._smi_once()
        reasons = [n for n, bit in self.REASONS if self.mask & bit]
        return {"sm_mhz": statistics.median(self.samples), "sm_min_mhz": min(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples), "source": "nvml, 5 ms polling over the timed region"}

    def _smi_once(self) -> dict:
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b = [float(v) for v in out.strip().split(",")]
            return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "source": "nvidia-smi after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}


def synth_inputs(spec: ModelSpec, cfg, B: int, s_in: int, seed: int):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, (B, spec.n_tok, spec.n_joints, 2)).astype(np.float32)
    m1 = stride_mask.stride_mask(spec.n_tok, cfg.SEQUENCE_STRIDE, s_in)
    return x, np.ascontiguousarray(np.broadcast_to(m1, (B, spec.n_tok))).astype(np.uint8), int(m1.sum())


def cpu_reference_run(spec, cfg, s_in: int, sample_B: int, steps: int, warmup: int, budget_s: float):
    """The reference algorithm (oracle port, PyTorch CPU ops, all host threads) on a bounded sample."""
    from oracle import forward_torch as OT           # checker used as the CPU baseline (allowed here only)
    from uplift_upsample_3dhpe_b200 import weights
    torch.set_num_threads(os.cpu_count() or 1)
    w = OT.to_torch(weights.init_weights(spec, 1), torch.float32)
    x, m, _ = synth_inputs(spec, cfg, sample_B, s_in, 0)
    xt, mt = torch.from_numpy(x), torch.from_numpy(m.astype(bool))
    for _ in range(max(1, warmup)):
        OT.test_step(spec, w, xt, mt)
    times, t_start = [], time.time()
    for _ in range(steps):
        t0 = time.perf_counter()
        OT.test_step(spec, w, xt, mt)
        times.append(time.perf_counter() - t0)
        if time.time() - t_start > budget_s:
            break
    ms = 1e3 * sum(times) / len(times)
    return {"value": sample_B / (ms / 1e3), "ms_per_step": ms, "cores": torch.get_num_threads(), "steps": len(times),
            "sample": f"{sample_B} windows/step x {len(times)} steps, PyTorch-CPU fp32 restatement of the reference "
                      f"forward (TensorFlow not installable here)"}


def cpu_train_baseline(cfg, spec, sample_B: int, budget_s: float):
    """The reference's training arithmetic (oracle port: torch-CPU autograd over the forward restatement + AdamW) on a
    bounded sample, all host threads."""
    from oracle import train_torch as TT               # checker used as the CPU baseline (allowed here only)
    from uplift_upsample_3dhpe_b200 import weights
    torch.set_num_threads(os.cpu_count() or 1)
    w = weights.init_weights(spec, 1)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (sample_B, spec.n_tok, spec.n_joints, 2)).astype(np.float32)
    gt = rng.normal(0, 0.3, (sample_B, spec.n_tok, spec.n_joints, 3)).astype(np.float32)
    m = stride_mask.batch_stride_masks_train(spec.n_tok, cfg.SEQUENCE_STRIDE, cfg.MASK_STRIDE, sample_B, seed=0)
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    am = {k: np.zeros_like(v) for k, v in w64.items()}
    av = {k: np.zeros_like(v) for k, v in w64.items()}
    times, t_start = [], time.time()
    for it in range(3):
        t0 = time.perf_counter()
        _, g = TT.loss_and_grads(spec, w, x, gt, m, sample_B, dtype=torch.float32)
        TT.adamw_step(w64, {k: v.astype(np.float64) for k, v in g.items()}, am, av, 4e-5, 1e-6, it + 1)
        times.append(time.perf_counter() - t0)
        if time.time() - t_start > budget_s:
            break
    sec = min(times)
    return {"value": sample_B / sec, "unit": "windows/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_B} windows/step, best of {len(times)} steps: PyTorch-CPU fp32 autograd over the forward "
                      f"restatement + AdamW (TensorFlow not installable here)"}


def train_measure(a, rank, world, local_rank, dist, peaks, steps=5, warmup=3):
    """BASELINE config 4 (config/amass_351.json training step: fwd + bwd + AdamW, NCCL gradient all-reduce).  Strong
    scaling = the config's GLOBAL batch split over the ranks; weak = that batch per GPU.  Every rank runs uu_train_step
    (library-owned NCCL communicator, all-reduce bucketed and overlapped with the backward pass)."""
    from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer
    from uplift_upsample_3dhpe_b200.train import Trainer
    cfg = UpliftUpsampleConfig.preset(a.train_config)
    spec = spec_from_config(cfg)
    Bg = int(cfg.BATCH_SIZE)
    out = {"config": f"config/{a.train_config}.json training step (fwd+bwd+AdamW), mask strides {cfg.MASK_STRIDE} drawn per "
                     f"window, DropPath on, EMA {'on' if cfg.EMA_ENABLED else 'off'}",
           "math": a.train_math, "unit": "windows/s", "steps": steps, "warmup": warmup,
           "parity": "gradients vs fp32 torch autograd: tests/test_gpu_train.py (autodiff / AdamW arithmetic of TF is "
                     "parity-unpinned: no TensorFlow here)"}
    # algorithmic FLOPs: frames the stride mask drops are skipped; the training masks mix the
    # MASK_STRIDE values, so the valid-frame count is the mean over one drawn batch
    m_probe = stride_mask.batch_stride_masks_train(spec.n_tok, cfg.SEQUENCE_STRIDE, cfg.MASK_STRIDE, Bg, seed=0)
    valid_mean = float(m_probe.sum()) / Bg
    out["valid_frames_per_window_mean"] = round(valid_mean, 2)
    flop_per_window = 3 * 2 * sum(macs_by_kind(spec, valid_mean)[k] for k in ("spatial", "attention", "gemm_tc"))
    for mode in (("strong", "weak") if world > 1 else ("strong",)):
        B = Bg // world if mode == "strong" else Bg
        model = build_uplift_upsample_transformer(cfg, device=local_rank, precision="fp32")
        cfg_run = cfg if mode == "strong" else UpliftUpsampleConfig.preset(a.train_config, BATCH_SIZE=Bg * world)
        tr = Trainer(model, cfg_run, droppath=True, seed=rank, math=a.train_math)
        if dist is not None:
            tr.init_comm(dist)
        rng = np.random.default_rng(rank)
        x = torch.from_numpy(rng.uniform(-1, 1, (B, spec.n_tok, spec.n_joints, 2)).astype(np.float32)).cuda()
        gt = torch.from_numpy(rng.normal(0, 0.3, (B, spec.n_tok, spec.n_joints, 3)).astype(np.float32)).cuda()
        m = torch.from_numpy(stride_mask.batch_stride_masks_train(spec.n_tok, cfg.SEQUENCE_STRIDE, cfg.MASK_STRIDE, B,
                                                                  seed=rank)).cuda()
        for _ in range(warmup):
            tr.train_step(x, gt, m, dist)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = tr.train_step(x, gt, m, dist)
        e1.record()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        clocks = sampler.stop() if rank == 0 else None
        # AdamW alone (28 B / parameter, +8 with the EMA copy): the memory-bound kernel of the step
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it_save = tr.iterations
        a0.record()
        for _ in range(10):
            tr.apply_gradients()
        a1.record()
        torch.cuda.synchronize()
        tr.iterations = it_save
        adam_ms = a0.elapsed_time(a1) / 10
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        total = B * world
        tfl = flop_per_window * total / (ms * 1e-3) / 1e12
        bytes_pp = 36 if cfg.EMA_ENABLED else 28
        out[mode] = {"value": total / (ms / 1e3), "ms_per_step": ms, "global_batch": total, "batch_per_gpu": B,
                     "loss": float(loss.item()), "clocks": clocks,
                     "roofline": {"bound": "tensor", "achieved": round(tfl, 2), "unit": "TFLOP/s",
                                  "peak": peaks["tensor_burst"], "frac": round(tfl / peaks["tensor_burst"], 4),
                                  "frac_of_sustained": round(tfl / peaks["tensor_sustained"], 4),
                                  "algorithmic_gflop_per_step": round(flop_per_window * total / 1e9, 1),
                                  "note": "3 x forward FLOPs (fwd + dgrad + wgrad; masked frames skipped in the spatial stage) over the whole step"},
                     "adamw": {"ms": round(adam_ms, 4), "bytes_per_param": bytes_pp,
                               "achieved_gbs": round(bytes_pp * model.param_count / (adam_ms * 1e-3) / 1e9, 1),
                               "peak_gbs": peaks["hbm"],
                               "frac": round(bytes_pp * model.param_count / (adam_ms * 1e-3) / 1e9 / peaks["hbm"], 3),
                               "note": "k_adamw timed alone, 10 back-to-back launches (state 166 MB > 126 MB L2)"},
                     "parallelism": f"data-parallel x{world}: uu_train_step, library NCCL communicator, 3 gradient buckets "
                                    f"all-reduced on a side stream while the backward pass continues"}
        model.close()
        del tr, model
        torch.cuda.empty_cache()
    if rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_train_baseline(cfg, spec, 8, budget_s=20.0)
    return out


def quick_infer(cfg_name: str, s_in: int, B: int, local_rank: int, rank: int, steps: int, peaks, dist, world: int,
                precision: str = "bf16"):
    """Device-timed forward throughput of one more configuration (BASELINE configs 1, 3, 5): same method as the headline
    (rotating input pool larger than L2, CUDA events, max over ranks), fewer steps."""
    from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer
    cfg = UpliftUpsampleConfig.preset(cfg_name)
    spec = spec_from_config(cfg)
    model = build_uplift_upsample_transformer(cfg, device=local_rank, precision=precision)
    x_np, m_np, valid = synth_inputs(spec, cfg, B, s_in, seed=rank)
    pool_n = max(2, int(np.ceil(160e6 / (x_np.nbytes + m_np.nbytes))))
    pool_n = min(pool_n, 12)
    xs = [torch.from_numpy(x_np).cuda() + 0.001 * i for i in range(pool_n)]
    mk = torch.from_numpy(m_np).cuda()
    full = torch.empty((B, spec.n_tok, spec.n_joints, 3), dtype=torch.float32, device="cuda")
    central = torch.empty((B, spec.n_joints, 3), dtype=torch.float32, device="cuda")
    qs = torch.cuda.Stream()                    # a capturable stream: the library replays a CUDA graph after two calls
    qs.wait_stream(torch.cuda.current_stream())
    stream = qs.cuda_stream
    for i in range(2 * pool_n):
        model.forward_raw(xs[i % pool_n].data_ptr(), mk.data_ptr(), B, full.data_ptr(), central.data_ptr(), stream)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(qs)
    for i in range(steps):
        model.forward_raw(xs[i % pool_n].data_ptr(), mk.data_ptr(), B, full.data_ptr(), central.data_ptr(), stream)
    e1.record(qs)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # memory-bound kernels of this configuration (gather list, upsampling-token fill, residual / positional passes)
    model.set_profiling(True)
    model.forward_raw(xs[0].data_ptr(), mk.data_ptr(), B, full.data_ptr(), central.data_ptr(), stream)
    torch.cuda.synchronize()
    prof = model.get_profile()
    model.set_profiling(False)
    macs = macs_by_kind(spec, valid)
    flops = 2 * sum(macs[k] for k in ("spatial", "attention", "gemm_tc"))
    hb = hbm_kernel_bytes(spec, B, valid)
    hbm = {k: {"ms": round(prof[k][0], 4), "algorithmic_mb": round(hb[k] / 1e6, 2),
               "achieved_gbs": round(hb[k] / (prof[k][0] * 1e-3) / 1e9, 1), "peak_gbs": peaks["hbm"]}
           for k in hb if k in prof and prof[k][1] > 0 and prof[k][0] > 0} if precision == "bf16" else None
    model.close()
    del model, xs, full, central
    torch.cuda.empty_cache()
    val = world * B / (ms / 1e3)
    return {"workload": f"config/{cfg_name}.json forward, s_in={s_in}, {B} windows/GPU/step, {precision} schedule", "value": val, "unit": UNIT,
            "ms_per_step": ms, "steps": steps, "valid_tokens": valid, "n_tok": spec.n_tok,
            "whole_step_frac_of_tensor_peak_burst": round(flops * B / (ms * 1e-3) / 1e12 / peaks["tensor_burst"], 4),
            "whole_step_frac_of_tensor_peak_sustained": round(flops * B / (ms * 1e-3) / 1e12 / peaks["tensor_sustained"], 4),
            "hbm_kernels": hbm}


def hbm_kernel_bytes(spec: ModelSpec, B: int, valid: int) -> dict:
    """Algorithmic HBM bytes per step of the memory-bound kernel kinds of the bf16 schedule."""
    N, d = spec.n_tok, spec.d_temporal
    R = B * N
    row, stats = 2 * d, 8 * (d // 64)
    gather = R + 4 * B * valid + 8 * B                          # mask bytes in, ordered list out, per-window counts
    fill = (R - B * valid) * (row + stats) + R                  # masked rows written (token + PE) + statistics, mask read
    ln = R * (2 * row + stats)                                  # + PE of strided block 1 over the whole stream
    for i in range(len(spec.strides)):
        Ro = B * spec.seq_lens[i + 1]
        ln += Ro * (3 * row + stats)                            # identity rows + conv update in, next stream out
    return {"gather": gather, "token_fill": fill, "layernorm": ln}


def golden_accuracy(local_rank: int):
    """Error of the benchmarked bf16 schedule (and of the fp32 schedule) against the float64 outputs the REFERENCE's own
    model produced for the committed golden case (tests/golden/forward_h36m_351_sin5.npz, scripts/make_golden.py)."""
    from uplift_upsample_3dhpe_b200 import weights
    from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer, test_step
    path = os.path.join(ROOT, "tests", "golden", "forward_h36m_351_sin5.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path, allow_pickle=False)
    cfg = UpliftUpsampleConfig.preset(str(z["config"]), MASK_STRIDE=int(z["mask_stride"]))
    spec = spec_from_config(cfg)
    w = weights.init_weights(spec, seed=int(z["seed"]), perturb=True)
    out = {"case": "tests/golden/forward_h36m_351_sin5.npz (float64 outputs of the reference model, perturbed random weights)"}
    valid = z["mask"].sum(1) > 0
    ref = np.concatenate([z["full"][valid].ravel(), z["central"][valid].ravel()])
    out["rms_output"] = float(np.sqrt((ref ** 2).mean()))
    out["max_abs_output"] = float(np.abs(ref).max())
    for prec in ("bf16", "tf32", "fp32"):
        model = build_uplift_upsample_transformer(cfg, device=local_rank, precision=prec, weights=w)
        full, central = test_step(model, torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["mask"]).cuda())
        got = np.concatenate([full.cpu().numpy()[valid].ravel(), central.cpu().numpy()[valid].ravel()]).astype(np.float64)
        e = got - ref
        rms = float(np.sqrt((e ** 2).mean()))
        out[prec] = {"max_abs_err_vs_f64": float(np.abs(e).max()), "rms_err": rms,
                     "rel_rms_err": rms / out["rms_output"],
                     "mm_rms_at_1m_output_rms": 1000.0 * rms / out["rms_output"],
                     "mm_max_at_1m_output_rms": 1000.0 * float(np.abs(e).max()) / out["rms_output"]}
        model.close()
    return out


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything libraries print to fd 1 (NCCL's version banner, ...) goes to stderr; the JSON line is written to the
    saved descriptor, so stdout carries exactly ONE line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj) -> None:
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="h36m_351")
    ap.add_argument("--s-in", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4096, help="windows per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=64, help="windows per CPU-baseline step")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="train: amass_351-style training step (fwd+bwd+AdamW, NCCL gradient all-reduce), strong scaling")
    ap.add_argument("--train-math", default="tf32", choices=["tf32", "fp32"],
                    help="training GEMM arithmetic (uu_train_set_math)")
    ap.add_argument("--train-config", default="amass_351")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-configurations, accuracy and training objects")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = UpliftUpsampleConfig.preset(a.config)
    spec = spec_from_config(cfg)
    workload = (f"config/{a.config}.json forward (N={spec.receptive_field} frames = {spec.n_tok} tokens, "
                f"s_out={cfg.SEQUENCE_STRIDE}, s_in={a.s_in}), batched sliding-window inference")

    if a.mode == "train":
        assert torch.cuda.is_available()
        torch.cuda.set_device(local_rank)
        dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        tr = train_measure(a, rank, world, local_rank, dist, load_peaks(), steps=a.steps, warmup=a.warmup)
        if rank == 0:
            st = tr["strong"]
            emit({"metric": "training windows/sec (fwd+bwd+AdamW)", "value": st["value"], "unit": "windows/s",
                  "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": st["ms_per_step"],
                  "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": a.train_math,
                  "data": "synthetic", "config": {"workload": tr["config"], "parallelism": st["parallelism"]},
                  "train": tr})
        if dist is not None:
            dist.destroy_process_group()
        return

    if a.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(spec, cfg, a.s_in, a.cpu_sample, a.steps, a.warmup, budget_s=150.0)
        emit({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": r["steps"], "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "sample_windows_per_step": a.cpu_sample},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        })
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer
    model = build_uplift_upsample_transformer(cfg, device=local_rank, precision=a.precision)
    B = a.batch
    # input pool larger than L2 (126 MB): rotate buffers so no step finds its inputs cached
    x_np, m_np, valid = synth_inputs(spec, cfg, B, a.s_in, seed=rank)
    in_bytes = x_np.nbytes + m_np.nbytes
    pool_n = max(2, int(np.ceil(160e6 / in_bytes)))
    xs = [torch.from_numpy(x_np).cuda() + 0.001 * i for i in range(pool_n)]
    ms_ = [torch.from_numpy(m_np).cuda() for _ in range(pool_n)]
    full = torch.empty((B, spec.n_tok, spec.n_joints, 3), dtype=torch.float32, device="cuda")
    central = torch.empty((B, spec.n_joints, 3), dtype=torch.float32, device="cuda")
    # a dedicated (capturable) stream: the library replays a CUDA graph of the forward after its second call per buffer set
    bench_stream = torch.cuda.Stream()
    bench_stream.wait_stream(torch.cuda.current_stream())
    torch.cuda.set_stream(bench_stream)
    stream = bench_stream.cuda_stream

    def step(i):
        j = i % pool_n
        model.forward_raw(xs[j].data_ptr(), ms_[j].data_ptr(), B, full.data_ptr(), central.data_ptr(), stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    a.warmup = max(a.warmup, 2 * pool_n)             # every pool buffer is seen twice: eager run, then graph capture
    for i in range(a.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(i)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = model.last_launch_count * a.steps
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / a.steps
    value = world * B / (ms_step / 1e3)

    # ---- end to end through the host-buffer call: H2D of the step's inputs + D2H of its poses, every step
    hx = [torch.from_numpy(x_np + 0.001 * i).pin_memory() for i in range(2)]
    hm = torch.from_numpy(m_np).pin_memory()
    hc = torch.empty((B, spec.n_joints, 3), dtype=torch.float32).pin_memory()
    for i in range(2):
        model.forward_host(hx[i % 2].numpy(), hm.numpy(), None, hc.numpy())
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        model.forward_host(hx[i % 2].numpy(), hm.numpy(), None, hc.numpy())
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / a.steps
    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {"value": world * B / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(hx[0].numel() * 4 + hm.numel()), "d2h_bytes_per_step": int(hc.numel() * 4),
           "note": "uu_forward_host on pinned host buffers; ONLY the central poses (the metric's unit) are copied back — the "
                   "full-sequence head (net:421's first output) is computed on the device and not transferred"}

    # ---- the same windows cut on the device from one video: H2D of the video + centre list only.
    # Key-frame centres (multiples of s_out) so every window carries the same 71 valid tokens as the main workload.
    T = cfg.SEQUENCE_STRIDE * B
    hv = torch.from_numpy(np.random.default_rng(7).uniform(-1, 1, (T, spec.n_joints, 2)).astype(np.float32)).pin_memory()
    hcen = torch.arange(0, T, cfg.SEQUENCE_STRIDE, dtype=torch.int32).pin_memory()
    for i in range(2):
        model.forward_video_host(hv.numpy(), hcen.numpy(), cfg.SEQUENCE_STRIDE, a.s_in, hc.numpy())
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        model.forward_video_host(hv.numpy(), hcen.numpy(), cfg.SEQUENCE_STRIDE, a.s_in, hc.numpy())
    barrier()
    ev_ms = 1e3 * (time.perf_counter() - t0) / a.steps
    if dist is not None:
        t = torch.tensor([ev_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ev_ms = float(t.item())
    e2e_video = {"value": world * B / (ev_ms / 1e3), "unit": UNIT, "ms_per_step": ev_ms,
                 "h2d_bytes_per_step": int(hv.numel() * 4 + hcen.numel() * 4), "d2h_bytes_per_step": int(hc.numel() * 4),
                 "note": "uu_forward_video_host: sliding windows + globally aligned stride masks built on the device from a "
                         f"{T}-frame video, one window per key frame (centres = multiples of s_out)"}

    # ---- per-kernel roofline: a separate pass with events around every launch (not the timed region)
    peaks = load_peaks()
    model.set_profiling(True)
    acc = {}
    prof_steps = min(a.steps, 5)
    for i in range(prof_steps):
        step(i)
        torch.cuda.synchronize()
        for k, (ms, n) in model.get_profile().items():
            acc[k] = (acc.get(k, (0.0, 0))[0] + ms, n)
    model.set_profiling(False)
    macs = macs_by_kind(spec, valid)
    kernels = {}
    for k, (ms, n) in acc.items():
        if n == 0:
            continue
        ms_k = ms / prof_steps
        ent = {"ms_per_step": round(ms_k, 4), "launches_per_step": n}
        if k in macs:
            ent["tflops"] = round(2 * macs[k] * B / (ms_k * 1e-3) / 1e12, 2)
        kernels[k] = ent
    tot = sum(v["ms_per_step"] for v in kernels.values())
    for v in kernels.values():
        v["share"] = round(v["ms_per_step"] / tot, 4)
    dom = max((k for k in kernels if k in macs), key=lambda k: kernels[k]["ms_per_step"])
    n_dom = kernels[dom]["launches_per_step"]
    achieved = kernels[dom]["tflops"]
    traffic = None          # DRAM bytes per launch of the dominant kind, from the committed ncu capture (static evidence)
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if dom in tj.get("kinds", {}) and tj.get("batch") == B:
            traffic = tj["kinds"][dom]["dram_bytes_per_launch"]
    # burst peak when the timed region is short and the clocks stayed near the maximum, else the sustained one
    use_burst = (ms_total < 1000.0 and clocks is not None and (clocks.get("sm_mhz") or 0) >= 1900.0) if rank == 0 else True
    peak_used = peaks["tensor_burst"] if use_burst else peaks["tensor_sustained"]
    flops_step = 2 * sum(macs[k] for k in ("spatial", "attention", "gemm_tc")) * B
    roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": peak_used,
                "unit": "TFLOP/s", "frac": round(achieved / peak_used, 4), "traffic": traffic,
                "frac_of_burst": round(achieved / peaks["tensor_burst"], 4),
                "frac_of_sustained": round(achieved / peaks["tensor_sustained"], 4),
                "peak_source": f"{peaks['source']} bf16 {'burst' if use_burst else 'sustained'} (timed region "
                               f"{ms_total / 1e3:.2f} s at a median SM clock of "
                               f"{(clocks or {}).get('sm_mhz')} MHz); both fractions are given",
                "whole_step_frac_of_burst": round(flops_step / (ms_step * 1e-3) / 1e12 / peaks["tensor_burst"], 4),
                "whole_step_frac_of_sustained": round(flops_step / (ms_step * 1e-3) / 1e12 / peaks["tensor_sustained"], 4),
                "launches_per_step": n_dom,
                "avg_launch_ms": round(kernels[dom]["ms_per_step"] / n_dom, 5),
                "algorithmic_gflop_per_launch_avg": round(2 * macs[dom] * B / n_dom / 1e9, 3),
                "whole_step_frac_of_tensor_peak": round(2 * sum(macs[k] for k in ("spatial", "attention", "gemm_tc"))
                                                        * B / (ms_step * 1e-3) / 1e12 / peaks["tensor_sustained"], 4)}

    hb = hbm_kernel_bytes(spec, B, valid)
    hbm_kernels = {k: {"ms": kernels[k]["ms_per_step"], "algorithmic_mb": round(hb[k] / 1e6, 2),
                       "achieved_gbs": round(hb[k] / (kernels[k]["ms_per_step"] * 1e-3) / 1e9, 1), "peak_gbs": peaks["hbm"]}
                   for k in hb if k in kernels and kernels[k]["ms_per_step"] > 0}
    model.close()
    del model, xs, ms_, full, central
    torch.cuda.empty_cache()
    torch.cuda.set_stream(torch.cuda.default_stream())
    sub, accuracy, train = None, None, None
    if not a.no_extras:
        sub = {}
        for key, (cn, si, bb) in (("s_in20", ("h36m_351", 20, B)), ("h36m_81", ("h36m_81", 4, B)),
                                  ("b512", (a.config, a.s_in, 512))):
            sub[key] = quick_infer(cn, si, bb, local_rank, rank, 10, peaks, dist, world)
        # accuracy / throughput curve of the three schedules (same workload, 512 windows per step)
        for prec in ("tf32", "fp32"):
            sub["b512_" + prec] = quick_infer(a.config, a.s_in, 512, local_rank, rank, 5, peaks, dist, world, precision=prec)
        if rank == 0:
            accuracy = golden_accuracy(local_rank)
        train = train_measure(a, rank, world, local_rank, dist, peaks)
    cpu = None
    if rank == 0 and world == 1:
        r = cpu_reference_run(spec, cfg, a.s_in, a.cpu_sample, steps=5, warmup=1, budget_s=25.0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        if sub is not None:          # CPU baselines of BASELINE configs 1 and 3 (same port, smaller sample)
            for key, (cn, si) in (("s_in20", ("h36m_351", 20)), ("h36m_81", ("h36m_81", 4))):
                cfg_k = UpliftUpsampleConfig.preset(cn)
                rk = cpu_reference_run(spec_from_config(cfg_k), cfg_k, si, 32, steps=3, warmup=1, budget_s=10.0)
                sub[key]["cpu_baseline"] = {"value": rk["value"], "unit": UNIT, "cores": rk["cores"], "kind": "port",
                                            "sample": rk["sample"]}
    if rank == 0:
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.precision, "data": "synthetic",
            "config": {"workload": workload, "batch_per_gpu": B, "global_batch": world * B, "valid_tokens": valid,
                       "pose": "central-frame pose of one window (eval.py:189); full-sequence poses/s = n_tok x value",
                       "cache": f"inputs rotate through a {pool_n}-buffer pool ({pool_n * in_bytes / 1e6:.0f} MB > 126 MB L2); "
                                f"activations ({B * spec.n_tok * 8000 / 1e6:.0f} MB/step) exceed L2",
                       "parallelism": f"batch-sharded x{world}, no collective"},
            "clocks": clocks, "e2e": e2e, "e2e_video": e2e_video, "gpu_launches": launches, "roofline": roofline, "kernels": kernels,
            "hbm_kernels": hbm_kernels, "cpu_baseline": cpu, "accuracy": accuracy, "sub": sub, "train": train,
        })
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
