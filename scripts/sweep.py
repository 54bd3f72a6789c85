"""BASELINE config 5: large-batch throughput sweep over N in {81, 351} x s_in (valid cells only: s_in must be a
multiple of s_out, uplifiting_dataset.py:252-254) against the bf16 tensor-core roofline.  Prints a markdown table.
usage (on a B200): python scripts/sweep.py [--batch 4096] > profiles/sweep.md"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, forward_macs, spec_from_config, stride_mask  # noqa: E402
from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"bf16_tflops_sustained": 1400.0}
peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
B = a.batch
print(f"B = {B} windows per step, bf16 schedule, device-timed on a dedicated stream (CUDA-graph replay), centred stride masks; "
      f"roofline = algorithmic FLOPs (valid frames only in the spatial stages) / sustained bf16 peak {peak} TFLOP/s\n")
print("| config | N (frames) | tokens | s_in | valid tokens | ms / step | poses / s | algorithmic TFLOP/s | % of bf16 roofline |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
s = torch.cuda.Stream()
for name, s_ins in (("h36m_81", (4, 10, 20)), ("h36m_351", (5, 10, 20))):
    cfg = UpliftUpsampleConfig.preset(name)
    spec = spec_from_config(cfg)
    model = build_uplift_upsample_transformer(cfg, precision="bf16")
    for s_in in s_ins:
        m1 = stride_mask.stride_mask(spec.n_tok, cfg.SEQUENCE_STRIDE, s_in)
        x = torch.rand((B, spec.n_tok, 17, 2), device="cuda") * 2 - 1
        m = torch.from_numpy(np.stack([m1] * B)).cuda().to(torch.uint8)
        full = torch.empty((B, spec.n_tok, 17, 3), device="cuda")
        central = torch.empty((B, 17, 3), device="cuda")
        torch.cuda.synchronize()
        with torch.cuda.stream(s):
            for _ in range(4):
                model.forward_raw(x.data_ptr(), m.data_ptr(), B, full.data_ptr(), central.data_ptr(), s.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(a.steps):
                model.forward_raw(x.data_ptr(), m.data_ptr(), B, full.data_ptr(), central.data_ptr(), s.cuda_stream)
            e1.record(s)
            s.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        tf = 2 * forward_macs(spec, int(m1.sum())) * B / (ms * 1e-3) / 1e12
        print(f"| {name} | {spec.receptive_field} | {spec.n_tok} | {s_in} | {int(m1.sum())} | {ms:.3f} | {B / ms * 1e3:,.0f} | "
              f"{tf:.0f} | {100 * tf / peak:.1f} % |")
    model.close()
