"""Quick device-side timing of the forward pass (development aid, not the bench contract)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, forward_macs, spec_from_config, stride_mask  # noqa: E402
from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="h36m_351")
ap.add_argument("--s-in", type=int, default=5)
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--precisions", default="fp32,bf16")
a = ap.parse_args()

cfg = UpliftUpsampleConfig.preset(a.config)
spec = spec_from_config(cfg)
B = a.batch
x = torch.rand((B, spec.n_tok, 17, 2), device="cuda") * 2 - 1
m1 = stride_mask.stride_mask(spec.n_tok, cfg.SEQUENCE_STRIDE, a.s_in)
m = torch.from_numpy(np.stack([m1] * B)).cuda()
flops = 2 * forward_macs(spec, int(m1.sum()))
for prec in a.precisions.split(","):
    model = build_uplift_upsample_transformer(cfg, precision=prec)
    for _ in range(a.warmup):
        model([x, m])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        model([x, m])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"{a.config} s_in={a.s_in} B={B} {prec}: {ms:.3f} ms/step, {B / ms * 1e3:.0f} windows/s, "
          f"{B * flops / ms / 1e9:.1f} TFLOP/s algorithmic, launches={model.last_launch_count}")
    model.close()
