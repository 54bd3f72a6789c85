"""Where does a small-batch training step go?  Local batch B (default 64 = the 8-GPU strong-scaling point of global 512) on
one GPU: device time per step (CUDA events), host time to ENQUEUE a step (perf_counter around the call, no sync), so that a
launch-bound step shows as host time ~ device time.  usage: python scripts/train_small_batch.py [B]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/scripts/", 1)[0])
from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, spec_from_config, stride_mask  # noqa: E402
from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer  # noqa: E402
from uplift_upsample_3dhpe_b200.train import Trainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = UpliftUpsampleConfig.preset("amass_351")
spec = spec_from_config(cfg)
model = build_uplift_upsample_transformer(cfg, device=0, precision="fp32")
tr = Trainer(model, cfg, droppath=True, seed=0, math="tf32")
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.uniform(-1, 1, (B, spec.n_tok, spec.n_joints, 2)).astype(np.float32)).cuda()
gt = torch.from_numpy(rng.normal(0, 0.3, (B, spec.n_tok, spec.n_joints, 3)).astype(np.float32)).cuda()
m = torch.from_numpy(stride_mask.batch_stride_masks_train(spec.n_tok, cfg.SEQUENCE_STRIDE, cfg.MASK_STRIDE, B, seed=0)).cuda()
for _ in range(5):
    tr.train_step(x, gt, m, None)
torch.cuda.synchronize()
n = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
host = 0.0
e0.record()
for _ in range(n):
    t0 = time.perf_counter()
    tr.train_step(x, gt, m, None)
    host += time.perf_counter() - t0
e1.record()
torch.cuda.synchronize()
print(f"B={B}: device {e0.elapsed_time(e1) / n:.3f} ms/step, host enqueue {1e3 * host / n:.3f} ms/step")
