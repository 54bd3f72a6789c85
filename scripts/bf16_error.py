"""Development aid: error statistics of the bf16 schedule against the float64 golden outputs of the reference
(tests/golden/forward_*.npz): max |err|, rms err, output scale."""
import glob, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_golden import FORWARD, load_forward_case
from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer, test_step

print("| case | precision | max abs err (central / full) | rms err | rms output | max abs output |")
print("|---|---|---:|---:|---:|---:|")
for path in FORWARD:
    cfg, spec, w, z = load_forward_case(path)
    valid = z["mask"].sum(1) > 0
    for prec in ("fp32", "bf16"):
        model = build_uplift_upsample_transformer(cfg, precision=prec, weights=w)
        full, central = test_step(model, torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["mask"]).cuda())
        full, central = full.cpu().numpy()[valid], central.cpu().numpy()[valid]
        rf, rc = z["full"][valid], z["central"][valid]
        e = np.concatenate([(full - rf).ravel(), (central - rc).ravel()])
        o = np.concatenate([rf.ravel(), rc.ravel()])
        print(f"| {os.path.basename(path)[8:-4]} | {prec} | {np.abs(central - rc).max():.2e} / {np.abs(full - rf).max():.2e} | "
              f"{np.sqrt((e ** 2).mean()):.2e} | {np.sqrt((o ** 2).mean()):.2f} | {np.abs(o).max():.1f} |")
        model.close()
