"""CUDA path (through the C ABI) against the golden vectors the reference's own code produced
(tests/golden, scripts/make_golden.py).  fp32 path: <= 1e-4 of the reference's float32 run on every
window (incl. all-masked ones) and of its float64 run where a valid token exists; bf16 path: <= 0.25
absolute on O(10) outputs (bf16 operands, fp32 accumulation)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from test_golden import FORWARD, load_forward_case  # noqa: E402
from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer  # noqa: E402
from uplift_upsample_3dhpe_b200.model import test_step as run_test_step  # noqa: E402


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("path", FORWARD, ids=[os.path.basename(p)[8:-4] for p in FORWARD])
def test_cuda_forward_matches_reference_golden(path, precision):
    cfg, spec, w, z = load_forward_case(path)
    x, m = z["x"], z["mask"]
    model = build_uplift_upsample_transformer(cfg, precision=precision, weights=w)
    full, central = run_test_step(model, torch.from_numpy(x).cuda(), torch.from_numpy(m).cuda())
    torch.cuda.synchronize()
    full, central = full.cpu().numpy(), central.cpu().numpy()
    model.close()
    tol = 1e-4 if precision == "fp32" else 0.25
    e32 = max(np.abs(central - z["central_f32"]).max(), np.abs(full - z["full_f32"]).max())
    valid = m.sum(axis=1) > 0
    e64 = max(np.abs(central[valid] - z["central"][valid]).max(), np.abs(full[valid] - z["full"][valid]).max())
    print(f"{os.path.basename(path)} {precision}: max|err| vs reference f32 {e32:.3e}, vs f64 (valid windows) {e64:.3e}")
    assert e32 <= tol and e64 <= tol


def test_h5_written_by_us_loads_through_c_abi(tmp_path):
    """Same file path the reference's loader was fed in make_golden.py: write .h5, load, compare outputs."""
    cfg, spec, w, z = load_forward_case(FORWARD[0])
    from uplift_upsample_3dhpe_b200 import h5lite
    p = str(tmp_path / "w.h5")
    h5lite.save_keras_weights(p, spec, w)
    model = build_uplift_upsample_transformer(cfg, precision="fp32")
    model.load_weights(p)
    full, central = run_test_step(model, torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["mask"]).cuda())
    torch.cuda.synchronize()
    assert np.abs(central.cpu().numpy() - z["central_f32"]).max() <= 1e-4
    model.close()
