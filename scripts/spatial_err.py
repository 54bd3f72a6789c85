"""Development aid: error statistics of the bf16 spatial kernel against the fp64 oracle."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import forward_np as O
from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, spec_from_config, weights, _lib
from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer
lib = _lib.load()
cfg = UpliftUpsampleConfig.preset("h36m_351", MASK_STRIDE=5)
spec = spec_from_config(cfg)
B = 40
for perturb in (False, True):
    for seed in (3, 4):
        w = weights.init_weights(spec, seed, perturb=perturb)
        if perturb == "big":
            pass
        rng = np.random.default_rng(B + seed)
        x = rng.uniform(-1, 1, (B, spec.n_tok, 17, 2)).astype(np.float32)
        m = np.ones((B, spec.n_tok), dtype=bool)
        _, _, inter = O.forward(spec, w, x, m, dtype=np.float64, return_intermediates=True)
        want = inter["spatial"].reshape(B * spec.n_tok, 544)
        model = build_uplift_upsample_transformer(cfg, precision="bf16", weights=w)
        xd = torch.from_numpy(x).cuda()
        out = torch.zeros((B * spec.n_tok, 544), dtype=torch.bfloat16, device="cuda")
        n = ctypes.c_int32(-1)
        _lib.check(lib.uu_op_spatial(model._h, ctypes.c_void_p(xd.data_ptr()), None, B, ctypes.c_void_p(out.data_ptr()), ctypes.byref(n), None))
        got = out.float().cpu().numpy()
        e = got - want
        # joint 16 rows vs others
        e3 = e.reshape(-1, 17, 32)
        print(f"perturb={perturb} seed={seed}: max {np.abs(e).max():.4f} rms {np.sqrt((e**2).mean()):.5f} "
              f"rms joints0-15 {np.sqrt((e3[:, :16]**2).mean()):.5f} rms joint16 {np.sqrt((e3[:, 16]**2).mean()):.5f} "
              f"bf16-rounding-of-want rms {np.sqrt(((torch.from_numpy(want).bfloat16().double().numpy()-want)**2).mean()):.5f}")
        model.close()
