"""Development aid: eager launches vs CUDA-graph replay of the same forward (how much of the step is launch gaps)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uplift_upsample_3dhpe_b200 import UpliftUpsampleConfig, spec_from_config, stride_mask
from uplift_upsample_3dhpe_b200.model import build_uplift_upsample_transformer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = UpliftUpsampleConfig.preset("h36m_351"); spec = spec_from_config(cfg)
x = torch.rand((B, spec.n_tok, 17, 2), device="cuda") * 2 - 1
m = torch.from_numpy(np.stack([stride_mask.stride_mask(spec.n_tok, 5, 5)] * B)).cuda().to(torch.uint8)
full = torch.empty((B, spec.n_tok, 17, 3), device="cuda"); central = torch.empty((B, 17, 3), device="cuda")
model = build_uplift_upsample_transformer(cfg, precision="bf16")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        model.forward_raw(x.data_ptr(), m.data_ptr(), B, full.data_ptr(), central.data_ptr(), s.cuda_stream)
    torch.cuda.synchronize()
    def timeit(fn, n=20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(n): fn()
        e1.record(s); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_eager = timeit(lambda: model.forward_raw(x.data_ptr(), m.data_ptr(), B, full.data_ptr(), central.data_ptr(), s.cuda_stream))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        model.forward_raw(x.data_ptr(), m.data_ptr(), B, full.data_ptr(), central.data_ptr(), s.cuda_stream)
    t_graph = timeit(lambda: g.replay())
print(f"B={B}: eager {t_eager:.3f} ms, graph replay {t_graph:.3f} ms")
