"""Development aid: device time of the tcgen05 GEMM at the forward pass's shapes (bf16 out, bias)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uplift_upsample_3dhpe_b200 import _lib
lib = _lib.load()
P = lambda t: ctypes.c_void_p(t.data_ptr())
M = int(sys.argv[1]) if len(sys.argv) > 1 else 290816
for (name, N, K) in [("qkv", 1152, 384), ("proj", 384, 384), ("fc1", 768, 384), ("fc2", 384, 768)]:
    A = torch.randn(M, K, device="cuda").bfloat16()
    Wt = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.randn(N, device="cuda")
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    flush = torch.empty(64 << 20, device="cuda", dtype=torch.int32)
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.uu_op_gemm_bf16(P(A), K, M, K, P(Wt), N, N, None if os.environ.get("NOBIAS") else P(bias), 0, None, 0, P(C), 1, N, None))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[1:])[len(ts[1:]) // 2]
    gb = (M * K + M * N) * 2 / 1e9
    print(f"{name:5s} M={M} N={N} K={K}: {t*1e3:7.1f} us  {2*M*N*K/t/1e9:7.1f} TFLOP/s  {gb/t*1e3:6.2f} TB/s (A+C bytes)")
