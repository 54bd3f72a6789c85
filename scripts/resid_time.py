"""Development aid: run the residual-epilogue GEMM (uu_op_resid_gemm_bf16) at the projection / fc2 shapes, to be timed
under `ncu --metrics gpu__time_duration.sum -k regex:k_gemm_tc` (the op allocates temporaries, so events would lie)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uplift_upsample_3dhpe_b200 import _lib
lib = _lib.load()
P = lambda t: ctypes.c_void_p(t.data_ptr())
M, d = 290816, 384
for K in (384, 768):
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = torch.randn(K, d, device="cuda") / K ** 0.5
    bias = torch.randn(d, device="cuda")
    x = torch.randn(M, d, device="cuda").bfloat16()
    stats = torch.empty(M, d // 64, 2, device="cuda")
    flush = torch.empty(64 << 20, device="cuda", dtype=torch.int32)
    for it in range(3):
        flush.zero_()
        _lib.check(lib.uu_op_resid_gemm_bf16(P(A), M, K, P(W), P(bias), d, P(x), 1, None, 1, P(stats), None))
    torch.cuda.synchronize()
