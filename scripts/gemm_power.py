"""Development aid: SM clock and board power while one GEMM shape runs back to back (is the gap between the
loads-off and the full kernel a clock effect?).  usage: [UU_GEMM_NOLOAD=1 | UU_GEMM_NOSTORE=1] python scripts/gemm_power.py"""
import ctypes, os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uplift_upsample_3dhpe_b200 import _lib
import pynvml
lib = _lib.load()
P = lambda t: ctypes.c_void_p(t.data_ptr())
M, N, K = 290816, 1152, 384
A = torch.randn(M, K, device="cuda").bfloat16()
Wt = torch.randn(N, K, device="cuda").bfloat16()
bias = torch.randn(N, device="cuda")
C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False
def poll():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
        time.sleep(0.01)
th = threading.Thread(target=poll); th.start()
for _ in range(50):
    _lib.check(lib.uu_op_gemm_bf16(P(A), K, M, K, P(Wt), N, N, P(bias), 0, None, 0, P(C), 1, N, None))
torch.cuda.synchronize()
samples.clear()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 4000
for _ in range(n):
    _lib.check(lib.uu_op_gemm_bf16(P(A), K, M, K, P(Wt), N, N, P(bias), 0, None, 0, P(C), 1, N, None))
e1.record(); torch.cuda.synchronize()
stop = True; th.join()
us = e0.elapsed_time(e1) / n * 1e3
clk = sorted(s[0] for s in samples); pw = sorted(s[1] for s in samples)
print(f"qkv {us:.1f} us/launch, {2*M*N*K/us/1e6:.0f} TFLOP/s, SM clock median {clk[len(clk)//2]} MHz, power median {pw[len(pw)//2]:.0f} W, "
      f"cycles/launch {us * clk[len(clk)//2]:.0f}, env {[k for k in os.environ if k.startswith('UU_GEMM')]}")
