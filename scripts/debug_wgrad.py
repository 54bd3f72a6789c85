import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes
from uplift_upsample_3dhpe_b200 import _lib
lib = _lib.load()
P = lambda t: ctypes.c_void_p(t.data_ptr())
R, Kd, Nd = 512, 384, 384
def run(X, Y):
    Xd, Yd = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    W = torch.zeros(Kd, Nd, device="cuda")
    _lib.check(lib.uu_op_wgrad_tf32(P(Xd), Kd, P(Yd), Nd, R, Kd, Nd, P(W), 0, None))
    torch.cuda.synchronize()
    return W.cpu().numpy()
for (r0, c0, n0) in [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 5, 9), (0, 33, 0), (0, 0, 33), (1, 0, 0), (9, 3, 2), (40, 130, 200), (300, 383, 383)]:
    X = np.zeros((R, Kd), np.float32); Y = np.zeros((R, Nd), np.float32)
    X[r0, c0] = 1; Y[r0, n0] = 1
    W = run(X, Y)
    nz = np.argwhere(W != 0)
    print((r0, c0, n0), "->", [(int(a), int(b), float(W[a, b])) for a, b in nz[:6]], len(nz))
rng = np.random.default_rng(0)
X = rng.normal(size=(R, Kd)).astype(np.float32); Y = rng.normal(size=(R, Nd)).astype(np.float32)
W = run(X, Y); want = X.T.astype(np.float64) @ Y
print("random: max err", np.abs(W - want).max(), "corr", np.corrcoef(W.ravel(), want.ravel())[0, 1])
