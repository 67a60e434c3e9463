"""ZGEMM development aid: C -= A*B at the trailing-update shape through mfb_zgemm_minus; usage gpu_gemm.py [m] [k] [n]   (n defaults to m)
MFB_GEMM_TMA=0 selects the cp.async kernel of round 1, default is the TMA kernel (gemm_tma.cu) when k is a multiple of 16."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multifebe_b200 import capi
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = int(sys.argv[3]) if len(sys.argv) > 3 else m
ctx = capi.Context(0)
rng = np.random.default_rng(1)
A = np.asfortranarray(rng.standard_normal((m, k)) + 1j * rng.standard_normal((m, k)))
B = np.asfortranarray(rng.standard_normal((k, n)) + 1j * rng.standard_normal((k, n)))
C = np.asfortranarray(rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n)))
for rep in range(3):
    R, ms = ctx.zgemm_minus(C, A, B)
print("tma", os.environ.get("MFB_GEMM_TMA", "1"), "cfg", os.environ.get("MFB_GEMM_CFG"), "m n k", m, n, k, "ms", round(ms, 3), "algorithmic TFLOP/s", round(8.0 * m * n * k / ms / 1e9, 2),
      "executed (6mnk)", round(6.0 * m * n * k / ms / 1e9, 2))
if m * n <= 4_000_000:
    ref = C - A @ B
    print("   max rel err (all entries)", np.abs(ref - R).max() / np.abs(ref).max())
else:
    idx = np.stack([rng.integers(0, m, size=300), rng.integers(0, n, size=300)], axis=1)
    idx[:8] = [[0, 0], [m - 1, n - 1], [m - 1, 0], [0, n - 1], [63, 31], [64, 32], [m // 2, n // 2], [17, n - 2]]
    ref = np.array([C[i, j] - A[i, :] @ B[:, j] for i, j in idx]); got = np.array([R[i, j] for i, j in idx])
    print("   max rel err on 300 sampled entries", np.abs(ref - got).max() / np.abs(ref).max())
