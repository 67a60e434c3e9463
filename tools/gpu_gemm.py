"""ZGEMM development aid: C -= A*B at the trailing-update shape (M = N = n, K = nb) through mfb_zgemm_minus; usage gpu_gemm.py [n] [k]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multifebe_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
k = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ctx = capi.Context(0)
rng = np.random.default_rng(1)
A = np.asfortranarray(rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k)))
B = np.asfortranarray(rng.standard_normal((k, n)) + 1j * rng.standard_normal((k, n)))
C = np.asfortranarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
for rep in range(2):
    R, ms = ctx.zgemm_minus(C, A, B)
print("cfg", os.environ.get("MFB_GEMM_CFG"), "n", n, "k", k, "ms", round(ms, 3), "algorithmic TFLOP/s", round(8.0 * n * n * k / ms / 1e9, 2))
idx = rng.integers(0, n, size=(200, 2))
ref = np.array([C[i, j] - A[i, :] @ B[:, j] for i, j in idx]); got = np.array([R[i, j] for i, j in idx])
print("   max rel err on 200 sampled entries", np.abs(ref - got).max() / np.abs(ref).max())
