"""LU development aid: correctness against LAPACK at moderate n, then timings of the device-resident LU at full size.
usage: gpu_lu.py check | time [m]      (GEMM tile config via MFB_GEMM_CFG, sub-panel width via MFB_LU_IB)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multifebe_b200 import capi
from multifebe_b200.host import *

ctx = capi.Context(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "check"
if mode == "check":
    from scipy.linalg import lapack
    for m in (1, 3, 5, 9, 13):
        md = Model(cube_mesh(m, shape.TRI3), cube_bcs()); n = md.n_dof
        pr = capi.Problem(ctx, md)
        rng = np.random.default_rng(n)
        A = np.asfortranarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        A[:, 3] *= 1e-3; A[5, :] *= 40.0
        B = np.asfortranarray(rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2)))
        lu_ref, piv_ref, info = lapack.zgetrf(A); x_ref, info = lapack.zgetrs(lu_ref, piv_ref, B)
        Af = A.copy(order="F")
        x, ipiv = pr.solve_lse_c(Af, B, want_ipiv=True)
        print("n", n, "ipiv equal", np.array_equal(ipiv - 1, piv_ref), "LU err", np.abs(Af - lu_ref).max() / np.abs(lu_ref).max(),
              "x err", np.abs(x - x_ref).max() / np.abs(x_ref).max(), flush=True)
        pr.close()
else:
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    md = Model(cube_mesh(m, shape.TRI3), cube_bcs()); mat = Material(1, 1, 0.25, 0.03)
    pr = capi.Problem(ctx, md); n = md.n_dof
    for rep in range(2):
        x = pr.solve_frequency(9.0, mat)
        s = pr.stats()
        print("cfg", os.environ.get("MFB_GEMM_CFG"), "ib", os.environ.get("MFB_LU_IB"), "n", n, {k: round(s[k], 2) for k in ("MS_LU", "MS_PANEL", "MS_SWAP", "MS_TRSM", "MS_GEMM", "MS_SOLVE", "MS_ASSEMBLE", "LU_LAUNCHES")},
              "LU TF", round(8 / 3 * n ** 3 / s["MS_LU"] / 1e9, 2), "GEMM TF", round(s["GEMM_FLOPS"] / s["MS_GEMM"] / 1e9, 2), flush=True)
    pr.build_lse_mechanics_bem_harela(9.0, mat, want_host=False)
    print("   berr, rel", pr.residual(x), flush=True)
    pr.close()
