"""Full comparison of the TMA trailing-update kernel against the cp.async kernel of round 1 on the same operands (every entry, not a sample):
usage gpu_gemm_cmp.py m k n [reps].  Reports the entries that differ by more than 1e-12 relative and where they sit (tile coordinates)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multifebe_b200 import capi
m, k, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ctx = capi.Context(0)
rng = np.random.default_rng(3)
A = np.asfortranarray(rng.standard_normal((m, k)) + 1j * rng.standard_normal((m, k)))
B = np.asfortranarray(rng.standard_normal((k, n)) + 1j * rng.standard_normal((k, n)))
C = np.asfortranarray(rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n)))
os.environ["MFB_GEMM_TMA"] = "0"
R0, ms0 = ctx.zgemm_minus(C, A, B)
os.environ["MFB_GEMM_TMA"] = "1"
for rep in range(reps):
    R1, ms1 = ctx.zgemm_minus(C, A, B)
    d = np.abs(R1 - R0) / np.abs(R0).max()
    bad = np.argwhere(d > 1e-12)
    print("m n k", m, n, k, "rep", rep, "ms old %.3f tma %.3f" % (ms0, ms1), "max rel diff %.3e" % d.max(), "bad entries", len(bad), flush=True)
    if len(bad):
        tiles = {}
        for i, j in bad[:200000]:
            tiles.setdefault((int(i) // 64, int(j) // 32), 0); tiles[(int(i) // 64, int(j) // 32)] += 1
        print("   bad tiles (m-tile, n-tile): count", len(tiles), "first", sorted(tiles.items())[:12])
        i, j = bad[0]; print("   first bad entry", i, j, "old", R0[i, j], "tma", R1[i, j], "diff", d[i, j])
        rows = sorted(set(int(i) % 64 for i, j in bad[:5000])); cols = sorted(set(int(j) % 32 for i, j in bad[:5000]))
        print("   rows in tile", rows[:70], "cols in tile", cols)
