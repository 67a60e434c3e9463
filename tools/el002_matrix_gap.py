"""ME-TH-EL-002 at 0.32 Hz: |A_gpu - A_oracle| entry by entry, relative to the column scale.  Run on the GPU box; about 15 s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from multifebe_b200 import capi
from multifebe_b200.host.casefile import CaseFile
from oracle import oracle as orc

case = CaseFile(os.path.join(ROOT, "tests", "golden", "ME-TH-EL-002", "INPUT_DATA_FILE.txt"))
md = case.build_model()
om = case.omega[10]
ctx = capi.Context(0); pr = capi.Problem(ctx, md)
A, b = pr.build_lse_mechanics_bem_harela(om, case.material)
A0, b0, _ = orc.Oracle(md).assemble(om, case.material)
sc = np.abs(A0).max(axis=0)
D = np.abs(A - A0) / sc
i, j = np.unravel_index(np.argmax(D), D.shape)
print("max |dA| / column scale %.2e at (%d, %d); |A0| there / column scale %.2e; median over entries %.2e; 99.9 %% quantile %.2e" % (
    D.max(), i, j, abs(A0[i, j]) / sc[j], np.median(D), np.quantile(D, 0.999)), flush=True)
print("b: max |db| / max|b| %.2e" % (np.abs(b - b0).max() / np.abs(b0).max()), flush=True)
# rows of the node that owns the column (free term + singular integrals) against all other entries
node_of_row = -np.ones(md.n_dof, dtype=int); node_of_col = -np.ones(md.n_dof, dtype=int)
for v in range(md.n_node):
    for k in range(3):
        node_of_row[md.row[v, k]] = v
        c = md.col_u[v, k] if md.ctype[v, k] != 0 else md.col_t[v, k]
        node_of_col[c] = v
own = node_of_row[:, None] == node_of_col[None, :]
print("own-node blocks: max %.2e   other entries: max %.2e" % (D[own].max(), D[~own].max()), flush=True)
x, x0 = np.linalg.solve(A, b), np.linalg.solve(A0, b0)
print("solving both on the host with LAPACK: |x - x0| / max|x0| %.2e (the GPU's own solve gave 2.7e-8)" % (np.abs(x - x0).max() / np.abs(x0).max()), flush=True)
pr.close(); ctx.close()
