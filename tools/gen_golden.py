#!/usr/bin/env python3
"""Generate the committed fixtures under tests/golden/ (run in the build container, where /root/reference exists).

  quad_tables.json   per-rule exactly-rounded checksums (math.fsum) of the quadrature nodes/weights parsed from the
                     REFERENCE's own data statements (lib/fbem/src/resources_quad_rules/{gl11,gl01,gj01,wantri}.rc);
                     this is the only numeric golden material the reference ships for this path.
  oracle_pairs.npz   h,g blocks of selected (collocation point, element) pairs and small assembled systems computed by
                     the CPU ORACLE (oracle/harela3d_oracle.cpp).  These are regression vectors of the oracle, NOT
                     reference output: the reference (Fortran) cannot be compiled or run here.
  oracle_static.npz  the same for the static (Kelvin) path: small assembled real systems, their dgetrf/dgetrs solutions and one
                     pair per integration mode (run with the argument `static` to write only this file).
"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np


def table_checksums():
    from gen_quad_tables import parse_rc, REF
    out = {}
    for pre in ("gl11", "gl01", "gj01"):
        d = parse_rc(f"{REF}/{pre}.rc")
        n = d[pre + "_n"]
        for r in range(32):
            x = d[pre + "_xi"][r * 32: r * 32 + n[r]]; w = d[pre + "_w"][r * 32: r * 32 + n[r]]
            out[f"{pre}:{r + 1}"] = [math.fsum(v * (i + 1) for i, v in enumerate(x)).hex(), math.fsum(w).hex(),
                                     math.fsum(a * b for a, b in zip(x, w)).hex()]
    d = parse_rc(f"{REF}/wantri.rc")
    n = d["wantri_n"]
    for r in range(d["wantri_nr"][0]):
        x1 = d["wantri_xi1"][r * 176: r * 176 + n[r]]; x2 = d["wantri_xi2"][r * 176: r * 176 + n[r]]; w = d["wantri_w"][r * 176: r * 176 + n[r]]
        out[f"wantri:{r + 1}"] = [n[r], math.fsum(v * (i + 1) for i, v in enumerate(x1)).hex(), math.fsum(v * (i + 1) for i, v in enumerate(x2)).hex(),
                                  math.fsum(w).hex()]
    return out


def oracle_vectors():
    from multifebe_b200.host import Model, Material, cube_mesh, cube_bcs, shape
    from oracle import oracle as orc
    mat = Material(1.0, 1.0, 0.25, 0.03)
    out = {}
    for et, m in [(shape.TRI3, 2), (shape.TRI6, 1), (shape.QUAD4, 2), (shape.QUAD8, 1), (shape.QUAD9, 1)]:
        md = Model(cube_mesh(m, et), cube_bcs())
        o = orc.Oracle(md)
        for om in (0.7, 4.0):
            A, b, st = o.assemble(om, mat)
            x, _, _ = orc.lu_solve(A, b)
            if om == 4.0:
                out[f"A:{et}:{m}:{om}"] = A
            out[f"b:{et}:{m}:{om}"] = b; out[f"x:{et}:{m}:{om}"] = x
        # one pair per integration mode
        seen = {}
        for c in range(md.n_colloc):
            for e in range(md.n_elem):
                mode, d, bx = o.pair_mode(e, md.colloc_x[c])
                key = "reg" if mode < 100 else ("adp" if mode == 100 else "sing")
                if key not in seen:
                    h, g, mode2, _ = o.pair(e, md.colloc_x[c], 4.0, mat)
                    seen[key] = 1
                    out[f"pair:{et}:{key}"] = np.concatenate([[c, e, mode2], h.ravel().view(np.float64), g.ravel().view(np.float64)])
            if len(seen) == 3:
                break
    return out


def oracle_static_vectors():
    from multifebe_b200.host import Model, Material, cube_mesh, cube_bcs, shape
    from oracle import oracle as orc
    mat = Material(1.0, 1.3, 0.25, 0.0)
    out = {}
    for et, m in [(shape.TRI3, 2), (shape.TRI6, 1), (shape.QUAD4, 2), (shape.QUAD8, 1), (shape.QUAD9, 1)]:
        md = Model(cube_mesh(m, et), cube_bcs())
        o = orc.Oracle(md)
        A, b, st = o.assemble_static(mat)
        x, _, _ = orc.lu_solve_real(A, b)
        out[f"A:{et}:{m}"] = A; out[f"b:{et}:{m}"] = b; out[f"x:{et}:{m}"] = x
        seen = {}
        for c in range(md.n_colloc):
            for e in range(md.n_elem):
                mode, d, bx = o.pair_mode(e, md.colloc_x[c])
                key = "reg" if mode < 100 else ("adp" if mode == 100 else "sing")
                if key not in seen:
                    h, g, mode2 = o.pair_static(e, md.colloc_x[c], mat)
                    seen[key] = 1
                    out[f"pair:{et}:{key}"] = np.concatenate([[c, e, mode2], h.ravel(), g.ravel()])
            if len(seen) == 3:
                break
    return out


if __name__ == "__main__":
    gd = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gd, exist_ok=True)
    np.savez_compressed(os.path.join(gd, "oracle_static.npz"), **oracle_static_vectors())
    if len(sys.argv) > 1 and sys.argv[1] == "static":
        print("static fixtures written to", gd); sys.exit(0)
    json.dump(table_checksums(), open(os.path.join(gd, "quad_tables.json"), "w"), indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(gd, "oracle_pairs.npz"), **oracle_vectors())
    print("golden fixtures written to", gd)
