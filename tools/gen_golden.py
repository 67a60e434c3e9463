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
  oracle_widening.npz  the same for the paths of SURVEY.md 8f rank 3 (argument `widening`): acoustic and poroelastic single regions
                     (A, b, x) and coupled two-region systems of the multi-region oracle (solid-fluid, fluid-poroelastic, solid-poroelastic,
                     poroelastic-poroelastic).
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


# the cases of oracle_widening.npz, shared with tests/test_golden_widening.py
def widening_cases():
    from multifebe_b200.host import (Fluid, FluidModel, Poro, PoroModel, Material, MultiRegionModel, Region, SOLID, FLUID, cube_mesh, two_box_mesh, shape)
    from multifebe_b200.host.multiregion import PORO
    fl = Fluid(1.25, 343.0, 0.01); po = Poro(rhof=1.0, rhos=2.2, lam=1.2, mu=1.0, xi=0.02, phi=0.35, rhoa=0.15, R=0.8, Q=0.5, b=0.4)
    ms = Material(2.0, 1.5, 0.25, 0.03); fl2 = Fluid(1.0, 1.2, 0.01)
    po2 = Poro(rhof=1.0, rhos=2.6, lam=2.0, mu=1.5, xi=0.03, phi=0.2, rhoa=0.1, R=0.5, Q=0.7, b=0.8)
    cases = {}
    fbc = {1: (0, 0.3 - 0.1j), 2: (0, 1.0), 3: (1, 0.0), 4: (1, 2e-6 + 1e-6j), 5: (1, 0.0), 6: (0, -0.5)}
    for et, m in [(shape.TRI3, 2), (shape.TRI6, 1), (shape.QUAD4, 2), (shape.QUAD8, 1), (shape.QUAD9, 1)]:
        cases["acoustic:%d:%d" % (et, m)] = ("fluid", FluidModel(cube_mesh(m, et), fbc), fl, 2 * np.pi * 60.0)
    pbc = {1: ([1, 0, 0, 0], [0.05 + 0.02j, 0, 0, 0]), 2: ([0, 1, 1, 1], [0.2, 1.0, 0, 0])}
    for p_, free in ((3, 2), (4, 2), (5, 3), (6, 3)):
        ct = [1, 1, 1, 1]; ct[free] = 0; pbc[p_] = (ct, [0, 0, 0, 0])
    for et, m in [(shape.TRI3, 1), (shape.QUAD4, 1), (shape.QUAD9, 1)]:
        cases["poro:%d:%d" % (et, m)] = ("poro", PoroModel(cube_mesh(m, et), pbc), po, 1.9)
    bpart = {b: b for b in (1, 2, 3, 4, 5, 6, 7, 13, 14, 15, 16)}
    lat1, lat2 = (3, 4, 5, 6), (13, 14, 15, 16)

    def side_bcs(kind, lat, end, driven):
        out = {}
        for q in lat:
            if kind == SOLID:
                out[q] = ([1, 0, 1] if q in (3, 4, 13, 14) else [1, 1, 0], [0, 0, 0])
            elif kind == FLUID:
                out[q] = (1, 0.0)
            else:
                ct = [1, 1, 1, 1]; ct[2 if q in (3, 4, 13, 14) else 3] = 0; out[q] = (ct, [0, 0, 0, 0])
        if kind == SOLID:
            out[end] = ([0, 1, 0], [0.1, 0.2 - 0.1j, 0.0]) if driven else ([0, 0, 0], [0, 0, 0])
        elif kind == FLUID:
            out[end] = (0, 0.7 + 0.1j) if driven else (1, 0.0)
        else:
            out[end] = ([0, 1, 1, 1], [0.3, 1.0, 0.0, 0.2j]) if driven else ([1, 0, 0, 0], [0, 0, 0, 0])
        return out
    mats = {SOLID: ms, FLUID: fl2, PORO: po}
    for kinds, ict in (((SOLID, FLUID), 0), ((FLUID, PORO), 0), ((PORO, FLUID), 1), ((SOLID, PORO), 0), ((PORO, PORO), 0)):
        bcs = side_bcs(kinds[0], lat1, 1, True); bcs.update(side_bcs(kinds[1], lat2, 2, False))
        m2 = po2 if kinds == (PORO, PORO) else mats[kinds[1]]
        mrm = MultiRegionModel(two_box_mesh(1, shape.QUAD8), [Region(kinds[0], mats[kinds[0]], [1, 3, 4, 5, 6, 7]), Region(kinds[1], m2, [-7, 2, 13, 14, 15, 16])],
                               bpart, bcs, interface_ctype={7: ict})
        cases["coupled:%s-%s:%d" % (kinds[0], kinds[1], ict)] = ("coupled", mrm, None, 1.7)
    return cases


def widening_system(kind, model, mat, omega):
    from oracle import oracle as orc
    from oracle.multiregion import MultiRegionOracle
    if kind == "fluid":
        A, b, _ = orc.PotOracle(model).assemble(omega, mat)
    elif kind == "poro":
        A, b, _ = orc.PorOracle(model).assemble(omega, mat)
    else:
        A, b = MultiRegionOracle(model).assemble(omega)
    return np.asarray(A), np.asarray(b)


def oracle_widening_vectors():
    out = {}
    for key, (kind, model, mat, omega) in widening_cases().items():
        A, b = widening_system(kind, model, mat, omega)
        out["b:" + key] = b; out["x:" + key] = np.linalg.solve(A, b)
        out["Av:" + key] = A @ widening_probe(len(b))                 # every entry of A enters these two products
        out["vA:" + key] = widening_probe(len(b))[::-1] @ A
        if key in FULL_MATRICES:
            out["A:" + key] = A
    return out


FULL_MATRICES = ("acoustic:5:2", "poro:7:1", "coupled:fluid-poro:0")


def widening_probe(n):
    k = np.arange(n)
    return np.cos(0.37 * k + 0.1) + 1j * np.sin(0.91 * k - 0.3)


if __name__ == "__main__":
    gd = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gd, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "widening":
        np.savez_compressed(os.path.join(gd, "oracle_widening.npz"), **oracle_widening_vectors())
        print("widening fixtures written to", gd); sys.exit(0)
    np.savez_compressed(os.path.join(gd, "oracle_static.npz"), **oracle_static_vectors())
    if len(sys.argv) > 1 and sys.argv[1] == "static":
        print("static fixtures written to", gd); sys.exit(0)
    json.dump(table_checksums(), open(os.path.join(gd, "quad_tables.json"), "w"), indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(gd, "oracle_pairs.npz"), **oracle_vectors())
    print("golden fixtures written to", gd)
