"""The ctypes glue of capi.py dry-run without a GPU: a stub library accepts every call and checks only that the NUMBER of arguments equals the
prototype in include/mfb.h.  Catches the slips a blind-written binding can hide until its first hardware run (a missing argument shifts every
pointer after it); the numerics are the GPU tests' business."""
import os
import re
import sys
import numpy as np
import pytest
from multifebe_b200 import capi
from multifebe_b200.host import (Model, Material, FluidModel, Fluid, PoroModel, MultiRegionModel, Region, FLUID, cube_mesh, two_box_mesh, shape,
                                 room_bcs, InternalPointsModel)
from multifebe_b200.host.multiregion import PORO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def prototypes():
    text = open(os.path.join(ROOT, "include", "mfb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S); text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for m in re.finditer(r"\b(mfb_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


class StubLibrary:
    def __init__(self):
        self.proto = prototypes(); self.called = {}

    def __getattr__(self, name):
        if name.startswith("_") or name in ("proto", "called"):
            raise AttributeError(name)
        assert name in self.proto, "capi.py calls %s, which include/mfb.h does not declare" % name

        def fn(*args):
            assert len(args) == self.proto[name], "%s called with %d arguments, include/mfb.h declares %d" % (name, len(args), self.proto[name])
            self.called[name] = self.called.get(name, 0) + 1
            return 0
        return fn


@pytest.fixture()
def stub(monkeypatch):
    s = StubLibrary()
    monkeypatch.setattr(capi, "lib", lambda: s)
    return s


def test_header_is_parsed(stub):
    assert stub.proto["mfb_init"] == 2 and stub.proto["mfb_zsolve"] == 8 and len(stub.proto) > 40


def test_single_region_paths(stub):
    ctx = capi.Context(0)
    mat = Material(2.0, 1.0, 0.25, 0.02)
    bc = {q: ([1, 1, 1], [0, 0, 0]) for q in range(1, 7)}; bc[1] = ([0, 0, 0], [0, 0, 0]); bc[2] = ([1, 1, 1], [1.0, 0, 0])
    md = Model(cube_mesh(1, shape.QUAD8), bc)
    pr = capi.Problem(ctx, md)
    pr.build_lse_mechanics_bem_harela(1.1, mat); pr.solve_frequency(1.1, mat); pr.build_lse_mechanics_bem_staela(mat); pr.solve_static(mat)
    pr.stats(); pr.residual(np.zeros(md.n_dof, dtype=np.complex128)); pr.get_solution()
    pr.solve_lse_c(np.asfortranarray(np.eye(md.n_dof, dtype=np.complex128)), np.zeros(md.n_dof, dtype=np.complex128))
    pts = np.array([[0.5, 0.5, 0.5], [0.2, 0.3, 0.4]])
    ip = capi.InternalPoints(ctx, md, pts)
    x = np.zeros(md.n_dof, dtype=np.complex128)
    ip.displacements(1.1, mat, x); ip.stresses(1.1, mat, x); ip.displacements_static(mat, x); ip.stresses_static(mat, x); ip.close()
    pr.close()
    fl = Fluid(1.25, 343.0)
    fm = FluidModel(cube_mesh(1, shape.TRI3, L=3.0), room_bcs(1.0))
    pf = capi.Problem(ctx, fm)
    pf.build_lse_mechanics_bem_harpot(30.0, fl); xf = pf.solve_frequency_fluid(30.0, fl)
    ipf = capi.InternalPoints(ctx, fm, pts); ipf.pressures(30.0, fl, np.zeros(fm.n_dof, dtype=np.complex128)); ipf.close()
    pf.close(); ctx.close()
    for name in ("mfb_harela3d_setup", "mfb_harela3d_assemble", "mfb_staela3d_assemble", "mfb_harpot3d_setup", "mfb_harpot3d_assemble", "mfb_zsolve"):
        assert stub.called.get(name), name


def test_poroelastic_and_coupled_paths(stub):
    from test_oracle_multiregion import PO, poro_bcs_side, BPART, LAT1, LAT2
    from test_coupled_from_single_region import bcs_for, FL
    ctx = capi.Context(0)
    bcs = {1: ([1, 0, 0, 0], [0, 0, 0, 0]), 2: ([0, 1, 1, 1], [0, 1.0, 0, 0])}; bcs.update(poro_bcs_side((3, 4, 5, 6)))
    pm = PoroModel(cube_mesh(1, shape.QUAD9), bcs)
    pr = capi.Problem(ctx, pm)
    A, b = pr.build_lse_mechanics_bem_harpor(1.3, PO); assert A.shape == (pm.n_dof, pm.n_dof)
    assert pr.solve_frequency_poro(1.3, PO).shape == (pm.n_dof,)
    pr.close()
    b2 = bcs_for(FLUID, LAT1, 1, True); b2.update(bcs_for(PORO, LAT2, 2, False))
    mrm = MultiRegionModel(two_box_mesh(2, shape.TRI3), [Region(FLUID, FL, [1, 3, 4, 5, 6, 7]), Region(PORO, PO, [-7, 2, 13, 14, 15, 16])], BPART, b2,
                           interface_ctype={7: 0})
    cp = capi.CoupledProblem(ctx, mrm)
    A, b = cp.assemble(1.7); assert A.shape == (mrm.n_dof, mrm.n_dof)
    cp.solve_frequency(1.7)
    assert cp.solve_frequency_resident(1.7).shape == (mrm.n_dof,) and cp.solve_frequency_resident(2.1).shape == (mrm.n_dof,)
    cp.close(); ctx.close()
    for name in ("mfb_harpor3d_setup", "mfb_harpor3d_assemble", "mfb_harpor3d_solve_frequency", "mfb_system_zero", "mfb_combine_columns", "mfb_add_entries", "mfb_freeterm_terms"):
        assert stub.called.get(name), name


def test_driver_with_coupled_regions_through_the_stub(stub, tmp_path, monkeypatch):
    """driver.run -> GpuSolver -> capi.CoupledProblem on a two-region case file : the glue runs end to end and writes the *.nso rows."""
    import io
    from multifebe_b200 import driver
    from multifebe_b200.host.mesh import write_gmsh22
    from test_casefile_driver import TWO_REGION_DAT
    write_gmsh22(two_box_mesh(1, shape.QUAD9), str(tmp_path / "boxes.msh"))
    path = str(tmp_path / "two.dat")
    open(path, "w").write(TWO_REGION_DAT)
    nso = driver.run(path, log=io.StringIO())
    rows = [s for s in open(nso) if s.strip() and not s.startswith("#")]
    assert len(rows) > 0 and stub.called.get("mfb_harela3d_setup") and stub.called.get("mfb_harpot3d_setup") and stub.called.get("mfb_zsolve")


def test_bench_device_arms_of_the_secondary_workloads_through_the_stub(stub, monkeypatch, capsys):
    """bench.py --workload acoustic / coupled, device arms: the glue from the workload to the JSON line runs with the
    stub library (statistics and peaks read as 1.0) and prints the keys of the contract."""
    import argparse
    import ctypes as C
    import json
    import torch
    sys.path.insert(0, ROOT)
    import bench

    def get_stats(h, ptr):
        C.memmove(ptr, (C.c_double * capi.STAT_COUNT)(*([1.0] * capi.STAT_COUNT)), 8 * capi.STAT_COUNT); return 0

    def measure_peaks(h, a, b, c):
        for r in (a, b, c):
            r._obj.value = 1.0
        return 0
    stub.mfb_get_stats = get_stats; stub.mfb_measure_peaks = measure_peaks
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    args = argparse.Namespace(impl="ours", gpus=1, steps=1, warmup=1, no_cpu_baseline=True, acoustic_etype="tri3", acoustic_m=2, coupled_etype="tri3", coupled_m=1)
    for fn in (bench.run_acoustic, bench.run_coupled):
        fn(args)
        line = [s for s in capsys.readouterr().out.splitlines() if s.startswith("{")][-1]
        d = json.loads(line)
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                    "clocks", "e2e", "gpu_launches", "roofline"):
            assert key in d, (fn.__name__, key)
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and "workload" in d["config"]
