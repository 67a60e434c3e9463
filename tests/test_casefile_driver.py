"""Host formats either side of the hot path: the reference's case file + Gmsh 2.2 mesh in, its *.nso file out,
and the stand-alone driver (python -m multifebe_b200 -i case.dat) run here with the ORACLE as the solver (no GPU in the CPU suite;
tests/test_gpu_driver.py runs the same cases through the CUDA path)."""
import io
import os
import sys
import numpy as np
import pytest

from multifebe_b200.host import cube_mesh, write_gmsh22, shape, Fluid, room_analytic
from multifebe_b200.host.casefile import CaseFile, CaseFileError, read_frequencies, elastic_constants
from multifebe_b200.host.export import read_nso
from multifebe_b200.host.fortran_format import RealFormat, fmt_real, fmt_int, int_width
from multifebe_b200 import driver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/docs/examples"

# numbers as the reference printed them with real_format = eng_double (en27.16e3):
# docs/examples/ME-ST-EL-002/doc_src/ME-ST-EL-002.tex:461-476 (its *.tot listing) -- golden vectors of the export format
GOLDEN_EN27 = ["416.6665124818034194E-003", "-21.7024668712401838E-009", "0.0000000000000000E+000", "-646.6494794654632685E-012",
               "89.2479633341829967E-009", "302.9093621814738513E-003", "-308.8234562630653990E-003", "149.1666871671048083E-009",
               "2.2524860476320668E-009", "833.3332552543941674E-003", "15.2971017006141144E-009", "-302.9094707058440639E-003"]


def test_engineering_format_reproduces_the_reference_listing():
    rf = RealFormat("eng_double")
    for s in GOLDEN_EN27:
        out = rf(float(s))
        assert len(out) == 27 and out.strip() == s


def test_other_edit_descriptors():
    assert RealFormat()(0.1) == "  100.00000000E-03"                      # default en18.8e2 (src/read_export.f90:66)
    assert RealFormat()(-1234.5678) == "   -1.23456780E+03"
    assert RealFormat()(999.999999996) == "    1.00000000E+03"              # rounding carries into the next group of three
    assert RealFormat("sci_double")(0.41666651248180342) == "  0.4166665124818034E+000"
    assert RealFormat("sci_simple")(-15.0) == " -0.15000000E+02"
    assert RealFormat("sci_less")(0.0) == "  0.000E+00"
    assert fmt_real(1.5e-7, "es", 12, 3, 2) == "   1.500E-07"
    assert fmt_int(42, 5) == "   42" and fmt_int(123456, 4) == "****"
    assert int_width(300, 1, 6, 744, 462) == 4                              # i4: digits of the largest id + 1 (:61)


def test_frequency_lists():
    om, u = read_frequencies(["rad/s", "lin", "300", "0.01", "15."])           # t3.dat
    assert u == "w" and len(om) == 300 and om[0] == 0.01 and om[-1] == 15.0
    delta = (15.0 - 0.01) / 299.0
    assert om[7] == 0.01 + delta * 7.0
    om, u = read_frequencies(["Hz", "lin", "50", "1.", "300.000000"])          # room.dat
    assert u == "f" and np.allclose(om / (2 * np.pi), np.linspace(1.0, 300.0, 50), rtol=1e-14)
    om, _ = read_frequencies(["rad/s", "log", "4", "1.", "1000."])
    assert np.allclose(om, [1.0, 10.0, 100.0, 1000.0], rtol=1e-14)
    om, _ = read_frequencies(["Hz", "list", "3", "5.", "2.", "9."])
    assert np.allclose(om, 2 * np.pi * np.array([5.0, 2.0, 9.0]))
    with pytest.raises(CaseFileError):
        read_frequencies(["rpm", "lin", "3", "1.", "2."])
    with pytest.raises(CaseFileError):
        read_frequencies(["Hz", "lin", "3", "2.", "1."])


def test_elastic_constants_all_pairs_agree():
    full = elastic_constants({"E": 2.5, "nu": 0.25})
    names = ["E", "nu", "lambda", "mu", "K"]
    for i in range(5):
        for j in range(i + 1, 5):
            got = elastic_constants({names[i]: full[names[i]], names[j]: full[names[j]]})
            for k in names:
                assert abs(got[k] - full[k]) < 1e-12, (names[i], names[j], k)
    with pytest.raises(CaseFileError):
        elastic_constants({"E": 1.0})


SOLID_DAT = """[problem]
n = 3D
type = mechanics
analysis = %(analysis)s
%(freq)s
[settings]
mesh_file_mode = 2 "cube.msh"

[materials]
1
1 elastic_solid rho 1. mu 1. nu 0.2 xi 0.02

[boundaries]
6
1 1 ordinary
2 2 ordinary
3 3 ordinary
4 4 ordinary
5 5 ordinary
6 6 ordinary

[regions]
1

1 be
6 1 2 3 4 5 6
material 1
0
0

[export]
real_format = eng_double

[conditions over be boundaries]
boundary 1: 0 %(z)s
            0 %(z)s
            0 %(z)s
boundary 2: 1 %(one)s
            1 %(z)s
            1 %(z)s
boundary 3: 1 %(z)s
            0 %(z)s
            1 %(z)s
boundary 4: 1 %(z)s
            0 %(z)s
            1 %(z)s
boundary 5: 1 %(z)s
            1 %(z)s
            0 %(z)s
boundary 6: 1 %(z)s
            1 %(z)s
            0 %(z)s
"""
FLUID_DAT = """[problem]
type = mechanics
analysis = harmonic
n = 3D

[frequencies]
Hz
list
2
20.
45.

[settings]
mesh_file_mode = 2 "cube.msh"

[boundaries]
6
1 1 ordinary
2 2 ordinary
3 3 ordinary
4 4 ordinary
5 5 ordinary
6 6 ordinary

[materials]
1
1 fluid c 343. rho 1.25

[regions]
1
1 be
6 1 2 3 4 5 6
material 1
0
0

[conditions over be boundaries]
boundary 1: 0 (0.,0.)
boundary 2: 0 (1.,0.)
boundary 3: 1 (0.,0.)
boundary 4: 1 (0.,0.)
"""


def _write_case(tmp_path, text, et=shape.QUAD9, m=2):
    write_gmsh22(cube_mesh(m, et), str(tmp_path / "cube.msh"))
    p = tmp_path / "case.dat"
    p.write_text(text)
    return str(p)


class OracleSolver:
    """The driver's solver interface backed by the CPU oracle (tests only)."""

    def __init__(self, case, model):
        from oracle import oracle as orc
        self.orc, self.case, self.model = orc, case, model
        self.o = orc.PotOracle(model) if case.region_type == 1 else orc.Oracle(model)

    def set_incident(self, arrays):
        self.o.set_incident(*arrays.get(0, (None, None)))

    def harmonic(self, omega):
        A, b, _ = self.o.assemble(omega, self.case.material)
        return np.linalg.solve(A, b)

    def static(self):
        A, b, _ = self.o.assemble_static(self.case.material)
        return np.linalg.solve(A, b).astype(np.complex128)

    def interior_static(self, x, points):
        """u and sigma at interior points from oracle pair integrals (Somigliana's identity and its hypersingular counterpart)."""
        u, t = self.model.nodal_solution(np.asarray(x))
        uu = np.zeros((len(points), 3)); sg = np.zeros((len(points), 3, 3))
        for ip, xp in enumerate(points):
            for e in range(self.model.n_elem):
                nodes = self.model.mesh.conn[e]
                h, g, _ = self.o.pair_static(e, xp, self.case.material)
                uu[ip] += np.einsum("jlk,jk->l", g, t[nodes].real) - np.einsum("jlk,jk->l", h, u[nodes].real)
                for kc in range(3):
                    n_i = np.zeros(3); n_i[kc] = 1.0
                    m_, l_, _ = self.o.pair_hbie_static(e, xp, n_i, self.case.material)
                    sg[ip, :, kc] += np.einsum("jlk,jk->l", l_, t[nodes].real) - np.einsum("jlk,jk->l", m_, u[nodes].real)
        return uu, sg

    def close(self):
        pass


def _run_with_oracle(path, **kw):
    case = CaseFile(path)
    return driver.run(path, solver=OracleSolver(case, case.build_model()), log=io.StringIO(), **kw), case


def test_static_case_to_nso(tmp_path):
    path = _write_case(tmp_path, SOLID_DAT % dict(analysis="static", freq="", z="0.", one="1."))
    nso, case = _run_with_oracle(path)
    assert nso == path + ".nso" and case.analysis == "static"
    rows = read_nso(nso)
    md = case.build_model()
    assert rows.shape == (md.n_node, 12 + 6)
    assert (rows[:, 0] == 0).all() and (rows[:, 1] == 0).all() and (rows[:, 2:5] == [1, 1, 2]).all() and (rows[:, 6:8] == 1).all()
    # exact solution of the column: u1 = P x1 / (lambda + 2 mu)
    mat = case.material
    lam2mu = 2.0 * mat.mu_r * mat.nu_r / (1.0 - 2.0 * mat.nu_r) + 2.0 * mat.mu_r      # xi only enters the harmonic analysis
    assert np.abs(rows[:, 12] - rows[:, 9] / lam2mu).max() < 5e-6                      # qsi_relative_error = 1e-6 quadrature
    # every node of the mesh appears once, with its Gmsh id and coordinates
    assert sorted(rows[:, 8].astype(int)) == list(range(1, md.n_node + 1))
    assert np.abs(rows[:, 9:12] - md.node_x[rows[:, 8].astype(int) - 1]).max() < 1e-15
    head = [s for s in open(nso) if s.startswith("#")]
    assert head[0] == "# Program      : multifebe\n" and head[2] == "# File_format  : nso\n" and "# C1-C2    Step index and value.\n" in head


def test_harmonic_fluid_case_to_nso(tmp_path):
    path = _write_case(tmp_path, FLUID_DAT)
    nso, case = _run_with_oracle(path)
    rows = read_nso(nso)
    md = case.build_model()
    assert case.region_type == 1 and rows.shape == (2 * md.n_node, 12 + 4 + 4)
    assert (rows[:md.n_node, 0] == 1).all() and (rows[md.n_node:, 0] == 2).all()
    assert np.allclose(rows[:md.n_node, 1], 20.0, rtol=1e-8) and np.allclose(rows[md.n_node:, 1], 45.0, rtol=1e-8)   # printed in Hz
    assert (rows[:, 4] == 1).all()                                                                                # region type 1 = fluid
    for kf, f in enumerate((20.0, 45.0)):
        r = rows[kf * md.n_node:(kf + 1) * md.n_node]
        p = r[:, 12] + 1j * r[:, 13]
        p_ex, _ = room_analytic(r[:, 9], 2 * np.pi * f, Fluid(1.25, 343.0))
        assert np.abs(p - p_ex).max() < 5e-4            # default en18.8e2 keeps 9 significant digits; discretisation error 2e-4 (2 x 2 quad9)
    assert (rows[:, 16:] == 0).all()                    # no incident field
    # defaults: boundaries 5 and 6 are not listed -> Un = 0 prescribed
    assert case.bcs[5] == ([1], [0j]) and case.bcs[6] == ([1], [0j])
    hdr = [s for s in open(nso) if s.startswith("#_")]
    assert len(hdr) == 1 and hdr[0].rstrip("\n").endswith("C44") and "C1-C2    Frequency index and value f (Hz)." in open(nso).read()
    w = len(rows) and len(open(nso).read().splitlines()[-1])
    assert w == 9 * 0 + 8 * int_width(2, 1, 6, md.n_elem, md.n_node) + (4 + 8) * 18     # 8 integer columns, 4 + 2*(2+2) real columns


def test_harmonic_solid_case_polar_notation(tmp_path):
    freq = "\n[frequencies]\nrad/s\nlin\n2\n0.5\n3.0\n"
    text = (SOLID_DAT % dict(analysis="harmonic", freq=freq, z="(0.,0.)", one="(1.,0.)")).replace("real_format = eng_double", "real_format = sci_double\ncomplex_notation = polar")
    path = _write_case(tmp_path, text, et=shape.QUAD8, m=1)
    nso, case = _run_with_oracle(path, output=str(tmp_path / "out"))
    assert nso == str(tmp_path / "out.nso")
    rows = read_nso(nso)
    md = case.build_model()
    assert rows.shape == (2 * md.n_node, 12 + 12 + 12)
    from multifebe_b200.host import column_analytic_u
    r = rows[md.n_node:]
    u1 = r[:, 12] * np.exp(1j * r[:, 13])
    assert np.abs(u1 - column_analytic_u(r[:, 9], 3.0, case.material)).max() < 2e-2 * np.abs(u1).max()   # one quad8 per face: 1.2 % discretisation error
    assert (r[:, 12] >= 0).all() and (np.abs(r[:, 13]) <= np.pi).all()


def test_unsupported_features_are_named(tmp_path):
    base = SOLID_DAT % dict(analysis="static", freq="", z="0.", one="1.")
    for old, new, word in [("1 1 ordinary", "1 1 crack-like", "ordinary"), ("[regions]\n1\n", "[regions]\n2\n", "regions announced"),
                           ("boundary 2: 1 1.", "boundary 2: 5 1.", "condition type 5"), ("n = 3D", "n = 2D", "3D"),
                           ("1 be\n", "1 fe\n", "`be`"), ('mesh_file_mode = 2 "cube.msh"', "mesh_file_mode = 0", "[nodes]"), ('mesh_file_mode = 2 "cube.msh"', "mesh_file_mode = 3", "wrong type of mesh mode"),
                           ("6 1 2 3 4 5 6", "6 1 2 3 4 5 -6", "reversed")]:
        assert old in base
        path = _write_case(tmp_path, base.replace(old, new, 1))
        with pytest.raises(CaseFileError) as ei:
            CaseFile(path)
        assert word in str(ei.value), (word, str(ei.value))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_reference_tutorial_case_files_parse():
    c = CaseFile(os.path.join(REF, "ME-TH-EL-001/case_files/t3.dat"))
    md = c.build_model()
    assert (c.analysis, len(c.omega), c.region_type, md.n_node, md.n_elem, md.n_dof) == ("harmonic", 300, 2, 462, 744, 1386)
    assert c.omega[0] == 0.01 and c.omega[-1] == 15.0 and c.material.nu_r == 0.2 and c.material.xi == 0.02
    assert c.bcs[4] == ([0, 0, 0], [0j, 0j, 0j]) and c.bcs[2] == ([1, 1, 1], [1 + 0j, 0j, 0j])
    c = CaseFile(os.path.join(REF, "ME-ST-EL-002/case_files/t2.dat"))
    md = c.build_model()
    assert (c.analysis, c.region_type, md.n_dof) == ("static", 2, 3 * md.n_node) and abs(c.material.mu_r - 0.4) < 1e-15 and c.material.nu_r == 0.25


def _gloo_worker(rank, world, port, path, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = CaseFile(path)
    nso = driver.run(path, output=path + ".w2", solver=OracleSolver(case, case.build_model()), rank=rank, world=world, dist=dist, log=io.StringIO())
    out[rank] = nso
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_driver_writes_the_same_file(tmp_path):
    """Frequency shard over two ranks (gloo): rank 0 writes the frequencies in order; the rows equal the single-rank file's."""
    import torch.multiprocessing as mp
    text = FLUID_DAT.replace("list\n2\n20.\n45.\n", "list\n3\n20.\n45.\n70.\n")
    path = _write_case(tmp_path, text, et=shape.QUAD4, m=2)
    nso1, _ = _run_with_oracle(path)
    port = 31500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gloo_worker, args=(2, port, path, out), nprocs=2, join=True)
        assert out[0] == path + ".w2.nso" and out[1] is None
    a = [s for s in open(nso1) if s.startswith("#") and not s.startswith("# Timestamp")]
    b = [s for s in open(path + ".w2.nso") if s.startswith("#") and not s.startswith("# Timestamp")]
    assert a == b
    # the oracle scatters under OpenMP in a nondeterministic order (like the reference): the last printed digit may differ
    ra, rb = read_nso(nso1), read_nso(path + ".w2.nso")
    assert ra.shape == rb.shape and np.array_equal(ra[:, :12], rb[:, :12]) and np.allclose(ra, rb, rtol=1e-7, atol=1e-12)


def test_two_rank_driver_with_an_incident_wave(tmp_path):
    """The same with an [incident waves] field: every rank sets the arrays of the frequency it is about to solve (they depend on omega), and the writer
    rank prints node()%incident_c of every frequency, also of those another rank solved."""
    import torch.multiprocessing as mp
    text = FLUID_DAT.replace("list\n2\n20.\n45.\n", "list\n3\n20.\n45.\n70.\n")
    text = text.replace("0\n0\n\n[conditions", "0\n1 2\n\n[incident waves]\n1\n2\nplane\nfull-space\n0 (1.,0.5) 0. 0. 0. 40. 10.\n0. 0. 0. 0. 0. 0.\nfluid p\n\n[conditions")
    assert "[incident waves]" in text
    path = _write_case(tmp_path, text, et=shape.QUAD4, m=2)
    nso1, case = _run_with_oracle(path)
    port = 33500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gloo_worker, args=(2, port, path, out), nprocs=2, join=True)
    ra, rb = read_nso(nso1), read_nso(path + ".w2.nso")
    assert ra.shape == rb.shape and np.array_equal(ra[:, :12], rb[:, :12]) and np.allclose(ra, rb, rtol=1e-7, atol=1e-12)
    inc_cols = ra[:, 16:20]
    assert all(np.abs(inc_cols[ra[:, 0] == kf]).max() > 0.5 for kf in (1, 2, 3))                 # the incident pressure (amplitude |1 + 0.5i|) at every frequency
    assert not np.allclose(inc_cols[ra[:, 0] == 1], inc_cols[ra[:, 0] == 2])                     # and it changes with the frequency


def test_default_solver_is_the_gpu_and_fails_loudly_without_one(tmp_path):
    """No CPU fallback: without a CUDA device the driver stops in mfb_init (the CPU suite runs on a box without a GPU)."""
    from multifebe_b200 import capi
    path = _write_case(tmp_path, FLUID_DAT, et=shape.QUAD4, m=1)
    with pytest.raises(capi.MfbError) as e:
        driver.run(path, log=io.StringIO())
    assert e.value.code == -2


TWO_REGION_DAT = """[problem]
n = 3D
type = mechanics
analysis = harmonic

[frequencies]
rad/s
list
1
2.5

[settings]
mesh_file_mode = 2 "boxes.msh"

[materials]
2
1 elastic_solid rho 1. mu 1. nu 0.25 xi 0.02
2 fluid rho 1. c 1.2 xi 0.01

[boundaries]
11
1 1 ordinary
2 2 ordinary
3 3 ordinary
4 4 ordinary
5 5 ordinary
6 6 ordinary
7 7 ordinary
13 13 ordinary
14 14 ordinary
15 15 ordinary
16 16 ordinary

[regions]
2

1 be
6 1 3 4 5 6 7
material 1
0
0

2 be
6 -7 2 13 14 15 16
material 2
0
0

[export]
real_format = sci_double

[conditions over be boundaries]
boundary 1: 0 (0.,0.)
            0 (0.,0.)
            0 (0.,0.)
boundary 3: 1 (0.,0.)
            0 (0.,0.)
            1 (0.,0.)
boundary 4: 1 (0.,0.)
            0 (0.,0.)
            1 (0.,0.)
boundary 5: 1 (0.,0.)
            1 (0.,0.)
            0 (0.,0.)
boundary 6: 1 (0.,0.)
            1 (0.,0.)
            0 (0.,0.)
boundary 2: 0 (1.,0.)
"""


def test_two_region_case_parses_numbers_and_exports(tmp_path):
    """A solid and a fluid region sharing boundary 7 (listed as -7 by the second region): the case file gives the MultiRegionModel of
    tests/test_oracle_multiregion.py; the *.nso file lists the interface nodes twice, face 1 (solid side: u, t = -p n) and face 2 (fluid
    side: p, Un = u.n).  The GPU solver refuses coupled regions by name."""
    from multifebe_b200.host import two_box_mesh, MultiRegionModel
    from oracle.multiregion import MultiRegionOracle
    write_gmsh22(two_box_mesh(1, shape.QUAD9), str(tmp_path / "boxes.msh"))
    path = str(tmp_path / "two.dat")
    open(path, "w").write(TWO_REGION_DAT)
    case = CaseFile(path)
    md = case.build_model()
    assert case.multi and case.interfaces == [7] and isinstance(md, MultiRegionModel) and [r[1] for r in case.regions] == [2, 1]
    assert 7 not in case.bcs and case.bcs[13] == ([1], [0j]) and case.bcs[2] == ([0], [1 + 0j])

    class Solver:
        def harmonic(self, omega):
            A, b = MultiRegionOracle(md).assemble(omega)
            return np.linalg.solve(A, b)

        def close(self):
            pass
    nso = driver.run(path, solver=Solver(), log=io.StringIO())
    n1 = sum(len(set(int(v) for e in md.elems_of_boundary[abs(b)] for v in md.mesh.conn[e])) for b in case.regions[0][3])
    n2 = sum(len(set(int(v) for e in md.elems_of_boundary[abs(b)] for v in md.mesh.conn[e])) for b in case.regions[1][3])
    lines = [s for s in open(nso) if s.strip() and not s.startswith("#")]
    assert len(lines) == n1 + n2
    solid, fluid = rows_of(lines, 2), rows_of(lines, 1)
    assert len(solid) == n1 and solid.shape[1] == 12 + 24 and len(fluid) == n2 and fluid.shape[1] == 12 + 8
    # interface rows: boundary 7, face 1 in the solid region and face 2 in the fluid region, same nodes; sigma_xx = -p, u_x = Un2 * (-1)
    s7 = solid[solid[:, 5] == 7]; f7 = fluid[fluid[:, 5] == 7]
    assert (s7[:, 7] == 1).all() and (f7[:, 7] == 2).all() and np.array_equal(s7[:, 8], f7[:, 8])
    t1 = s7[:, 18] + 1j * s7[:, 19]; p = f7[:, 12] + 1j * f7[:, 13]
    assert np.abs(t1 + p).max() < 1e-12 * np.abs(p).max()
    u1 = s7[:, 12] + 1j * s7[:, 13]; un2 = f7[:, 14] + 1j * f7[:, 15]
    assert np.abs(u1 + un2).max() < 1e-12 * np.abs(u1).max()


def rows_of(lines, rtype):
    return np.array([[float(t) for t in s.split()] for s in lines if int(s.split()[4]) == rtype])


def test_internal_points_of_an_elastic_region(tmp_path):
    """[internal points] of the case file -> rows with boundary columns 0 0 0, the point id, u_k and the tractions on the three coordinate
    planes; ME-ST-EL-002's exact field (u1 = x1/(lambda+2mu), sigma_11 = 1, sigma_22 = sigma_33 = nu/(1-nu)) at the points."""
    text = (SOLID_DAT % dict(analysis="static", freq="", z="0.", one="1.")).replace("eng_double", "sci_double")
    text += "\n[internal points]\n3\n1 1 0.5 0.5 0.5\n2 1 0.2 0.7 0.4\n7 1 0.8 0.3 0.6\n"
    path = _write_case(tmp_path, text, et=shape.QUAD4, m=3)
    case = CaseFile(path)
    md = case.build_model()
    assert [p[0] for p in case.internal_points] == [1, 2, 7]

    nso = driver.run(path, solver=OracleSolver(case, md), log=io.StringIO())
    lines = [s for s in open(nso) if s.strip() and not s.startswith("#")]
    ip_rows = np.array([[float(t) for t in s.split()] for s in lines[md.n_node:]])
    assert len(lines) == md.n_node + 3 and ip_rows.shape == (3, 12 + 3 + 9)
    assert (ip_rows[:, 5:8] == 0).all() and list(ip_rows[:, 8]) == [1, 2, 7] and np.allclose(ip_rows[:, 9:12], [[0.5, 0.5, 0.5], [0.2, 0.7, 0.4], [0.8, 0.3, 0.6]])
    mat = case.material
    lam2mu = 2.0 * mat.mu_r * mat.nu_r / (1.0 - 2.0 * mat.nu_r) + 2.0 * mat.mu_r
    assert np.abs(ip_rows[:, 12] - ip_rows[:, 9] / lam2mu).max() < 1e-4
    sig = ip_rows[:, 15:24].reshape(3, 3, 3)                     # [point][plane kc][component k]
    exact = np.diag([1.0, mat.nu_r / (1.0 - mat.nu_r), mat.nu_r / (1.0 - mat.nu_r)])
    assert np.abs(sig - exact).max() < 5e-4
    # a fluid region has no internal-point support
    with pytest.raises(CaseFileError):
        CaseFile(_write_case(tmp_path, FLUID_DAT + "\n[internal points]\n1\n1 1 0.5 0.5 0.5\n"))


def test_shipped_examples_parse_and_the_static_one_solves():
    ex = os.path.join(ROOT, "examples")
    sizes = {}
    for name in ("room", "column_harmonic", "column_static"):
        c = CaseFile(os.path.join(ex, name, name + ".dat"))
        sizes[name] = (c.analysis, c.region_type, c.build_model().n_dof, len(c.omega), len(c.internal_points))
    assert sizes == {"room": ("harmonic", 1, 486, 12, 0), "column_harmonic": ("harmonic", 2, 1458, 8, 3), "column_static": ("static", 2, 1170, 0, 3)}
    c = CaseFile(os.path.join(ex, "inclusion_p_wave", "inclusion_p_wave.dat"))
    md = c.build_model()
    assert (c.multi, md.n_dof, c.region_incident, c.formulation[6], c.incident_fields[1]["wave"]) == (True, 2916, [[], [1]], ("sbie_boundary_mca", 0.05), "p")
    arr = c.incident_arrays(md, 2.0)
    assert list(arr) == [1] and arr[1][0].shape == (96 * 9, 3) and abs(np.linalg.norm(arr[1][0], axis=1).max() - 0.5) < 0.02      # |u| = 1/2: the reference halves the elastic field
    c = CaseFile(os.path.join(ex, "room", "room.dat"))
    assert abs(c.omega[0] / (2 * np.pi) - 5.0) < 1e-12 and abs(c.omega[-1] / (2 * np.pi) - 115.0) < 1e-9 and c.description.startswith("pressure waves")


PORO_DAT = """[problem]
n = 3D
type = mechanics
analysis = harmonic

[frequencies]
rad/s
list
1
2.0

[settings]
mesh_file_mode = 2 "cube.msh"

[materials]
1
1 biot_poroelastic_medium phi 0.35 lambda 1.2 mu 1.0 Q 0.5 R 0.8 rho_f 1.0 rho_s 2.2 rho_a 0.15 xi 0.02 b 0.4

[boundaries]
6
1 1 ordinary
2 2 ordinary
3 3 ordinary
4 4 ordinary
5 5 ordinary
6 6 ordinary

[regions]
1
1 be
6 1 2 3 4 5 6
material 1
0
0

[export]
real_format = sci_double

[conditions over be boundaries]
boundary 1: 1 (0.,0.)
            0 (0.,0.)
            0 (0.,0.)
            0 (0.,0.)
boundary 2: 0 (0.,0.)
            1 (1.,0.)
            1 (0.,0.)
            1 (0.,0.)
boundary 3: 1 (0.,0.)
            1 (0.,0.)
            0 (0.,0.)
            1 (0.,0.)
boundary 4: 1 (0.,0.)
            1 (0.,0.)
            0 (0.,0.)
            1 (0.,0.)
boundary 5: 1 (0.,0.)
            1 (0.,0.)
            1 (0.,0.)
            0 (0.,0.)
boundary 6: 1 (0.,0.)
            1 (0.,0.)
            1 (0.,0.)
            0 (0.,0.)
"""


def test_poroelastic_case_to_nso(tmp_path):
    """A saturated column from its case file (biot_poroelastic_medium, four conditions per boundary): region type 3, rows with tau, u_k, Un, t_k;
    the oracle-solved field follows the exact Biot solution of tests/test_oracle_poroelastic.py."""
    from test_oracle_poroelastic import biot_column
    path = _write_case(tmp_path, PORO_DAT, et=shape.QUAD9, m=2)
    case = CaseFile(path)
    md = case.build_model()
    assert case.region_type == 3 and md.ndof == 4 and md.n_dof == 4 * md.n_node
    po = case.material
    assert (po.phi, po.rho1, po.rho2, po.rhoa, po.b) == (0.35, (1 - 0.35) * 2.2, 0.35 * 1.0, 0.15, 0.4) and abs(po.lam - 1.2 * (1 + 0.04j)) < 1e-15

    class Solver:
        def harmonic(self, omega):
            from oracle import oracle as orc
            A, b, _ = orc.PorOracle(md).assemble(omega, po)
            return np.linalg.solve(A, b)

        def close(self):
            pass
    nso = driver.run(path, solver=Solver(), log=io.StringIO())
    rows = read_nso(nso)
    assert rows.shape == (md.n_node, 12 + 16 + 16) and (rows[:, 4] == 3).all()
    field, _ = biot_column(2.0, po)
    ua, Ua, sa, ta = field(rows[:, 9])
    u1 = rows[:, 14] + 1j * rows[:, 15]
    assert np.abs(u1 - ua).max() < 4e-3 * np.abs(ua).max()
    side = rows[:, 5] >= 3
    tau = rows[:, 12] + 1j * rows[:, 13]
    assert np.abs(tau[side] - ta[side]).max() < 4e-3 * np.abs(ta).max()


# ---- [symmetry planes] (src/read_symmetry_planes.f90): the P-wave column modelled as a quarter, its lateral conditions on y = 0 and z = 0 replaced by planes ----
SYM_DAT = """[problem]
n = 3D
type = mechanics
analysis = %(analysis)s
%(freq)s
[settings]
mesh_file_mode = 2 "quarter.msh"

[materials]
1
1 elastic_solid rho 1. mu 1. nu 0.2 xi 0.02

[boundaries]
4
1 1 ordinary
2 2 ordinary
4 4 ordinary
6 6 ordinary

[regions]
1
1 be
4 1 2 4 6
material 1
0
0

[symmetry planes]
%(planes)s

[export]
real_format = eng_double

[conditions over be boundaries]
boundary 1: 0 %(z)s
            0 %(z)s
            0 %(z)s
boundary 2: 1 %(one)s
            1 %(z)s
            1 %(z)s
boundary 4: 1 %(z)s
            0 %(z)s
            1 %(z)s
boundary 6: 1 %(z)s
            1 %(z)s
            0 %(z)s
"""


def _write_quarter(tmp_path, text, et=shape.QUAD9, m=2):
    from multifebe_b200.host import without_parts
    write_gmsh22(without_parts(cube_mesh(m, et), {3, 5}), str(tmp_path / "quarter.msh"))
    p = tmp_path / "case.dat"
    p.write_text(text)
    return str(p)


def test_symmetry_planes_section_static_column(tmp_path):
    path = _write_quarter(tmp_path, SYM_DAT % dict(analysis="static", freq="", z="0.", one="1.", planes="plane_n2: symmetry\nplane_xy : symmetry"))
    nso, case = _run_with_oracle(path)
    assert case.symmetry == [("y", "symmetry"), ("z", "symmetry")]
    md = case.build_model()
    assert list(md.symplane_eid) == [2, 3] and np.array_equal(md.symplane_t, [[1, -1, 1], [1, 1, -1]])
    rows = read_nso(nso)
    assert rows.shape == (md.n_node, 12 + 6)
    mat = case.material
    lam2mu = 2.0 * mat.mu_r * mat.nu_r / (1.0 - 2.0 * mat.nu_r) + 2.0 * mat.mu_r
    assert np.abs(rows[:, 12] - rows[:, 9] / lam2mu).max() < 5e-6          # u1 = P x1 / (lambda + 2 mu), as the full column of test_static_case_to_nso
    assert np.abs(rows[:, 13:15]).max() < 5e-6                             # no lateral displacement anywhere: the planes hold the column


def test_symmetry_planes_section_forms_and_errors(tmp_path):
    kw = dict(analysis="harmonic", freq="\n[frequencies]\nrad/s\nlist\n1\n2.0\n", z="(0.,0.)", one="(1.,0.)")
    a = CaseFile(_write_quarter(tmp_path, SYM_DAT % dict(planes="plane_zx: symmetry\nplane_n3: antisymmetry", **kw)))
    assert a.symmetry == [("y", "symmetry"), ("z", "antisymmetry")]
    # the explicit form: scalar multiplier, then the three translation multipliers
    b = CaseFile(_write_quarter(tmp_path, SYM_DAT % dict(planes="y = 1 1 -1 1\nz = -1 -1 -1 1", **kw)))
    ma, mb = a.build_model(), b.build_model()
    assert np.array_equal(ma.symplane_eid, mb.symplane_eid) and np.array_equal(ma.symplane_t, mb.symplane_t)
    for planes, word in [("plane_n2: mirror", "symmetry or antisymmetry"), ("y = 1 1 2 1", "+1 or -1"), ("plane_n2: symmetry\ny = 1 1 -1 1", "twice")]:
        with pytest.raises(CaseFileError) as ei:
            CaseFile(_write_quarter(tmp_path, SYM_DAT % dict(planes=planes, **kw)))
        assert word in str(ei.value)
    # a mesh on both sides of a plane is refused (fbem_check_nodes_symplanes_configuration)
    c = CaseFile(_write_quarter(tmp_path, SYM_DAT % dict(planes="plane_n1: symmetry", **kw)))
    c.mesh.nodes[:, 0] -= 0.5
    with pytest.raises(ValueError):
        c.build_model()


def test_symmetry_planes_on_a_fluid_region(tmp_path):
    """The acoustic room of ME-TH-AC-001 as a quarter model: its rigid walls y = 0, z = 0 replaced by [symmetry planes] (scalar multiplier symplane_s = +1)."""
    from multifebe_b200.host import without_parts
    write_gmsh22(without_parts(cube_mesh(2, shape.QUAD9), {3, 5}), str(tmp_path / "quarter.msh"))
    text = """[problem]
type = mechanics
analysis = harmonic
n = 3D

[frequencies]
Hz
list
1
45.

[settings]
mesh_file_mode = 2 "quarter.msh"

[boundaries]
4
1 1 ordinary
2 2 ordinary
4 4 ordinary
6 6 ordinary

[materials]
1
1 fluid c 343. rho 1.25

[regions]
1
1 be
4 1 2 4 6
material 1
0
0

[symmetry planes]
plane_n2: symmetry
z = 1 1 1 -1

[conditions over be boundaries]
boundary 1: 0 (0.,0.)
boundary 2: 0 (1.,0.)
boundary 4: 1 (0.,0.)
boundary 6: 1 (0.,0.)
"""
    p = tmp_path / "case.dat"; p.write_text(text)
    nso, case = _run_with_oracle(str(p))
    md = case.build_model()
    assert case.region_type == 1 and list(md.symplane_eid) == [2, 3] and list(md.symplane_s) == [1.0, 1.0]
    rows = read_nso(nso)
    pr = rows[:, 12] + 1j * rows[:, 13]
    p_ex, _ = room_analytic(rows[:, 9], 2 * np.pi * 45.0, Fluid(1.25, 343.0))
    assert np.abs(pr - p_ex).max() < 5e-4


def test_local_axes_rotation_field_and_pressure_conditions_in_a_case_file(tmp_path):
    """Condition types 2 / 3 (local axes), 4 (infinitesimal rotation field -> a ctype-0 condition with values per node) and the refusal of mixed local /
    global types (read_conditions_bem_boundaries_mechanics_harmonic.f90:124-162)."""
    base = SOLID_DAT % dict(analysis="static", freq="", z="0.", one="1.")
    walls = ("boundary 3: 1 0.\n            0 0.\n            1 0.\n", "boundary 3: 2 0.\n            3 0.\n            3 0.\n")
    assert walls[0] in base
    path = _write_case(tmp_path, base.replace(walls[0], walls[1]))
    case = CaseFile(path); md = case.build_model()
    assert case.bcs[3] == ([2, 3, 3], [0j, 0j, 0j])
    w3 = md.node_part == 3
    assert (md.ctype[w3] == [2, 3, 3]).all() and (md.row_bc[w3] >= 0).all() and md.n_dof == 3 * md.n_node + 3 * w3.sum()
    # the sliding wall in local axes is the sliding wall of the global pattern: same exact column solution
    nso, _ = _run_with_oracle_local(path)
    rows = read_nso(nso)
    mat = case.material
    lam2mu = 2.0 * mat.mu_r * mat.nu_r / (1.0 - 2.0 * mat.nu_r) + 2.0 * mat.mu_r
    assert np.abs(rows[:, 12] - rows[:, 9] / lam2mu).max() < 5e-6
    with pytest.raises(CaseFileError) as ei:
        CaseFile(_write_case(tmp_path, base.replace(walls[0], "boundary 3: 2 0.\n            0 0.\n            3 0.\n")))
    assert "can not be mixed" in str(ei.value)
    # rotation field about the z axis through the centre of the face x = 0
    rot = "boundary 1: 4 0. 0.5 0.5  0. 0. 2.  0.01\n            4 0. 0.5 0.5  0. 0. 2.  0.01\n            4 0. 0.5 0.5  0. 0. 2.  0.01\n"
    old1 = "boundary 1: 0 0.\n            0 0.\n            0 0.\n"
    assert old1 in base
    md = CaseFile(_write_case(tmp_path, base.replace(old1, rot))).build_model()
    f1 = md.node_part == 1
    assert (md.ctype[f1] == 0).all()
    want = 0.01 * np.cross([0.0, 0.0, 1.0], md.node_x[f1] - np.array([0.0, 0.5, 0.5]))
    assert np.abs(md.cvalue[f1] - want).max() < 1e-15


def _run_with_oracle_local(path):
    """The oracle as the solver of a single-region case with local-axes rows (the host adds them, as the reference's build_lse_mechanics_* does)."""
    case = CaseFile(path)

    class S(OracleSolver):
        def static(self):
            A, b, _ = self.o.assemble_static(self.case.material)
            A = A.astype(np.complex128); b = b.astype(np.complex128)
            self.model.add_condition_rows(A, b)
            return np.linalg.solve(A, b)
    return driver.run(path, solver=S(case, case.build_model()), log=io.StringIO()), case


def _native_sections(mesh):
    names = {shape.TRI3: "tri3", shape.TRI6: "tri6", shape.QUAD4: "quad4", shape.QUAD8: "quad8", shape.QUAD9: "quad9"}
    out = ["[nodes]", str(len(mesh.nodes))] + ["%d %.17g %.17g %.17g" % (k + 1, x[0], x[1], x[2]) for k, x in enumerate(mesh.nodes)]
    out += ["", "[elements]", str(mesh.n_elem)]
    out += ["%d %s 1 %d %s" % (k + 1, names[int(mesh.etype[k])], mesh.part[k], " ".join(str(int(v) + 1) for v in mesh.conn[k])) for k in range(mesh.n_elem)]
    parts = sorted(set(int(p) for p in mesh.part))
    out += ["", "[parts]", str(len(parts))] + ["%d face%d" % (p, p) for p in parts]
    return "\n".join(out) + "\n"


def test_mesh_inside_the_case_file_and_in_a_native_file(tmp_path):
    """mesh_file_mode 0 / absent ([nodes], [elements], [parts] of the case file) and 1 (the same sections in an auxiliary file) give the model of the Gmsh file."""
    base = SOLID_DAT % dict(analysis="static", freq="", z="0.", one="1.")
    path = _write_case(tmp_path, base, et=shape.QUAD8, m=2)
    ref = CaseFile(path).build_model()
    mesh = cube_mesh(2, shape.QUAD8)
    p0 = str(tmp_path / "inline.dat")
    open(p0, "w").write(base.replace('mesh_file_mode = 2 "cube.msh"', "") + "\n" + _native_sections(mesh))
    (tmp_path / "cube.native").write_text(_native_sections(mesh))
    p1 = str(tmp_path / "native.dat")
    open(p1, "w").write(base.replace('mesh_file_mode = 2 "cube.msh"', 'mesh_file_mode = 1 "cube.native"'))
    for p, mode in ((p0, 0), (p1, 1)):
        c = CaseFile(p)
        md = c.build_model()
        assert c.mesh_file_mode == mode and np.array_equal(md.node_x, ref.node_x) and np.array_equal(md.elem_node, ref.elem_node)
        assert np.array_equal(md.row, ref.row) and np.array_equal(md.colloc_x, ref.colloc_x) and np.array_equal(md.mesh.node_ids, ref.mesh.node_ids)
    # a part that no boundary uses is dropped with its nodes, as the reference drops it
    extra = _native_sections(mesh).replace("[elements]\n%d\n" % mesh.n_elem, "[elements]\n%d\n%d quad4 1 99 1 2 3 4\n" % (mesh.n_elem + 1, mesh.n_elem + 1))
    open(p0, "w").write(base.replace('mesh_file_mode = 2 "cube.msh"', "") + "\n" + extra)
    assert CaseFile(p0).build_model().n_elem == ref.n_elem
    open(p0, "w").write(base.replace('mesh_file_mode = 2 "cube.msh"', ""))
    with pytest.raises(CaseFileError) as ei:
        CaseFile(p0)
    assert "[nodes]" in str(ei.value)


def test_bem_formulation_section_and_selected_export_nodes(tmp_path):
    """[bem formulation over boundaries]: the MCA displacement per boundary; the exact column solution survives every choice.  [export] nso_nodes: rows of those nodes only."""
    from multifebe_b200.host.model import MCA_BOUNDARY_DELTA
    base = (SOLID_DAT % dict(analysis="static", freq="", z="0.", one="1.")).replace("eng_double", "sci_double")
    default = CaseFile(_write_case(tmp_path, base, et=shape.QUAD9, m=2)).build_model()
    assert set(np.unique(default.mca_delta)) == {0.0, MCA_BOUNDARY_DELTA}
    text = base + "\n[bem formulation over boundaries]\nboundary 1: sbie_boundary_mca 0.01\nboundary 2: sbie_mca 0.\nboundary 3: sbie_mca 0.3\nboundary 4: sbie_boundary_mca -1.\n"
    text = text.replace("real_format = sci_double", "real_format = sci_double\nnso_nodes = 3 5 1 40")
    path = _write_case(tmp_path, text, et=shape.QUAD9, m=2)
    case = CaseFile(path)
    md = case.build_model()
    assert case.formulation == {1: ("sbie_boundary_mca", 0.01), 2: ("sbie_mca", 0.0), 3: ("sbie_mca", 0.3), 4: ("sbie_boundary_mca", -1.0)} and case.nso_nodes == {1, 5, 40}
    for part, rim, inner in ((1, 0.01, 0.0), (2, -1.0, -1.0), (3, 0.3, 0.3), (4, MCA_BOUNDARY_DELTA, 0.0), (5, MCA_BOUNDARY_DELTA, 0.0)):
        sel = md.node_part == part
        assert set(md.mca_delta[sel & md.in_boundary]) == {rim} and set(md.mca_delta[sel & ~md.in_boundary]) == {inner}
    # sbie_mca: one collocation point per element node of the boundary, moved by the default of the element order (quadratic: 0.2254) or by the given delta
    e2 = [e for e in range(md.n_elem) if int(md.mesh.part[e]) == 2]
    assert sum((md.colloc_elem == e).sum() for e in e2) == 9 * len(e2)
    k = int(np.flatnonzero((md.colloc_elem == e2[0]) & (md.colloc_kn == 0))[0])
    assert np.allclose(md.colloc_xi[k], np.array([-1.0, -1.0]) * (1.0 - 0.22540333))
    assert md.n_colloc > default.n_colloc and md.n_dof == default.n_dof
    nso, _ = _run_with_oracle(path)
    rows = read_nso(nso)
    assert sorted(rows[:, 8]) == [1, 5, 40]
    lam2mu = 2.0 * case.material.mu_r * case.material.nu_r / (1.0 - 2.0 * case.material.nu_r) + 2.0 * case.material.mu_r
    assert np.abs(rows[:, 12] - rows[:, 9] / lam2mu).max() < 2e-4                      # u1 = x1 / (lambda + 2 mu), ME-ST-EL-002's exact field
    for old, new, word in [("boundary 3: sbie_mca 0.3", "boundary 3: sbie", "open rim"), ("boundary 3: sbie_mca 0.3", "boundary 3: hbie 0.2", "not covered"),
                           ("boundary 3: sbie_mca 0.3", "boundary 3: sbie_mca", "needs its delta"), ("nso_nodes = 3 5 1 40", "nso_nodes = 3 5 5 40", "repeated"),
                           ("[bem formulation", "[element options]\nx\n\n[bem formulation", "element options")]:
        assert old in text
        p = _write_case(tmp_path, text.replace(old, new, 1), et=shape.QUAD9, m=2)
        with pytest.raises(CaseFileError) as ei:
            CaseFile(p).build_model()
        assert word in str(ei.value), (word, str(ei.value))
