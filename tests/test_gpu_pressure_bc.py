"""ctype = 10 (normal pressure known) on the GPU path: parity with the oracle (pinned by the hydrostatic solution in tests/test_oracle_pressure_bc.py)
and the exact solution itself, through the C ABI (mfb_set_node_normals) and through the case-file driver."""
import io
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from multifebe_b200.host import Material, Model, cube_mesh, shape, write_gmsh22, without_parts  # noqa: E402
from test_oracle_pressure_bc import octant  # noqa: E402

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_pressure_condition_parity_and_exact_solution(gpu_ctx, oracle_lib):
    from multifebe_b200 import capi
    p = 0.7
    smat = Material(rho=1.0, mu=1.3, nu=0.2, xi=0.0)
    for etype, m in [(shape.QUAD4, 3), (shape.TRI6, 2), (shape.QUAD9, 2)]:
        md = octant(etype, m, p)
        pr = capi.Problem(gpu_ctx, md); o = oracle_lib.Oracle(md)
        A, b = pr.build_lse_mechanics_bem_staela(smat)
        Ao, bo, _ = o.assemble_static(smat)
        assert relerr(A, Ao) < 1e-11 and relerr(b, bo) < 1e-11
        u, _ = md.nodal_solution(np.asarray(pr.solve_static(smat), dtype=np.complex128))
        eps = p * (1.0 - 2.0 * smat.nu_r) / (2.0 * smat.mu_r * (1.0 + smat.nu_r))
        assert np.abs(u.real - eps * md.node_x).max() < 2e-5 * eps
        pr.close()


def test_pressure_on_a_cavity_and_mixed_with_other_conditions(gpu_ctx, oracle_lib):
    """Harmonic; a reversed boundary (pressure inside a cavity of the full space: the sign follows the orientation), parts with other conditions beside it."""
    from multifebe_b200 import capi
    mat = Material(rho=1.0, mu=1.0, nu=0.25, xi=0.02)
    cases = [Model(cube_mesh(3, shape.TRI3), {q: ([10, 10, 10], [0.3 + 0.1j] * 3) for q in range(1, 7)}, reversed_parts=(1, 2, 3, 4, 5, 6)),
             Model(cube_mesh(2, shape.QUAD8), {1: ([0, 0, 0], [0.1, 0, 0]), 2: ([10, 10, 10], [1.0] * 3), 3: ([1, 0, 1], [0, 0, 0]), 4: ([1, 1, 1], [0, 0.2, 0]),
                                               5: ([10, 10, 10], [-0.5j] * 3), 6: ([1, 1, 1], [0, 0, 0])})]
    for md in cases:
        pr = capi.Problem(gpu_ctx, md); o = oracle_lib.Oracle(md)
        for omega in (1.0, 5.0):
            A, b = pr.build_lse_mechanics_bem_harela(omega, mat)
            Ao, bo, _ = o.assemble(omega, mat)
            assert relerr(A, Ao) < 1e-11 and relerr(b, bo) < 1e-11
        xo, _, _ = oracle_lib.lu_solve(Ao, bo)
        assert relerr(pr.solve_frequency(5.0, mat), xo) < 1e-8
        pr.close()


def test_pressure_condition_in_a_case_file(tmp_path):
    """`boundary <id>: 10 <P>` (one record for the three components) + [symmetry planes]: the hydrostatic octant through the stand-alone driver."""
    from multifebe_b200 import driver
    from multifebe_b200.host.casefile import CaseFile
    from multifebe_b200.host.export import read_nso
    write_gmsh22(without_parts(cube_mesh(2, shape.QUAD9), {1, 3, 5}), str(tmp_path / "octant.msh"))
    text = """[problem]
n = 3D
type = mechanics
analysis = static

[settings]
mesh_file_mode = 2 "octant.msh"

[materials]
1
1 elastic_solid rho 1. mu 1.3 nu 0.2

[boundaries]
3
2 2 ordinary
4 4 ordinary
6 6 ordinary

[regions]
1
1 be
3 2 4 6
material 1
0
0

[symmetry planes]
plane_n1: symmetry
plane_n2: symmetry
plane_n3: symmetry

[export]
real_format = sci_double

[conditions over be boundaries]
boundary 2: 10 0.7
boundary 4: 10 0.7
boundary 6: 10 0.7
"""
    path = tmp_path / "case.dat"; path.write_text(text)
    case = CaseFile(str(path))
    assert case.bcs[2] == ([10, 10, 10], [0.7 + 0j] * 3)
    rows = read_nso(driver.run(str(path), log=io.StringIO()))
    eps = 0.7 * (1.0 - 0.4) / (2.0 * 1.3 * 1.2)
    assert np.abs(rows[:, 12:15] - eps * rows[:, 9:12]).max() < 2e-5 * eps       # u = eps x at every node
