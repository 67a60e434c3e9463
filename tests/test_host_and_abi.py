"""Host-side logic (mesh generators, DOF numbering, collocation points) and the C-ABI boundary: libmfb.so must load,
export every symbol include/mfb.h declares and refuse to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import numpy as np
import pytest
from multifebe_b200.host import Model, Material, cube_mesh, cube_bcs, halfspace_patch, shape, read_gmsh22, write_gmsh22

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    from multifebe_b200 import build, capi
    build.build()
    L = capi.lib()
    hdr = open(os.path.join(ROOT, "include", "mfb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(mfb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), "libmfb.so does not export %s" % n
    assert L.mfb_version() >= 100


@pytest.mark.skipif(_has_cuda(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_a_device():
    from multifebe_b200 import capi
    with pytest.raises(capi.MfbError) as e:
        capi.Context(0)
    assert e.value.code == -2 and "no CUDA device" in str(e.value)
    L = capi.lib()
    assert L.mfb_harela3d_assemble(None, C.c_double(1.0), None, None, C.c_double(1.0), None, None, None, None) == -1
    assert L.mfb_zsolve(None, 4, None, 4, None, None, 1, 1) == -1
    assert L.mfb_measure_peaks(None, None, None, None) == -1


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "multifebe_b200")
    for dp, dn, fn in os.walk(pkg):
        if "build" in dp.split(os.sep)[-1:]:
            continue
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle/" not in txt.replace("see oracle/", "") or f == "plan_host.h", (f, "mentions oracle/")
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f


@pytest.mark.parametrize("et,nn_face,ne_face", [(shape.TRI3, lambda m: (m + 1) ** 2, lambda m: 2 * m * m), (shape.QUAD4, lambda m: (m + 1) ** 2, lambda m: m * m),
                                               (shape.TRI6, lambda m: (2 * m + 1) ** 2, lambda m: 2 * m * m), (shape.QUAD8, lambda m: (2 * m + 1) ** 2 - m * m, lambda m: m * m),
                                               (shape.QUAD9, lambda m: (2 * m + 1) ** 2, lambda m: m * m)])
def test_cube_mesh_and_numbering(et, nn_face, ne_face):
    m = 3
    mesh = cube_mesh(m, et)
    assert len(mesh.nodes) == 6 * nn_face(m) and mesh.n_elem == 6 * ne_face(m)
    md = Model(mesh, cube_bcs())
    assert md.n_dof == 3 * md.n_node
    # rows: three consecutive per node in first-visit order; one column per unknown; square system
    assert sorted(md.row.ravel().tolist()) == list(range(md.n_dof))
    cols = np.where(md.ctype == 0, md.col_t, md.col_u).ravel()
    assert sorted(cols.tolist()) == list(range(md.n_dof))
    assert np.all(md.row[:, 1] == md.row[:, 0] + 1) and np.all(md.row[:, 2] == md.row[:, 0] + 2)
    # outward normals: element normal . (centroid - cube centre) > 0
    for e in range(0, mesh.n_elem, 7):
        x = mesh.nodes[mesh.conn[e]]
        nv = 3 if et in (shape.TRI3, shape.TRI6) else 4
        n = np.cross(x[1] - x[0], x[nv - 1] - x[0])
        assert n @ (x[:nv].mean(axis=0) - 0.5) > 0
    # collocation: rim nodes of every face are MCA points (one per incident element), interior nodes nodal (once)
    rim = md.in_boundary
    inc = np.zeros(md.n_node, dtype=int)
    for c in mesh.conn:
        inc[c] += 1
    assert md.n_colloc == int(inc[rim].sum() + (~rim).sum())
    mca = md.colloc_xi[:, 0] != -9.0
    assert np.all(rim[md.colloc_node[mca]]) and not np.any(rim[md.colloc_node[~mca]])
    # MCA points lie strictly inside their element, 5 % away from the node (assign_default_bem_formulation.f90:40)
    for c in np.where(mca)[0][:20]:
        e, kn = md.colloc_elem[c], md.colloc_kn[c]
        x = shape.position(et, md.node_x[mesh.conn[e]], md.colloc_xi[c])
        assert np.allclose(x, md.colloc_x[c]) and np.linalg.norm(x - md.node_x[mesh.conn[e][kn]]) > 1e-3


def test_gmsh22_roundtrip(tmp_path):
    mesh = halfspace_patch(4, shape.QUAD9)
    p = str(tmp_path / "m.msh")
    write_gmsh22(mesh, p)
    m2 = read_gmsh22(p)
    assert np.array_equal(mesh.nodes, m2.nodes) and np.array_equal(mesh.etype, m2.etype) and np.array_equal(mesh.part, m2.part)
    assert all(np.array_equal(a, b) for a, b in zip(mesh.conn, m2.conn))
    assert set(mesh.part.tolist()) == {1, 2}


def test_material_follows_read_regions():
    mat = Material(2.0, 3.0, 0.3, 0.05)
    assert mat.mu == 3.0 * (1 + 0.1j) and abs(mat.lam - 2 * mat.mu * 0.3 / 0.4) < 1e-15
    assert abs(mat.c2 ** 2 * 2.0 - mat.mu) < 1e-14 and abs(mat.c1 ** 2 * 2.0 - (mat.lam + 2 * mat.mu)) < 1e-14
