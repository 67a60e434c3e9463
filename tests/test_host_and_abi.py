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


# ---- host logic of the single-frequency multi-GPU mode (no device needed) ----
@pytest.mark.parametrize("n,nb,nranks", [(30258, 256, 8), (1386, 256, 2), (1000, 32, 3), (64, 32, 4), (31, 32, 2)])
def test_block_cyclic_layout_covers_every_column_once(n, nb, nranks):
    from multifebe_b200 import capi
    seen = np.zeros(n, dtype=int)
    for r in range(nranks):
        cols = capi.dist_layout(n, nb, nranks, r)
        assert np.all(np.diff(cols) > 0)                       # local order = global order
        assert np.all((cols // nb) % nranks == r)              # block j lives on rank j % P
        seen[cols] += 1
    assert np.all(seen == 1)


def test_tile_partition_keeps_row_blocks_whole_and_balanced():
    from multifebe_b200 import capi
    # 40 row blocks of 32 nodes (96 rows), the last 6 with two layers (MCA rim nodes), then 2 loose tiles
    row0, nbytes = [], []
    for i in range(40):
        for _ in range(2 if i >= 34 else 1):
            row0.append(96 * i); nbytes.append(24 * 32)
    row0 += [0, 0]; nbytes += [0, 0]
    n_dof = 96 * 40 + 30
    for nranks in (1, 2, 3, 8):
        tr, rb = capi.dist_partition_tiles(row0, nbytes, n_dof, nranks)
        assert rb[0] == 0 and rb[-1] == n_dof and np.all(np.diff(rb) >= 0)
        assert np.all(tr[-2:] == nranks - 1)                   # loose tiles -> last rank, whose rows run to n_dof
        for t in range(len(row0) - 2):
            assert rb[tr[t]] <= row0[t] and row0[t] + 96 <= rb[tr[t] + 1]          # a tile's rows belong to its rank
        for t in range(1, len(row0) - 2):
            if row0[t] == row0[t - 1]:
                assert tr[t] == tr[t - 1]                      # layers of one row block stay together
        counts = np.bincount(tr, minlength=nranks)
        assert counts.max() - counts.min() <= 4
    tr, rb = capi.dist_partition_tiles([0, 0], [24 * 4, 0], 20, 4)      # fewer row blocks than ranks
    assert list(tr) == [0, 3] and list(rb) == [0, 12, 12, 12, 20]


def test_plain_c_program_builds_against_the_header_and_runs(tmp_path):
    """include/mfb.h is a C header (not C++ in disguise): a C99 program compiled with gcc -std=c99 -pedantic links against libmfb.so, computes through the
    host-only entry points and sees MFB_ERR_NO_DEVICE from mfb_init on a box without a GPU (tests/native/abi_c_program.c)."""
    import subprocess
    from multifebe_b200 import build as b
    so = b.build()
    exe = str(tmp_path / "abi_c_program")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "native", "abi_c_program.c"),
                           "-o", exe, so, "-lm", "-Wl,-rpath," + os.path.dirname(so)])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
