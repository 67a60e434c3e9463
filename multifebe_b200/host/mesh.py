"""Meshes for the harmonic BEM hot path: Gmsh 2.2 reader/writer and deterministic synthetic generators.

Stand-in for the reference's host-side mesh input (src/read_elements.f90, lib/fbem/src/gmsh.f90); the
synthetic generators are the S-cube / S-halfspace inputs fixed in SURVEY.md section 8(d).
A mesh is: nodes (n,3) float64; elements list of (etype, part_id, node ids 0-based); parts = boundaries.
"""
import numpy as np
from .shape import TRI3, TRI6, QUAD4, QUAD8, QUAD9, GMSH_TYPE, GMSH_CODE, N_NODES


class Mesh:
    def __init__(self, nodes, etype, part, conn, node_ids=None, elem_ids=None):
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.etype = np.asarray(etype, dtype=np.int32)
        self.part = np.asarray(part, dtype=np.int32)
        self.conn = [np.asarray(c, dtype=np.int32) for c in conn]
        # identifiers of the mesh file (node()%id, element()%id of the reference: what the result files print); default 1..n
        self.node_ids = np.arange(1, len(self.nodes) + 1, dtype=np.int64) if node_ids is None else np.asarray(node_ids, dtype=np.int64)
        self.elem_ids = np.arange(1, len(self.conn) + 1, dtype=np.int64) if elem_ids is None else np.asarray(elem_ids, dtype=np.int64)

    @property
    def n_elem(self):
        return len(self.conn)


def read_gmsh22(path):
    """Gmsh MSH 2.2 ASCII ($Nodes / $Elements); the first tag (physical entity) is the part/boundary id.  Only surface elements are
    kept (the boundary-element mesh); nodes that no kept element uses (geometry points, nodes of line elements) are dropped, the
    identifiers of the file are kept in Mesh.node_ids / Mesh.elem_ids."""
    lines = open(path).read().split("\n")
    i = 0
    ids, xyz, et, part, conn, nid, eid = {}, [], [], [], [], [], []
    while i < len(lines):
        s = lines[i].strip()
        if s == "$Nodes":
            n = int(lines[i + 1])
            for k in range(n):
                t = lines[i + 2 + k].split()
                ids[int(t[0])] = k
                nid.append(int(t[0]))
                xyz.append([float(t[1]), float(t[2]), float(t[3])])
            i += n + 2
        elif s == "$Elements":
            n = int(lines[i + 1])
            for k in range(n):
                t = [int(v) for v in lines[i + 2 + k].split()]
                gt, ntags = t[1], t[2]
                if gt in GMSH_TYPE:
                    et.append(GMSH_TYPE[gt])
                    part.append(t[3])
                    eid.append(t[0])
                    conn.append([ids[v] for v in t[3 + ntags:3 + ntags + N_NODES[GMSH_TYPE[gt]]]])
            i += n + 2
        else:
            i += 1
    used = np.zeros(len(xyz), dtype=bool)
    for c in conn:
        used[c] = True
    remap = np.cumsum(used) - 1
    conn = [[int(remap[v]) for v in c] for c in conn]
    return Mesh(np.array(xyz)[used], et, part, conn, node_ids=np.array(nid)[used], elem_ids=eid)


NATIVE_TYPE = {"tri3": TRI3, "tri6": TRI6, "quad4": QUAD4, "quad8": QUAD8, "quad9": QUAD9}


def read_native_mesh(nodes, elements):
    """The reference's own mesh sections (mesh_file_mode 0: inside the case file; 1: in an auxiliary file): `[nodes]` = count, then `<id> <x1> <x2> <x3>`
    (src/read_nodes.f90:148-160); `[elements]` = count, then `<id> <type> <number of tags> <tag 1 = part> ... <node 1> ... <node N>` with the type by
    name or by its Gmsh code and the node order of Gmsh (src/read_elements.f90:30-36, :236-248).  Arguments: the non-empty lines of the two sections.
    Surface elements only; unused nodes are dropped, identifiers kept, as read_gmsh22 does."""
    ids, xyz, nid = {}, [], []
    n = int(nodes[0].split()[0])
    if len(nodes) < n + 1:
        raise ValueError("[nodes]: %d nodes announced, %d records found" % (n, len(nodes) - 1))
    for k in range(n):
        t = nodes[1 + k].replace(",", " ").split()
        ids[int(t[0])] = k; nid.append(int(t[0]))
        xyz.append([float(v.lower().replace("d", "e")) for v in t[1:4]])
    et, part, conn, eid = [], [], [], []
    n = int(elements[0].split()[0])
    if len(elements) < n + 1:
        raise ValueError("[elements]: %d elements announced, %d records found" % (n, len(elements) - 1))
    for k in range(n):
        t = elements[1 + k].split()
        ty = NATIVE_TYPE.get(t[1].lower())
        if ty is None and t[1].isdigit():
            ty = GMSH_TYPE.get(int(t[1]))
        if ty is None:
            raise ValueError("element %s: type %r is not a surface boundary element (tri3, tri6, quad4, quad8, quad9)" % (t[0], t[1]))
        ntags = int(t[2])
        et.append(ty); part.append(int(t[3])); eid.append(int(t[0]))
        conn.append([ids[int(v)] for v in t[3 + ntags:3 + ntags + N_NODES[ty]]])
    used = np.zeros(len(xyz), dtype=bool)
    for c in conn:
        used[c] = True
    remap = np.cumsum(used) - 1
    conn = [[int(remap[v]) for v in c] for c in conn]
    return Mesh(np.array(xyz)[used], et, part, conn, node_ids=np.array(nid)[used], elem_ids=eid)


def write_gmsh22(mesh, path, names=None):
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
        parts = sorted(set(int(p) for p in mesh.part))
        f.write("$PhysicalNames\n%d\n" % len(parts))
        for p in parts:
            f.write('2 %d "%s"\n' % (p, (names or {}).get(p, "part%d" % p)))
        f.write("$EndPhysicalNames\n$Nodes\n%d\n" % len(mesh.nodes))
        for k, x in enumerate(mesh.nodes):
            f.write("%d %.17g %.17g %.17g\n" % (k + 1, x[0], x[1], x[2]))
        f.write("$EndNodes\n$Elements\n%d\n" % mesh.n_elem)
        for k in range(mesh.n_elem):
            f.write("%d %d 2 %d %d %s\n" % (k + 1, GMSH_CODE[int(mesh.etype[k])], mesh.part[k], mesh.part[k],
                                           " ".join(str(int(v) + 1) for v in mesh.conn[k])))
        f.write("$EndElements\n")


def _face_grid(origin, eu, ev, m, etype, part, nodes, et, pt, conn):
    """m x m cells on the parallelogram origin + u*eu + v*ev (u,v in [0,1]); normal = eu x ev."""
    quad = etype in (TRI6, QUAD8, QUAD9)
    n1 = (2 * m + 1) if quad else (m + 1)
    base = len(nodes)
    idx = -np.ones((n1, n1), dtype=np.int64)
    for j in range(n1):
        for i in range(n1):
            if etype == QUAD8 and (i % 2 == 1) and (j % 2 == 1):
                continue
            idx[i, j] = base + len(nodes) - base
            nodes.append(origin + (i / (n1 - 1.0)) * eu + (j / (n1 - 1.0)) * ev)
    s = 2 if quad else 1
    for j in range(m):
        for i in range(m):
            a, b, c, d = idx[s * i, s * j], idx[s * i + s, s * j], idx[s * i + s, s * j + s], idx[s * i, s * j + s]
            if etype == QUAD4:
                cs = [[a, b, c, d]]
            elif etype == QUAD8:
                cs = [[a, b, c, d, idx[2 * i + 1, 2 * j], idx[2 * i + 2, 2 * j + 1], idx[2 * i + 1, 2 * j + 2], idx[2 * i, 2 * j + 1]]]
            elif etype == QUAD9:
                cs = [[a, b, c, d, idx[2 * i + 1, 2 * j], idx[2 * i + 2, 2 * j + 1], idx[2 * i + 1, 2 * j + 2], idx[2 * i, 2 * j + 1], idx[2 * i + 1, 2 * j + 1]]]
            elif etype == TRI3:
                cs = [[a, b, c], [a, c, d]]
            else:  # TRI6: split along the a-c diagonal
                mab, mbc, mca = idx[2 * i + 1, 2 * j], idx[2 * i + 2, 2 * j + 1], idx[2 * i + 1, 2 * j + 1]
                mcd, mda = idx[2 * i + 1, 2 * j + 2], idx[2 * i, 2 * j + 1]
                cs = [[a, b, c, mab, mbc, mca], [a, c, d, mca, mcd, mda]]
            for cc in cs:
                et.append(etype); pt.append(part); conn.append(cc)


def cube_mesh(m, etype=TRI3, L=1.0):
    """S-cube(m, etype): unit cube, 6 faces = 6 parts with *unshared* rim nodes (as in the reference's
    docs/examples/ME-TH-EL-001/case_files/t3.msh), m x m cells per face, outward normals.
    Part ids: 1 x=0, 2 x=L, 3 y=0, 4 y=L, 5 z=0, 6 z=L."""
    nodes, et, pt, conn = [], [], [], []
    ex, ey, ez, o = np.array([L, 0, 0.]), np.array([0, L, 0.]), np.array([0, 0, L]), np.zeros(3)
    _face_grid(o, ez, ey, m, etype, 1, nodes, et, pt, conn)            # x=0, normal -x
    _face_grid(o + ex, ey, ez, m, etype, 2, nodes, et, pt, conn)       # x=L, normal +x
    _face_grid(o, ex, ez, m, etype, 3, nodes, et, pt, conn)            # y=0, normal -y
    _face_grid(o + ey, ez, ex, m, etype, 4, nodes, et, pt, conn)       # y=L, normal +y
    _face_grid(o, ey, ex, m, etype, 5, nodes, et, pt, conn)            # z=0, normal -z
    _face_grid(o + ez, ex, ey, m, etype, 6, nodes, et, pt, conn)       # z=L, normal +z
    return Mesh(np.array(nodes), et, pt, conn)


def two_box_mesh(m, etype=TRI3, xs=0.5, L=1.0):
    """The cube [0,L]^3 cut by the plane x = xs*L into two boxes that share ONE interface part (for two coupled BE regions): every face
    of every box is its own part with unshared rim nodes, m x m cells per face, normals outward from the box that owns the face; the
    interface is meshed once, with the normal +x (outward from the box x < xs*L = its region 1).
    Parts: 1 x=0; 3, 4, 5, 6 = y=0, y=L, z=0, z=L of the first box; 7 interface; 2 x=L; 13, 14, 15, 16 the lateral faces of the second box."""
    nodes, et, pt, conn = [], [], [], []
    a, b = xs * L, (1.0 - xs) * L
    ex, ey, ez, o = np.array([1.0, 0, 0]), np.array([0, L, 0.]), np.array([0, 0, L]), np.zeros(3)
    _face_grid(o, ez, ey, m, etype, 1, nodes, et, pt, conn)                   # x=0, normal -x
    _face_grid(o, a * ex, ez, m, etype, 3, nodes, et, pt, conn)               # y=0, normal -y
    _face_grid(o + ey, ez, a * ex, m, etype, 4, nodes, et, pt, conn)          # y=L, normal +y
    _face_grid(o, ey, a * ex, m, etype, 5, nodes, et, pt, conn)               # z=0, normal -z
    _face_grid(o + ez, a * ex, ey, m, etype, 6, nodes, et, pt, conn)          # z=L, normal +z
    _face_grid(o + a * ex, ey, ez, m, etype, 7, nodes, et, pt, conn)          # interface x=xs L, normal +x
    o2 = o + a * ex
    _face_grid(o2 + b * ex, ey, ez, m, etype, 2, nodes, et, pt, conn)         # x=L, normal +x
    _face_grid(o2, b * ex, ez, m, etype, 13, nodes, et, pt, conn)
    _face_grid(o2 + ey, ez, b * ex, m, etype, 14, nodes, et, pt, conn)
    _face_grid(o2, ey, b * ex, m, etype, 15, nodes, et, pt, conn)
    _face_grid(o2 + ez, b * ex, ey, m, etype, 16, nodes, et, pt, conn)
    return Mesh(np.array(nodes), et, pt, conn)


def halfspace_patch(m, etype=TRI3, L=1.0, footing=0.25):
    """S-halfspace(m): flat free-surface patch z=0 of side L (soil below, outward normal +z); part 1 = free
    surface, part 2 = central square footing of half-width `footing`*L (cells whose centre lies inside)."""
    nodes, et, pt, conn = [], [], [], []
    _face_grid(np.array([-L / 2, -L / 2, 0.]), np.array([L, 0, 0.]), np.array([0, L, 0.]), m, etype, 1, nodes, et, pt, conn)
    nodes = np.array(nodes)
    for k, c in enumerate(conn):
        ctr = nodes[np.array(c[:4 if etype in (QUAD4, QUAD8, QUAD9) else 3])].mean(axis=0)
        if abs(ctr[0]) < footing * L and abs(ctr[1]) < footing * L:
            pt[k] = 2
    # parts must not share nodes (each boundary owns its nodes): duplicate nodes used by both parts
    used = {}
    nodes = list(nodes)
    for k, c in enumerate(conn):
        if pt[k] == 2:
            for j, v in enumerate(c):
                if v not in used:
                    used[v] = len(nodes); nodes.append(nodes[v].copy())
                c[j] = used[v]
    # drop now-unreferenced originals
    ref = sorted(set(int(v) for c in conn for v in c))
    remap = {v: i for i, v in enumerate(ref)}
    conn = [[remap[int(v)] for v in c] for c in conn]
    return Mesh(np.array(nodes)[ref], et, pt, conn)


_REVERSED_ORDER = {TRI3: [0, 2, 1], TRI6: [0, 2, 1, 5, 4, 3], QUAD4: [0, 3, 2, 1], QUAD8: [0, 3, 2, 1, 7, 6, 5, 4], QUAD9: [0, 3, 2, 1, 7, 6, 5, 4, 8]}


def without_parts(mesh, parts):
    """The mesh without the elements of the given parts (and without the nodes only they use)."""
    keep = [e for e in range(mesh.n_elem) if int(mesh.part[e]) not in parts]
    used = sorted(set(int(v) for e in keep for v in mesh.conn[e]))
    new = {v: i for i, v in enumerate(used)}
    return Mesh(mesh.nodes[used], [mesh.etype[e] for e in keep], [mesh.part[e] for e in keep], [[new[int(v)] for v in mesh.conn[e]] for e in keep],
                node_ids=np.asarray(mesh.node_ids)[used], elem_ids=[mesh.elem_ids[e] for e in keep])


def mirror_mesh(mesh, axis, part_offset=100, tol=1e-9):
    """The union of `mesh` and its mirror image across the plane through the origin normal to `axis` (0, 1, 2): the FULL model that a
    half model with a [symmetry planes] entry stands for.  Nodes lying in the plane are shared by both halves (the open edge there
    closes); a mirrored element lists its nodes in the reversed order, so that its normal stays outward.  A part that touches the
    plane keeps its id on both sides (one boundary crossing the plane), any other mirrored part gets id + part_offset.
    Returns (full mesh, image_of_node: index of every original node's mirror node in the full mesh)."""
    n = len(mesh.nodes)
    on_plane = np.abs(mesh.nodes[:, axis]) <= tol
    image = np.arange(n)
    extra = np.flatnonzero(~on_plane)
    image[extra] = n + np.arange(len(extra))
    mirrored = mesh.nodes[extra].copy(); mirrored[:, axis] *= -1.0
    nodes = np.vstack([mesh.nodes, mirrored])
    touching = set(int(mesh.part[e]) for e in range(mesh.n_elem) if on_plane[mesh.conn[e]].any())
    et, pt, conn = list(mesh.etype), list(mesh.part), [list(c) for c in mesh.conn]
    for e in range(mesh.n_elem):
        t, p = int(mesh.etype[e]), int(mesh.part[e])
        et.append(t); pt.append(p if p in touching else p + part_offset)
        conn.append([int(image[mesh.conn[e][k]]) for k in _REVERSED_ORDER[t]])
    return Mesh(nodes, et, pt, conn), image
