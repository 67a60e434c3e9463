"""Lagrange shape functions of the 2D boundary elements (continuous, delta = 0).

Host-side (Python) evaluation used only to place the collocation points exactly like the
Fortran host does (src/build_data_at_collocation_points.f90:124-160,285-310).  Node order and
reference coordinates follow lib/fbem/src/resources_shape_functions/xi_{tri3,tri6,quad4,quad8,quad9}.rc.
"""
import numpy as np

TRI3, TRI6, QUAD4, QUAD8, QUAD9 = 5, 6, 7, 8, 9          # lib/fbem/src/shape_functions.f90:198-206
N_NODES = {TRI3: 3, TRI6: 6, QUAD4: 4, QUAD8: 8, QUAD9: 9}
N_VERTICES = {TRI3: 3, TRI6: 3, QUAD4: 4, QUAD8: 4, QUAD9: 4}
GMSH_TYPE = {2: TRI3, 9: TRI6, 3: QUAD4, 16: QUAD8, 10: QUAD9}  # src/read_elements.f90:246-250
GMSH_CODE = {v: k for k, v in GMSH_TYPE.items()}

XI_NODES = {
    TRI3: np.array([[1., 0.], [0., 1.], [0., 0.]]),
    TRI6: np.array([[1., 0.], [0., 1.], [0., 0.], [.5, .5], [0., .5], [.5, 0.]]),
    QUAD4: np.array([[-1., -1.], [1., -1.], [1., 1.], [-1., 1.]]),
    QUAD8: np.array([[-1., -1.], [1., -1.], [1., 1.], [-1., 1.], [0., -1.], [1., 0.], [0., 1.], [-1., 0.]]),
    QUAD9: np.array([[-1., -1.], [1., -1.], [1., 1.], [-1., 1.], [0., -1.], [1., 0.], [0., 1.], [-1., 0.], [0., 0.]]),
}


def edges_of(etype):
    """Local node ids (vertex a, vertex b[, mid]) of each edge: shape_functions.f90:841-905."""
    nv = N_VERTICES[etype]
    quad = etype in (TRI6, QUAD8, QUAD9)
    return [(k, (k + 1) % nv) + ((nv + k,) if quad else ()) for k in range(nv)]


def phi(etype, xi):
    a1, a2 = float(xi[0]), float(xi[1])
    if etype == TRI3:
        return np.array([a1, a2, 1.0 - a1 - a2])
    if etype == TRI6:
        a3 = 1.0 - a1 - a2
        a4 = 4.0 * a1
        return np.array([a1 * (2.0 * a1 - 1.0), a2 * (2.0 * a2 - 1.0), a3 * (2.0 * a3 - 1.0), a4 * a2, 4.0 * a2 * a3, a4 * a3])
    if etype == QUAD4:
        a3, a4, a5, a6 = 0.25 * (1.0 + a1), 0.25 * (1.0 - a1), 1.0 + a2, 1.0 - a2
        return np.array([a4 * a6, a3 * a6, a3 * a5, a4 * a5])
    if etype == QUAD8:
        a3, a4, a5, a6 = 0.25 * (1.0 + a1), 0.25 * (1.0 - a1), 1.0 + a2, 1.0 - a2
        a7, a8 = 1.0 - a1 * a1, 1.0 - a2 * a2
        return np.array([a4 * a6 * (-a1 - a5), a3 * a6 * (a1 - a5), a3 * a5 * (a1 - a6), a4 * a5 * (-a1 - a6),
                         0.5 * a6 * a7, 2.0 * a3 * a8, 0.5 * a5 * a7, 2.0 * a4 * a8])
    if etype == QUAD9:
        a3, a4 = 0.25 * a1 * (a1 + 1.0), 0.25 * a1 * (a1 - 1.0)
        a5, a6 = a2 * (a2 + 1.0), a2 * (a2 - 1.0)
        a7, a8 = 1.0 - a1 * a1, 1.0 - a2 * a2
        return np.array([a4 * a6, a3 * a6, a3 * a5, a4 * a5, 0.5 * a6 * a7, 2.0 * a3 * a8, 0.5 * a5 * a7, 2.0 * a4 * a8, a7 * a8])
    raise ValueError("unsupported element type %r" % etype)


def position(etype, x_nodes, xi):
    """x(xi) = sum_k phi_k x_k accumulated in node order (fbem_position3d)."""
    p = phi(etype, xi)
    x = np.zeros(3)
    for k in range(N_NODES[etype]):
        x = x + p[k] * x_nodes[k]
    return x


def move_xi_from_edge(etype, xi, delta):
    """MCA shift towards the element interior: lib/fbem/src/geometry.f90:5631-5658."""
    if etype in (TRI3, TRI6):
        return np.array([xi[0] * (1.0 - delta) + 0.333333333333333333 * delta, xi[1] * (1.0 - delta) + 0.333333333333333333 * delta])
    return np.array([xi[0] * (1.0 - delta), xi[1] * (1.0 - delta)])


def dphi(etype, xi):
    """(dphi/dxi1, dphi/dxi2): the derivatives of `phi` (resources_shape_functions/dphidxi{1,2}_*.rc of the reference)."""
    a1, a2 = float(xi[0]), float(xi[1])
    if etype == TRI3:
        return np.array([1.0, 0.0, -1.0]), np.array([0.0, 1.0, -1.0])
    if etype == TRI6:
        return (np.array([4.0 * a1 - 1.0, 0.0, 4.0 * a1 + 4.0 * a2 - 3.0, 4.0 * a2, -4.0 * a2, -4.0 * (a2 + 2.0 * a1 - 1.0)]),
                np.array([0.0, 4.0 * a2 - 1.0, 4.0 * a1 + 4.0 * a2 - 3.0, 4.0 * a1, -4.0 * (2.0 * a2 + a1 - 1.0), -4.0 * a1]))
    if etype == QUAD4:
        a3, a4, a5, a6 = 0.25 * (1.0 + a1), 0.25 * (1.0 - a1), 1.0 + a2, 1.0 - a2
        return np.array([-0.25 * a6, 0.25 * a6, 0.25 * a5, -0.25 * a5]), np.array([-a4, -a3, a3, a4])
    if etype == QUAD8:
        b3, b4, b5, b6 = a2 + 1.0, a2 - 1.0, a2 + 2.0 * a1, a2 - 2.0 * a1
        d1 = np.array([-0.25 * b4 * b5, 0.25 * b4 * b6, 0.25 * b3 * b5, -0.25 * b3 * b6, a1 * b4, -0.5 * b3 * b4, -a1 * b3, 0.5 * b3 * b4])
        c3, c4, c5, c6 = a1 + 1.0, a1 - 1.0, 2.0 * a2 + a1, 2.0 * a2 - a1
        d2 = np.array([-0.25 * c4 * c5, 0.25 * c3 * c6, 0.25 * c3 * c5, -0.25 * c4 * c6, 0.5 * c3 * c4, -a2 * c3, -0.5 * c3 * c4, a2 * c4])
        return d1, d2
    if etype == QUAD9:
        b3, b4, b5, b6, b7 = 2.0 * a1 + 1.0, 2.0 * a1 - 1.0, a2 + 1.0, a2 - 1.0, 0.25 * a2
        b8 = b5 * b6; b9 = -0.5 * b8; b10 = -a1 * a2
        d1 = np.array([b7 * b4 * b6, b7 * b3 * b6, b7 * b3 * b5, b7 * b4 * b5, b10 * b6, b9 * b3, b10 * b5, b9 * b4, 2.0 * a1 * b8])
        c3, c4, c5, c6, c7 = 2.0 * a2 + 1.0, 2.0 * a2 - 1.0, a1 + 1.0, a1 - 1.0, 0.25 * a1
        c8 = c5 * c6; c9 = -0.5 * c8
        d2 = np.array([c7 * c6 * c4, c7 * c5 * c4, c7 * c5 * c3, c7 * c6 * c3, c9 * c4, b10 * c5, c9 * c3, b10 * c6, 2.0 * a2 * c8])
        return d1, d2
    raise ValueError("unsupported element type %r" % etype)


def unit_normal(etype, x_nodes, xi):
    """Unit normal T1 x T2 / |T1 x T2| of the element at xi (fbem_unormal3d)."""
    d1, d2 = dphi(etype, xi)
    t1 = np.zeros(3); t2 = np.zeros(3)
    for k in range(N_NODES[etype]):
        t1 = t1 + d1[k] * x_nodes[k]; t2 = t2 + d2[k] * x_nodes[k]
    n = np.cross(t1, t2)
    return n / np.sqrt(n @ n)


def node_normal_tangents(etype, x_nodes, node):
    """(n, t+, t-) at element node `node`: unit normal and the unit tangents of the two element edges that meet there (pointing away from
    the node along the boundary in the positive / negative sense), or along / against the edge for a mid-side node
    (src/build_data_at_geometrical_nodes.f90:198-222, fbem_utangents_at_boundary, lib/fbem/src/geometry.f90:657-914)."""
    xi = XI_NODES[etype][node]
    d1, d2 = dphi(etype, xi)
    T1 = np.zeros(3); T2 = np.zeros(3)
    for k in range(N_NODES[etype]):
        T1 = T1 + d1[k] * x_nodes[k]; T2 = T2 + d2[k] * x_nodes[k]
    N = np.cross(T1, T2)
    n = N / np.sqrt(N @ N)
    t1 = T1 / np.sqrt(T1 @ T1); t2 = T2 / np.sqrt(T2 @ T2)
    if etype in (TRI3, TRI6):
        if etype == TRI3:
            d3 = np.array([1.0, -1.0, 0.0])
        else:
            d3 = np.array([4.0 * xi[0] - 1.0, 4.0 * xi[0] - 3.0, 0.0, 4.0 * (1.0 - 2.0 * xi[0]), 0.0, 0.0])
        T3 = np.zeros(3)
        for k in range(N_NODES[etype]):
            T3 = T3 + d3[k] * x_nodes[k]
        t3 = T3 / np.sqrt(T3 @ T3)
        table = {0: (-t3, -t1), 1: (-t2, t3), 2: (t1, t2), 3: (-t3, t3), 4: (-t2, t2), 5: (t1, -t1)}
    else:
        z = np.zeros(3)
        table = {0: (t1, t2), 1: (t2, -t1), 2: (-t1, -t2), 3: (-t2, t1), 4: (t1, -t1), 5: (t2, -t2), 6: (-t1, t1), 7: (-t2, t2), 8: (z, z)}
    tp, tm = table[node]
    return n, tp, tm
