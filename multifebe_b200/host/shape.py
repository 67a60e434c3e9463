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
