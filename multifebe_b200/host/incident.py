"""Incident wave fields for the harmonic elastic path: what the host hands to mfb_harela3d_set_incident.

In the reference the host fills element()%incident_c with u_inc, t_inc at the nodes of every element
(src/calculate_incident_mechanics_harmonic.f90:420-500: the field evaluated at x_fn with the region's outward normal, n_fn negated on a
reversed boundary) and the assembly adds hp u_inc - gp t_inc to b (src/assemble_bem_harela_equation.f90:651-666).  The fields themselves
(half-space reflections, Rayleigh waves, layered soils: lib/fbem/src/harela_incident_field.f90) stay with the Fortran host -- the library
takes the arrays.  This module builds them for the simplest member of that family, written from the wave equation itself: a plane
P or S wave of a FULL space, time factor exp(i omega t),

    u(x) = A p exp(-i k d.x),   sigma = lambda (div u) I + mu (grad u + grad u^T),   t = sigma n,

with d the unit propagation direction, p the unit polarisation (P: p = d, k = omega / c1; S: p perpendicular to d, k = omega / c2).
"""
import numpy as np

from . import shape as sh


def plane_wave(kind, direction, mat, omega, polarisation=None, amplitude=1.0):
    """-> field(x, n) = (u (3,), t (3,)) complex of a plane wave in the full space with the properties of `mat`."""
    d = np.asarray(direction, dtype=np.float64)
    d = d / np.linalg.norm(d)
    if kind == "P":
        p, k = d, omega / mat.c1
    elif kind == "S":
        p = np.asarray(polarisation, dtype=np.float64)
        p = p - d * np.dot(p, d)
        if np.linalg.norm(p) < 1e-12:
            raise ValueError("S wave: the polarisation must not be parallel to the direction")
        p, k = p / np.linalg.norm(p), omega / mat.c2
    else:
        raise ValueError("kind: 'P' or 'S'")

    def field(x, n):
        ph = amplitude * np.exp(-1j * k * np.dot(d, x))
        u = p * ph
        grad = np.outer(p, d) * (-1j * k * ph)                 # grad[i, j] = du_i / dx_j
        sigma = mat.lam * np.trace(grad) * np.eye(3) + mat.mu * (grad + grad.T)
        return u, sigma @ np.asarray(n, dtype=np.float64)
    return field


def element_incident(model, field):
    """u_inc, t_inc ((sum nn, 3) complex each, element order) = element()%incident_c(1:3,kn,1), (4:6,kn,1): `field` at the nodes of every
    element with the region's outward normal there (the element's normal at the node, negated on a reversed boundary)."""
    n_rows = int(model.elem_ptr[-1])
    u = np.zeros((n_rows, 3), dtype=np.complex128); t = np.zeros((n_rows, 3), dtype=np.complex128)
    for e in range(model.n_elem):
        et = int(model.etype[e]); c = model.mesh.conn[e]; xn = model.node_x[c]
        sgn = -1.0 if model.elem_reversed[e] else 1.0
        for kn in range(len(c)):
            n = sgn * sh.unit_normal(et, xn, sh.XI_NODES[et][kn])
            u[model.elem_ptr[e] + kn], t[model.elem_ptr[e] + kn] = field(xn[kn], n)
    return u, t


def plane_wave_fluid(direction, fluid, omega, amplitude=1.0):
    """-> field(x, n) = (p, Un) of a plane pressure wave p = A exp(-i k d.x), k = omega / c, in an inviscid fluid; Un = (dp/dn) / (rho omega^2) is the normal
    displacement, the flux variable of the reference's fluid regions (build_lse_mechanics_bem_harpot.f90:751)."""
    d = np.asarray(direction, dtype=np.float64)
    d = d / np.linalg.norm(d)
    k = omega / fluid.c

    def field(x, n):
        p = amplitude * np.exp(-1j * k * np.dot(d, x))
        return p, (-1j * k * np.dot(d, np.asarray(n, dtype=np.float64)) * p) / (fluid.rho * omega ** 2)
    return field


def element_incident_fluid(model, field):
    """p_inc, Un_inc ((sum nn,) complex each, element order) at the nodes of every element of a fluid region, with the region's outward normal."""
    n_rows = int(model.elem_ptr[-1])
    p = np.zeros(n_rows, dtype=np.complex128); un = np.zeros(n_rows, dtype=np.complex128)
    for e in range(model.n_elem):
        et = int(model.etype[e]); c = model.mesh.conn[e]; xn = model.node_x[c]
        sgn = -1.0 if model.elem_reversed[e] else 1.0
        for kn in range(len(c)):
            n = sgn * sh.unit_normal(et, xn, sh.XI_NODES[et][kn])
            p[model.elem_ptr[e] + kn], un[model.elem_ptr[e] + kn] = field(xn[kn], n)
    return p, un
