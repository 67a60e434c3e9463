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


# ---------------------------------------------------------------------------------------------------------------------------------------
# The incident fields of the [incident waves] section, in the REFERENCE's conventions (angles, normalisation, symmetry decomposition), so that
# a case file that names them gives the arrays the Fortran host would hand to the assembly.  Restated from the formulas of
# lib/fbem/src/harpot_incident_field.f90 (fluid: point wave :49-119, plane wave :122-290) and lib/fbem/src/harela_incident_field.f90:355-745
# (elastic plane P / SV / SH wave in a full space or in a half-space z <= z_fs with a free surface), called as
# src/calculate_incident_mechanics_harmonic.f90:366-378 and :434-441 call them.  There is no Fortran compiler in this image, so these are
# pinned by the physics they must satisfy (tests/test_incident_waves.py: wave equation, stress-free / pressure-release / rigid plane, the
# decomposition adding up to the whole field), not by outputs of the reference: "parity unpinned" for the constants read off the source.
#
# A field is held as a superposition of plane waves  f(x) = sum_w a_w pol_w exp(-i kv_w.(x - xs))  (pol_w scalar 1 for a fluid); the
# symmetric / antisymmetric part about a coordinate plane is the half sum / half difference with the mirrored superposition, which is
# what the W / Z products (fluid) and the eys / eya factors (elastic) of the reference spell out component by component.
# ---------------------------------------------------------------------------------------------------------------------------------------
def _direction(varphi, theta):
    """Unit propagation vector of the reference's angles (radians): varphi from the y axis in the xy plane, theta from the xy plane."""
    return np.array([np.cos(theta) * np.sin(varphi), np.cos(theta) * np.cos(varphi), np.sin(theta)])


def _mirror_parts(waves, symconf, vector):
    """waves [(a, pol, kv)] -> the superposition reduced to its symmetric (symconf[c] = 1) or antisymmetric (-1) part about the plane x_c = xs_c,
    for every axis c; 0 leaves the axis alone.  A mirrored elastic wave has its polarisation reflected too (vector = True)."""
    for c in range(3):
        s = int(round(symconf[c]))
        if s == 0:
            continue
        out = []
        for a, pol, kv in waves:
            m = np.ones(3); m[c] = -1.0
            out.append((0.5 * a, pol, kv))
            out.append((0.5 * a * s, pol * m if vector else pol, kv * m))
        waves = out
    return waves


def fluid_plane_wave_reference(fluid, omega, amplitude=1.0, x0=(0, 0, 0), varphi=0.0, theta=0.0, space="full-space", np_axis=3, xp=0.0, bc=1,
                               symconf=(0, 0, 0), xs=(0, 0, 0)):
    """-> field(x, n) = (p_inc, Un_inc) of `plane` / fluid / `p` (fbem_harpot_planewave): p = A exp(-i k q.(x - x0)), time factor exp(i omega t);
    half-space: plus the wave reflected at the plane x_np = xp, of amplitude -A (bc 0, p = 0 there) or +A (bc 1, Un = 0 there) times the phase
    that makes the two meet at the plane.  symconf / xs: the symmetric (+1) / antisymmetric (-1) part about the planes through xs.
    Kept as the reference has it: the phase of the reflected amplitude is written with the ALREADY reflected direction (harpot_incident_field.f90:227,
    :238-243), so the condition on the plane holds when the origin x0 lies on the plane (x0(np) = xp; the case of every shipped example) and is off by
    exp(4 i k q_np (xp - x0(np))) otherwise.  A drop-in gives what the reference gives; tests/test_incident_waves.py states the finding."""
    k = omega / fluid.c
    q = _direction(varphi, theta)
    x0, xs = np.asarray(x0, dtype=np.float64), np.asarray(xs, dtype=np.float64)
    A = complex(amplitude)
    waves = [(A * np.exp(1j * k * np.dot(q, x0 - xs)), 1.0, k * q)]
    if space == "half-space":
        c = int(np_axis) - 1
        if int(round(symconf[c])) != 0:
            raise ValueError("incident wave: the half-space plane cannot be a symmetry plane (symconf(np) must be 0)")
        qr = q.copy(); qr[c] = -qr[c]
        Ar = (-A if int(bc) == 0 else A) * np.exp(-2j * k * qr[c] * (xp - x0[c]))
        waves.append((Ar * np.exp(1j * k * np.dot(qr, x0 - xs)), 1.0, k * qr))
    elif space != "full-space":
        raise ValueError("incident wave: space %r of a fluid plane wave" % (space,))
    waves = _mirror_parts(waves, symconf, vector=False)
    rw2 = fluid.rho * omega ** 2

    def field(x, n):                                   # one point (3,), (3,) or a batch (m, 3), (m, 3)
        d = np.asarray(x, dtype=np.float64) - xs
        nn = np.asarray(n, dtype=np.float64)
        p, grad = np.zeros(d.shape[:-1], dtype=np.complex128), np.zeros(d.shape, dtype=np.complex128)
        for a, _, kv in waves:
            e = a * np.exp(-1j * (d @ kv))
            p = p + e; grad = grad + e[..., None] * (-1j * kv)
        un = np.sum(grad * nn, axis=-1) / rw2
        return (p, un) if d.ndim > 1 else (complex(p), complex(un))
    return field


def fluid_point_wave_reference(fluid, omega, amplitude=1.0, x0=(0, 0, 0)):
    """-> field(x, n) = (p_inc, Un_inc) of `point` / fluid in a full space (fbem_harpot_pointwave, 3D): a monopole at x0 whose pressure is
    `amplitude` at unit distance with the phase of that distance taken out, p = A exp(-i k (r - 1)) / r."""
    k = omega / fluid.c
    x0 = np.asarray(x0, dtype=np.float64)
    A = complex(amplitude)
    rw2 = fluid.rho * omega ** 2

    def field(x, n):                                   # one point or a batch, as above
        rv = np.asarray(x, dtype=np.float64) - x0
        r = np.linalg.norm(rv, axis=-1)
        p = A * np.exp(-1j * k * (r - 1.0)) / r
        dpdr = -p * (1.0 / r + 1j * k)
        un = dpdr * np.sum(rv * np.asarray(n, dtype=np.float64), axis=-1) / r / rw2
        return (p, un) if rv.ndim > 1 else (complex(p), complex(un))
    return field


def elastic_plane_wave_reference(wave, mat, omega, varphi=0.0, theta=np.pi / 2, space="full-space", z_fs=0.0, symconf_y=0):
    """-> field(x, n) = (u_inc (3,), t_inc (3,)) of `plane` / elastic / `p` | `sv` | `sh` | `rayleigh` (fbem_harela_incident_plane_wave) with the halving of
    its caller (calculate_incident_mechanics_harmonic.f90:439-440), which makes the free-field motion of the surface under vertical incidence 1.
    The wave travels in the vertical plane that makes the angle varphi with the yz plane, rising at the angle theta over the horizontal; in the
    half-space z <= z_fs (np = 3, bc = 1: stress-free surface) the reflected waves are added: SH -> SH; P -> P + SV; SV -> SV + P, the P wave
    evanescent beyond the critical angle (cos theta > kappa = c2 / c1 = sqrt((1 - 2 nu) / (2 (1 - nu))), nu real as the reference takes it).
    symconf_y = +1 / -1: only the part symmetric / antisymmetric about the plane y = 0 (the only decomposition the reference offers in 3D).
    The amplitude record of the section is not used for elastic waves in the reference either."""
    k1, k2 = omega / mat.c1, omega / mat.c2
    kap = np.sqrt((1.0 - 2.0 * mat.nu_r) / (2.0 * (1.0 - mat.nu_r)))
    c0, s0 = np.cos(theta), np.sin(theta)
    inc, rfl = np.array([0.0, c0, s0], dtype=np.complex128), np.array([0.0, c0, -s0], dtype=np.complex128)
    if wave == "sh":
        ex = np.array([1.0, 0, 0], dtype=np.complex128)
        local = [(1.0, ex, k2, inc), (1.0, ex, k2, rfl)]
    elif wave == "p":
        c2 = kap * c0; s2 = np.sqrt(complex(1.0 - c2 * c2))
        s20, s22, c22 = 2.0 * s0 * c0, 2.0 * s2 * c2, c2 * c2 - s2 * s2
        den = kap * kap * s20 * s22 + c22 * c22
        local = [(1.0, inc.copy(), k1, inc), ((kap * kap * s20 * s22 - c22 * c22) / den, rfl.copy(), k1, rfl),
                 (2.0 * kap * s20 * c22 / den, np.array([0.0, -s2, -c2]), k2, np.array([0.0, c2, -s2]))]
    elif wave == "sv":
        c2 = c0 / kap
        s2 = np.sqrt(complex(1.0 - c2 * c2)) if abs(c0) <= kap else -1j * np.sqrt(c2 * c2 - 1.0)
        s20, s22, c20 = 2.0 * s0 * c0, 2.0 * s2 * c2, c0 * c0 - s0 * s0
        den = kap * kap * s20 * s22 + c20 * c20
        local = [(1.0, np.array([0.0, s0, -c0], dtype=np.complex128), k2, inc), ((kap * kap * s20 * s22 - c20 * c20) / den, np.array([0.0, -s0, -c0], dtype=np.complex128), k2, rfl),
                 (-kap * 2.0 * s20 * c20 / den, np.array([0.0, c2, -s2]), k1, np.array([0.0, c2, -s2]))]
    elif wave == "rayleigh":
        # surface wave of the half-space (harela_incident_field.f90:576-611): gamma = (c_R / c_2)^2 from Rayleigh's cubic g^3 - 8 g^2 + 8 (3 - 2 kap^2) g - 16 (1 - kap^2) = 0
        # (the reference evaluates Cardano's closed form, `obtenerc`, and between nu = 0.395 and 0.405 takes the six-digit constant 0.887732: kept); a shear-type and
        # a dilatational-type partial wave with the horizontal wavenumber k_R = k_2 / sqrt(gamma), both decaying with depth, amplitudes 1 and -2 / (2 - gamma)
        if space != "half-space":
            raise ValueError("incident wave: a Rayleigh wave exists only in the half-space")
        if 0.395 < mat.nu_r < 0.405:
            gs = 0.887732
        else:
            roots = np.roots([1.0, -8.0, 8.0 * (3.0 - 2.0 * kap * kap), -16.0 * (1.0 - kap * kap)])
            gs = float(min(r.real for r in roots if abs(r.imag) < 1e-9 and 0.0 < r.real < 1.0))
        gp = kap * kap * gs
        kr = k2 / np.sqrt(gs)
        q1, q2 = 1j * np.sqrt(1.0 - gs), 1j * np.sqrt(1.0 - gp)
        local = [(1.0, np.array([0.0, c0, 1j / np.sqrt(1.0 - gs)]), kr, np.array([0.0, c0, q1])),
                 (-2.0 / (2.0 - gs), np.array([0.0, c0, q2]), kr, np.array([0.0, c0, q2]))]
    else:
        raise ValueError("incident wave: elastic wave type %r (p, sv, sh, rayleigh)" % (wave,))
    if space == "full-space":
        local = local[:1]
    elif space != "half-space":
        raise ValueError("incident wave: space %r of an elastic plane wave (the layered half-space is not covered)" % (space,))
    R = np.array([[np.cos(varphi), np.sin(varphi), 0.0], [-np.sin(varphi), np.cos(varphi), 0.0], [0.0, 0.0, 1.0]])     # local (in-plane y, z) -> global axes
    waves = [(0.5 * a, R @ pol, kk * (R @ pdir)) for a, pol, kk, pdir in local]
    waves = _mirror_parts(waves, (0, symconf_y, 0), vector=True)
    org = np.array([0.0, 0.0, z_fs])
    lam, mu = mat.lam, mat.mu

    def field(x, n):                                   # one point or a batch, as above
        d = np.asarray(x, dtype=np.float64) - org
        nn = np.asarray(n, dtype=np.float64)
        u, grad = np.zeros(d.shape, dtype=np.complex128), np.zeros(d.shape + (3,), dtype=np.complex128)
        for a, pol, kv in waves:
            e = a * np.exp(-1j * (d @ kv))
            u = u + e[..., None] * pol; grad = grad + e[..., None, None] * np.outer(pol, -1j * kv)     # grad[..., i, j] = du_i / dx_j
        tr = np.trace(grad, axis1=-2, axis2=-1)
        sigma = lam * tr[..., None, None] * np.eye(3) + mu * (grad + np.swapaxes(grad, -1, -2))
        return u, np.einsum("...ij,...j->...i", sigma, nn)
    return field


def element_incident_of(node_x, etype, elem_ptr, elem_node, elem_reversed, field, ndof):
    """element_incident / element_incident_fluid on flat arrays (a single-region model or the view of one region of a coupled model):
    -> (u_inc, t_inc), (sum nn, ndof) complex each."""
    X, N = element_node_geometry(node_x, etype, elem_ptr, elem_node, elem_reversed)
    return field_at(field, X, N, ndof)


def element_node_geometry(node_x, etype, elem_ptr, elem_node, elem_reversed):
    """(X, N), (sum nn, 3) each: position of every element node and the region's outward unit normal of its element there (element()%x_fn, n_fn with the
    reversal applied) -- the part of the incident arrays that does not depend on the frequency."""
    n_rows = int(elem_ptr[-1])
    X = np.zeros((n_rows, 3)); N = np.zeros((n_rows, 3))
    for e in range(len(etype)):
        et = int(etype[e]); c = np.asarray(elem_node[elem_ptr[e]:elem_ptr[e + 1]]); xn = node_x[c]
        sgn = -1.0 if elem_reversed[e] else 1.0
        for kn in range(len(c)):
            X[elem_ptr[e] + kn] = xn[kn]; N[elem_ptr[e] + kn] = sgn * sh.unit_normal(et, xn, sh.XI_NODES[et][kn])
    return X, N


def field_at(field, X, N, ndof):
    """`field` over the batch (the *_reference fields take batches; any other callable is applied point by point) -> (primary, secondary), (m, ndof) each."""
    try:
        u, t = field(X, N)
        u, t = np.asarray(u, dtype=np.complex128).reshape(len(X), ndof), np.asarray(t, dtype=np.complex128).reshape(len(X), ndof)
    except (ValueError, TypeError):
        u = np.zeros((len(X), ndof), dtype=np.complex128); t = np.zeros((len(X), ndof), dtype=np.complex128)
        for i in range(len(X)):
            u[i], t[i] = field(X[i], N[i])
    return u, t
