"""Nodal-solutions file (*.nso) of the reference for the regions this library covers.

Layout of src/export_solution_mechanics_harmonic_nso.f90 (header :62-150, rows :228-300) and
src/export_solution_mechanics_static_nso.f90 for one BE region with ordinary boundaries:
  harmonic row: kf, frequency (Hz or rad/s as in the case file), region id / class (1 = BE) / type (1 fluid, 2 elastic),
                boundary id / class (1 = ordinary) / face (1), node id, x1 x2 x3, then the total field value_c(k_start:k_end) as
                (Re, Im) or (|.|, arg) pairs -- fluid: p, Un; elastic: u1 u2 u3 t1 t2 t3 -- then the incident field in the same order (zero without [incident waves]).
  static row:   0, 0.0, the same identification columns, then u1 u2 u3 t1 t2 t3 (real).
Rows follow the region's boundary list; inside a boundary the nodes appear in first-visit order of the part's elements.
"""
import datetime
import numpy as np

from .fortran_format import RealFormat, fmt_int, int_width

MULTIFEBE_VERSION = "2.0.1"      # the reference release this layout follows
N_COLUMNS_3D = 44                # ncmax of the harmonic writer for problem%n = 3


class NsoWriter:
    def __init__(self, fh, case, model):
        self.f, self.case, self.m = fh, case, model
        self.rf = RealFormat(case.real_format)
        ids = [len(case.omega), max(r[0] for r in case.regions), max(b for b, _ in case.boundaries), max([i_ for i_, _, _ in case.internal_points] or [0]), int(model.mesh.elem_ids.max()), int(model.mesh.node_ids.max())]
        if case.integer_format in (None, "auto"):
            self.wi = int_width(*ids)
        elif case.integer_format == "max":
            self.wi = 11
        else:
            self.wi = int(case.integer_format.lstrip("i"))
        # rows: (region index, region id, region type, boundary id, face, node index) in the reference's order; face = 2 when the region is
        # region 2 of a be-be boundary (export_solution_mechanics_harmonic_nso.f90:326-333)
        part_of_boundary = dict(case.boundaries)
        only = getattr(case, "nso_nodes", None)
        self.rows = []
        for kr, (rid, rtype, _, rb) in enumerate(case.regions):
            for sb in rb:
                b = abs(sb)
                face = 2 if sb < 0 else 1
                seen = set()
                for e in range(model.n_elem):
                    if int(model.mesh.part[e]) != part_of_boundary[b]:
                        continue
                    for v in model.mesh.conn[e]:
                        if int(v) not in seen:
                            seen.add(int(v))
                            if only is None or int(model.mesh.node_ids[v]) in only:       # [export] nso_nodes: node()%export
                                self.rows.append((kr, rid, rtype, b, face, int(v)))

    def _i(self, n):
        return fmt_int(n, self.wi)

    def header(self):
        c, w = self.case, self.f.write
        w("# Program      : multifebe\n")
        w("# Version      : %-5s\n" % MULTIFEBE_VERSION[:5])
        w("# File_format  : nso\n")
        w("# Problem_dim  : 3\n")
        w("# Input_file   : %s\n" % c.filename)
        w("# Description  : %s\n" % c.description)
        w("# Timestamp    : %s\n" % datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S.%f")[:23])
        w("\n")
        w("# Columns  Description\n")
        if c.analysis == "harmonic":
            w("# C1-C2    Frequency index and value %s.\n" % ("f (Hz)" if c.frequency_units == "f" else "w (rad/s)"))
        else:
            w("# C1-C2    Step index and value.\n")
        w("# C3-C5    Region id, class and type.\n")
        w("# C6-C8    (if C4 == 1) Boundary id, class and face.\n" if c.analysis == "harmonic" else "# C6-C8    (if col C4 == 1) Boundary id, class and face.\n")
        w("# C6-C8    (if C4 == 2) Subregion id, number of DOF and 0.\n" if c.analysis == "harmonic" else "# C6-C8    (if col C4 == 2) Subregion id, number of DOF and 0.\n")
        w("# C9-C12   Node id, x1, x2 and x3.\n")
        w("# >=C13    Node variables. Depend on the region class and type (see documentation).\n")
        w("#\n")
        if c.analysis != "harmonic":
            return
        w("# Complex notation: %s\n" % c.complex_notation)
        w("#\n")
        line = "#" + "_" * (self.wi - 3) + "C1"
        for kc in range(2, N_COLUMNS_3D + 1):
            nc = self.wi if 3 <= kc <= 9 else self.rf.w
            line += "_" * (nc - len(str(kc)) - 1) + "C%d" % kc
        w(line + "\n")

    def _ident(self, kf, value, row):
        kr, rid, rtype, b, face, v = row
        x = self.m.node_x[v]
        return (self._i(kf) + self.rf(value) + self._i(rid) + self._i(1) + self._i(rtype) + self._i(b) + self._i(1) +
                self._i(face) + self._i(int(self.m.mesh.node_ids[v])) + "".join(self.rf(t) for t in x))

    def _nodal(self, x):
        """Per region: (primary, secondary) variables as (n_node, nvar) arrays."""
        out = []
        for kr in range(len(self.case.regions)):
            prim, sec = self.m.nodal_solution(np.asarray(x), kr) if self.case.multi else self.m.nodal_solution(np.asarray(x))
            out.append((np.asarray(prim).reshape(self.m.n_node, -1), np.asarray(sec).reshape(self.m.n_node, -1)))
        return out

    def _cpair(self, z):
        if self.case.complex_notation == "polar":
            return self.rf(abs(z)) + self.rf(float(np.angle(z)))
        return self.rf(z.real) + self.rf(z.imag)

    def frequency(self, kf, x):
        """Rows of frequency index kf (1-based) from the solution vector x of that frequency."""
        c = self.case
        omega = c.omega[kf - 1]
        value = omega * 0.159154943091895335768883763373 if c.frequency_units == "f" else omega   # c_1_2pi
        nodal = self._nodal(x)
        incident = self._nodal_incident(omega)
        out = []
        for row in self.rows:
            prim, sec = nodal[row[0]]
            v = row[5]
            vals = "".join(self._cpair(complex(z)) for z in list(prim[v]) + list(sec[v]))
            if row[0] in incident:
                pi, si = incident[row[0]]
                tail = "".join(self._cpair(complex(z)) for z in list(pi[v]) + list(si[v]))
            else:
                tail = self._cpair(0j) * (2 * prim.shape[1])
            out.append(self._ident(kf, value, row) + vals + tail + "\n")
        self.f.write("".join(out))

    def _nodal_incident(self, omega):
        """{region index: (primary, secondary)} incident field at the nodes, (n_node, nvar) each: node()%incident_c, the mean of element()%incident_c over
        the elements of the node (src/calculate_incident_mechanics_harmonic.f90:628-638); regions without incident fields are absent (zeros in the file)."""
        c = self.case
        if not any(getattr(c, "region_incident", ())):
            return {}
        out = {}
        for kr, (u, t) in c.incident_arrays(self.m, omega).items():
            v = self.m.views[kr] if c.multi else self.m
            nv = u.shape[1]
            acc = np.zeros((2, self.m.n_node, nv), dtype=np.complex128); cnt = np.zeros(self.m.n_node)
            nodes = np.asarray(v.elem_node[:int(v.elem_ptr[-1])])
            np.add.at(acc[0], nodes, u); np.add.at(acc[1], nodes, t); np.add.at(cnt, nodes, 1.0)
            cnt[cnt == 0] = 1.0
            out[kr] = (acc[0] / cnt[:, None], acc[1] / cnt[:, None])
        return out

    def _ip_ident(self, kf, value, pid, x):
        """Columns 1-12 of an internal-point row: no boundary (0 0 0), the point id and its position (export_solution_mechanics_harmonic_nso.f90:401-405)."""
        rid, rtype = self.case.regions[0][0], self.case.regions[0][1]
        return (self._i(kf) + self.rf(value) + self._i(rid) + self._i(1) + self._i(rtype) + self._i(0) + self._i(0) + self._i(0) + self._i(pid) +
                "".join(self.rf(t) for t in x))

    def frequency_internal(self, kf, u, sigma):
        """Internal-point rows of frequency kf: u_k, then t_k on the planes with normals e_1, e_2, e_3 (sigma[p, k, kc]), then the incident field
        (zero) in the same order (:407-447)."""
        c = self.case
        omega = c.omega[kf - 1]
        value = omega * 0.159154943091895335768883763373 if c.frequency_units == "f" else omega
        out = []
        for p, (pid, _, xp) in enumerate(c.internal_points):
            vals = "".join(self._cpair(complex(z)) for z in u[p])
            for kc in range(3):
                vals += "".join(self._cpair(complex(sigma[p, k, kc])) for k in range(3))
            out.append(self._ip_ident(kf, value, pid, xp) + vals + self._cpair(0j) * 12 + "\n")
        self.f.write("".join(out))

    def static_internal(self, u, sigma):
        out = []
        for p, (pid, _, xp) in enumerate(self.case.internal_points):
            vals = "".join(self.rf(float(np.real(z))) for z in u[p])
            for kc in range(3):
                vals += "".join(self.rf(float(np.real(sigma[p, k, kc]))) for k in range(3))
            out.append(self._ip_ident(0, 0.0, pid, xp) + vals + "\n")
        self.f.write("".join(out))

    def static(self, x):
        u, t = self.m.nodal_solution(np.asarray(x, dtype=np.complex128))
        out = []
        for row in self.rows:
            v = row[5]
            vals = "".join(self.rf(float(z.real)) for z in list(u[v]) + list(t[v]))
            out.append(self._ident(0, 0.0, row) + vals + "\n")
        self.f.write("".join(out))


def read_nso(path):
    """Numeric rows of an *.nso file as a float array (comment lines skipped): for tests and post-processing."""
    rows = [[float(t) for t in s.split()] for s in open(path) if s.strip() and not s.startswith("#")]
    return np.array(rows)
