"""Reader of the reference's input (case) file for the analyses this library covers.

Mirrors src/read_input_file.f90 for the sections the hot path consumes:
  [problem]      src/read_problem.f90        n = 3D, type = mechanics, analysis = harmonic | static
  [frequencies]  src/read_frequencies.f90:20-120  Hz | rad/s; list | lin | log(10)
  [settings]     src/read_settings.f90:74-216     mesh_file_mode = 2 "<gmsh 2.2 file>" | 1 "<native file>" | 0 / absent ([nodes], [elements], [parts] of the
                                             case file), qsi_relative_error, qsi_ns_max, precalsets,
                                                  geometric_tolerance
  [materials]    src/read_materials.f90      fluid (two of K, rho, c; xi) / elastic_solid (two of E, nu, lambda, mu, K; rho, xi) /
                                             biot_poroelastic_medium (phi, two elastic constants, Q, R, rho_f, rho_s, rho_a, xi, b)
  [boundaries]   src/read_boundaries.f90     `<id> <part> ordinary`
  [regions]      src/read_regions.f90        `be` regions, full space (several regions sharing be-be boundaries: negative id = reversed), `material <id>` or the legacy in-line forms
                                             `fluid rho c`, `viscoelastic rho mu nu xi`, `elastic rho mu nu`
  [conditions over be boundaries]            src/read_conditions_bem_boundaries_mechanics_{harmonic,static}.f90: global-axes
                                             conditions 0 / 1 per component; defaults (not listed) = 1 with value 0
  [internal points]  src/read_internal_points.f90    `<id> <region> x1 x2 x3` (one elastic region)
  [symmetry planes]  src/read_symmetry_planes.f90    `plane_n1|plane_yz : symmetry|antisymmetry` (also plane_n2|plane_zx, plane_n3|plane_xy) or the
                                             explicit form `x = <s> <t1> <t2> <t3>` (y, z alike); one elastic region
  [bem formulation over boundaries]  src/read_bem_formulation_boundaries.f90   `boundary <id>: sbie | sbie_boundary_mca <delta> | sbie_mca <delta>`
  [incident waves]   src/read_incident_mechanics_harmonic.f90   plane / point waves in fluids, plane P / SV / SH / Rayleigh waves in elastic solids, full space or
                                             homogeneous half-space; listed per region by the last record of [regions] (`<n> <id> ...`)
  [export]       src/read_export.f90:61-293  export_nso, real_format, integer_format, complex_notation, nso_nodes
Anything else the reference accepts (be-fe coupling, crack-like boundaries, close-pore conditions, local-axes or spring conditions,
half-space fundamental solutions, body loads, layered / poroelastic incident fields, internal points of fluid regions, FE regions ...) raises
CaseFileError naming the feature:
the Fortran host keeps those.
"""
import os
import re
import numpy as np

from .mesh import read_gmsh22, read_native_mesh, without_parts
from .model import Model, FluidModel, PoroModel, Material, Fluid, Poro
from .fortran_format import DEFAULT_REAL_FORMAT, REAL_FORMATS


class CaseFileError(ValueError):
    pass


def _sections(text):
    """{section name: [non-empty lines]} (fbem_search_section: a line `[name]` opens a section)."""
    out, cur = {}, None
    for raw in text.replace("﻿", "").splitlines():
        s = raw.strip()
        m = re.fullmatch(r"\[(.+?)\]", s)
        if m:
            cur = m.group(1).strip().lower()
            out[cur] = []
        elif cur is not None and s:
            out[cur].append(s)
    return out


def _keyword(lines, key):
    """Value string after `key =` (fbem_search_keyword), or None."""
    for s in lines:
        m = re.match(r"\s*%s\s*=\s*(.*)$" % re.escape(key), s)
        if m:
            return m.group(1).strip()
    return None


def _fortran_float(tok):
    return float(tok.lower().replace("d", "e"))


def _fortran_complex(s):
    """`(re,im)` list-directed complex, or a bare real."""
    m = re.match(r"\(\s*([^,\s]+)\s*,\s*([^)\s]+)\s*\)", s.strip())
    if m:
        return complex(_fortran_float(m.group(1)), _fortran_float(m.group(2)))
    return complex(_fortran_float(s.split()[0]))


def _logical(s):
    t = s.strip().lower().strip(".")
    if t in ("t", "true"):
        return True
    if t in ("f", "false"):
        return False
    raise CaseFileError("invalid logical value %r" % s)


def elastic_constants(given):
    """Two of E, nu, lambda, mu, K -> all five (fbem_ela_properties, lib/fbem/src/harela_incident_field.f90:889-966)."""
    g = dict(given)
    if len(g) != 2:
        raise CaseFileError("only 2 elastic constants are needed")
    E, nu, lam, mu, K = (g.get(k) for k in ("E", "nu", "lambda", "mu", "K"))
    if K is not None and E is not None:
        lam = 3.0 * K * (3.0 * K - E) / (9.0 * K - E); mu = 3.0 * K * E / (9.0 * K - E); nu = (3.0 * K - E) / (6.0 * K)
    elif K is not None and lam is not None:
        E = 9.0 * K * (K - lam) / (3.0 * K - lam); mu = 1.5 * (K - lam); nu = lam / (3.0 * K - lam)
    elif K is not None and mu is not None:
        E = 9.0 * K * mu / (3.0 * K + mu); lam = K - 2.0 / 3.0 * mu; nu = 0.5 * (3.0 * K - 2.0 * mu) / (3.0 * K + mu)
    elif K is not None and nu is not None:
        E = 3.0 * K * (1.0 - 2.0 * nu); lam = 3.0 * K * nu / (1.0 + nu); mu = 1.5 * K * (1.0 - 2.0 * nu) / (1.0 + nu)
    elif E is not None and lam is not None:
        r = np.sqrt(E ** 2 + 9.0 * lam ** 2 + 2.0 * E * lam)
        K = (E + 3.0 * lam + r) / 6.0; mu = (E - 3.0 * lam + r) / 4.0; nu = 2.0 * lam / (E + lam + r)
    elif E is not None and mu is not None:
        K = E * mu / (3.0 * mu - E) / 3.0; lam = mu * (E - 2.0 * mu) / (3.0 * mu - E); nu = 0.5 * E / mu - 1.0
    elif E is not None and nu is not None:
        K = E / (1.0 - 2.0 * nu) / 3.0; lam = E * nu / (1.0 + nu) / (1.0 - 2.0 * nu); mu = 0.5 * E / (1.0 + nu)
    elif lam is not None and mu is not None:
        K = lam + 2.0 * mu / 3.0; E = mu * (3.0 * lam + 2.0 * mu) / (lam + mu); nu = 0.5 * lam / (lam + mu)
    elif lam is not None and nu is not None:
        K = lam * (1.0 + nu) / (3.0 * nu); E = lam / nu * (1.0 + nu) * (1.0 - 2.0 * nu); mu = 0.5 * lam * (1.0 - 2.0 * nu) / nu
    else:
        K = 2.0 * mu * (1.0 + nu) / (1.0 - 2.0 * nu) / 3.0; E = 2.0 * mu * (1.0 + nu); lam = 2.0 * mu * nu / (1.0 - 2.0 * nu)
    return {"E": E, "nu": nu, "lambda": lam, "mu": mu, "K": K}


def read_frequencies(lines):
    """(omega[rad/s] array, units 'f' | 'w'): src/read_frequencies.f90.  The list is stored in rad/s (Hz values times 2 pi)."""
    units = {"hz": "f", "rad/s": "w"}.get(lines[0].lower())
    if units is None:
        raise CaseFileError('the frequency units can be only "Hz" or "rad/s"')
    mode = {"list": 1, "lin": 2, "log10": 3, "log": 3}.get(lines[1].lower(), 0)
    if mode == 0:
        raise CaseFileError('the frequency specification mode is "list", "lin" or "log"')
    n = int(lines[2].split()[0])
    if mode == 1:
        if n <= 0:
            raise CaseFileError("the number of frequencies must be >=1")
        # list-directed reads: one value per record
        f = np.array([_fortran_float(lines[3 + i].split()[0]) for i in range(n)])
    else:
        if n < 2:
            raise CaseFileError("the number of frequencies must be >=2")
        f = np.zeros(n)
        f[0] = _fortran_float(lines[3].split()[0]); f[-1] = _fortran_float(lines[4].split()[0])
        if f[0] >= f[-1]:
            raise CaseFileError("the range of frequencies is invalid")
        if mode == 2:
            delta = (f[-1] - f[0]) / float(n - 1)
            for i in range(1, n - 1):
                f[i] = f[0] + delta * float(i)
        else:
            delta = (np.log10(f[-1]) - np.log10(f[0])) / float(n - 1)
            for i in range(1, n - 1):
                f[i] = 10.0 ** (np.log10(f[0]) + delta * float(i))
    if units == "f":
        f = 6.28318530717958623199592693709 * f      # c_2pi of lib/fbem/src/numerical.f90
    return f, units


class CaseFile:
    """Parsed case: .analysis ('harmonic' | 'static'), .omega, .frequency_units, .mesh, .material (Material | Fluid), .region_type
    (1 fluid, 2 elastic: the region type codes of the result files), .boundaries [(id, part)], .bcs, settings and export options."""

    def __init__(self, path):
        self.path = os.path.abspath(path)
        self.dir = os.path.dirname(self.path)
        self.filename = os.path.basename(path)
        sec = _sections(open(path, encoding="utf-8", errors="replace").read())
        self._sec = sec
        for unsupported in ("fe subregions", "be body loads", "be bodyloads", "internal elements", "groups", "cross sections", "sensitivity",
                            "conditions over nodes", "conditions over be bodyloads", "conditions over fe elements", "discontinuous be boundaries",
                            "discontinuous be elements", "element options", "fem node options", "special be elements", "commands", "export sif",
                            "internal points from mesh", "geometry"):
            if sec.get(unsupported):
                raise CaseFileError("section [%s] is outside the path this library covers" % unsupported)
        # ---- [problem]
        pb = sec.get("problem")
        if pb is None:
            raise CaseFileError("[problem]: this section is required")
        if (_keyword(pb, "n") or "").upper() != "3D":
            raise CaseFileError("[problem] n: only 3D problems are covered")
        if (_keyword(pb, "type") or "mechanics").lower() != "mechanics":
            raise CaseFileError("[problem] type: only `mechanics` is covered")
        self.analysis = (_keyword(pb, "analysis") or "").lower()
        if self.analysis not in ("harmonic", "static"):
            raise CaseFileError("[problem] analysis: `harmonic` or `static`")
        self.description = _keyword(pb, "description") or ""
        # ---- [settings]
        st = sec.get("settings", [])
        # mesh_file_mode (src/read_settings.f90:219-246): absent or 0 = the [nodes] / [elements] / [parts] sections of the case file itself, 1 = the same
        # sections in an auxiliary file (native format), 2 = a Gmsh 2.2 file
        v = _keyword(st, "mesh_file_mode")
        self.mesh_file_mode, self.mesh_file = 0, None
        if v:
            m = re.match(r'(\d+)\s*(?:"([^"]*)"|(\S+))?', v)
            if not m or int(m.group(1)) not in (0, 1, 2):
                raise CaseFileError('[settings] mesh_file_mode = <0, 1, 2> "<mesh file>": wrong type of mesh mode')
            self.mesh_file_mode = int(m.group(1))
            if self.mesh_file_mode:
                name = (m.group(2) or m.group(3) or "").strip()
                if not name:
                    raise CaseFileError('[settings] mesh_file_mode = %d needs the name of the mesh file' % self.mesh_file_mode)
                self.mesh_file = os.path.join(self.dir, name)
        self.qsi_relative_error = _fortran_float(_keyword(st, "qsi_relative_error") or "1e-6")
        self.qsi_ns_max = int(_keyword(st, "qsi_ns_max") or 16)
        pc = _keyword(st, "precalsets")
        self.precalset_gln = tuple(int(t) for t in pc.split()[1:1 + int(pc.split()[0])]) if pc else (2, 3, 4, 5, 6, 7, 8, 9)
        self.geometric_tolerance = _fortran_float(_keyword(st, "geometric_tolerance") or "1e-6")
        self.collapse_nodal_pos = _logical(_keyword(st, "collapse_nodal_pos") or "T")      # default T (src/read_settings.f90:168-177)
        for k in ("lse_scaling", "lse_condition", "lse_refine"):
            if _keyword(st, k) and _logical(_keyword(st, k)):
                raise CaseFileError("[settings] %s = T is not covered (plain zgesv / dgesv)" % k)
        # ---- [frequencies]
        self.omega, self.frequency_units = (np.zeros(0), "w")
        if self.analysis == "harmonic":
            if _keyword(st, "frequencies_file"):
                raise CaseFileError("[settings] frequencies_file is not covered: put the list in [frequencies]")
            if "frequencies" not in sec:
                raise CaseFileError("[frequencies]: this section is required")
            self.omega, self.frequency_units = read_frequencies(sec["frequencies"])
        # ---- [materials]
        self.materials = {}
        ml = sec.get("materials", [])
        if ml:
            for s in ml[1:1 + int(ml[0].split()[0])]:
                w = s.split()
                props = {w[2 + 2 * j]: _fortran_float(w[3 + 2 * j]) for j in range((len(w) - 2) // 2)}
                if (len(w) - 2) % 2:
                    raise CaseFileError("material %s: wrong number of arguments" % w[0])
                self.materials[int(w[0])] = (w[1], props)
        # ---- [boundaries]
        bl = sec.get("boundaries")
        if not bl:
            raise CaseFileError("[boundaries]: this section is required")
        self.boundaries = []
        for s in bl[1:1 + int(bl[0].split()[0])]:
            w = s.split()
            if len(w) < 3 or w[2].lower() != "ordinary":
                raise CaseFileError("boundary %s: only `ordinary` boundaries are covered (crack-like boundaries stay with the Fortran host)" % w[0])
            self.boundaries.append((int(w[0]), int(w[1])))
        # ---- [bem formulation over boundaries]: `boundary <id>: sbie | sbie_boundary_mca <delta> | sbie_mca <delta>` (read_bem_formulation_selected_nodes.f90:74-160)
        self.formulation = {}
        for s_ in sec.get("bem formulation over boundaries", []):
            m = re.match(r"(boundary|part)\s+(\d+)\s*:\s*(\S+)\s*(\S+)?", s_, re.I)
            if not m:
                raise CaseFileError("[bem formulation over boundaries]: cannot parse %r" % s_)
            if m.group(1).lower() != "boundary":
                raise CaseFileError("[bem formulation over boundaries]: only `boundary <id>: ...` records are covered")
            bid, kind = int(m.group(2)), m.group(3).lower()
            if bid not in dict(self.boundaries):
                continue                                 # the reference looks the listed boundaries up and ignores anything else
            if kind not in ("sbie", "sbie_boundary_mca", "sbie_mca"):
                raise CaseFileError("boundary %d: BEM formulation %r is not covered (sbie, sbie_boundary_mca, sbie_mca; the hypersingular and dual formulations "
                                    "stay with the Fortran host)" % (bid, kind))
            if kind != "sbie" and not m.group(4):
                raise CaseFileError("boundary %d: %s needs its delta (<= 0: the default)" % (bid, kind))
            self.formulation[bid] = (kind, _fortran_float(m.group(4)) if kind != "sbie" else 0.0)
        # ---- [regions]
        rl = sec.get("regions")
        if not rl:
            raise CaseFileError("[regions]: this section is required")
        n_regions = int(rl[0].split()[0])
        self.regions = []                     # (id, type code 1 fluid / 2 elastic, material, signed boundary ids)
        self.region_incident = []             # per region: ids of its incident fields
        k = 1
        for kr_ in range(n_regions):
            if k + 2 >= len(rl):
                raise CaseFileError("[regions]: %d regions announced, the records of region number %d are missing" % (n_regions, kr_ + 1))
            w = rl[k].split()
            rid = int(w[0])
            if w[1].lower() != "be":
                raise CaseFileError("region %d: only `be` regions are covered" % rid)
            if len(w) > 2 and w[2].lower() != "full-space":
                raise CaseFileError("region %d: only the full-space fundamental solution is covered" % rid)
            w = [int(t) for t in rl[k + 1].split()]
            rb = w[1:1 + w[0]]
            material, rtype = self._material(rid, rl[k + 2].split())
            k += 3
            # remaining records of the region: number of BE body loads (must be 0) and, in the harmonic analysis, `<n> <id> ...` of its incident
            # fields (src/read_regions.f90:1327-1330)
            n_tail = 2 if self.analysis == "harmonic" else 1
            if k < len(rl) and not re.fullmatch(r"[\s0]+", rl[k]):
                raise CaseFileError("region %d: BE body loads are not covered" % rid)
            inc_ids = []
            if n_tail == 2 and k + 1 < len(rl):
                w = rl[k + 1].split()
                try:
                    inc_ids = [int(t) for t in w[1:1 + int(w[0])]]
                    if len(inc_ids) != int(w[0]):
                        raise ValueError
                except ValueError:
                    raise CaseFileError("region %d: the incident fields record is `<n> <id 1> ... <id n>`" % rid)
            k += n_tail
            self.regions.append((rid, rtype, material, rb))
            self.region_incident.append(inc_ids)
        self.multi = n_regions > 1
        listed = [abs(b) for r in self.regions for b in r[3]]
        if not self.multi and any(b < 0 for b in self.regions[0][3]):
            # a negative id = the boundary is seen reversed from this region (region 2 of a be-be boundary)
            raise CaseFileError("region %d: reversed boundaries belong to multi-region models" % self.regions[0][0])
        if self.multi and self.analysis == "static":
            raise CaseFileError("[regions]: one BE region is covered in the static analysis")
        self.region_id, self.region_type, self.material, self.region_boundaries = self.regions[0]
        self.incident_fields = self._incident_waves(sec.get("incident waves", []))
        self.interfaces = sorted(b for b in set(listed) if listed.count(b) == 2)
        if self.analysis == "static" and self.region_type != 2:
            raise CaseFileError("static analysis: only elastic solids are covered")
        # ---- [conditions over be boundaries]
        region_of_boundary = {}
        for rid, rtype, _, rb in self.regions:
            for b in rb:
                if b > 0:
                    region_of_boundary[b] = rtype
        for b, _ in self.boundaries:
            if b not in region_of_boundary:
                raise CaseFileError("boundary %d is not used with a positive id by any region" % b)
        ndof_of = {b: {1: 1, 2: 3, 3: 4}[region_of_boundary[b]] for b, _ in self.boundaries}
        self.bcs = {bid: ([1] * ndof_of[bid], [0j] * ndof_of[bid]) for bid, _ in self.boundaries if bid not in self.interfaces}   # defaults: t / Un = 0
        cl = sec.get("conditions over be boundaries", [])
        i = 0
        while i < len(cl):
            m = re.match(r"boundary\s+(\d+)\s*:\s*(.*)$", cl[i], re.I)
            if not m:
                raise CaseFileError("[conditions over be boundaries]: cannot parse %r" % cl[i])
            bid = int(m.group(1))
            if bid in self.interfaces:
                raise CaseFileError("boundary %d: conditions over be-be boundaries other than perfect bonding / contact (the default) are not covered" % bid)
            if bid not in self.bcs:
                raise CaseFileError("[conditions over be boundaries]: unknown boundary %d" % bid)
            ndof = ndof_of[bid]
            recs = [m.group(2)] + cl[i + 1:i + ndof]
            ct, cv = [], []
            w0 = recs[0].split(None, 1)
            if int(w0[0]) == 10:      # one record: normal pressure t_k = P n_k on the three components (read_conditions_bem_boundaries_mechanics_harmonic.f90:144-149)
                if ndof != 3 or self.multi:
                    raise CaseFileError("boundary %d: condition type 10 (normal pressure) is covered for one elastic region" % bid)
                pval = _fortran_complex(w0[1])
                self.bcs[bid] = ([10, 10, 10], [pval, pval, pval])
                i += 1
                continue
            for k in range(ndof):
                w = recs[k].split(None, 1)
                t = int(w[0])
                solid_single = ndof == 3 and not self.multi
                if t == 4 and solid_single:
                    # 4: infinitesimal rotation field u_k = theta (axis x (x - center)) . e_k: `4 cx cy cz ax ay az theta` (read_conditions_bem_boundaries_mechanics_harmonic.f90:129-143)
                    q = w[1].replace(",", " ").replace("(", " ").replace(")", " ").split()
                    if len(q) < 7:
                        raise CaseFileError("boundary %d: condition type 4 needs center(3), axis(3), theta" % bid)
                    num = [_fortran_float(z) for z in q]
                    theta = complex(num[6], num[7]) if len(num) >= 8 else complex(num[6])
                    ct.append(4); cv.append((num[0:3], num[3:6], theta))
                    continue
                if t not in (0, 1) and not (t in (2, 3) and solid_single):
                    raise CaseFileError("boundary %d: condition type %d is not covered (0: primary variable known, 1: secondary variable known; on one elastic region also "
                                        "2 / 3: local axes, 4: rotation field, 10: normal pressure)" % (bid, t))
                ct.append(t); cv.append(_fortran_complex(w[1]))
            if any(t in (2, 3) for t in ct) and not all(t in (2, 3) for t in ct):
                raise CaseFileError("boundary %d: local axes B.C. and global axes B.C. can not be mixed" % bid)
            self.bcs[bid] = (ct, cv)
            i += ndof
        # ---- [internal points] (src/read_internal_points.f90: `<id> <region> x1 x2 x3`): displacements and stresses inside an elastic region
        self.internal_points = []             # (id, region index, x)
        il = sec.get("internal points", [])
        if il:
            if self.multi or self.region_type != 2:
                raise CaseFileError("[internal points]: covered for one elastic region (a fluid region would need the hypersingular fluid kernels)")
            if any(self.region_incident):
                # the interior identities would need the incident field at the points and its terms on the interior rows (calculate_internal_points_mechanics_bem_harela.f90
                # adds them); not wired, and silently wrong otherwise
                raise CaseFileError("[internal points] together with [incident waves] is not covered")
            for s_ in il[1:1 + int(il[0].split()[0])]:
                w = s_.split()
                if int(w[1]) != self.region_id:
                    raise CaseFileError("internal point %s: the indicated region does not exist" % w[0])
                self.internal_points.append((int(w[0]), 0, np.array([_fortran_float(t) for t in w[2:5]])))
            if len(set(i_ for i_, _, _ in self.internal_points)) != len(self.internal_points) or min(i_ for i_, _, _ in self.internal_points) <= 0:
                raise CaseFileError("[internal points]: identifiers must be positive and unique")
        # ---- [export]
        ex = sec.get("export", [])
        self.export_nso = _logical(_keyword(ex, "export_nso") or "T")
        self.real_format = (_keyword(ex, "real_format") or DEFAULT_REAL_FORMAT).strip("'\" ").lower()
        self.real_format = REAL_FORMATS.get(self.real_format, self.real_format)
        self.integer_format = (_keyword(ex, "integer_format") or "").strip("'\" ").lower() or None
        cn = (_keyword(ex, "complex_notation") or "cartesian").strip("'\" ").lower()
        if cn not in ("polar", "cartesian"):
            raise CaseFileError("[export] complex_notation: polar or cartesian")
        self.complex_notation = cn
        # nso_nodes = <n> <id 1> ... <id n>: rows only for these nodes; n <= 0 or absent: every node (src/read_export.f90:254-293)
        self.nso_nodes = None
        v = _keyword(ex, "nso_nodes")
        if v:
            w = [int(t) for t in v.split()]
            if w[0] > 0:
                if len(w) < 1 + w[0]:
                    raise CaseFileError("[export] nso_nodes: %d nodes announced, %d given" % (w[0], len(w) - 1))
                if len(set(w[1:1 + w[0]])) != w[0]:
                    raise CaseFileError("[export] nso_nodes: there are repeated nodes for export")
                self.nso_nodes = set(w[1:1 + w[0]])
        # ---- [symmetry planes] (src/read_symmetry_planes.f90:76-228): the explicit multipliers first, then the named kinds, x before y before z
        self.symmetry = []
        sp = sec.get("symmetry planes") or []
        if sp:
            for ax, names in (("x", ("plane_n1", "plane_yz")), ("y", ("plane_n2", "plane_zx")), ("z", ("plane_n3", "plane_xy"))):
                given = []
                v = _keyword(sp, ax)
                if v is not None:
                    w = v.split()
                    try:
                        t = [int(q) for q in w[0:4]]
                    except ValueError:
                        t = []
                    if len(t) != 4 or any(abs(q) != 1 for q in t):
                        raise CaseFileError("[symmetry planes] %s = <s> <t1> <t2> <t3>: multipliers must be +1 or -1" % ax)
                    given.append(tuple(float(q) for q in t))
                for nm in names:
                    for line in sp:
                        mm = re.match(r"\s*%s\s*:\s*(\S+)" % nm, line)
                        if mm:
                            kind = mm.group(1).strip("'\"").lower()
                            if kind not in ("symmetry", "antisymmetry"):
                                raise CaseFileError("[symmetry planes] %s: symmetry or antisymmetry" % nm)
                            given.append(kind)
                    if given and isinstance(given[-1], str):
                        break      # plane_yz is looked for only when plane_n1 is absent
                if len(given) > 1:
                    raise CaseFileError("[symmetry planes]: plane %s is given twice" % ax)
                if given:
                    self.symmetry.append((ax, given[0]))
        # ---- mesh: parts are the physical groups of the Gmsh file; a boundary is one part
        if self.mesh_file_mode == 2:
            self.mesh = read_gmsh22(self.mesh_file)
        else:
            msec = sec if self.mesh_file_mode == 0 else _sections(open(self.mesh_file, encoding="utf-8", errors="replace").read())
            if not msec.get("nodes") or not msec.get("elements"):
                raise CaseFileError("the mesh: sections [nodes] and [elements] are required (%s)" % ("case file" if self.mesh_file_mode == 0 else self.mesh_file))
            try:
                self.mesh = read_native_mesh(msec["nodes"], msec["elements"])
            except (ValueError, KeyError, IndexError) as e:
                raise CaseFileError("the mesh: %s" % e)
        # elements of parts that no boundary uses are left out, as the reference does (src/read_elements.f90:211-215: part()%entity = 0)
        part_of_boundary = dict(self.boundaries)
        used_parts = set(part_of_boundary[abs(b)] for r in self.regions for b in r[3])
        unused = set(int(p_) for p_ in self.mesh.part) - used_parts
        if unused:
            self.mesh = without_parts(self.mesh, unused)
        if self.mesh.n_elem == 0:
            raise CaseFileError("the mesh holds no surface element of the parts of the boundaries")

    def _incident_waves(self, lines):
        """[incident waves] (src/read_incident_mechanics_harmonic.f90:22-42): per field `<id>`, `<class>`, `<space> [np xp bc]`,
        `<variable> <amplitude> <x0(3)> <varphi> <theta>` (degrees), `<xs(3)> <symconf(3)>`, `<region type> <wave type>`.  Covered: plane and
        point waves in fluids, plane P / SV / SH waves in elastic solids, full space or homogeneous half-space; the restrictions of the reference
        (:290-347) are kept.  -> {id: dict}"""
        fields = {}
        if not lines or self.analysis != "harmonic":
            if any(self.region_incident):
                raise CaseFileError("[regions]: incident fields are listed, but there is no [incident waves] section (harmonic analysis)")
            return fields
        n = int(lines[0].split()[0])
        k = 1
        for _ in range(n):
            if k + 6 > len(lines):
                raise CaseFileError("[incident waves]: %d fields announced, records are missing" % n)
            fid = int(lines[k].split()[0])
            if fid <= 0:
                raise CaseFileError("incident wave %d: the identifier must be greater than 0" % fid)
            cls = lines[k + 1].split()[0].lower()
            if cls not in ("point", "plane"):
                raise CaseFileError("incident wave %d: class %r is not covered (point, plane)" % (fid, cls))
            w = lines[k + 2].split()
            f = dict(cls=cls, space=w[0].lower(), np=3, xp=0.0, bc=1)
            if f["space"] == "half-space":
                f["np"], f["xp"], f["bc"] = int(w[1]), _fortran_float(w[2]), int(w[3])
                if not 1 <= f["np"] <= 3:
                    raise CaseFileError("incident wave %d: 1 <= np <= 3" % fid)
            elif f["space"] != "full-space":
                raise CaseFileError("incident wave %d: space %r is not covered (full-space, half-space)" % (fid, w[0]))
            m = re.match(r"\s*(\d+)\s+(\([^)]*\)|\S+)\s+(.*)$", lines[k + 3])
            if not m:
                raise CaseFileError("incident wave %d: cannot parse %r" % (fid, lines[k + 3]))
            f["variable"], f["amplitude"] = int(m.group(1)), _fortran_complex(m.group(2))
            num = [_fortran_float(t) for t in m.group(3).split()]
            if len(num) < 5:
                raise CaseFileError("incident wave %d: <variable> <amplitude> <x0(3)> <varphi> <theta> expected" % fid)
            f["x0"], f["varphi"], f["theta"] = np.array(num[0:3]), np.deg2rad(num[3]), np.deg2rad(num[4])
            num = [_fortran_float(t) for t in lines[k + 4].split()]
            if len(num) < 6:
                raise CaseFileError("incident wave %d: <xs(3)> <symconf(3)> expected" % fid)
            f["xs"], f["symconf"] = np.array(num[0:3]), tuple(int(round(t)) for t in num[3:6])
            if f["space"] == "half-space" and f["symconf"][f["np"] - 1] != 0:
                raise CaseFileError("incident wave %d: symconf(np) must be 0" % fid)
            w = lines[k + 5].split()
            rt = {"fluid": 1, "elastic": 2, "viscoelastic": 2, "poroelastic": 3}.get(w[0].lower())
            if rt is None:
                raise CaseFileError('incident wave %d: the region type must be "fluid", "elastic", "viscoelastic" or "poroelastic"' % fid)
            f["region_type"], f["wave"] = rt, w[1].lower()
            if rt == 3:
                raise CaseFileError("incident wave %d: incident fields of poroelastic media are not covered -- the reference stops there too ('please, check poroelastic free-field', "
                                    "calculate_incident_mechanics_harmonic.f90:511); arrays of the caller's own can be given through set_incident" % fid)
            if rt == 1:
                if f["wave"] != "p":
                    raise CaseFileError('incident wave %d: the wave type for a fluid can be only "p"' % fid)
                if f["variable"] != 0:
                    raise CaseFileError("incident wave %d: a fluid field in terms of normal displacements is not implemented in the reference" % fid)
                if cls == "point" and f["space"] != "full-space":
                    raise CaseFileError("incident wave %d: a point wave in a half-space is not implemented in the reference" % fid)
            else:
                if cls != "plane":
                    raise CaseFileError("incident wave %d: only plane waves are covered in elastic solids" % fid)
                if f["wave"] not in ("p", "sv", "sh", "rayleigh"):
                    raise CaseFileError('incident wave %d: the wave type for a viscoelastic solid can be "p", "sv", "sh" or "rayleigh"' % fid)
                if f["wave"] == "rayleigh" and (f["space"] != "half-space" or f["variable"] != 0):
                    raise CaseFileError("incident wave %d: a Rayleigh wave needs the half-space and variable 0 (in terms of stresses it is not implemented in the reference)" % fid)
                if np.any(f["x0"] != 0) or np.any(f["xs"] != 0):
                    raise CaseFileError("incident wave %d: all components of x0 and xs can be only 0." % fid)
                if f["space"] == "half-space" and (f["np"] != 3 or f["bc"] != 1):
                    raise CaseFileError("incident wave %d: np can be only 3 and bc must be 1." % fid)
                if f["symconf"][0] != 0 or f["symconf"][2] != 0:
                    raise CaseFileError("incident wave %d: the symmetry/anti-symmetry decomposition can not be done for the x and z directions" % fid)
                if f["variable"] not in (0, 1):
                    raise CaseFileError("incident wave %d: variable 0 (displacements) or 1 (stresses)" % fid)
            if fid in fields:
                raise CaseFileError("incident wave %d is repeated" % fid)
            fields[fid] = f
            k += 6
        for (rid, rtype, _, _), ids in zip(self.regions, self.region_incident):
            for fid in ids:
                if fid not in fields:
                    raise CaseFileError("region %d: incident field %d does not exist" % (rid, fid))
                if fields[fid]["region_type"] != rtype:
                    raise CaseFileError("region %d: incident field %d is of a different type of region" % (rid, fid))
        return fields

    def incident_arrays(self, model, omega):
        """{region index: (u_inc, t_inc)} at the frequency omega: the fields of every region summed at the nodes of its elements with the region's
        outward normal (src/calculate_incident_mechanics_harmonic.f90:326-470), ready for Problem.set_incident / CoupledProblem.set_incident."""
        from . import incident as inc
        out = {}
        for kr, ((rid, rtype, mat, _), ids) in enumerate(zip(self.regions, self.region_incident)):
            if not ids:
                continue
            v = model.views[kr] if self.multi else model
            tot = None
            for fid in ids:
                f = self.incident_fields[fid]
                if rtype == 1:
                    if f["cls"] == "point":
                        fld = inc.fluid_point_wave_reference(mat, omega, f["amplitude"], f["x0"])
                    else:
                        fld = inc.fluid_plane_wave_reference(mat, omega, f["amplitude"], f["x0"], f["varphi"], f["theta"], f["space"], f["np"], f["xp"],
                                                             f["bc"], f["symconf"], f["xs"])
                    scale = 1.0
                else:
                    fld = inc.elastic_plane_wave_reference(f["wave"], mat, omega, f["varphi"], f["theta"], f["space"], f["xp"], f["symconf"][1])
                    scale = 1.0
                    if f["variable"] == 1:           # the field in terms of stresses (calculate_incident_mechanics_harmonic.f90:456-470)
                        scale = 1.0 / (-1j * (omega / mat.c1) * (mat.lam + 2.0 * mat.mu)) if f["wave"] == "p" else 1.0 / (-1j * (omega / mat.c2) * mat.mu)
                geom = self.__dict__.setdefault("_incident_geometry", {})            # positions and normals: once per model, not per frequency
                if (id(model), kr) not in geom:
                    geom[(id(model), kr)] = inc.element_node_geometry(model.node_x, v.etype, v.elem_ptr, v.elem_node, v.elem_reversed)
                u, t = inc.field_at(fld, *geom[(id(model), kr)], 1 if rtype == 1 else 3)
                tot = (scale * u, scale * t) if tot is None else (tot[0] + scale * u, tot[1] + scale * t)
            out[kr] = tot
        return out

    def _material(self, rid, w):
        """(material, region type code) of the material record of a region: `material <id>` or the legacy in-line forms."""
        kind = w[0].lower()
        if kind == "material":
            if int(w[1]) not in self.materials:
                raise CaseFileError("region %d: unknown material %s" % (rid, w[1]))
            mtype, props = self.materials[int(w[1])]
            mtype = mtype.lower()
            if mtype in ("fluid", "inviscid_fluid"):
                kk = {k: props[k] for k in ("K", "rho", "c") if k in props}
                if len(kk) != 2:
                    raise CaseFileError("material %s: only 2 properties are needed" % w[1])
                rho = kk["rho"] if "rho" in kk else kk["K"] / kk["c"] ** 2
                c = kk["c"] if "c" in kk else np.sqrt(kk["K"] / kk["rho"])
                return Fluid(rho=rho, c=c, xi=props.get("xi", 0.0)), 1
            if mtype == "elastic_solid":
                ec = elastic_constants({k: props[k] for k in ("E", "nu", "lambda", "mu", "K") if k in props})
                if self.analysis == "harmonic" and ("rho" not in props or "xi" not in props):
                    raise CaseFileError("material %s: rho and xi are required for the material of this region" % w[1])
                return Material(rho=props.get("rho", 1.0), mu=ec["mu"], nu=ec["nu"], xi=props.get("xi", 0.0)), 2
            if mtype == "biot_poroelastic_medium":      # src/read_materials.f90:351-520, src/read_regions.f90:737-860
                need = [k for k in ("phi", "Q", "R", "rho_f", "rho_s") if k not in props]
                if need:
                    raise CaseFileError("material %s: %s required for the material of this region" % (w[1], ", ".join(need)))
                ec = elastic_constants({k: props[k] for k in ("E", "nu", "lambda", "mu", "K") if k in props})
                return Poro(rhof=props["rho_f"], rhos=props["rho_s"], lam=ec["lambda"], mu=ec["mu"], xi=props.get("xi", 0.0), phi=props["phi"],
                            rhoa=props.get("rho_a", 0.0), R=props["R"], Q=props["Q"], b=props.get("b", 0.0)), 3
            raise CaseFileError("material type %r is not covered (fluid, elastic_solid, biot_poroelastic_medium)" % mtype)
        if kind in ("fluid", "inviscid_fluid"):
            return Fluid(rho=_fortran_float(w[1]), c=_fortran_float(w[2])), 1
        if kind == "viscoelastic":
            return Material(rho=_fortran_float(w[1]), mu=_fortran_float(w[2]), nu=_fortran_float(w[3]), xi=_fortran_float(w[4])), 2
        if kind == "elastic":
            return Material(rho=_fortran_float(w[1]), mu=_fortran_float(w[2]), nu=_fortran_float(w[3]), xi=0.0), 2
        raise CaseFileError("region %d: material specification %r is not covered" % (rid, kind))

    def build_model(self):
        """The flat model (multifebe_b200.host.Model / FluidModel, or MultiRegionModel for coupled regions) with the reference's numbering:
        boundaries in the order of the regions' lists (build_auxiliary_variables_mechanics_harmonic.f90:151-198)."""
        try:
            return self._build_model()
        except ValueError as e:
            if isinstance(e, CaseFileError):
                raise
            raise CaseFileError(str(e))

    def _build_model(self):
        part_of_boundary = dict(self.boundaries)
        kw = dict(qsi_relative_error=self.qsi_relative_error, qsi_ns_max=self.qsi_ns_max, precalset_gln=self.precalset_gln,
                  geometric_tolerance=self.geometric_tolerance)
        if self.multi:
            from .multiregion import MultiRegionModel, Region, SOLID, FLUID
            from .multiregion import PORO
            regs = [Region({1: FLUID, 2: SOLID, 3: PORO}[rtype], mat, rb) for _, rtype, mat, rb in self.regions]
            bcs = {b: ((ct[0], cv[0]) if len(ct) == 1 else (ct, cv)) for b, (ct, cv) in self.bcs.items()}
            return MultiRegionModel(self.mesh, regs, part_of_boundary, bcs, symmetry=self.symmetry, formulation=self.formulation, **kw)
        kw["part_order"] = [part_of_boundary[b] for b in self.region_boundaries]
        kw["formulation"] = {part_of_boundary[b]: f for b, f in self.formulation.items()}
        kw["symmetry"] = self.symmetry
        kw["collapse_nodal_pos"] = self.collapse_nodal_pos
        if self.region_type == 1:
            bcs = {part_of_boundary[b]: (ct[0], cv[0]) for b, (ct, cv) in self.bcs.items()}
            return FluidModel(self.mesh, bcs, **kw)
        if self.region_type == 3:
            return PoroModel(self.mesh, {part_of_boundary[b]: (ct, cv) for b, (ct, cv) in self.bcs.items()}, **kw)
        bcs = {part_of_boundary[b]: (ct, cv) for b, (ct, cv) in self.bcs.items()}
        return Model(self.mesh, bcs, **kw)
