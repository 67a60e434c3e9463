"""Stand-alone driver: the reference's main program (src/multifebe.f90) for the analyses this library covers.

    python -m multifebe_b200 -i case.dat [-o output] [-b verbose]          (options of src/process_command_line_options.f90)
    torchrun --nproc-per-node N -m multifebe_b200 -i case.dat               (frequency shard over N GPUs)

reads the reference's case file and Gmsh 2.2 mesh (host/casefile.py), runs `do kf = 1, n_frequencies: build_lse_mechanics_harmonic,
solve_lse_c, assign_solution, export_solution` (src/multifebe.f90:107-124) -- or the static sequence build_lse_mechanics_static,
solve_lse_r -- on the GPU through the C ABI, and writes `<output>.nso` as the reference does (host/export.py).  There is no CPU path:
without a CUDA device the default solver fails in mfb_init.  `solver` is injectable so that the CPU tests can drive the file
handling with the oracle.
"""
import argparse
import os
import sys
import time
import numpy as np

from .host.casefile import CaseFile, CaseFileError
from .host.export import NsoWriter
from .sweep import FrequencySweep


class GpuSolver:
    """One mfb_problem on one GPU: harmonic(omega) / static() -> solution vector in the reference's column order."""

    def __init__(self, case, model, device=0):
        self.cp = None
        if case.multi:
            from . import capi
            self.capi, self.case = capi, case
            self.ctx = capi.Context(device)
            self.cp = capi.CoupledProblem(self.ctx, model)
            return
        from . import capi
        self.capi, self.case = capi, case
        self.ctx = capi.Context(device)
        self.pr = capi.Problem(self.ctx, model)

    def harmonic(self, omega):
        if self.cp is not None:
            return self.cp.solve_frequency(omega)
        if self.case.region_type == 1:
            return self.pr.solve_frequency_fluid(omega, self.case.material)
        if self.case.region_type == 3:
            return self.pr.solve_frequency_poro(omega, self.case.material)
        return self.pr.solve_frequency(omega, self.case.material)

    def set_incident(self, arrays):
        """arrays = CaseFile.incident_arrays(model, omega): the incident fields of the regions at the frequency about to be solved."""
        if self.cp is not None:
            for kr in range(len(self.case.regions)):
                self.cp.set_incident(kr, *arrays.get(kr, (None, None)))
        else:
            self.pr.set_incident(*arrays.get(0, (None, None)))

    def static(self):
        return self.pr.solve_static(self.case.material)

    # internal points of an elastic region (calculate_internal_points_mechanics_bem_harela / _staela): u (n,3) and sigma (n,3,3) from the boundary solution
    def _ip(self, points):
        if getattr(self, "_ipo", None) is None:
            self._ipo = self.capi.InternalPoints(self.ctx, self.pr.m, points)
        return self._ipo

    def interior_harmonic(self, omega, x, points):
        ip = self._ip(points)
        return ip.displacements(omega, self.case.material, x), ip.stresses(omega, self.case.material, x)

    def interior_static(self, x, points):
        ip = self._ip(points)
        return ip.displacements_static(self.case.material, x), ip.stresses_static(self.case.material, x)

    def stats(self):
        return self.pr.stats()

    def close(self):
        if self.cp is not None:
            self.cp.close(); self.ctx.close()
            return
        if getattr(self, "_ipo", None) is not None:
            self._ipo.close()
        self.pr.close(); self.ctx.close()


def run(case_path, output=None, solver=None, verbose=1, rank=0, world=1, dist=None, device=None, device_index=0, log=sys.stdout):
    """Runs the case; the writer rank (0) returns the path of the *.nso file, the others None.  solver = None: the GPU (GpuSolver on
    CUDA device `device_index`); `device` = torch device of the gather buffers when world > 1 (None: host tensors, gloo)."""
    t0 = time.time()
    case = CaseFile(case_path)
    model = case.build_model()
    out_base = output or case_path
    if verbose >= 1 and rank == 0:
        log.write("multifebe_b200: %s analysis, %d region(s) of type %s, %d nodes, %d elements, %d DOF, %d frequencies, %d rank(s)\n" % (
            case.analysis, len(case.regions), "/".join(str(r[1]) for r in case.regions), model.n_node, model.n_elem, model.n_dof, len(case.omega), world))
    own = solver is None
    if own:
        solver = GpuSolver(case, model, device_index)
    ip_x = np.array([xp for _, _, xp in case.internal_points]) if case.internal_points else None
    nso = None
    fh = None
    if rank == 0 and case.export_nso:
        nso = out_base + ".nso"
        fh = open(nso, "w")
        wr = NsoWriter(fh, case, model)
        wr.header()
    try:
        if case.analysis == "static":
            if rank == 0:
                x = solver.static()
                if fh:
                    wr.static(x)
                    if case.internal_points:
                        wr.static_internal(*solver.interior_static(x, ip_x))
        else:
            has_inc = any(case.region_incident)

            def one_frequency(kf, om):
                if has_inc:                     # [incident waves]: the arrays depend on the frequency (calculate_incident_mechanics_harmonic(kf))
                    solver.set_incident(case.incident_arrays(model, om))
                return solver.harmonic(om)
            sweep = FrequencySweep(case.omega, model.n_dof, one_frequency, rank=rank, world=world, dist=dist, device=device)
            for r in range(sweep.n_rounds()):
                sweep.round(r)
                if rank == 0:
                    for kf in range(r * world, min((r + 1) * world, len(case.omega))):   # in-order export, one frequency at a time
                        xk = sweep.results.pop(kf)
                        if fh:
                            wr.frequency(kf + 1, xk)
                            if case.internal_points:
                                # the writer rank evaluates the interior identities of every frequency with its own problem objects
                                wr.frequency_internal(kf + 1, *solver.interior_harmonic(case.omega[kf], xk, ip_x))
                        if verbose >= 2:
                            log.write("  frequency %d / %d done\n" % (kf + 1, len(case.omega)))
    finally:
        if fh:
            fh.close()
        if own:
            solver.close()
    if verbose >= 1 and rank == 0:
        log.write("multifebe_b200: done in %.2f s%s\n" % (time.time() - t0, (", results in " + nso) if nso else ""))
    return nso


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m multifebe_b200", description="B200-native driver for MultiFEBE case files (3D BEM hot path)")
    ap.add_argument("-i", "--input", required=True, help="input (case) file")
    ap.add_argument("-o", "--output", default=None, help="output files base name (default: the input file name)")
    ap.add_argument("-b", "--verbose", type=int, default=1)
    args = ap.parse_args(argv)
    if not os.path.exists(args.input):
        print("Input file does not exist, check the given path.")
        return 2
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dist = device = None
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
        device = torch.device("cuda", local)
    try:
        run(args.input, args.output, verbose=max(args.verbose, 0), rank=rank, world=world, dist=dist, device=device, device_index=local)
    except CaseFileError as e:
        print("multifebe_b200: %s" % e)
        return 1
    except RuntimeError as e:          # capi.MfbError: no CUDA device, CUDA failure, singular system ... (there is no CPU path to fall back to)
        print("multifebe_b200: %s" % e)
        return 3
    finally:
        if world > 1:
            dist.destroy_process_group()
    return 0
