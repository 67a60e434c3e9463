"""Frequency-sharded sweep: the multi-GPU form of the reference's serial frequency loop (src/multifebe.f90:107-124).

Iterations of that loop are independent (A_c, b_c are re-zeroed at src/build_lse_mechanics_harmonic.f90:73-74; mesh, plan and
precalculated point sets do not depend on omega), so frequency kf is owned by rank kf % world_size: one process per GPU,
mesh + plan replicated, NO collective on the data path.  The only communication is the gather of each solution vector
(n_dof complex numbers) to the writer rank, which keeps the reference's in-order, per-frequency export
(export_solution_mechanics_harmonic(kf), src/multifebe.f90:116).

`solve(kf, omega) -> x` is any callable; on a GPU box it is Problem.solve_frequency, in the CPU tests a stub.
"""
import numpy as np


def owned_frequencies(n_freq, rank, world):
    """Round-robin shard: the kf (0-based) this rank processes, in sweep order."""
    return list(range(rank, n_freq, world))


def linear_frequencies(omega_min, omega_max, n):
    """`lin` frequency list of the reference's [frequencies] section (src/read_frequencies.f90)."""
    return [omega_min + (omega_max - omega_min) * k / (n - 1) for k in range(n)] if n > 1 else [omega_min]


class FrequencySweep:
    def __init__(self, frequencies, n_dof, solve, rank=0, world=1, dist=None, device=None, writer=0):
        self.freq = list(frequencies)
        self.n_dof = n_dof
        self.solve = solve
        self.rank, self.world, self.dist, self.device, self.writer = rank, world, dist, device, writer
        self.results = {}   # kf -> x, filled on the writer rank only (all ranks when world == 1)

    def round(self, r):
        """Process sweep round r: rank q solves frequency kf = r*world + q (if it exists); the writer receives the
        `world` solutions of the round in kf order.  Returns the kf solved by this rank (or None)."""
        kf = r * self.world + self.rank
        mine = kf < len(self.freq)
        x = self.solve(kf, self.freq[kf]) if mine else np.zeros(self.n_dof, dtype=np.complex128)
        if self.world == 1:
            if mine:
                self.results[kf] = x
            return kf if mine else None
        import torch
        t = torch.from_numpy(np.ascontiguousarray(x).view(np.float64).copy())
        if self.device is not None:
            t = t.to(self.device, non_blocking=True)
        bufs = [torch.empty_like(t) for _ in range(self.world)] if self.rank == self.writer else None
        self.dist.gather(t, bufs, dst=self.writer)
        if self.rank == self.writer:
            for q in range(self.world):
                k = r * self.world + q
                if k < len(self.freq):
                    self.results[k] = bufs[q].cpu().numpy().view(np.complex128)
        return kf if mine else None

    def n_rounds(self):
        return (len(self.freq) + self.world - 1) // self.world

    def run(self):
        for r in range(self.n_rounds()):
            self.round(r)
        return self.results
