"""Frequency shard of the sweep (multifebe_b200/sweep.py) on world_size 2 with the gloo backend: ownership, ordered gather
to the writer rank, ragged last round, no collective other than the gather."""
import os
import sys
import numpy as np
import pytest
from multifebe_b200.sweep import FrequencySweep, owned_frequencies, linear_frequencies

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ownership_is_a_partition():
    for n, w in ((64, 8), (7, 2), (3, 4), (1, 1)):
        all_k = sorted(k for r in range(w) for k in owned_frequencies(n, r, w))
        assert all_k == list(range(n))
    f = linear_frequencies(0.01, 15.0, 300)      # docs/examples/ME-TH-EL-001/case_files/t3.dat:6-11
    assert len(f) == 300 and f[0] == 0.01 and abs(f[-1] - 15.0) < 1e-12


def _worker(rank, world, port, n_freq, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 6
    solved = []

    def solve(kf, om):
        solved.append(kf)
        return (om + 1j * kf) * np.arange(1, n + 1)
    sw = FrequencySweep(linear_frequencies(1.0, 2.0, n_freq), n, solve, rank=rank, world=world, dist=dist)
    res = sw.run()
    out[rank] = (solved, {k: v.copy() for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_freq", [5, 4])
def test_two_rank_sweep_gathers_in_order(n_freq):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + n_freq
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, n_freq, out), nprocs=2, join=True)
        s0, r0 = out[0]
        s1, r1 = out[1]
    assert s0 == owned_frequencies(n_freq, 0, 2) and s1 == owned_frequencies(n_freq, 1, 2)
    assert r1 == {}                                   # only the writer holds results
    assert sorted(r0) == list(range(n_freq))
    f = linear_frequencies(1.0, 2.0, n_freq)
    for k in range(n_freq):
        assert np.array_equal(r0[k], (f[k] + 1j * k) * np.arange(1, 7))
