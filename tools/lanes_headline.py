"""Two frequencies of the headline mesh in flight on one GPU (capi.ProblemLanes): does the assembly / panel / solve of one hide behind the LU of the other?
   python tools/lanes_headline.py [n_lanes] [n_freq]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import argparse
import bench
from multifebe_b200 import capi

n_lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_freq = int(sys.argv[2]) if len(sys.argv) > 2 else 6
args = argparse.Namespace(workload="headline", m=40, etype="tri3", gpus=1)
try:
    md, mat, freqs, name = bench.workload(args)
except Exception as e:
    print("workload() failed:", e); raise
print(name, "n_dof", md.n_dof)
lanes = capi.ProblemLanes(md, 0, n_lanes)
om = [float(freqs[(s * 21) % len(freqs)]) for s in range(n_freq)]
lanes.run(om[:n_lanes], mat, host=True)          # warm-up: one frequency per lane
t0 = time.time()
X = lanes.run(om, mat, host=True)
dt = time.time() - t0
print("lanes %d: %d frequencies in %.3f s -> %.4f solves/s (%.1f ms per frequency)" % (n_lanes, n_freq, dt, n_freq / dt, 1e3 * dt / n_freq))
lanes.close()
