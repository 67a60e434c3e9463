"""Build libmfb.so (the C-ABI shared library) in-tree with nvcc for sm_100a.

    python -m multifebe_b200.build [--force]

nvcc cross-compiles without a GPU.  The host planner (plan_host.cpp) is compiled by the system g++ with
-ffp-contract=off (its arithmetic feeds discrete quadrature decisions, see csrc/plan_host.h).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmfb.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
GXX = "/usr/bin/g++"   # the image's $CXX wrapper lacks libgomp.spec
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
SOURCES_CU = ["api.cu", "assembly.cu", "potential.cu", "poro.cu", "combine.cu", "lu.cu", "gemm_tma.cu", "solve_ex.cu", "dist.cu"]
SOURCES_CPP = ["plan_host.cpp", "plan_values.cpp"]   # host planner: decision core (follows the reference op by op) and value geometry (independent derivations)
DEPS = SOURCES_CU + SOURCES_CPP + ["plan_host.h", "plan_values.h", "assembly.cuh", "potential.cuh", "pot_math.cuh", "poro.cuh", "por_math.cuh", "por_pair.cuh", "combine.cuh", "lu.cuh", "solve_ex.cuh", "dist.cuh", "bem_math.cuh",
                     os.path.join("..", "..", "include", "mfb.h"), os.path.join("..", "..", "data", "quad_tables.h")]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


HEADERS = [d for d in DEPS if not d.endswith((".cu", ".cpp"))]


def _obj_stale(obj, src):
    """An object is rebuilt when its source or any header is newer (headers are few: no per-file dependency scan)."""
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in [src] + HEADERS)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    objs = []
    for src in SOURCES_CPP:
        o = os.path.join(bdir, src.replace(".cpp", ".o"))
        cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-c", os.path.join(CSRC, src), "-o", o]
        if force or _obj_stale(o, src):
            subprocess.check_call(cmd)
        objs.append(o)
    for src in SOURCES_CU:
        o = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [NVCC, "-ccbin", GXX, "-O3", "-std=c++17", "-lineinfo"] + os.environ.get("MFB_NVCC_FLAGS", "").split() + ARCH + [
            "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off,-fcx-fortran-rules",
            "-Xptxas", "-v" if verbose else "-O3", "-c", os.path.join(CSRC, src), "-o", o]
        if verbose:
            print(" ".join(cmd))
        if force or verbose or _obj_stale(o, src):
            subprocess.check_call(cmd)
        objs.append(o)
    cmd = [NVCC, "-ccbin", GXX, "-shared", "-o", OUT] + objs + ["-Xcompiler", "-fopenmp", "-lquadmath", "-lcudart", "-lgomp", "-ldl"] + ARCH
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
