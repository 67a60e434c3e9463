#!/bin/bash
# oracle/build_ref.sh -- builds the UNMODIFIED reference (mmc-siani-es/MultiFEBE, Fortran 2003 + OpenBLAS) into oracle/_ref/ when a Fortran compiler and
# OpenBLAS exist on the machine, runs it on a case file with full-precision export and leaves <case>.nso for tests/test_reference_binary.py to diff against
# the GPU (tolerance 1e-8 on the nodal solutions).  TEST INFRASTRUCTURE ONLY.  Never copies reference sources into the repository: it compiles them where
# they lie (REF, default /root/reference) with the Release flags of the reference's CMakeLists.txt:13 (-O3 -march=core2 -fopenmp -cpp).
#
# Probe results on record (round 2): this container and the B200 boxes of the pool have NO Fortran compiler (gfortran, flang, ifx, nvfortran: absent;
# gpurun_out/probe_fortran.log) -- the script then prints why it stops and exits 3, and parity stays pinned by tests/test_oracle_bruteforce.py instead.
set -u
REF=${REF:-/root/reference}
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
FC=""
for c in gfortran flang ifx nvfortran; do if command -v $c >/dev/null 2>&1; then FC=$c; break; fi; done
if [ -z "$FC" ]; then echo "build_ref: no Fortran compiler on this machine (looked for gfortran, flang, ifx, nvfortran): oracle/_ref cannot be built"; exit 3; fi
if [ ! -d "$REF/src" ]; then echo "build_ref: reference tree not found at $REF"; exit 3; fi
BLAS=""
for l in /usr/lib/x86_64-linux-gnu/libopenblas.so /usr/lib64/libopenblas.so /usr/lib/libopenblas.so; do [ -e $l ] && BLAS=$l; done
if [ -z "$BLAS" ]; then echo "build_ref: OpenBLAS not found (the reference links it for zgetrf/zgetrs)"; exit 3; fi
mkdir -p "$OUT/obj" && cd "$OUT/obj" || exit 1
FLAGS="-O3 -march=core2 -fopenmp -cpp -ffree-line-length-none -J$OUT/obj"
# modules first (lib/fbem, then the program's module files), then everything else; two passes resolve the remaining module order
SRC_LIB=$(ls $REF/lib/fbem/src/*.f90); SRC_APP=$(ls $REF/src/*.f90)
for pass in 1 2 3; do
  for f in $SRC_LIB $SRC_APP; do o=$(basename ${f%.f90}).o; [ -e $o ] || $FC $FLAGS -I$REF/lib/fbem/src -c $f -o $o 2>/dev/null; done
done
missing=0; for f in $SRC_LIB $SRC_APP; do [ -e $(basename ${f%.f90}).o ] || { echo "build_ref: failed to compile $f"; missing=1; }; done
[ $missing = 0 ] || exit 4
$FC -fopenmp *.o $BLAS -o "$OUT/multifebe" || exit 4
echo "build_ref: built $OUT/multifebe with $FC"
if [ $# -ge 1 ]; then   # run a case: the case file must ask for real_format = sci_double and export_nso = T (src/read_export.f90:158-166)
  (cd "$(dirname "$1")" && OMP_NUM_THREADS=$(nproc) "$OUT/multifebe" -i "$(basename "$1")" -b 2) || exit 5
fi
