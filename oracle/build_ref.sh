#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Compiles the UNMODIFIED reference kernel
# (/root/reference/lib/rhf_response.cc) where it lies, into oracle/_ref/ (git-ignored,
# travels to the GPU box with the snapshot).  No reference source is copied.
#   libqcdmet_ref_mkl.so      : MKL inside torch/lib/libtorch_cpu.so + dsyev_ shim (import torch first)
#   libqcdmet_ref_openblas.so : pthread OpenBLAS bundled with opencv (run with OPENBLAS_NUM_THREADS=1)
# The reference's own CMake build is not used: icc flags / find_package(LAPACK).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF=${QCDMET_REFERENCE_SRC:-/root/reference/lib/rhf_response.cc}
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -f "$REF" ]; then
  echo "build_ref: $REF absent (GPU box?) - keeping prebuilt files in $OUT" >&2
  exit 0
fi
SP="$(python - <<'PY'
import sysconfig; print(sysconfig.get_paths()["purelib"])
PY
)"
CXXFLAGS="-O3 -g -march=x86-64-v3 -fopenmp -shared -fPIC"
# MKL flavour
gcc -O2 -fPIC -c "$HERE/dsyev_shim.c" -o "$OUT/dsyev_shim.o"
g++ $CXXFLAGS "$REF" "$OUT/dsyev_shim.o" -o "$OUT/libqcdmet_ref_mkl.so" \
    -L"$SP/torch/lib" -ltorch_cpu -lc10 -Wl,--disable-new-dtags,-rpath,"$SP/torch/lib"
# OpenBLAS flavour
OB="$(ls "$SP"/opencv_python_headless.libs/libopenblasp-*.so 2>/dev/null | head -1 || true)"
if [ -n "$OB" ]; then
  g++ $CXXFLAGS "$REF" -o "$OUT/libqcdmet_ref_openblas.so" \
      -L"$(dirname "$OB")" -l:"$(basename "$OB")" -Wl,--disable-new-dtags,-rpath,"$(dirname "$OB")"
fi
rm -f "$OUT/dsyev_shim.o"
ls -la "$OUT"
