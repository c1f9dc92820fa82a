#!/usr/bin/env bash
# Build the UNMODIFIED reference C++/OpenMP backend (pybind11 module `_kernels`)
# straight from the sources where they lie under $REF (default /root/reference),
# writing ONLY into oracle/_ref/.  No reference source is copied into this repo.
#
# Mirrors the flags of the reference's own CMake recipe
# (src/strata_fdtd/_kernels/CMakeLists.txt:5-7 C++17, :55-57 -march=x86-64-v2,
#  :70 -O3 -DNDEBUG, :73-85 source list, :93-105 OpenMP + version defines).
# Test infrastructure only: the product path never loads this module.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${STRATA_REFERENCE:-/root/reference}"
K="$REF/src/strata_fdtd/_kernels"
OUT="$HERE/_ref"
if [ ! -d "$K" ]; then
  echo "[oracle/build_ref] reference sources not found at $K (expected on the GPU box); keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
PY="${PYTHON:-python3}"
EXT="$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')"
TARGET="$OUT/_kernels$EXT"
if [ -f "$TARGET" ] && [ "$TARGET" -nt "$K/kernels.cpp" ] && [ "$TARGET" -nt "$K/src/fdtd_step.cpp" ] && [ -z "${FORCE:-}" ]; then
  echo "[oracle/build_ref] up to date: $TARGET"; exit 0
fi
g++ -O3 -DNDEBUG -std=c++17 -fPIC -shared -fopenmp -march=x86-64-v2 \
    -DFDTD_HAS_OPENMP=1 -DFDTD_KERNELS_VERSION='"0.1.0"' \
    -I"$K/include" $($PY -m pybind11 --includes) \
    "$K/kernels.cpp" "$K"/src/*.cpp -o "$TARGET"
echo "[oracle/build_ref] built $TARGET"
