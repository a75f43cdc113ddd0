#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference Cython core
# (raypier/core/*.pyx under /root/reference) into oracle/_ref/ so that
#   (1) the C restatement in oracle/rpx_oracle.c can be pinned against it,
#   (2) tests/golden/ fixtures can be generated from the real thing,
#   (3) bench.py --impl reference / cpu_baseline(kind="reference") can time it.
#
# Nothing is copied from the reference: cython reads the .pyx/.pxd where they lie
# (read-only) and writes generated C into oracle/_ref/build/; gcc writes the .so
# files into oracle/_ref/<flavour>/raypier/core/.  The two __init__.py files we
# create are EMPTY (the reference's raypier/__init__.py needs Traits, absent here).
#
# Flavours:
#   parity : -O2 -ffp-contract=off -fopenmp   (no FMA contraction; the parity pin; -fopenmp only because
#            cdistortions calls omp_get_num_procs at import -- no prange sits on the trace path)
#   timing : -O2 -fopenmp -march=x86-64-v3   (the reference uses -march=native, setup.py:35-46; a .so built
#            with the build box's native ISA may SIGILL on the GPU box, so the portable AVX2+FMA level is used)
#
# Usage: oracle/build_ref.sh [parity|timing|all]   (default: all)
set -euo pipefail
REF=${RPX_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
WHAT="${1:-all}"
if [ ! -d "$REF/raypier/core" ]; then
    echo "reference not present at $REF; keeping whatever is in $OUT" >&2
    exit 0
fi
PY=${PYTHON:-python}
CC=${RPX_REF_CC:-/usr/bin/gcc}
PYINC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
NPINC=$($PY -c "import numpy; print(numpy.get_include())")
EXT=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
MODS="ctracer cfaces cmaterials cshapes cdistortions cimplicit_surfs cfields obbtree cbezier"
mkdir -p "$OUT/build"
# 1. cythonize (once; shared by both flavours)
for m in $MODS; do
    if [ ! -s "$OUT/build/$m.c" ] || [ "$REF/raypier/core/$m.pyx" -nt "$OUT/build/$m.c" ]; then
        ( cd "$REF" && $PY -m cython -3 -I "$NPINC" -o "$OUT/build/$m.c" "raypier/core/$m.pyx" ) &
    fi
done
wait
build_flavour() {
    local name="$1"; shift
    local dst="$OUT/$name/raypier/core"
    mkdir -p "$dst"
    : > "$OUT/$name/raypier/__init__.py"
    : > "$dst/__init__.py"
    for m in $MODS; do
        if [ ! -s "$dst/$m$EXT" ] || [ "$OUT/build/$m.c" -nt "$dst/$m$EXT" ]; then
            $CC -shared -fPIC -fwrapv -fno-strict-aliasing -w "$@" \
                -DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION \
                -I"$PYINC" -I"$NPINC" "$OUT/build/$m.c" -o "$dst/$m$EXT" -lm &
        fi
    done
    wait
}
case "$WHAT" in
    parity|all) build_flavour parity -O2 -ffp-contract=off -fopenmp ;;
esac
case "$WHAT" in
    timing|all) build_flavour timing -O2 -fopenmp -march=x86-64-v3 ;;
esac
echo "reference core built into $OUT ($WHAT)"
