#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY (never part of the product path).
#
# Compiles the UNMODIFIED reference (STMMQR C99 sources) from where they lie
# under $REF (default /root/reference) into oracle/_ref/libstmmqr_ref.so.
# Nothing is copied out of the reference tree except the bundled Data/*.mtx
# test matrices (data, not source), which go to oracle/_ref/data/ so that they
# travel to the GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored).
#
# Recipe follows the reference's own Makefile.option / src/*/Makefile flags
# (STMMQR/Makefile.option:8-75: gcc -std=c99 -O2 -fPIC -DPRINT_TIME -DDLONG
#  -DBACKUP -DSETSTACKSIZE -DDEFAULT) with three host adaptations:
#   * -std=gnu99 : CPU_SET / pthread affinity macros (tpsm_threads.c:113-160)
#   * -fcommon   : tentative globals FCHUNK/SMALL/... live in a header
#                  (STMMQR/include/SparseQR.h:16-19)
#   * -include oracle/shim/tpsm_platform.h, -Ioracle/shim (numa.h): see shims.
# BLAS/LAPACK (dlarfg/dlarf/dlarft/dlarfb/dnrm2, not vendored by the
# reference): the only LP64 OpenBLAS+LAPACK in this image is OpenBLAS 0.3.15
# inside the opencv wheel; it is linked by absolute path + rpath.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
S="$REF/STMMQR"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
JOBS="${JOBS:-$(nproc)}"

if [ ! -d "$S/src/qr" ]; then
    echo "build_ref.sh: reference tree not found at $REF (expected on the GPU box); keeping prebuilt oracle/_ref" >&2
    exit 0
fi

PYSITE="$(python -c 'import sysconfig; print(sysconfig.get_paths()["purelib"])')"
BLASDIR="$PYSITE/opencv_python_headless.libs"
BLAS="$(ls "$BLASDIR"/libopenblasp-*.so | head -1)"
[ -f "$BLAS" ] || { echo "OpenBLAS not found under $BLASDIR" >&2; exit 1; }

mkdir -p "$OBJ/base" "$OBJ/core" "$OBJ/chol" "$OBJ/qr" "$OBJ/ss" "$OBJ/metis" "$OUT/data"

# Up-to-date check: rebuild only when a reference source or this recipe is newer.
STAMP="$OUT/.stamp"
if [ -f "$OUT/libstmmqr_ref.so" ] && [ -f "$STAMP" ] && [ "$STAMP" -nt "$HERE/build_ref.sh" ] \
   && [ "$STAMP" -nt "$HERE/shim/numa.h" ] && [ "$STAMP" -nt "$HERE/shim/tpsm_platform.h" ] \
   && [ -z "$(find "$S/src" "$S/include" -newer "$STAMP" -name '*.[ch]' | head -1)" ]; then
    echo "oracle/_ref up to date"
    exit 0
fi

CC="gcc -std=gnu99 -fcommon -w -O2 -fPIC -DDLONG -DPRINT_TIME -DBACKUP -DSETSTACKSIZE -DDEFAULT -DORACLE_NPROC=256"
INC="-I$S/include -I$S/include/tpsm -I$HERE/shim -I$S/CAMD/Include -I$S/CCOLAMD/Include -I$S/SuiteSparse_config"
PRE="-include $HERE/shim/tpsm_platform.h"

CMDS="$OUT/compile_cmds.txt"
: > "$CMDS"
emit() { # emit <objdir> <src> [extra flags]
    local o="$OBJ/$1/$(basename "${2%.c}").o"
    echo "$CC $INC $PRE ${3:-} -c $2 -o $o" >> "$CMDS"
}
for f in SparseBase_config xerbla amd colamd tpsm_base tpsm_distribution tpsm_buffer \
         tpsm_synchronization tpsm_tcb tpsm_barrier tpsm_threads tpsm_auxiliary tpsm_main; do
    emit base "$S/src/base/$f.c"
done
for f in matrixops change_factor common matrix_type check read_write norm metis nesdis csymamd camd ccolamd; do
    emit core "$S/src/core/SparseCore_$f.c"
done
for f in analyze factorize super_numeric super_symbolic solve super_solve; do
    emit chol "$S/src/chol/SparseChol_$f.c"
done
for f in SparseQR SparseQR_analyze SparseQR_factorize SparseQR_multithreads SparseLQ; do
    emit qr "$S/src/qr/$f.c"
done
emit ss "$S/SuiteSparse_config/SuiteSparse_config.c"
emit ss "$S/CCOLAMD/Source/ccolamd.c"
for f in "$S"/CAMD/Source/camd_*.c; do emit ss "$f"; done

# METIS 5.1.0 with the reference's 64-bit idx header (include/metis.h:69) first on the path.
MCC="gcc -std=gnu99 -w -O2 -fPIC -DLINUX -D_FILE_OFFSET_BITS=64 -DNDEBUG -DNDEBUG2 -DHAVE_EXECINFO_H -DHAVE_GETLINE"
MINC="-I$S/include -I$S/metis-5.1.0/GKlib -I$S/metis-5.1.0/libmetis"
for f in "$S"/metis-5.1.0/GKlib/*.c; do
    echo "$MCC $MINC -c $f -o $OBJ/metis/gk_$(basename "${f%.c}").o" >> "$CMDS"
done
for f in "$S"/metis-5.1.0/libmetis/*.c; do
    echo "$MCC $MINC -c $f -o $OBJ/metis/lm_$(basename "${f%.c}").o" >> "$CMDS"
done

echo "compiling $(wc -l < "$CMDS") reference translation units with $JOBS jobs ..."
xargs -P "$JOBS" -I{} sh -c '{}' < "$CMDS"

gcc -shared -o "$OUT/libstmmqr_ref.so" \
    "$OBJ"/base/*.o "$OBJ"/core/*.o "$OBJ"/chol/*.o "$OBJ"/qr/*.o "$OBJ"/ss/*.o "$OBJ"/metis/*.o \
    "$BLAS" -Wl,--disable-new-dtags -Wl,-rpath,"$BLASDIR" -lpthread -lm -lrt

# The reference's own acceptance driver, unmodified (STMMQR/test/qrtest.c).
gcc -std=gnu99 -fcommon -w -O2 $INC $PRE "$S/test/qrtest.c" -o "$OUT/qrtest" \
    -L"$OUT" -lstmmqr_ref -Wl,--disable-new-dtags -Wl,-rpath,"$OUT" -Wl,-rpath,"$BLASDIR" "$BLAS" -lpthread -lm -lrt

# Bundled matrices (data) so the GPU box has them.
cp -f "$REF"/Data/*.mtx "$OUT/data/" 2>/dev/null || true
echo "$BLAS" > "$OUT/blas_path.txt"
touch "$STAMP"
echo "built $OUT/libstmmqr_ref.so and $OUT/qrtest"
