#!/usr/bin/env bash
# oracle/build_oracle.sh -- TEST INFRASTRUCTURE ONLY.
# Builds (1) the plain-C restatement oracle/libstmqr_oracle.so and, when the reference tree is
# present, (2) oracle/_ref/libstmmqr_ref.so via build_ref.sh and (3) the ctypes harness
# oracle/_ref/libref_harness.so compiled against the reference headers in place.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
S="$REF/STMMQR"

gcc -std=gnu99 -O2 -fPIC -shared -Wall -Wno-unused-variable "$HERE/stmqr_oracle.c" \
    -o "$HERE/libstmqr_oracle.so" -lm
echo "built $HERE/libstmqr_oracle.so"

if [ -d "$S/src/qr" ]; then
    bash "$HERE/build_ref.sh"
    INC="-I$S/include -I$S/include/tpsm -I$HERE/shim -I$S/CAMD/Include -I$S/CCOLAMD/Include -I$S/SuiteSparse_config"
    BLAS="$(cat "$HERE/_ref/blas_path.txt")"
    gcc -std=gnu99 -fcommon -w -O2 -fPIC -shared -include "$HERE/shim/tpsm_platform.h" $INC \
        "$HERE/ref_harness.c" -o "$HERE/_ref/libref_harness.so" \
        -L"$HERE/_ref" -lstmmqr_ref "$BLAS" -Wl,--disable-new-dtags -Wl,-rpath,'$ORIGIN' \
        -Wl,-rpath,"$(dirname "$BLAS")" -ldl -lpthread -lm
    echo "built $HERE/_ref/libref_harness.so"
else
    echo "reference tree absent: keeping prebuilt oracle/_ref/*.so" >&2
fi
