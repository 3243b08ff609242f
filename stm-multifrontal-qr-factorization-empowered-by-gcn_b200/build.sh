#!/usr/bin/env bash
# Builds the product: lib/libstmqr_b200.so (CUDA engine + C ABI, sm_100a) and, when the
# reference headers are available, lib/libstmqr_dropin.so (the host-side qr_factorize).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
mkdir -p "$HERE/lib"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC -Xcompiler -Wall -shared ${NVCC_EXTRA:-} \
    "$HERE/csrc/stmqr_b200.cu" -o "$HERE/lib/libstmqr_b200.so"
echo "built $HERE/lib/libstmqr_b200.so"
if [ -d "$REF/STMMQR/include" ] && [ -f "$HERE/host/qr_factorize_b200.c" ]; then
    S="$REF/STMMQR"
    gcc -std=gnu99 -fcommon -w -O2 -fPIC -shared \
        -I"$S/include" -I"$S/include/tpsm" -idirafter "$HERE/host/compat" -I"$HERE/../include" \
        "$HERE/host/qr_factorize_b200.c" -o "$HERE/lib/libstmqr_dropin.so" \
        -L"$HERE/lib" -lstmqr_b200 -Wl,-rpath,'$ORIGIN'
    echo "built $HERE/lib/libstmqr_dropin.so"
fi
