#!/bin/sh
# Builds libmsgwam_b200.so (sm_100a only) in-tree next to the Python package.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/msgwam_b200/libmsgwam_b200.so"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -ccbin /usr/bin/g++ -Xcompiler -fPIC -Xcompiler -fopenmp -shared -cudart static \
    ${MSGWAM_NVCC_EXTRA} \
    -o "$OUT" "$HERE"/csrc/column_step.cu "$HERE"/csrc/general.cu "$HERE"/csrc/compact.cu "$HERE"/csrc/host_path.cu -lgomp
echo "$OUT"
