#!/bin/sh
# Developer tool: build tuning variants of the library into gpurun-visible build/ (not product).
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$HERE/tools/_variants"
for cfg in "$@"; do
  nt=${cfg%x*}; r=${cfg#*x}
  out="$HERE/tools/_variants/lib_${nt}x${r}.so"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ \
    -Xcompiler -fPIC -shared -cudart static -DMSGWAM_COL_NT=$nt -DMSGWAM_COL_R=$r \
    -o "$out" "$HERE"/python-msgwam_b200/csrc/*.cu
  echo "$out"
done
