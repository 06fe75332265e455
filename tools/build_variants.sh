#!/bin/sh
# Developer tool: build tuning variants of the library into gpurun-visible build/ (not product).
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$HERE/tools/_variants"
for cfg in "$@"; do
  nt=${cfg%%x*}; rest=${cfg#*x}; r=${rest%%x*}; pf=1; case "$rest" in *x*) pf=${rest#*x};; esac
  out="$HERE/tools/_variants/lib_${cfg}.so"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ \
    -Xcompiler -fPIC -shared -cudart static -DMSGWAM_COL_NT=$nt -DMSGWAM_COL_R=$r -DMSGWAM_COL_PREFETCH=$pf -Xptxas -v \
    -o "$out" "$HERE"/python-msgwam_b200/csrc/*.cu
  echo "$out"
done
