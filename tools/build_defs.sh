#!/bin/sh
# Developer tool: build a tuning variant of the library with extra -D flags.  usage: build_defs.sh NAME -DX=1 -DY=2 ...
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
mkdir -p "$HERE/tools/_variants"
name=$1; shift
out="$HERE/tools/_variants/lib_${name}.so"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC -shared -cudart static "$@" -Xptxas -v \
  -o "$out" "$HERE"/python-msgwam_b200/csrc/*.cu 2>&1 | grep -A1 "column_pass" | grep -B1 "registers" | grep -o "column_pass[A-Za-z_0-9]*\|Used [0-9]* registers\|[0-9]* bytes spill stores" | paste - - | sort | uniq -c
echo "$out"
