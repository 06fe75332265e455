#!/bin/sh
# Developer tool (on the GPU box): tools/var_timing.py for the product library and every variant in tools/_variants
python tools/var_timing.py $1
for f in tools/_variants/lib_*.so; do
  [ -f "$f" ] && MSGWAM_B200_LIB=$PWD/$f python tools/var_timing.py $1 2>&1 | tail -1
done
