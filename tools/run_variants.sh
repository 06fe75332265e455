for v in 512x1 512x2 384x2 256x3; do echo "== $v"; MSGWAM_B200_LIB=$PWD/tools/_variants/lib_$v.so python tools/kernel_timing.py 0 1e6 1e7 2>&1 | grep -v shuffled..true; done
