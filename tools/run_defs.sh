#!/bin/sh
# Developer tool (on the GPU box): time every library variant in tools/_variants on the ordered / shuffled / evolved ensembles.
for f in tools/_variants/lib_*.so; do
  echo "== $f"
  MSGWAM_B200_LIB=$PWD/$f python tools/kernel_timing.py ${SIZES:-1e6} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   n=%g shuffled=%s A %.1f  B %.1f  step %.1f fused %.1f us' % (d['n'], d['shuffled'], d['pass_a_us'], d['pass_b_us'], d['step_us'], d['fused_us']))
"
  MSGWAM_B200_LIB=$PWD/$f python tools/col_timing.py 1e6 ${NSTEPS:-40} 2>&1 | cut -c1-400
done
