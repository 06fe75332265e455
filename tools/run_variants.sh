for f in tools/_variants/lib_*.so; do echo "== $f"; MSGWAM_B200_LIB=$PWD/$f python tools/kernel_timing.py 1e6 1e7 2>&1 | grep -v "shuffled.: true" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   n=%g A %.1f  B %.1f  fin %.1f  step %.1f us  -> %.3e ray-steps/s' % (d['n'], d['pass_a_us'], d['pass_b_us'], d['finish_us'], d['step_us'], d.get('ray_steps_per_s', 0)))
"; done
