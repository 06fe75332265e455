#!/bin/sh
# Developer tool (on a 2-GPU box): bench.py at N = 2 with the product library and with a variant library, alternating,
# so that both are timed on the same box (box-to-box spread is ~0.5 %).   usage: sh tools/ab_two_libs.sh <variant.so>
# (How the ticketed mean-flow chain was compared with the previous one-slice-per-CTA form: the variant was the library
# built from the commit before it.)
V=$(readlink -f "$1")
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 50 --warmup 10 --no-extras --no-cpu-baseline --no-parity $2 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"; }
for i in 1 2; do
  echo "product c1:"; run 2953$i "--workload c1"
  echo "variant c1:"; MSGWAM_B200_LIB=$V run 2954$i "--workload c1"
done
echo "product c2:"; run 29511 ""
echo "variant c2:"; MSGWAM_B200_LIB=$V run 29521 ""
