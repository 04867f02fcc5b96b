#!/bin/bash
# Step-kernel time against the env count (decorrelated start, L2 flushed between steps); run on the GPU box.
cd "$(dirname "$0")/.."
echo "# envs  ms/step  env-steps/s  useful-FLOP frac of the FFMA peak   (task ${TASK:-Env01}, flags ${FLAGS:-0})"
for n in ${SIZES:-8192 16384 32768 49152 65536 75776 98304 131072 151552 196608 262144 524288 1048576}; do
  python bench.py --envs-per-gpu $n --steps ${STEPS:-100} --warmup 10 --task ${TASK:-Env01} --flags ${FLAGS:-0} --no-cpu-baseline --no-e2e --no-tasks --no-ppo 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('%8d  %.4f  %.3e  %.3f' % ($n, d['ms_per_step'], d['value'], d['roofline']['frac']))"
done
