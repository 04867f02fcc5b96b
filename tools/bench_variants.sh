#!/bin/bash
# Times every build/libso100_*.so with bench.py (kernel-only leg); run on the GPU box.
cd "$(dirname "$0")/.."
for f in build/libso100_*.so; do
  SO100_B200_LIB=$PWD/$f python bench.py --steps ${STEPS:-60} --warmup 10 --no-cpu-baseline --no-e2e ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('%-28s %.4f ms/step  %.3e env-steps/s  fp32 frac %.3f  clk %s' % ('$f'.split('libso100_')[1], d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks']['sm_mhz']))"
done
