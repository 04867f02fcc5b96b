#!/bin/bash
# e2e (host-buffer C-ABI call) vs the number of env chunks of the pipelined host path; run on the GPU box.
cd "$(dirname "$0")/.."
for c in ${CHUNKS:-1 2 4 8 16}; do
  SO100_HOST_CHUNKS=$c python bench.py --steps ${STEPS:-200} --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('chunks %-3s kernel %.4f ms/step  e2e %.3e env-steps/s (%.4f ms/step)' % ('$c', d['ms_per_step'], d['e2e']['value'], 1e3*65536/d['e2e']['value']))"
done
