#!/bin/bash
# e2e (host-buffer C-ABI call): zero-copy single launch vs the chunked copy pipeline at several chunk counts; GPU box.
cd "$(dirname "$0")/.."
run() { # label env...
  env "${@:2}" python bench.py --steps ${STEPS:-200} --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('%-22s kernel %.4f ms/step  e2e %.3e env-steps/s (%.4f ms/step)' % ('$1', d['ms_per_step'], d['e2e']['value'], 1e3*65536/d['e2e']['value']))"
}
run "zero-copy" SO100_HOST_ZEROCOPY=1
for c in ${CHUNKS:-1 2 4 8}; do run "copy pipeline x$c" SO100_HOST_ZEROCOPY=0 SO100_HOST_CHUNKS=$c; done
