#!/bin/bash
# e2e through the host-buffer C ABI: the pipelined call (so100_step_host_async over G env groups) against the
# synchronous one, and the one-CTA-per-SM variant of the synchronous launch.  GPU box.
cd "$(dirname "$0")/.."
run() { # label, bench args..., env via ENVV
  env $ENVV python bench.py --steps ${STEPS:-200} --warmup 20 --no-cpu-baseline --no-tasks --no-ppo "${@:2}" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); n=d['config']['envs_per_gpu']
    print('%-28s kernel %.4f ms/step | e2e pipelined %.3e (%.4f ms/step) | e2e sync %.3e (%.4f ms/step)' % ('$1', d['ms_per_step'], d['e2e']['value'], 1e3*n/d['e2e']['value'], d['e2e_sync']['value'], 1e3*n/d['e2e_sync']['value']))"
}
for g in ${GROUPS_LIST:-1 2 4 6 8 12 16}; do ENVV="" run "groups=$g" --e2e-groups $g; done
for w in ${DEVBUF:-in out both}; do for g in ${DEVBUF_GROUPS:-1 4 8 12}; do ENVV="SO100_HOST_ALLOW_DEVICE=1" run "device buffers: $w, groups=$g" --e2e-groups $g --e2e-device-buffers $w; done; done
for k in ${COPY_KNOBS:-"SO100_ASYNC_H2D_COPY=1" "SO100_ASYNC_D2H_COPY=1" "SO100_ASYNC_H2D_COPY=1 SO100_ASYNC_D2H_COPY=1"}; do for g in ${COPY_GROUPS:-4 8}; do ENVV="$k" run "$k groups=$g" --e2e-groups $g; done; done
