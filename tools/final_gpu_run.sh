#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): GPU tests, the default bench line, the reference arm, the ncu
# launch list of the bench command and one full capture of the shipped step kernel.  Everything lands in gpurun_out/.
set -u
tag=${1:-r2z}
python -c "from so100_mujoco_rl_b200 import _native; print(_native.lib().so100_build_id().decode())" > gpurun_out/${tag}_build_id.txt 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 400 gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
# ncu: only after the plain commands above exited; numbers printed under ncu are never bench values
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_ncu_launches.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-tasks --no-ppo --no-contact > gpurun_out/${tag}_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 70 -c 1 -f -o gpurun_out/prof_${tag}_step_65536 \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-tasks --no-ppo --no-contact > gpurun_out/${tag}_ncu_step.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:act_kernel_tc -c 1 -f -o /tmp/prof_act_tc \
  python tools/bench_ppo.py > gpurun_out/${tag}_ncu_act.log 2>&1
ncu -i /tmp/prof_act_tc.ncu-rep --page raw --csv > gpurun_out/${tag}_act_tc_ncu_raw.csv 2>/dev/null   # gpurun_out/ is capped at 64 MiB: one .ncu-rep only
timeout 120 python tools/bench_ppo.py > gpurun_out/${tag}_bench_ppo.json 2>&1
timeout 200 python tools/bench_contact.py --envs 65536 262144 > gpurun_out/${tag}_contact.txt 2>&1
ls -la gpurun_out | tail -5
