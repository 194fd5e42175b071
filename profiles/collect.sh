#!/bin/bash
# Run on the GPU box (gpurun): launch list of the bench command + one full capture of the dominant kernel, both in the
# steady-state episode mix (after the 400-step pre-roll).
mkdir -p gpurun_out
TAG=${1:-r1}
# (1) launch list: per-launch durations of the timed region of `bench.py --steps 12 --warmup 3` (pre-roll launches skipped)
ncu --metrics gpu__time_duration.sum --clock-control none -s 6200 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 12 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1 || true
# (2) full captures: k_env and k_ik of step 405 (steady state)
ncu --set full --clock-control none --import-source on -k regex:k_env -s 404 -c 1 -o gpurun_out/prof_kenv_$TAG -f python profiles/profile_step.py 4096 8 400 > gpurun_out/prof_kenv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ik -s 404 -c 1 -o gpurun_out/prof_kik_$TAG -f python profiles/profile_step.py 4096 8 400 > gpurun_out/prof_kik.log 2>&1
ls -la gpurun_out | tail -8
