#!/bin/bash
# Run on the GPU box (gpurun): launch list of the bench command + one full capture of the dominant kernel.
set -e
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 12 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:k_env -s 8 -c 1 -o gpurun_out/prof_kenv_r1 -f python profiles/profile_step.py 4096 12 > gpurun_out/prof_kenv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ik -s 8 -c 1 -o gpurun_out/prof_kik_r1 -f python profiles/profile_step.py 4096 12 > gpurun_out/prof_kik.log 2>&1
