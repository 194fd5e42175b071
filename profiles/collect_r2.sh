#!/bin/bash
# Round-2 profile collection (run on the GPU box through gpurun): launch list of the bench command in the steady-state mix,
# one full ncu capture of the dominant kernel, and the records of the wider BASELINE configs on one GPU.
mkdir -p gpurun_out/r2prof
# (1) launch list: per-launch durations inside `bench.py --steps 12 --warmup 3 --no-graph` after the 400-step pre-roll
#     (~27 launches per env step incl. the torch element-wise kernels of the synthetic action stream)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 11000 -c 400 --csv --log-file gpurun_out/r2prof/launches_r2.csv \
    python bench.py --steps 12 --warmup 3 --no-graph > gpurun_out/r2prof/bench_under_ncu.log 2>&1 || true
# (2) full capture of k_env at step 405 (steady state), with source
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env -s 404 -c 1 -o gpurun_out/r2prof/prof_kenv_r2 -f \
    python profiles/profile_step.py 4096 8 400 > gpurun_out/r2prof/prof_kenv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ik -s 404 -c 1 -o gpurun_out/r2prof/prof_kik_r2 -f \
    python profiles/profile_step.py 4096 8 400 > gpurun_out/r2prof/prof_kik.log 2>&1
# (3) BASELINE config 3 on one GPU (Sorting-4, 8192 envs, DDPM-MLP in the loop) and the headline once more
python bench.py --workload sorting4-ddpm --steps 40 --warmup 5 > gpurun_out/r2prof/sorting4_ddpm_n1.json 2> gpurun_out/r2prof/sorting4_ddpm_n1.err
python bench.py --steps 200 --warmup 20 > gpurun_out/r2prof/pushing_n1.json 2> gpurun_out/r2prof/pushing_n1.err
ls -la gpurun_out/r2prof | tail -12
