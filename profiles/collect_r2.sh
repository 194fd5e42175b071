#!/bin/bash
# Round-2 profile collection (run on the GPU box through gpurun): launch list of the bench command in the steady-state mix,
# one full ncu capture of the dominant kernel and of k_ik, and the headline / Stacking records on one GPU.
#   python profiles/summarize.py r2 r2prof     (afterwards, in the dev container) -> profiles/r2_summary.md, r2_launches.csv
mkdir -p gpurun_out/r2prof
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 11000 -c 400 --csv --log-file gpurun_out/r2prof/launches_r2.csv \
    python bench.py --steps 12 --warmup 3 --no-graph > gpurun_out/r2prof/bench_under_ncu.log 2>&1 || true
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env -s 404 -c 1 -o gpurun_out/r2prof/prof_kenv_r2 -f \
    python profiles/profile_step.py 4096 8 400 > gpurun_out/r2prof/prof_kenv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ik -s 404 -c 1 -o gpurun_out/r2prof/prof_kik_r2 -f \
    python profiles/profile_step.py 4096 8 400 > gpurun_out/r2prof/prof_kik.log 2>&1
python bench.py --workload stacking --steps 100 --warmup 10 > gpurun_out/r2prof/stacking_n1.json 2> gpurun_out/r2prof/stacking_n1.err
python bench.py --steps 200 --warmup 20 > gpurun_out/r2prof/pushing_n1.json 2> gpurun_out/r2prof/pushing_n1.err
ls -la gpurun_out/r2prof | tail -12
