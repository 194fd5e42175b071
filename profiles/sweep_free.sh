#!/bin/bash
# Diagnostic build (-DD3IL_DIAG): how many free-running CTAs (D3IL_N_FREE) with how many envs each (D3IL_FPC)?
mkdir -p gpurun_out/r2e
for cfg in "16 4" "8 8" "32 2" "64 1" "32 4" "16 8" "8 1" "24 4" "0 4"; do set -- $cfg
  D3IL_VARIANT=diag D3IL_N_FREE=$1 D3IL_FPC=$2 python profiles/run_variant.py bench.py --steps 60 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n_free', $1, 'fpc', $2, 'ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value', round(d['value']))"
done | tee gpurun_out/r2e/sweep_free.log
D3IL_VARIANT=timing python profiles/timeline.py 4096 520 stagger > gpurun_out/r2e/timeline_free_520.log 2>&1
