#!/bin/bash
# Which phase barriers of the tick pay for themselves with 16 lock-step warps?  Variants built with
#   for v in 1 2 4 8 16 32 64; do bash profiles/build_variant.sh skip$v "-DD3IL_SKIP_BARS=$v"; done; bash profiles/build_variant.sh base ''
mkdir -p gpurun_out/r2e
run() { python profiles/run_variant.py bench.py --steps 60 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value']))"; }
( for v in base skip1 skip2 skip4 skip8 skip16 skip32 skip64 base; do D3IL_VARIANT=$v run $v; done ) | tee gpurun_out/r2e/sweep_barriers.log
