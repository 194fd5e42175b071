#!/bin/bash
# Diagnostic build (-DD3IL_DIAG): how many of the most expensive envs should run in single-env CTAs (D3IL_N_SINGLE)?
for ns in ${@:-1 8 15 22 29 43}; do
  D3IL_VARIANT=diag D3IL_N_SINGLE=$ns python profiles/run_variant.py bench.py --steps 60 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n_single', $ns, 'ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value']))"
done
