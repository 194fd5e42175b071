mkdir -p gpurun_out/r2e
run() { python bench.py --steps 100 --warmup 10 $2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value', round(d['value']))"; }
( run graph
run nograph --no-graph
run graph2
run nograph2 --no-graph ) | tee gpurun_out/r2e/graphmode.log
