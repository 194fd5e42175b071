mkdir -p gpurun_out/r2e
run() { python profiles/run_variant.py bench.py --steps 80 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value', round(d['value']))"; }
( D3IL_VARIANT=diag run base
D3IL_VARIANT=diag D3IL_NO_PDL=1 run nopdl
D3IL_VARIANT=ik64 run ik64
D3IL_VARIANT=ik256 run ik256
D3IL_VARIANT=diag run base2 ) | tee gpurun_out/r2e/sweep_pdl.log
