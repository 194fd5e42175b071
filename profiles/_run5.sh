mkdir -p gpurun_out/r2e
python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e/pytest.log
for i in 1 2; do
python bench.py --steps 100 --warmup 20 > gpurun_out/r2e/bench_b$i.json 2> gpurun_out/r2e/bench_b.err
D3IL_VARIANT=nowood python profiles/run_variant.py bench.py --steps 100 --warmup 20 > gpurun_out/r2e/bench_nowood$i.json 2> gpurun_out/r2e/bench_nowood.err
done
D3IL_VARIANT=timingfree python profiles/steady_phase.py 40 > gpurun_out/r2e/phase_free.log 2>&1
for cfg in "16 4" "32 2" "64 1" "24 4" "32 4"; do set -- $cfg
  D3IL_VARIANT=diag D3IL_N_FREE=$1 D3IL_FPC=$2 python profiles/run_variant.py bench.py --steps 60 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n_free', $1, 'fpc', $2, 'ms_per_step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value', round(d['value']))"
done | tee gpurun_out/r2e/sweep_free2.log
tail -3 gpurun_out/r2e/pytest.log
for f in gpurun_out/r2e/bench_b?.json gpurun_out/r2e/bench_nowood?.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))"; done
