mkdir -p gpurun_out/r2e
python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e/pytest.log
D3IL_SEG=1 python bench.py --steps 100 --warmup 10 --no-graph 2> gpurun_out/r2e/bench_e.err > gpurun_out/r2e/bench_e_ng.json
python bench.py --steps 100 --warmup 10 > gpurun_out/r2e/bench_e.json 2>/dev/null
D3IL_VARIANT=timingdense python profiles/steady_phase.py 60 > gpurun_out/r2e/phase_dense.log 2>&1
tail -3 gpurun_out/r2e/pytest.log; grep seg gpurun_out/r2e/bench_e.err
for f in gpurun_out/r2e/bench_e_ng.json gpurun_out/r2e/bench_e.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))"; done
grep "per Newton" gpurun_out/r2e/phase_dense.log
