mkdir -p gpurun_out/r2e
python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e/pytest.log
for i in 1 2; do python bench.py --steps 100 --warmup 20 > gpurun_out/r2e/bench_d$i.json 2> gpurun_out/r2e/bench_d.err; done
python bench.py --workload stacking --steps 40 --warmup 10 > gpurun_out/r2e/bench_stacking.json 2> gpurun_out/r2e/bench_stacking.err
python bench.py --workload mixed7 --steps 40 --warmup 10 > gpurun_out/r2e/bench_mixed7.json 2> gpurun_out/r2e/bench_mixed7.err
python bench.py --workload sorting4-ddpm --steps 40 --warmup 5 > gpurun_out/r2e/bench_s4.json 2> gpurun_out/r2e/bench_s4.err
tail -5 gpurun_out/r2e/pytest.log
for f in gpurun_out/r2e/bench_d?.json gpurun_out/r2e/bench_stacking.json gpurun_out/r2e/bench_mixed7.json gpurun_out/r2e/bench_s4.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), d.get('env_step_fault_bit_counts'))"; done
