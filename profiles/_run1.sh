mkdir -p gpurun_out/r2e
python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e/pytest.log
for i in 1 2; do
python bench.py --steps 100 --warmup 20 > gpurun_out/r2e/bench_e8_$i.json 2> gpurun_out/r2e/bench_e8_$i.err
D3IL_VARIANT=e7 python profiles/run_variant.py bench.py --steps 100 --warmup 20 > gpurun_out/r2e/bench_e7_$i.json 2> gpurun_out/r2e/bench_e7_$i.err
done
D3IL_VARIANT=timing0 python profiles/steady_phase.py 20 > gpurun_out/r2e/phase0.log 2>&1
D3IL_VARIANT=timing python profiles/timeline.py > gpurun_out/r2e/timeline.log 2>&1
python bench.py --workload mixed7 --steps 40 --warmup 10 > gpurun_out/r2e/bench_mixed7.json 2> gpurun_out/r2e/bench_mixed7.err
python bench.py --workload stacking --steps 40 --warmup 10 > gpurun_out/r2e/bench_stacking.json 2> gpurun_out/r2e/bench_stacking.err
tail -3 gpurun_out/r2e/pytest.log; cat gpurun_out/r2e/bench_e8_1.json | cut -c1-300
