mkdir -p gpurun_out/r2e
python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e/pytest.log
python bench.py --steps 100 --warmup 20 > gpurun_out/r2e/bench_a.json 2> gpurun_out/r2e/bench_a.err
D3IL_VARIANT=timingfree python profiles/steady_phase.py 40 > gpurun_out/r2e/phase_free.log 2>&1
D3IL_VARIANT=timingmed python profiles/steady_phase.py 40 > gpurun_out/r2e/phase_med.log 2>&1
D3IL_VARIANT=timing python profiles/timeline.py 4096 520 stagger > gpurun_out/r2e/timeline_free_520.log 2>&1
tail -3 gpurun_out/r2e/pytest.log; cut -c1-200 gpurun_out/r2e/bench_a.json
