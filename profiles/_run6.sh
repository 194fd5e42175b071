mkdir -p gpurun_out/r2e
python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e/pytest.log
for i in 1 2; do python bench.py --steps 100 --warmup 20 > gpurun_out/r2e/bench_c$i.json 2> gpurun_out/r2e/bench_c.err; done
D3IL_VARIANT=timingmed python profiles/steady_phase.py 40 > gpurun_out/r2e/phase_med.log 2>&1
for w in 430 520 610; do D3IL_VARIANT=timing python profiles/timeline.py 4096 $w stagger > gpurun_out/r2e/timeline_c_$w.log 2>&1; done
tail -3 gpurun_out/r2e/pytest.log
for f in gpurun_out/r2e/bench_c?.json; do python -c "
import json; d=json.load(open('$f')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))"; done
