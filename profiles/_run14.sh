mkdir -p gpurun_out/r2e
python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2e/pytest.log
tail -3 gpurun_out/r2e/pytest.log
for wl in inserting sorting2 pushing; do python bench.py --workload $wl --steps 60 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', 'ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value']), d.get('env_step_fault_bit_counts'))"; done
