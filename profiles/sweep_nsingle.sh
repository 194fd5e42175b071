for ns in 0 1 15 50; do
  D3IL_N_SINGLE=$ns python bench.py --steps 60 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n_single $ns', round(d['value']), round(d['ms_per_step'],3))"
done
