for t in 32 64 128 256; do
  make -C d3il_b200/csrc -B EXTRA="-DIK_THREADS=$t" > /dev/null 2>&1
  python bench.py --steps 60 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('IK_THREADS $t', round(d['value']), round(d['ms_per_step'],3))"
done
