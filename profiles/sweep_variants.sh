set -e
cd d3il_b200/csrc
for cfg in "32 7" "32 14" "32 16" "16 16" "16 14"; do
  set -- $cfg
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ftz=true -prec-div=false -prec-sqrt=false -Xcompiler -fPIC -DG_LANES=$1 -DENVS_PER_CTA=$2 -shared -o libd3il.so d3il_capi.cu 2>/dev/null
  cd ../..
  echo "== G=$1 EPC=$2"
  python bench.py --steps 20 --warmup 4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'k_env',round(d['roofline']['kernel_ms'],2),'k_ik',round(d['roofline']['ik_kernel_ms'],2))" || echo failed
  cd d3il_b200/csrc
done
