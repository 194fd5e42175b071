mkdir -p gpurun_out/r2e
for st in 50 63 64 65; do
D3IL_TL=1 D3IL_VARIANT=timing python profiles/run_variant.py bench.py --steps $st --warmup 10 > gpurun_out/r2e/bench_tl.json 2> gpurun_out/r2e/bench_tl_$st.err
grep "timeline. k_env\|last k_env\|mean period" gpurun_out/r2e/bench_tl_$st.err
done
