mkdir -p gpurun_out/r2e
for w in 430 520 610; do D3IL_VARIANT=timing python profiles/timeline.py 4096 $w stagger > gpurun_out/r2e/timeline_mix_$w.log 2>&1; done
tail -8 gpurun_out/r2e/timeline_mix_430.log
