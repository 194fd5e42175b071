mkdir -p gpurun_out/r2e
D3IL_VARIANT=timingfree python profiles/steady_phase.py 40 > gpurun_out/r2e/phase_free.log 2>&1
D3IL_VARIANT=timingmed python profiles/steady_phase.py 40 > gpurun_out/r2e/phase_med.log 2>&1
head -24 gpurun_out/r2e/phase_free.log
