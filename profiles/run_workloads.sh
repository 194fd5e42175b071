#!/bin/bash
# Run on the GPU box: one short bench line per workload (BASELINE.json configs 2-4 + the other scenes).
mkdir -p gpurun_out
for w in "$@"; do
  python bench.py --workload $w --steps 40 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || tail -3 gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$w.json"))
    print("$w", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "k_env ms", round(d["roofline"]["kernel_ms"], 2), "cpu(all cores)", round(d["cpu_baseline"]["value"]), "cores", d["cpu_baseline"]["cores"], "clk", d["clocks"]["sm_mhz"], "faults", d.get("env_step_faults"), "bits", d.get("env_step_fault_bits"))
except Exception as e:
    print("$w failed", e)
PY
done
