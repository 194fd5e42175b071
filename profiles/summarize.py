"""Turn gpurun_out/<sub>/{launches_<tag>.csv, prof_kenv_<tag>.ncu-rep, prof_kik_<tag>.ncu-rep} into profiles/<tag>_*.{csv,md}
(run in the dev container, no GPU):  python profiles/summarize.py r2 r2prof"""
import csv
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_md = []


def raw(rep):
    o = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(o.splitlines()))
    return dict(zip(rows[0], rows[-1]))


# ---- launch list
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
SUB = sys.argv[2] if len(sys.argv) > 2 else ""
lp = os.path.join(ROOT, "gpurun_out", SUB, f"launches_{TAG}.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            agg[r[ik].split("(")[0]][0] += 1
            agg[r[ik].split("(")[0]][1] += float(r[iv].replace(",", "")) / 1e6
        except (ValueError, IndexError):
            pass
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{TAG}_launches.csv"), "w") as f:
        f.write("kernel,launches,total_ms,share\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k},{n},{ms:.3f},{ms / tot:.4f}\n")
    out_md.append("## Launch list (ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 12 --warmup 3 --no-graph`, timed-region launches after the pre-roll)\n")
    out_md.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1])[:12]:
        out_md.append(f"| `{k}` | {n} | {ms:.2f} | {100 * ms / tot:.1f} % |")
    out_md.append("")

keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for name in (f"prof_kenv_{TAG}", f"prof_kik_{TAG}"):
    rep = os.path.join(ROOT, "gpurun_out", SUB, name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    d = raw(rep)
    out_md.append(f"## `{name}` (ncu --set full --clock-control none, 4096 envs, one launch)\n")
    out_md.append("| metric | value |\n|---|---|")
    for k in keys:
        if k in d:
            out_md.append(f"| {k} | {d[k]} |")
    stalls = sorted(((float(v), k) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)[:8]
    out_md.append("\nTop warp stall reasons (warps per issue-active cycle):\n")
    for v, k in stalls:
        out_md.append(f"* {k.split('issue_stalled_')[1].split('_per_issue')[0]}: {v:.2f}")
    out_md.append("")
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_lines.py"), rep, "25"], capture_output=True, text=True).stdout
    out_md.append("Hottest source lines (warp instructions executed):\n\n```\n" + lines + "```\n")
open(os.path.join(ROOT, "profiles", f"{TAG}_summary.md"), "w").write(f"# Round {TAG[1:]} ncu summary (B200)\n\n" + "\n".join(out_md))
print("\n".join(out_md)[:3000])
