"""Turn the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

usage: python tools/summarize_profiles.py <tag> <launches.csv> <full.ncu-rep> [rows_in_capture]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
rows_in_capture = int(sys.argv[4]) if len(sys.argv) > 4 else 16384
os.makedirs("profiles", exist_ok=True)
V = 151936


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    return name.replace("kd::<unnamed>::", "kd::")


out = [f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none)", "",
       "Per-launch times are cold-cache and serialised by ncu: compare SHARES, not absolutes.", ""]
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.defaultdict(lambda: [0, 0.0, "", ""])
ours = 0.0
for r in data:
    if len(r) <= iv:
        continue
    k, v = short(r[ik]), float(r[iv].replace(",", ""))
    if not k.startswith("kd::"):
        continue
    a = agg[k]
    a[0] += 1; a[1] += v; a[2] = r[ig]; a[3] = r[ib]
    ours += v
out += ["| kernel | launches | grid | block | avg us | share of our kernels |", "|---|---|---|---|---|---|"]
for k, (c, v, g, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"| `{k}` | {c} | {g} | {b} | {v / c / 1e3:.1f} | {100 * v / ours:.1f} % |")
open(f"profiles/{tag}_launches.md", "w").write("\n".join(out) + "\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
d = dict(zip(rr[0], rr[2]))
u = dict(zip(rr[0], rr[1]))
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_max_active", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "inst_executed", "thread_inst_executed", "smsp__inst_executed_op_tma_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
lines = [f"# {tag}: ncu --set full, dominant kernel `{short(rr[2][rr[0].index('Kernel Name')])}`", "",
         f"capture: {rows_in_capture} rows x {V} columns bf16 (one launch)", "", "| metric | value | unit |", "|---|---|---|"]
for k in keys:
    if k in d:
        lines.append(f"| {k} | {d[k]} | {u.get(k, '')} |")
def f(x):
    return float(d[x].replace(",", ""))
unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
traffic = f("dram__bytes_read.sum") * unit[u["dram__bytes_read.sum"]] + f("dram__bytes_write.sum") * unit[u["dram__bytes_write.sum"]]
per_row = traffic / rows_in_capture
lines += ["", f"DRAM traffic per launch: {traffic / 1e9:.3f} GB = {per_row:.0f} B/row (algorithmic 6*V = {6 * V} B/row)",
          f"thread instructions per (row, column) cell: {f('thread_inst_executed') / (rows_in_capture * V):.2f}"]
open(f"profiles/{tag}_emit_full.md", "w").write("\n".join(lines) + "\n")
json.dump({"kernel": short(rr[2][rr[0].index('Kernel Name')]), "dram_bytes_per_row": per_row, "rows_in_capture": rows_in_capture,
           "source": f"profiles/{tag}_emit_full.md"}, open("profiles/emit_traffic.json", "w"))
print("\n".join(out[-8:]))
print("\n".join(lines[-3:]))
