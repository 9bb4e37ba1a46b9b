"""One `ncu --set full` report -> a tracked markdown summary under profiles/ (and, for the emit kernel, profiles/emit_traffic.json).

usage: python tools/summarize_ncu.py <out.md> <report.ncu-rep> "<what was captured>" <algorithmic bytes or flops per launch> <bytes|flops> [--emit-traffic ROWS]
"""
import csv
import json
import subprocess
import sys

out_md, rep, what, alg, kind = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, rows = rr[0], rr[1], rr[2:]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_max_active", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "thread_inst_executed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}
import os
commit = os.environ.get("KD_COMMIT") or subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or "?"
lines = [f"# ncu --set full --clock-control none: {what}", "", f"report: `{rep}` (scratch), summarised at commit {commit}; numbers under the profiler are not bench values", ""]
first = None
for r in rows:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    name = d.get("Kernel Name", "?").replace("kd::<unnamed>::", "kd::")
    lines += [f"## `{name[:150]}`", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in d and d[k] not in ("", "n/a"):
            lines.append(f"| {k} | {d[k]} | {u.get(k, '')} |")

    def val(k):
        return float(d[k].replace(",", "")) * UNIT.get(u.get(k, ""), 1.0)
    try:
        traffic = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        dur = val("gpu__time_duration.sum")
        lines += ["", f"DRAM traffic {traffic / 1e9:.3f} GB per launch; duration under ncu {dur * 1e3:.3f} ms"]
        if kind == "bytes":
            lines.append(f"algorithmic bytes per launch {alg / 1e9:.3f} GB -> traffic / algorithmic = {traffic / alg:.3f}; "
                         f"algorithmic GB/s under ncu {alg / dur / 1e9:.0f}")
        else:
            lines.append(f"algorithmic flops per launch {alg / 1e12:.3f} TFLOP -> {alg / dur / 1e12:.0f} TFLOP/s under ncu (cold, serialised)")
        if first is None:
            first = (name, traffic)
    except Exception as exc:  # noqa: BLE001
        lines.append(f"(summary unavailable: {exc})")
    lines.append("")
open(out_md, "w").write("\n".join(lines) + "\n")
if "--emit-traffic" in sys.argv and first:
    n = int(sys.argv[sys.argv.index("--emit-traffic") + 1])
    json.dump({"kernel": first[0][:120], "dram_bytes_per_row": first[1] / n, "rows": n, "source": out_md, "commit": commit,
               "made_by": "tools/summarize_ncu.py from the ncu --set full capture of tools/profile_ops.py loca_reference"},
              open(os.path.join(os.path.dirname(out_md), "emit_traffic.json"), "w"))
print("\n".join(lines[:60]))
