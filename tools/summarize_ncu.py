"""Developer tool: turn gpurun_out/*.ncu-rep and launch lists into the text summaries kept under profiles/."""
import csv, json, os, subprocess, sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
    "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
]


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def summarize_report(rep, title, note=""):
    hdr, units, rows = raw_page(rep)
    lines = [f"## {title}", "", f"source: `{os.path.basename(rep)}` (ncu --set full --clock-control none --import-source on); {note}", ""]
    for r in rows:
        d = dict(zip(hdr, r))
        lines.append(f"### {d['Kernel Name'][:90]}  grid={d['Grid Size']} block={d['Block Size']}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in KEYS:
            if k in d:
                lines.append(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
        try:
            rd = float(d["dram__bytes_read.sum"]); wr = float(d["dram__bytes_write.sum"]); t = float(d["gpu__time_duration.sum"])
            u = units[hdr.index("dram__bytes_read.sum")]
            lines.append(f"| **dram read+write** | {rd + wr:.4f} | {u} |")
        except Exception:
            pass
        lines.append("")
    return "\n".join(lines)


def summarize_launches(path, title):
    rows = list(csv.reader(open(path, errors="replace")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    agg = defaultdict(lambda: [0, 0.0])
    order = []
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        short = name.split("(")[0][-70:] if "tfx" in name else name.split("<")[0][-60:]
        key = (short, d["Grid Size"], d["Block Size"])
        if key not in agg:
            order.append(key)
        agg[key][0] += 1
        agg[key][1] += float(d["Metric Value"]) / 1e6
    total = sum(v[1] for v in agg.values())
    lines = [f"## {title}", "", f"source: `{os.path.basename(path)}` (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: compare shares)", "",
             "| kernel | grid | block | launches | total ms | share |", "|---|---|---|---|---|---|"]
    for key in sorted(order, key=lambda k: -agg[k][1]):
        n, ms = agg[key]
        lines.append(f"| `{key[0]}` | {key[1]} | {key[2]} | {n} | {ms:.3f} | {100 * ms / total:.1f} % |")
    lines.append("")
    return "\n".join(lines)


if __name__ == "__main__":
    spec = json.load(open(sys.argv[1]))
    out = []
    for item in spec["items"]:
        p = os.path.join(ROOT, item["path"])
        if item["kind"] == "report":
            out.append(summarize_report(p, item["title"], item.get("note", "")))
        else:
            out.append(summarize_launches(p, item["title"]))
    open(os.path.join(ROOT, spec["out"]), "w").write(spec.get("header", "") + "\n\n" + "\n\n".join(out) + "\n")
    print("wrote", spec["out"])
