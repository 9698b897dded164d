"""ncu raw pages -> the metric table of profiles/r02*_ncu_summary.md.

  ncu -i gpurun_out/prof_X.ncu-rep --page raw --csv > /tmp/raw_X.csv        (one per capture)
  python profiles/r02_summary_table.py "column title=/tmp/raw_X.csv[:kernel regex]" ...
"""
import csv
import re
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "pcie__write_bytes.sum.per_second",
           "syslts__t_sectors_srcunit_tex_aperture_sysmem_op_write_lookup_miss.sum",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__warps_eligible.avg.per_cycle_active"]


def load(path, pattern):
    rows = list(csv.reader(open(path, newline="")))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[h], rows[h + 1]
    data = [dict(zip(names, r)) for r in rows[h + 2:] if len(r) == len(names)]
    data = [r for r in data if re.search(pattern, r["Kernel Name"])]
    return data[-1], dict(zip(names, units))


def fmt(v, unit):
    try:
        x = float(str(v).replace(",", ""))
    except ValueError:
        return str(v)
    if unit in ("ns",):
        return f"{x / 1e6:.2f} ms"
    if unit == "us":
        return f"{x / 1e3:.2f} ms"
    if unit in ("inst", "sector") or x >= 1e6:
        return f"{x:,.0f}"
    return f"{x:.2f}"


cols = []
for arg in sys.argv[1:]:
    title, rest = arg.split("=", 1)
    path, _, pattern = rest.partition(":")
    row, units = load(path, pattern or "k_whitted|k_montecarlo")
    cols.append((title + ": `" + row["Kernel Name"].split("(")[0].replace("void ", "").replace("rtb::", "") + "`", row, units))
print("| metric | " + " | ".join(c[0] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for m in METRICS:
    cells = []
    for _, row, units in cols:
        cells.append(fmt(row.get(m, ""), units.get(m, "")) + (" " + units.get(m, "") if units.get(m, "") not in ("", "ns", "us", "inst", "%") else ("" if units.get(m, "") != "%" else " %")) if m in row else "-")
    print(f"| `{m}` | " + " | ".join(cells) + " |")
