"""ncu report -> profiles/r02_issue_counts.json (what bench.py's roofline reads).

  ncu -i gpurun_out/prof_<workload>.ncu-rep --page raw --csv > /tmp/raw.csv
  python profiles/r02_collect.py <workload> /tmp/raw.csv [--frame-kernels N] [--commit HASH]

Takes the render kernels (k_whitted_* / k_montecarlo) of the LAST profiled frame -- tools/profile_frame.py renders a few
warm frames of one view, so that is a steady-state frame -- and records, per kernel and summed: warp instructions
(smsp__inst_executed.sum), thread instructions, DRAM bytes, duration under ncu, issue-active %, IPC, L1 hit rate.
"""
import csv
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "r02_issue_counts.json")
RENDER = re.compile(r"k_whitted|k_montecarlo")


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return None


def main():
    wl, path = sys.argv[1], sys.argv[2]
    n_last = int(sys.argv[sys.argv.index("--frame-kernels") + 1]) if "--frame-kernels" in sys.argv else None
    commit = sys.argv[sys.argv.index("--commit") + 1] if "--commit" in sys.argv else "?"
    rows = list(csv.reader(open(path, newline="")))
    header = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names = rows[header]
    units = rows[header + 1]
    data = [dict(zip(names, r)) for r in rows[header + 2:] if len(r) == len(names)]
    data = [r for r in data if RENDER.search(r["Kernel Name"])]
    if n_last is None:  # the kernels after the last k_whitted_chain / k_montecarlo launch that STARTS a frame: take distinct names from the tail
        seen, tail = set(), []
        for r in reversed(data):
            short = r["Kernel Name"].split("<")[0].split("(")[0]
            if short in seen:
                break
            seen.add(short)
            tail.append(r)
        data = list(reversed(tail))
    else:
        data = data[-n_last:]
    unit_of = dict(zip(names, units))

    def get(r, key, scale_bytes=False):
        v = num(r.get(key))
        if v is None:
            return None
        if scale_bytes:
            u = unit_of.get(key, "")
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return v

    kernels, total = [], dict(warp_inst=0.0, thread_inst=0.0, dram=0.0, dur=0.0)
    for r in data:
        dur = get(r, "gpu__time_duration.sum")
        dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit_of.get("gpu__time_duration.sum", "ns"), 1e-6) if dur is not None else None
        k = {"kernel": r["Kernel Name"].split("(")[0][:120], "grid": r.get("Grid Size"), "block": r.get("Block Size"),
             "warp_inst": get(r, "smsp__inst_executed.sum"),
             # --set full carries the lanes-per-instruction ratio, not the thread-instruction sum
             "thread_inst": get(r, "smsp__thread_inst_executed.sum") if get(r, "smsp__thread_inst_executed.sum") is not None else
             ((get(r, "smsp__inst_executed.sum") or 0) * get(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
              if get(r, "smsp__thread_inst_executed_per_inst_executed.ratio") is not None else None),
             "dram_read": get(r, "dram__bytes_read.sum", True), "dram_write": get(r, "dram__bytes_write.sum", True),
             "duration_ms_under_ncu": dur_ms,
             "issue_active_pct": get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "ipc_per_sm": get(r, "sm__inst_executed.avg.per_cycle_active"),
             "l1_hit_pct": get(r, "l1tex__t_sector_hit_rate.pct"),
             "registers": get(r, "launch__registers_per_thread"),
             "local_bytes_per_thread": None,
             "achieved_occupancy_pct": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active")}
        kernels.append(k)
        total["warp_inst"] += k["warp_inst"] or 0
        total["thread_inst"] += k["thread_inst"] or 0
        total["dram"] += (k["dram_read"] or 0) + (k["dram_write"] or 0)
        total["dur"] += k["duration_ms_under_ncu"] or 0
    top = max(kernels, key=lambda k: k["warp_inst"] or 0)
    entry = {"warp_inst": total["warp_inst"], "thread_inst": total["thread_inst"],
             "thread_inst_per_warp_inst": total["thread_inst"] / total["warp_inst"] if total["warp_inst"] else None,
             "dram_bytes": total["dram"], "duration_ms_under_ncu": total["dur"], "issue_active_pct": top["issue_active_pct"],
             "ipc_per_sm": top["ipc_per_sm"], "l1_hit_pct": top["l1_hit_pct"], "kernels": kernels}
    prof = {"source": "", "workloads": {}}
    if os.path.exists(OUT):
        prof = json.load(open(OUT))
    prof["workloads"][wl] = entry
    prof["source"] = f"ncu --set full --clock-control none of tools/profile_frame.py (last frame's render kernels), commit {commit}; profiles/r02_collect.py"
    json.dump(prof, open(OUT, "w"), indent=1)
    print(json.dumps({k: v for k, v in entry.items() if k != "kernels"}, indent=1))
    for k in kernels:
        print(k["kernel"][:70], k["warp_inst"], k["duration_ms_under_ncu"], k["issue_active_pct"])


if __name__ == "__main__":
    main()
