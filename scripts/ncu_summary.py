#!/usr/bin/env python
"""Summarise an ncu report (--set full) into a small markdown table for profiles/.

  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep "title / command" > profiles/rN_ncu_<kernel>.md
  python scripts/ncu_summary.py --traffic profiles/r2_traffic.json name=report.ncu-rep[:kernel-regex] ...
      per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches of the kernel)
      keyed by the names of bench.py's per-kernel profile; bench.py reads that file for `roofline.traffic`.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (of elapsed)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2->SM rate"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
]


def _raw(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[1], rows[2:]


def traffic(out_path, specs):
    import json
    import re
    out = {}
    for spec in specs:
        name, rest = spec.split("=", 1)
        rep, _, rx = rest.partition(":")
        hdr, units, data = _raw(rep)
        ix = {h: i for i, h in enumerate(hdr)}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals, durs = [], []
        for r in data:
            if rx and not re.search(rx, r[ix["Kernel Name"]]):
                continue
            b = sum(float(r[ix[k]]) * scale[units[ix[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            vals.append(b)
            durs.append(float(r[ix["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[ix["gpu__time_duration.sum"]]])
        if vals:
            out[name] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches_captured": len(vals),
                         "duration_us_under_ncu": sum(durs) / len(durs), "kernel": rx or "all",
                         "source": "profiles/" + rep.split("/")[-1].replace(".ncu-rep", ".md").replace("r2_prof_", "r2_ncu_")}
    with open(out_path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")


def main():
    if sys.argv[1] == "--traffic":
        return traffic(sys.argv[2], sys.argv[3:])
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary: {title}\n")
    print(f"source report: `{rep}` (scratch; numbers below are per launch, profiler-serialised, cold cache)\n")
    names = [r[ix["Kernel Name"]][:60] for r in data]
    print("| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
    print("|---|" + "---|" * len(data))
    print("| kernel | " + " | ".join(names) + " |")
    for key, label in KEYS:
        if key not in ix:
            continue
        i = ix[key]
        vals = []
        for r in data:
            try:
                vals.append(f"{float(r[i]):.4g}")
            except ValueError:
                vals.append(r[i])
        print(f"| {label} [{units[i]}] | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
