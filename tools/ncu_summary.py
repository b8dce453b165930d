#!/usr/bin/env python
"""Turn an Nsight Compute report (.ncu-rep, captured under gpurun with --set full) into the two files kept under
profiles/: a readable summary (the metrics DESIGN.md quotes) and a JSON record that bench.py reads for
roofline.traffic and roofline.executed.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r2_name "command line that was profiled"
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in data:
        name = r[col["Kernel Name"]]
        m = {}
        for h in hdr:
            if h in KEEP or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio")):
                v = r[col[h]].replace(",", "")
                try:
                    m[h] = float(v)
                except ValueError:
                    m[h] = v
                m[h + "#unit"] = units[col[h]]
        kernels.append({"name": name, "metrics": m})

    def val(k, h, scale_units=True):
        v = k["metrics"].get(h)
        u = k["metrics"].get(h + "#unit", "")
        if v is None or isinstance(v, str):
            return None
        if scale_units:
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0,
                    "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(u)
            if mult is not None:
                return v * mult
        return v

    js = {"command": cmd, "report": rep, "kernels": []}
    with open(out + ".txt", "w") as f:
        f.write(f"{cmd}\n(ncu --set full --clock-control none; cold-cache, serialised launches)\n")
        for k in kernels:
            f.write(f"\n==== {k['name']}\n")
            for h in sorted(k["metrics"]):
                if h.endswith("#unit"):
                    continue
                f.write(f"   {h:90s} {k['metrics'][h]!s:>16} {k['metrics'].get(h + '#unit', '')}\n")
            t = val(k, "gpu__time_duration.sum")
            js["kernels"].append({
                "name": k["name"], "time_s": t,
                "dram_bytes": (val(k, "dram__bytes_read.sum") or 0) + (val(k, "dram__bytes_write.sum") or 0),
                "dram_read_bytes": val(k, "dram__bytes_read.sum"), "dram_write_bytes": val(k, "dram__bytes_write.sum"),
                "fp64_pipe_active_pct": val(k, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", False),
                "fp64_pipe_elapsed_pct": val(k, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", False),
                "fma_pipe_active_pct": val(k, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False),
                "issue_active_pct": val(k, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
                "warps_active_pct": val(k, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
                "inst_executed": val(k, "smsp__inst_executed.sum", False),
                "registers": val(k, "launch__registers_per_thread", False),
            })
    tot = sum(k["time_s"] or 0 for k in js["kernels"])
    if tot > 0:
        js["time_weighted"] = {
            "fp64_pipe_elapsed_pct": sum((k["fp64_pipe_elapsed_pct"] or 0) * (k["time_s"] or 0) for k in js["kernels"]) / tot,
            "fp64_pipe_active_pct": sum((k["fp64_pipe_active_pct"] or 0) * (k["time_s"] or 0) for k in js["kernels"]) / tot,
            "issue_active_pct": sum((k["issue_active_pct"] or 0) * (k["time_s"] or 0) for k in js["kernels"]) / tot,
            "time_s": tot, "dram_bytes": sum(k["dram_bytes"] for k in js["kernels"]),
        }
    with open(out + ".json", "w") as f:
        json.dump(js, f, indent=1)
    print(json.dumps(js.get("time_weighted"), indent=1))


if __name__ == "__main__":
    main()
