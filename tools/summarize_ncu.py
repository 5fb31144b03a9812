#!/usr/bin/env python
"""Reduce the two ncu captures of tools/profile_round.sh to profiles/rNN_ncu_top_kernels.json.

  python tools/summarize_ncu.py gpurun_out r01     # needs the `ncu` CLI to export the reports as CSV
"""
import collections
import csv
import json
import re
import subprocess
import sys

SC = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1, "tbyte": 1e12}


def export(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def num(x):
    try:
        return float(x)
    except ValueError:
        return None


def main(d, r):
    out = {"source": "ncu --clock-control none on `python bench.py --steps 1 --warmup 1 --no-cpu-baseline` (Taobao workload, "
                     "S=4096); tools/profile_round.sh + tools/summarize_ncu.py", "kernels": [], "traffic": {}}
    hdr, units, data = export("%s/%s_full_misc.ncu-rep" % (d, r))
    idx = {h: i for i, h in enumerate(hdr)}
    U = {h: units[i] for i, h in enumerate(hdr)}
    gb = lambda row, k: num(row[idx[k]]) * SC[U[k].lower()]
    agg = collections.defaultdict(list)
    for row in data:
        name = re.sub(r"\(.*", "", row[idx["Kernel Name"]]).replace("void ", "")
        k = dict(name=name, set="full", dur_us=num(row[idx["gpu__time_duration.sum"]]),
                 dram_read_bytes=gb(row, "dram__bytes_read.sum"), dram_write_bytes=gb(row, "dram__bytes_write.sum"),
                 dram_pct=num(row[idx["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                 warps_active_pct=num(row[idx["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
                 issue_active_pct=num(row[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                 regs=num(row[idx["launch__registers_per_thread"]]), grid=num(row[idx["launch__grid_size"]]))
        k["dram_GBps"] = (k["dram_read_bytes"] + k["dram_write_bytes"]) / 1e9 / (k["dur_us"] / 1e6)
        out["kernels"].append(k)
        agg[name].append(k)
    for name, ks in agg.items():
        tot = sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in ks)
        out["traffic"][name] = {"launches_captured": len(ks), "dram_bytes_per_launch_mean": tot / len(ks),
                                "dram_bytes_sum": tot, "dur_us_sum": sum(k["dur_us"] for k in ks)}
    ad = [k for n, ks in agg.items() if "adam_sweep" in n for k in ks]
    out["traffic"]["adam_sweep (4 tables, one step)"] = {
        "dram_bytes": sum(k["dram_read_bytes"] + k["dram_write_bytes"] for k in ad), "dur_us": sum(k["dur_us"] for k in ad)}
    for src in list(out["traffic"]):
        if "gather_hist" in src or "scatter_hist" in src:
            out["traffic"][src]["dram_bytes"] = out["traffic"][src]["dram_bytes_per_launch_mean"]
    hdr, units, data = export("%s/%s_tc.ncu-rep" % (d, r))
    idx = {h: i for i, h in enumerate(hdr)}
    U = {h: units[i] for i, h in enumerate(hdr)}
    for row in data:
        name = re.sub(r"\(.*", "", row[idx["Kernel Name"]]).replace("void ", "")
        bps = num(row[idx["dram__bytes.sum.per_second"]]) * SC[U["dram__bytes.sum.per_second"].lower().split("/")[0]]
        k = dict(name=name, set="sections", dur_us=num(row[idx["gpu__time_duration.sum"]]), dram_GBps=bps / 1e9,
                 dram_pct=num(row[idx["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                 warps_active_pct=num(row[idx["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
                 issue_active_pct=num(row[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                 regs=num(row[idx["launch__registers_per_thread"]]), grid=num(row[idx["launch__grid_size"]]),
                 smem_dyn_KB=num(row[idx["launch__shared_mem_per_block_dynamic"]]),
                 lsu_wavefront_pct=num(row[idx["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]]),
                 tensor_pipe_pct=(num(row[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
                                  if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in idx else None))
        k["dram_bytes"] = bps * k["dur_us"] / 1e6
        out["kernels"].append(k)
    tc = [k for k in out["kernels"] if k["set"] == "sections"]
    for nm in sorted({k["name"] for k in tc}):
        ks = [k for k in tc if k["name"] == nm]
        out["traffic"][nm] = {"launches_captured": len(ks), "dram_bytes_sum": sum(k["dram_bytes"] for k in ks),
                              "dur_us_sum": sum(k["dur_us"] for k in ks)}
    with open("profiles/%s_ncu_top_kernels.json" % r, "w") as f:
        json.dump(out, f, indent=1)
    for k in sorted(tc, key=lambda k: -k["dur_us"])[:10]:
        print(k["name"], k["dur_us"], round(k["dram_GBps"]), round(k["dram_pct"], 1), round(k["issue_active_pct"], 1),
              round(k["lsu_wavefront_pct"], 1), k["smem_dyn_KB"])
    for n, t in out["traffic"].items():
        print(n, {a: (round(b) if isinstance(b, float) else b) for a, b in t.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out", sys.argv[2] if len(sys.argv) > 2 else "r01")
