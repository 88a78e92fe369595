#!/usr/bin/env python
"""Condense ncu output into the small text summaries committed under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches_X.csv          per-kernel share of the timed region
  python scripts/ncu_summary.py kernel gpurun_out/prof.ncu-rep [n_units]    headline metrics + stall reasons per launch
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_shared_atom.sum",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "")
        e = d.setdefault(name, [0, 0.0])
        e[0] += 1
        e[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in d.values())
    print("%-52s %7s %14s %8s %12s" % ("kernel", "launches", "total us", "share", "avg us"))
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print("%-52s %7d %14.1f %7.1f%% %12.1f" % (k[:52], v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / 1e3 / v[0]))
    print("%-52s %7d %14.1f" % ("total", sum(v[0] for v in d.values()), tot / 1e3))


def kernel(path, units=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0]
        print("== %s  (launch id %s)" % (name, r[h.index("ID")]))
        vals = {k: (r[i], u[i]) for i, k in enumerate(h)}
        for k in KEYS:
            if k in vals and vals[k][0] != "":
                print("  %-70s %16s %s" % (k, vals[k][0], vals[k][1]))
        try:
            rd = float(vals["dram__bytes_read.sum"][0].replace(",", ""))
            wr = float(vals["dram__bytes_write.sum"][0].replace(",", ""))
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tr = rd * scale[vals["dram__bytes_read.sum"][1]] + wr * scale[vals["dram__bytes_write.sum"][1]]
            print("  %-70s %16.4e byte" % ("dram traffic (read+write)", tr))
            if units:
                print("  %-70s %16.2f byte/unit (units = %d)" % ("dram traffic per unit", tr / units, units))
        except Exception as ex:  # metric missing in this capture
            print("  traffic: n/a (%r)" % (ex,))
        stalls = []
        for k, (v, _) in vals.items():
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(v.replace(",", "")), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            print("  top warp stall reasons (warps stalled per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:6]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], int(float(sys.argv[3])) if len(sys.argv) > 3 else None)
