#!/usr/bin/env python3
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the JSON summaries kept under profiles/.
Usage: ncu_summary.py raw.csv kernel_name utterances_in_capture utterances_per_bench_launch out_prefix"""
import csv, json, sys
raw, kernel, n_cap, n_bench, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_op_read_hit_rate.pct"]
d = {}
stalls = {}
for h, u, v in zip(hdr, units, vals):
    if h in keep: d[h] = {"value": v, "unit": u}
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v)
d["stall_warps_per_issue"] = stalls
d["kernel"] = kernel
json.dump(d, open(out + "_full.json", "w"), indent=1)
def num(k):
    x = d[k]; f = float(x["value"]); u = x["unit"].lower()
    return f * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}[u]
rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
json.dump({"kernel": kernel, "capture": f"ncu --set full, {n_cap}-utterance launch (bench.py --utts {n_cap})",
           "dram_bytes_read": rd, "dram_bytes_write": wr, "utterances_in_capture": n_cap,
           "dram_bytes_per_launch": (rd + wr) * n_bench / n_cap,
           "note": f"scaled x{n_bench / n_cap:g} to the {n_bench}-utterance bench launch (per-utterance traffic is constant)"},
          open(out + "_traffic.json", "w"), indent=1)
print(json.dumps({k: d[k]["value"] for k in keep if k in d}, indent=1))
