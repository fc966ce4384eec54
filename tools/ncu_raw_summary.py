#!/usr/bin/env python3
"""Condense `ncu -i X.ncu-rep --page raw --csv` (one or more launches) into a JSON summary per launch.
Usage: ncu_raw_summary.py raw.csv out.json [note]"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_op_read_hit_rate.pct",
        "lts__t_bytes.sum", "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_op_shared_ld.sum", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active"]
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k); d[k] = {"value": r[i], "unit": units[i]}
    d["stall_warps_per_issue"] = {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: round(float(v), 3)
                                  for h, v in zip(hdr, r) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v}
    out.append(d)
json.dump({"capture": sys.argv[3] if len(sys.argv) > 3 else "", "launches": out}, open(sys.argv[2], "w"), indent=1)
for d in out:
    print(d["kernel"][:80], {k: d[k]["value"] for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") if k in d})
