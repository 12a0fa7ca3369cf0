#!/usr/bin/env python
"""Prints the key metrics of every kernel in an `ncu --page raw --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("----")
    for k in keys:
        if k in d:
            print(k, d[k])
    st = []
    for k in hdr:
        if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
            try:
                v = float(d[k].replace(",", ""))
            except ValueError:
                continue
            if v > 0:
                st.append((v, k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
    tot = sum(v for v, _ in st) or 1
    print("  stalls:", ", ".join(f"{n} {100*v/tot:.1f}%" for v, n in sorted(st, reverse=True)[:9]))
