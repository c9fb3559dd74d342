#!/bin/bash
# usage: tools/ncu_key_metrics.sh file.ncu-rep  -> the metrics the profiling recipe names
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c '
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]; vals=rows[2] if len(rows)>2 else []
want=["Kernel Name","gpu__time_duration.sum","launch__grid_size","launch__cluster","launch__registers_per_thread","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","dram__throughput.avg.pct_of_peak_sustained_elapsed","sm__pipe_tensor_cycles_active","sm__pipe_tensor_subpipe","sm__mem_tensor","sm__warps_active.avg.pct_of_peak_sustained_active","sm__throughput.avg.pct","lts__t_bytes.sum","lts__t_sector_hit_rate.pct","lts__throughput.avg.pct","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__issue_active.avg.pct","sm__inst_executed_pipe_uniform","smsp__warp_issue_stalled","lts__t_sectors_srcunit_tex_op_read.sum"]
for i,h in enumerate(hdr):
  if any(h.startswith(w) or w in h for w in want):
    v=vals[i] if i<len(vals) else ""
    if "smsp__warp_issue_stalled" in h and "per_warp_active.pct" not in h: continue
    print("%-86s %-14s %s"%(h[:86],units[i][:14],v[:60]))
'
