"""Print the key counters of an `ncu --page raw --csv` export, one column per captured kernel. Usage: ncu_raw.py file.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
        "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "sm__sass_inst_executed_op_global_ld.sum", "smsp__cycles_active.avg"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-75s %-12s %s" % (w, units[i], " | ".join(r[i][:60] for r in data)))
