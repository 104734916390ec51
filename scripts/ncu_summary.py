"""Selected counters of one kernel from `ncu -i <rep> --page raw --csv` as a small metric,unit,value table.
usage: ncu -i X.ncu-rep --page raw --csv | python scripts/ncu_summary.py [kernel name substring] > profiles/....csv"""
import csv, sys
WANT = ["Kernel Name", "dram__bytes.sum.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
sub = sys.argv[1] if len(sys.argv) > 1 else ""
kn = hdr.index("Kernel Name")
row = next(r for r in rows[2:] if sub in r[kn])
out = csv.writer(sys.stdout)
out.writerow(["metric", "unit", "value"])
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        out.writerow([w, units[i], row[i]])
