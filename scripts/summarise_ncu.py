"""`ncu -i X.ncu-rep --page raw --csv` -> a compact (metric, unit, value) table of the metrics DESIGN.md / bench.py quote.
usage: python scripts/summarise_ncu.py raw.csv out.csv"""
import csv
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum", "sm__cycles_elapsed.avg"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "metric", "unit", "value"])
    for li, vals in enumerate(rows[2:]):
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                w.writerow([li, h, u, v])
print("wrote", sys.argv[2])
