"""`ncu -i X.ncu-rep --page raw --csv` -> the columns DESIGN.md discusses (one row per profiled launch).

    python scripts/reduce_ncu_raw.py gpurun_out/r02_kernels_raw.csv profiles/r02_kernels_ncu_full.csv
"""
import csv
import sys

KEEP = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
idx = [h.index(k) for k in KEEP if k in h]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch"] + [h[i] for i in idx])
    w.writerow([""] + [rows[hdr + 1][i] for i in idx])
    for n, r in enumerate(rows[hdr + 2:]):
        if len(r) > max(idx):
            w.writerow([n] + [r[i] for i in idx])
print("wrote", sys.argv[2])
