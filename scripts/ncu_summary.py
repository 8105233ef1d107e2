"""Condenses `ncu -i <rep> --page raw --csv` of the loop capture into the per-launch summary kept under profiles/.
usage: python scripts/ncu_summary.py gpurun_out/prof_r1c_raw.csv profiles/r1_ncu_full_loop_summary.csv"""
import csv, sys
COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = [hdr.index(c) for c in COLS if c in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        out = [r[i] for i in idx]
        out[0] = out[0].split("(")[0].replace("hq::", "")
        w.writerow(out)
print(f"{len(rows) - 2} launches -> {sys.argv[2]}")
