"""Turn `ncu -i X.ncu-rep --page raw --csv` into the small tracked summary under profiles/:
selected metrics per captured kernel, plus (optionally) the DRAM bytes per launch of the fused
kernel as JSON for bench.py's roofline.traffic.
Usage: python tools/ncu_summary.py raw.csv out.csv [dram.json "source note"]"""
import csv
import json
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1]
idx = [hdr.index(k) for k in KEEP if k in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
if len(sys.argv) > 3:
    def col(name, r):
        v = float(r[hdr.index(name)].replace(",", ""))
        u = units[hdr.index(name)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    fq = [r for r in rows[2:] if "fused_query" in r[hdr.index("Kernel Name")]]
    r = fq[-1]
    rd, wr = col("dram__bytes_read.sum", r), col("dram__bytes_write.sum", r)
    json.dump({"dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
               "source": sys.argv[4] if len(sys.argv) > 4 else "", "algorithmic_bytes_per_launch": 187500000},
              open(sys.argv[3], "w"), indent=1)
