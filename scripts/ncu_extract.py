"""ncu -i <rep> --page raw --csv | python scripts/ncu_extract.py out.csv [note]  -> the key metrics of every captured launch as a small CSV."""
import csv
import sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = ["ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg"]
cols = [c for c in cols if c in idx]
note = sys.argv[2] if len(sys.argv) > 2 else ""
out = [["note"] + cols, [""] + [units[idx[c]] for c in cols]]
for r in rows[2:]:
    out.append([note] + [r[idx[c]] for c in cols])
csv.writer(open(sys.argv[1], "w", newline="")).writerows(out)
for o in out[2:]:
    d = dict(zip(out[0], o))
    print(d["Kernel Name"][:60].ljust(62), d.get("gpu__time_duration.sum"), "us  tensor%", d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
          " dram%", d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), " regs", d.get("launch__registers_per_thread"))
