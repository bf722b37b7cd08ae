"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` (file or stdin)."""
import csv
import re
import sys

pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|tensor.*pct|lts__t_bytes.sum$|"
                 r"lts__throughput.avg.pct|l1tex__throughput|sm__throughput.avg.pct|registers_per_thread|grid_size|"
                 r"sm__warps_active|lts__t_sector_hit_rate|smsp__inst_executed.sum$|dram__cycles_active|sm__cycles_elapsed.max")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("----", r[hdr.index("Kernel Name")][:60], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if pat.search(h):
            print(f"  {h:95s} {r[i]:>18s} {units[i]}")
