"""profiles/fill_traffic.json from an ncu --set full report of the dominant fill kernel: DRAM bytes
(read + written) of the captured launch divided by the cells it processed.

    python tools/ncu_traffic.py gpurun_out/<capture>.ncu-rep <cells of the launch> profiles/<summary>.txt
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

rep, cells, summary = sys.argv[1], float(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, row = rows[0], rows[1], rows[2]


def value(name):
    k = hdr.index(name)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[k]]
    return float(row[k].replace(",", "")) * scale


total = value("dram__bytes_read.sum") + value("dram__bytes_write.sum")
out = {"kernel": row[hdr.index("Kernel Name")], "source": summary, "cells_of_the_launch": cells,
       "dram_bytes": total, "dram_bytes_per_cell": total / cells,
       "note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch / its cells"}
Path(__file__).resolve().parent.parent.joinpath("profiles", "fill_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out))
