#!/usr/bin/env python
"""profiles/rNN_ncu_traffic.json from an `ncu --set full` capture of the level-0 kernels (y = Ax, r = b - Ax, L1 sweep of
the 7-pt 256^3 matrix): DRAM bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum. bench.py reads the newest
file for `roofline.traffic` and compares the recorded hash of spmv.cu with the current one.
    python scripts/make_traffic_json.py gpurun_out/r02z/level0_kernels.ncu-rep profiles/r02_ncu_traffic.json"""
import csv, hashlib, io, json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
def to_us(v, unit):
    f = float(v.replace(",", ""))
    return f * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
mode_of = {"<0,": "mxv", "<2,": "resid", "<4,": "l1"}
kern = {}
for r in data:   # the last capture of each mode wins (warm caches do not matter: the matrix is 10x the L2)
    name = r[col["Kernel Name"]]
    if "csr_pipe_kernel" not in name:
        continue
    m = next((v for k, v in mode_of.items() if ("csr_pipe_kernel" + k) in name.replace("(int)", "").replace(" ", "")), None)
    if m is None:
        continue
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    kern[m] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr,
               "duration_us_under_ncu": to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]]),
               "dram_pct_of_peak": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]])}
sha = hashlib.sha256((ROOT / "faspsolver_b200" / "csrc" / "spmv.cu").read_bytes()).hexdigest()[:16]
json.dump({"source": "ncu --set full --clock-control none -k regex:csr_pipe_kernel (python scripts/level_sweep.py --n 256 --levels 0 "
                     "--ops A --kernels 0,2,11), " + Path(rep).name,
           "matrix": "7-pt Poisson 256^3 level-0 A: 16777216 rows, 117047296 nnz", "spmv_cu_sha16": sha, "kernels": kern},
          open(out, "w"), indent=1)
print(json.dumps(kern, indent=1))
