#!/usr/bin/env python
"""Host AMG setup time of the n^3 7-point problem with FASP's own fasp_amg_setup_rs: plain CPU, and with
libfasp_cuda_setup.so interposed (transpose + Galerkin product of every level on the device). Prints the time
and a checksum of the hierarchy (must be identical).  Run twice:
    python scripts/setup_timing.py 256
    LD_PRELOAD=faspsolver_b200/lib/libfasp_cuda_setup.so python scripts/setup_timing.py 256
"""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
from faspsolver_b200 import api, problems as PB, fasp_types as T

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
hf = api.HostFasp(str(ROOT / "oracle" / "_ref" / "libfasp_seq.so"))
A = PB.poisson7(n)
amg = hf.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
if os.environ.get("LD_PRELOAD"):
    api.lib().fasp_cuda_init(0)       # context creation outside the timed region
t = time.time()
mgl = hf.amg_setup(A, amg)
dt = time.time() - t
nl = mgl[0].num_levels
chk = []
for l in range(nl):
    m = mgl[l].A
    ja = np.ctypeslib.as_array(m.JA, shape=(m.nnz,))
    va = np.ctypeslib.as_array(m.val, shape=(m.nnz,))
    w = np.arange(1, 1 + min(m.nnz, 1 << 22), dtype=np.float64)
    chk.append([int(m.row), int(m.nnz), float(np.dot(ja[:w.size].astype(np.float64), w)), float(np.dot(va[:w.size], w))])
print(json.dumps({"n": n, "interposed": bool(os.environ.get("LD_PRELOAD")), "setup_s": round(dt, 2), "levels": nl,
                  "checksum": chk}))
