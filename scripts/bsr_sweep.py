#!/usr/bin/env python
"""BSR kernel-shape sweep on the config-5 matrix (3x3-block 7-point): block rows per CTA, blocks in flight per
thread, pipeline stages. Times the resident kernels with CUDA events (30 back-to-back launches).

    python scripts/bsr_sweep.py [--n 160]
"""
import argparse, json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from faspsolver_b200 import api
import bench as B
sys.path.insert(0, str(ROOT / "scripts"))
from bench_configs import blockoil7_lean


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=160)
    ap.add_argument("--quick", action="store_true", help="only the shapes worth comparing at large sizes")
    a = ap.parse_args()
    L = api.lib(); api.check(L.fasp_cuda_init(0))
    peak, _ = B.peaks()
    A, _b = blockoil7_lean(a.n)
    by = (8.0 * 9 + 4) * A.NNZ + 4.0 * (A.ROW + 1) + 8.0 * 3 * (A.COL + A.ROW)
    print("# 3x3-block 7-point %d^3: %d block rows, %d blocks, %.2f GB per pass; peak %.0f GB/s" % (a.n, A.ROW, A.NNZ, by / 1e9, peak))
    combos = [(32, 4, 2), (64, 4, 2), (64, 8, 2)] if a.quick else [(rb, u, st) for rb in (32, 64) for u in (4, 8) for st in (2, 3)]
    if True:
        if True:
            for rb, u, st in combos:
                for k, v in (("bsr_rb", rb), ("bsr_u", u), ("bsr_stages", st)):
                    L.fasp_cuda_set_option(k.encode(), float(v))
                h = L.fasp_cuda_dbsr_upload(A.ptr())
                if not h:
                    print("rb %d u %d stages %d: %s" % (rb, u, st, api.last_error())); continue
                row = {"rb": rb, "u": u, "stages": st}
                for what, nm, extra in ((0, "mxv", 0.0), (2, "resid", 24.0 * A.ROW), (10, "jacobi", 96.0 * A.ROW)):
                    ms = L.fasp_cuda_dbsr_time_kernel(h, what, 5, 30, 0)
                    row[nm] = "%.4f ms %.0f GB/s %.3f" % (ms, (by + extra) / ms * 1e-6, (by + extra) / ms * 1e-6 / peak) if ms > 0 else api.last_error()
                L.fasp_cuda_dbsr_free(h)
                print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
