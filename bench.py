#!/usr/bin/env python
"""bench.py — AMG-PCG solve time on the 3-D 7-point Poisson 256^3 system (BASELINE.json
configs[1]) on B200, with the SpMV/smoother kernel roofline and the reference's CPU path beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 256]

A "step" is ONE full AMG-PCG solve (x0 = 0, relative residual 1e-8, classical RS hierarchy,
V(1,1) L1-Jacobi) on a hierarchy that is already resident in HBM; `value` is the mean solve
time in ms over K steps measured with CUDA events (max over ranks), `e2e` the same solve through
the host-pointer C-ABI call fasp_cuda_krylov_amg_solve (b and x0 copied H2D, x copied D2H inside
the timed region). The hierarchy is built once, before timing, by FASP's own host setup
(north star: "built by FASP's own classical/SA setup on the host and uploaded once").

Multi-GPU (N > 1, launched by torchrun, one rank per GPU): see DESIGN.md §multi-GPU.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "amg_pcg_solve_time_poisson3d_7pt"
UNIT = "ms"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.lines = []
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------
# problem + hierarchy
# ---------------------------------------------------------------------------------------
def host_fasp():
    """FASP's host library for the SETUP phase (hierarchy construction, not timed)."""
    from faspsolver_b200.api import HostFasp
    path = os.environ.get("FASP_CUDA_HOST_LIBFASP") or str(ROOT / "oracle" / "_ref" / "libfasp_seq.so")
    return HostFasp(path)


def build_problem(n):
    from faspsolver_b200 import problems as PB
    t = time.time()
    A = PB.poisson7(n)
    b = np.ones(A.shape[0])
    log("[bench] 7-pt Poisson %d^3: %d rows, %d nnz (generated in %.1fs)" % (n, A.shape[0], A.nnz, time.time() - t))
    return A, b


def amg_recipe(hf):
    from faspsolver_b200 import fasp_types as T
    amg = hf.amg_param(print_level=0, AMG_type=T.CLASSIC_AMG, coarsening_type=T.COARSE_RS,
                       interpolation_type=T.INTERP_DIR, smoother=T.SMOOTHER_L1DIAG,
                       cycle_type=T.V_CYCLE, presmooth_iter=1, postsmooth_iter=1)
    it = hf.its_param(itsolver_type=T.SOLVER_CG, tol=1e-8, maxit=500, print_level=0,
                      stop_type=T.STOP_REL_RES)
    return amg, it


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def run_ours(args):
    rank, world, local = dist_env()
    if args.gpus > 1 or world > 1:
        from faspsolver_b200 import multigpu
        return multigpu.bench_main(args)

    from faspsolver_b200 import api
    from faspsolver_b200 import fasp_types as T
    L = api.lib()
    api.check(L.fasp_cuda_init(local))
    for kv in args.opt:
        k, v = kv.split("=")
        api.check(L.fasp_cuda_set_option(k.encode(), float(v)))

    hf = host_fasp()
    A, b = build_problem(args.n)
    n = A.shape[0]
    amg, it = amg_recipe(hf)
    t = time.time()
    mgl = hf.amg_setup(A, amg)
    t_setup = time.time() - t
    info = api.hierarchy_info(mgl)
    log("[bench] host AMG setup (FASP, sequential): %.1fs, %d levels, rows %s" %
        (t_setup, len(info), [r for r, _, _ in info]))
    t = time.time()
    solver = api.KrylovAmgSolver(mgl, amg)
    t_upload = time.time() - t
    log("[bench] upload + coarse factor: %.2fs" % t_upload)

    d_b = L.fasp_cuda_dvec_alloc(n)
    d_x = L.fasp_cuda_dvec_alloc(n)
    api.check(L.fasp_cuda_dvec_h2d(d_b, T.as_preal(b), n))
    zero = np.zeros(n)

    def dev_solve():
        api.check(L.fasp_cuda_dvec_h2d(d_x, T.as_preal(zero), n))   # x0 = 0 (outside the events)
        st = solver.solve_dev(d_b, d_x, it)
        if st < 0:
            raise RuntimeError("device solve failed: %d %s" % (st, api.last_error()))
        return st, solver.stat(2)

    for _ in range(args.warmup):
        iters, _ms = dev_solve()
    sampler = ClockSampler(local)
    sampler.start()
    L.fasp_cuda_launch_count_reset()
    times = []
    for _ in range(args.steps):
        iters, ms = dev_solve()
        times.append(ms)
    launches = int(L.fasp_cuda_launch_count())
    clocks = sampler.stop()
    ms_step = float(np.mean(times))
    relres = solver.stat(1)
    log("[bench] device-resident solve: %d iterations, relres %.3e, %.3f ms (min %.3f max %.3f)" %
        (iters, relres, ms_step, min(times), max(times)))

    # ---- e2e: host-pointer call, H2D(b, x0) + solve + D2H(x) inside the timed region
    e2e_times = []
    x_host = None
    x_buf = np.zeros(n)
    api.pin_host(b)        # the application's own arrays, page-locked once (pinned host memory)
    api.pin_host(x_buf)
    for k in range(max(1, args.warmup // 2) + args.steps):
        st, x_host = solver.solve(b, zero, it, out=x_buf)
        if st < 0:
            raise RuntimeError("host solve failed: %d %s" % (st, api.last_error()))
        if k >= max(1, args.warmup // 2):
            e2e_times.append(solver.stat(4))
    api.unpin_host(b)
    api.unpin_host(x_buf)
    e2e_ms = float(np.mean(e2e_times))
    # true residual of the returned solution (size-independent correctness check)
    r = b - A.to_scipy() @ x_host
    true_rel = float(np.linalg.norm(r) / np.linalg.norm(b))
    log("[bench] e2e solve %.3f ms, true relres of returned x %.3e" % (e2e_ms, true_rel))
    if not (true_rel <= 1e-8 * 1.001):
        raise RuntimeError("solution does not meet the tolerance: %g" % true_rel)

    # ---- roofline of the dominant kernel: CSR row kernel on the level-0 matrix, timed live
    # with CUDA events around every launch during one more real solve (graphs off)
    L.fasp_cuda_set_option(b"profile", 1.0)
    L.fasp_cuda_profile_dump(None, 0)   # clear
    dev_solve()
    buf = C.create_string_buffer(64 << 20)
    L.fasp_cuda_profile_dump(buf, len(buf))
    L.fasp_cuda_set_option(b"profile", 0.0)
    recs = [ln.split() for ln in buf.value.decode().splitlines()]
    recs = [(int(k), int(r), int(z), float(ms), float(by)) for k, r, z, ms, by in recs]
    # drop launches that returned at once: branch-gated kernels (kind >= 50 = conditional tag) and
    # the look-ahead iterations enqueued after convergence (device `done` flag set)
    groups = {}
    for rec in recs:
        groups.setdefault(rec[:3], []).append(rec[3])
    med = {k: float(np.percentile(v, 90)) for k, v in groups.items()}   # gated launches can be the majority
    recs = [rec for rec in recs if rec[0] < 50 and rec[3] >= 0.25 * med[rec[:3]]]
    peak, peak_src = peaks()
    lvl0 = [x for x in recs if x[1] == n and x[2] == A.nnz]
    tot_ms = sum(x[3] for x in recs)
    l0_ms = sum(x[3] for x in lvl0)
    by_kind = {}
    for k, r_, z, ms, by in lvl0:
        by_kind.setdefault(k, []).append((ms, by))
    kind_names = {0: "mxv", 1: "aAxpy", 2: "resid", 3: "jacobi", 4: "l1", 7: "resid_dinv"}
    per_kind = {kind_names.get(k, str(k)): {"launches": len(v), "ms": float(np.mean([m for m, _ in v])),
                                            "GBps": float(np.mean([b_ / m * 1e-6 for m, b_ in v]))}
                for k, v in by_kind.items()}
    ach = float(sum(by for *_, by in lvl0) / l0_ms * 1e-6) if l0_ms > 0 else 0.0
    # DRAM traffic per launch of the same kernel from the committed ncu --set full capture
    traffic, traffic_src = None, None
    tfs = sorted((ROOT / "profiles").glob("r*_ncu_traffic.json"))
    if tfs and args.n == 256:
        tj = json.loads(tfs[-1].read_text())
        ks = tj["kernels"]
        per = {"resid": ks.get("resid"), "l1": ks.get("l1"), "mxv": ks.get("mxv") or ks.get("resid")}
        num = sum(per[kind_names[k]]["traffic"] * len(v) for k, v in by_kind.items() if per.get(kind_names.get(k)))
        den = sum(len(v) for k, v in by_kind.items() if per.get(kind_names.get(k)))
        traffic = num / den if den else None
        traffic_src = tfs[-1].name + ": " + tj["source"]
        # the capture belongs to one version of the kernel source: say so when the kernel has changed since
        import hashlib
        sha = hashlib.sha256((ROOT / "faspsolver_b200" / "csrc" / "spmv.cu").read_bytes()).hexdigest()[:16]
        if tj.get("spmv_cu_sha16") and tj["spmv_cu_sha16"] != sha:
            traffic_src += " [captured before the last change of spmv.cu]"
    alg_per_launch = float(np.mean([by for *_, by in lvl0])) if lvl0 else None
    roofline = {"bound": "hbm", "kernel": "csr_pipe_kernel on the level-0 matrix (SpMV / residual / L1-Jacobi sweep), "
                                          "CUDA events around every launch of one solve",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "frac_of_nominal_8000": ach / 8000.0, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg_per_launch,
                "launch_ms": l0_ms / len(lvl0) if lvl0 else None,
                "peak_source": peak_src, "share_of_matrix_kernel_time": l0_ms / tot_ms if tot_ms else None,
                "per_mode": per_kind}
    # per-level table for profiles/
    levels = {}
    for k, r_, z, ms, by in recs:
        e = levels.setdefault((r_, z), [0, 0.0, 0.0])
        e[0] += 1
        e[1] += ms
        e[2] += by
    level_table = [{"rows": r_, "nnz": z, "launches": c, "ms": ms, "GBps": by / ms * 1e-6 if ms else 0}
                   for (r_, z), (c, ms, by) in sorted(levels.items(), key=lambda kv: -kv[0][1])]

    # ---- CPU baseline: the reference's own PCG + V-cycle on the same hierarchy, bounded sample
    cpu = None if args.no_cpu_baseline else cpu_baseline_sample(hf, A, b, mgl, amg, it, iters, args)

    out = {
        "metric": METRIC, "value": ms_step, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: 3D 7-point Poisson %d^3 (%d rows, %d nnz), rhs=1, AMG-PCG tol 1e-8, "
                               "classical RS + direct interpolation (FASP host setup), V(1,1) L1-Jacobi" %
                               (args.n, n, A.nnz),
                   "levels": len(info), "iterations": int(iters), "final_relres": relres,
                   "true_relres": true_rel, "hierarchy_bytes": int(solver.stat(5)),
                   "l2_policy": "inputs larger than L2 (hierarchy >> 126 MB); no flush needed",
                   "setup_s_host": t_setup, "upload_s": t_upload},
        "e2e": {"value": e2e_ms, "unit": UNIT, "h2d_bytes_per_step": int(16 * n),
                "d2h_bytes_per_step": int(8 * n)},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "levels": level_table,
    }
    solver.close()
    hf.amg_free(mgl, amg)
    if args.extras and args.n == 256:
        out["extra"] = run_extras(args, local)
    return out


def run_extras(args, local):
    """BASELINE configs[3] (conv-diff 256^3 GMRES(30) + polynomial smoother) and configs[4] (BSR 272^3) in two
    separate processes after the headline measurement is complete: their host setups (single-threaded FASP
    code, ~100 s and ~20 s) overlap, their GPU phases are serialised by a lock file. A config that fails or runs
    out of its time budget is reported as unavailable; the headline line is never at risk."""
    import tempfile
    script = str(ROOT / "scripts" / "bench_configs.py")
    lock = tempfile.NamedTemporaryFile(prefix="fasp_bench_gpu_", suffix=".lock", delete=False).name
    env = dict(os.environ, LOCAL_RANK=str(local))
    jobs = {}
    for key, cfg, n, budget in (("config4", 4, args.c4_n, 420), ("config5", 5, args.c5_n, 420),
                                ("config3_proxy", 3, args.c3_n, 480)):
        if cfg == 3:   # configs[2] at the size one dCSRmat (and the sequential oracle) can hold: 27-pt 256^3
            if n <= 0:
                continue
            cmd = [sys.executable, str(ROOT / "scripts" / "bench_config3.py"), "--size", str(n), "--steps", "5",
                   "--warmup", "3", "--lock", lock]
        else:
            cmd = [sys.executable, script, "--config", str(cfg), "--n", str(n), "--lock", lock]
        for kv in args.opt:
            cmd += ["--opt", kv]
        jobs[key] = (subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True),
                     time.time(), budget)
    res = {}
    for key, (proc, t0, budget) in jobs.items():
        try:
            so, se = proc.communicate(timeout=max(1.0, budget - (time.time() - t0)))
        except subprocess.TimeoutExpired:
            proc.kill()
            so, se = proc.communicate()
            res[key] = {"unavailable": "did not finish within %d s" % budget}
            continue
        sys.stderr.write(se[-3000:])
        line = [ln for ln in so.splitlines() if ln.startswith("{")]
        if proc.returncode != 0 or not line:
            res[key] = {"unavailable": "rc=%d: %s" % (proc.returncode, se.strip().splitlines()[-1][:300] if se.strip() else "")}
        else:
            res[key] = json.loads(line[-1])
            res[key]["wall_s"] = round(time.time() - t0, 1)
    try:
        os.unlink(lock)
    except OSError:
        pass
    return res


def cpu_baseline_sample(hf, A, b, mgl, amg, it, full_iters, args):
    """Times `sample_it` PCG iterations of the SAME solve with the reference's own code
    (sequential libfasp: fasp_solver_dcsr_pcg + fasp_precond_amg on the same hierarchy) and
    scales to the full iteration count. 1 core; reported baseline, not the target."""
    from oracle.ref import RefFasp
    from faspsolver_b200 import fasp_types as T
    ref = RefFasp()
    n = A.shape[0]
    sample_it = max(1, min(int(full_iters), args.cpu_sample_iters))
    # precond_data for fasp_precond_amg (fasp.h:894-981): filled by the reference itself
    pcdata = T.precond_data()
    ref.L.fasp_param_amg_to_prec(C.byref(pcdata), C.byref(amg))
    pcdata.max_levels = mgl[0].num_levels   # SolCSR.c:531-533
    pcdata.mgl_data = mgl
    pc = T.precond(C.cast(C.byref(pcdata), C.c_void_p), C.cast(ref.L.fasp_precond_amg, T.PRECOND_FCT))
    vb, vx = T.Vec(b), T.Vec(np.zeros(n))
    t = time.perf_counter()
    st = ref.L.fasp_solver_dcsr_pcg(A.ptr(), vb.ptr(), vx.ptr(), C.byref(pc), it.tol, it.abstol, sample_it,
                                    it.stop_type, 0)
    dt = time.perf_counter() - t
    per_it = dt / sample_it
    # a k-iteration solve applies the preconditioner and A k+1 times (KryPcg.c:125-131)
    est = dt * (full_iters + 1) / (sample_it + 1)
    log("[bench] CPU reference sample: %d PCG iterations in %.2fs (%.2f s/it) -> %.1f s per solve (1 core)" %
        (sample_it, dt, per_it, est))
    return {"value": est * 1e3, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": "%d of %d PCG iterations (fasp_solver_dcsr_pcg + fasp_precond_amg, sequential "
                      "libfasp, same hierarchy) timed = %.2f s, scaled by (%d+1)/(%d+1); status %d" %
                      (sample_it, full_iters, dt, full_iters, sample_it, st),
            "host_cores_available": os.cpu_count()}


# ---------------------------------------------------------------------------------------
# reference arm: the reference's OpenMP CPU path on the same config
# ---------------------------------------------------------------------------------------
def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return None
    exe = ROOT / "oracle" / "_ref" / "fasp_ref_bench"
    if not exe.exists():
        return {"impl": "reference", "unavailable": "oracle/_ref/fasp_ref_bench not built (oracle/build_ref.sh)"}
    env = dict(os.environ)
    ncores = os.cpu_count() or 1
    try:
        ncores = len(os.sched_getaffinity(0)) or ncores
    except Exception:
        pass
    # assignment, not setdefault: torchrun exports OMP_NUM_THREADS=1 to its workers, which would time the
    # reference on one thread. Rank 0 runs alone (the other ranks exit), so it takes every host core.
    env["OMP_NUM_THREADS"] = str(int(os.environ.get("FASP_REF_THREADS", ncores)))
    env["OMP_PROC_BIND"] = "true"
    env["OMP_PLACES"] = "cores"
    cmd = [str(exe), str(args.n), str(args.steps), str(args.warmup), str(args.ref_sample_iters)]
    log("[bench] reference arm:", " ".join(cmd), "OMP_NUM_THREADS=%s" % env["OMP_NUM_THREADS"])
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    sys.stderr.write(r.stderr[-4000:])
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not line:
        return {"impl": "reference", "unavailable": "fasp_ref_bench failed rc=%d" % r.returncode}
    d = json.loads(line[-1])
    n = args.n ** 3
    val = d["ms_per_solve"]
    return {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": val, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: 3D 7-point Poisson %d^3 (%d rows), rhs=1, AMG-PCG tol 1e-8, classical RS, "
                               "OpenMP FASP (runs multicolour GS regardless of the requested smoother)" % (args.n, n),
                   "levels": d.get("levels"), "iterations": d.get("iterations"), "setup_s": d.get("setup_s"),
                   "extrapolated": d.get("extrapolated"), "ms_first_solve": d.get("ms_first_solve")},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": d.get("threads"), "kind": "reference",
                         "sample": d.get("sample")},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("FASP_BENCH_N", "256")))
    ap.add_argument("--cpu-sample-iters", type=int, default=3)
    ap.add_argument("--ref-sample-iters", type=int, default=0,
                    help="reference arm: 0 = every step is a full solve (default); k > 0 = k-iteration samples, scaled")
    ap.add_argument("--extras", type=int, default=int(os.environ.get("FASP_BENCH_EXTRAS", "1")),
                    help="N = 1: also measure BASELINE configs 4 and 5 (separate processes, bounded) -> extra.config4/5")
    ap.add_argument("--c4-n", type=int, default=256)
    ap.add_argument("--c5-n", type=int, default=272)
    ap.add_argument("--c3-n", type=int, default=int(os.environ.get("FASP_BENCH_C3N", "256")),
                    help="extra.config3_proxy: 27-point n^3 through the slab path at every N (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU sample (profiling runs)")
    ap.add_argument("--opt", action="append", default=[], help="libfasp_cuda option key=value")
    ap.add_argument("--agg-rows", type=int, default=8000,
                    help="multi-GPU: levels with fewer global rows are replicated instead of partitioned")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
