"""Multi-GPU host plumbing: one process per GPU (torchrun), torch.distributed only bootstraps the
NCCL communicator of libfasp_cuda (unique id broadcast) and reduces timings; the data path is
entirely inside the library (halo send/recv, all-reduce, all-gather over NVLink / NVSwitch).

Also holds the CPU-side (numpy) restatement of the slab extraction used by the gloo tests.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time

import numpy as np

from . import api
from . import fasp_types as T


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_comm(backend=None):
    """Initialise torch.distributed (env:// rendezvous) and the library's NCCL communicator."""
    import torch
    import torch.distributed as dist
    rank, world, local = dist_env()
    L = api.lib()
    api.check(L.fasp_cuda_init(local))
    if world == 1:
        return rank, world, local
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        import datetime
        # a rank that fails must not leave its peers waiting for the default 30 minutes
        dist.init_process_group(backend=backend or "gloo", rank=rank, world_size=world,
                                timeout=datetime.timedelta(minutes=10))
    ident = (C.c_ubyte * 128)()
    if rank == 0:
        api.check(L.fasp_cuda_comm_unique_id(ident))
    box = [bytes(ident)]
    dist.broadcast_object_list(box, src=0)
    buf = (C.c_ubyte * 128).from_buffer_copy(box[0])
    # NCCL may print a version banner on stdout; bench.py's stdout must stay one JSON line
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        api.check(L.fasp_cuda_comm_init(buf, rank, world))
        api.check(L.fasp_cuda_sync())
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    return rank, world, local


def allreduce_max(x: float) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        return x
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def allreduce_sum(x: float) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        return x
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


def barrier():
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()


class DistSolver(api.KrylovAmgSolver):
    """Row-partitioned solver: b / x are this rank's local slices (rows [row0, row1))."""

    def __init__(self, mgl, amgparam, agg_rows=8000):
        L = api.lib()
        self.h = L.fasp_cuda_dist_krylov_amg_create(mgl, C.byref(amgparam), int(agg_rows))
        if not self.h:
            raise api.FaspCudaError(-1, api.last_error())
        b, e = C.c_int(0), C.c_int(0)
        api.check(L.fasp_cuda_dist_row_range(self.h, C.byref(b), C.byref(e)))
        self.row0, self.row1 = b.value, e.value


class SlabSolver(api.KrylovAmgSolver):
    """The same solver built from a slabsetup.SlabHierarchy (no global matrix on any rank)."""

    def __init__(self, sh):
        L = api.lib()
        self.h = sh.create_solver()
        b, e = C.c_int(0), C.c_int(0)
        api.check(L.fasp_cuda_dist_row_range(self.h, C.byref(b), C.byref(e)))
        self.row0, self.row1 = b.value, e.value


def plane_partition(n_planes, world):
    """z-planes per rank, as even as possible; returns plane offsets (world + 1)."""
    return [(n_planes * r) // world for r in range(world + 1)]


# ---------------------------------------------------------------------------------------
# one host hierarchy per node
# ---------------------------------------------------------------------------------------
class SharedHierarchy:
    """FASP's host setup runs on rank 0 only; the CSR arrays of every level (A, P, R) are written once to a
    shared-memory directory and mapped read-only by the other ranks, which wrap them in an AMG_data array
    (fasp.h:804-888) for fasp_cuda_dist_krylov_amg_create. Falls back to one setup per rank when no shared
    directory has room. The AMG_param mutations of the setup (PreAMGSetupRS.c:83) are broadcast with it."""

    def __init__(self, hf, A, amg, rank, world, root=None):
        import shutil
        import tempfile
        import torch.distributed as dist
        self.hf, self.amg, self.rank, self.owner, self.dir = hf, amg, rank, False, None
        self._keep = []
        if world == 1:
            self.mgl, self.how, self.owner = hf.amg_setup(A, amg), "FASP sequential setup", True
            return
        box = [None]
        if rank == 0:
            need = 16.0 * A.nnz * 6.0   # generous bound on the bytes of all levels
            for cand in ([root] if root else []) + ["/dev/shm", tempfile.gettempdir()]:
                try:
                    if shutil.disk_usage(cand).free > 1.3 * need:
                        box[0] = tempfile.mkdtemp(prefix="fasp_hier_", dir=cand)
                        break
                except OSError:
                    continue
        dist.broadcast_object_list(box, src=0)
        if box[0] is None:   # no room anywhere: every rank runs the (deterministic) setup itself
            self.mgl, self.how, self.owner = hf.amg_setup(A, amg), "FASP sequential setup on every rank", True
            return
        self.dir = box[0]
        meta = [None]
        if rank == 0:
            self.mgl, self.owner = hf.amg_setup(A, amg), True
            nl = self.mgl[0].num_levels
            shapes = []
            for l in range(nl):
                lv = {}
                for nm in ("A", "P", "R"):
                    if nm != "A" and l == nl - 1:
                        continue
                    m = getattr(self.mgl[l], nm)
                    lv[nm] = (int(m.row), int(m.col), int(m.nnz), bool(m.val))
                    np.save(os.path.join(self.dir, "%d%s_ia.npy" % (l, nm)), np.ctypeslib.as_array(m.IA, shape=(m.row + 1,)))
                    if m.nnz:
                        np.save(os.path.join(self.dir, "%d%s_ja.npy" % (l, nm)), np.ctypeslib.as_array(m.JA, shape=(m.nnz,)))
                        if m.val:
                            np.save(os.path.join(self.dir, "%d%s_val.npy" % (l, nm)),
                                    np.ctypeslib.as_array(m.val, shape=(m.nnz,)))
                shapes.append(lv)
            meta[0] = (nl, shapes, bytes(amg))
        dist.broadcast_object_list(meta, src=0)
        nl, shapes, amg_bytes = meta[0]
        C.memmove(C.byref(amg), amg_bytes, C.sizeof(amg))
        self.how = "FASP sequential setup on rank 0, mapped read-only by the other ranks from %s" % os.path.dirname(self.dir)
        if rank != 0:
            arr = (T.AMG_data * max(int(amg.max_levels), nl))()
            for l, lv in enumerate(shapes):
                for nm, (row, col, nnz, has_val) in lv.items():
                    ia = np.load(os.path.join(self.dir, "%d%s_ia.npy" % (l, nm)), mmap_mode="r")
                    ja = np.load(os.path.join(self.dir, "%d%s_ja.npy" % (l, nm)), mmap_mode="r") if nnz else None
                    va = np.load(os.path.join(self.dir, "%d%s_val.npy" % (l, nm)), mmap_mode="r") if (nnz and has_val) else None
                    self._keep += [ia, ja, va]
                    m = T.dCSRmat(row, col, nnz, C.cast(ia.ctypes.data, T.PINT),
                                  C.cast(ja.ctypes.data, T.PINT) if ja is not None else None,
                                  C.cast(va.ctypes.data, T.PREAL) if va is not None else None)
                    setattr(arr[l], nm, m)
            arr[0].num_levels = nl
            arr[0].max_levels = max(int(amg.max_levels), nl)
            self.mgl = arr

    def close(self):
        import shutil
        import torch.distributed as dist
        if self.owner:
            self.hf.amg_free(self.mgl, self.amg)
        self.mgl = None
        self._keep = []
        if self.dir is not None:
            if dist.is_initialized():
                dist.barrier()
            if self.rank == 0:
                shutil.rmtree(self.dir, ignore_errors=True)


# ---------------------------------------------------------------------------------------
# bench.py --gpus N
# ---------------------------------------------------------------------------------------
def bench_main(args):
    import bench as B   # repo-root bench.py (helpers: problem, recipe, clocks, peaks)
    rank, world, local = init_comm()
    L = api.lib()
    for kv in getattr(args, "opt", []):
        k, v = kv.split("=")
        api.check(L.fasp_cuda_set_option(k.encode(), float(v)))
    log = (lambda *a: print(*a, file=sys.stderr, flush=True)) if rank == 0 else (lambda *a: None)

    hf = B.host_fasp()
    if rank == 0:
        A, b = B.build_problem(args.n)
        n = A.shape[0]
    else:   # only rank 0 needs the global matrix (setup, true residual); b = 1 everywhere
        A, n = None, args.n ** 3
        b = np.ones(n)
    amg, it = B.amg_recipe(hf)
    t = time.time()
    # the host hierarchy is built ONCE per node: rank 0 runs FASP's setup, the other ranks map its arrays
    # read-only from shared memory (each rank only uploads its own row slabs of it)
    shared = SharedHierarchy(hf, A, amg, rank, world)
    mgl = shared.mgl
    t_setup = time.time() - t
    info = api.hierarchy_info(mgl)
    log("[bench] host AMG setup (%s): %.1fs, %d levels" % (shared.how, t_setup, len(info)))
    t = time.time()
    solver = DistSolver(mgl, amg, agg_rows=args.agg_rows)
    t_upload = time.time() - t
    r0, r1 = solver.row0, solver.row1
    nloc = r1 - r0
    b_loc = b[r0:r1].copy()   # this rank's own array (page-locked below), not a view of the global vector
    zero = np.zeros(nloc)
    log("[bench] rank 0 owns rows [%d, %d) of %d; upload %.2fs" % (r0, r1, n, t_upload))

    x_buf = np.zeros(nloc)
    api.pin_host(b_loc)    # the application's own slices, page-locked once
    api.pin_host(x_buf)

    def host_solve():
        st, x = solver.solve(b_loc, zero, it, out=x_buf)
        if st < 0:
            raise RuntimeError("solve failed on rank %d: %d %s" % (rank, st, api.last_error()))
        return st, x

    for _ in range(args.warmup):
        iters, _x = host_solve()
    barrier()
    sampler = B.ClockSampler(local)
    sampler.start()
    L.fasp_cuda_launch_count_reset()
    dev_ms, e2e_ms = [], []
    for _ in range(args.steps):
        barrier()
        iters, x_loc = host_solve()
        dev_ms.append(allreduce_max(solver.stat(2)))     # device time of the Krylov loop, max over ranks
        e2e_ms.append(allreduce_max(solver.stat(4)))     # incl. staging + H2D/D2H of the local slices
    launches = int(L.fasp_cuda_launch_count())
    clocks = sampler.stop()
    barrier()
    # roofline of the dominant kernel on this rank's slab: CUDA events around every launch of one
    # more solve (graphs off); level-0 local matrix = the largest (rows, nnz) record
    L.fasp_cuda_set_option(b"profile", 1.0)
    L.fasp_cuda_profile_dump(None, 0)
    host_solve()
    buf = C.create_string_buffer(64 << 20)
    L.fasp_cuda_profile_dump(buf, len(buf))
    L.fasp_cuda_set_option(b"profile", 0.0)
    recs = [ln.split() for ln in buf.value.decode().splitlines()]
    recs = [(int(k), int(r), int(z_), float(ms_), float(by)) for k, r, z_, ms_, by in recs]
    mat = [r for r in recs if r[0] < 50 and r[1] == nloc]
    roofline = None
    if mat:
        top = max(r[2] for r in mat)
        l0 = [r for r in mat if r[2] == top]
        med = float(np.median([r[3] for r in l0]))
        l0 = [r for r in l0 if r[3] >= 0.25 * med]
        t_ms = sum(r[3] for r in l0)
        ach = sum(r[4] for r in l0) / t_ms * 1e-6
        peak, src = B.peaks()
        comm_ms = sum(r[3] for r in recs if r[0] >= 400)
        roofline = {"bound": "hbm", "kernel": "csr_pipe_kernel on rank 0's level-0 slab (%d rows, %d nnz)" % (nloc, top),
                    "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "frac_of_nominal_8000": ach / 8000.0, "traffic": None,
                    "peak_source": src, "launch_ms": t_ms / len(l0),
                    "comm_ms_per_solve_profiled": comm_ms, "comm_ops_per_solve": len([r for r in recs if r[0] >= 400])}
    # true residual of the assembled solution (rank 0)
    import torch
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, (r0, x_loc))
    out = None
    if rank == 0:
        x = np.empty(n)
        for p0, xp in parts:
            x[p0:p0 + xp.size] = xp
        true_rel = float(np.linalg.norm(b - A.to_scipy() @ x) / np.linalg.norm(b))
        if not true_rel <= 1e-8 * 1.001:
            raise RuntimeError("assembled solution misses the tolerance: %g" % true_rel)
        # N > 1 parity, visible to the driver: rank 0 also runs the ONE-GPU solve on the same hierarchy and
        # compares (the other ranks wait at the barrier below); the bench fails if the solutions differ
        s1 = api.KrylovAmgSolver(mgl, amg)
        it1, x1 = s1.solve(b, np.zeros(n), it)
        s1.close()
        dx = float(np.linalg.norm(x - x1) / np.linalg.norm(x1))
        parity = {"iters_1": int(it1), "iters_N": int(iters), "dx_rel": dx, "bar": "|iters_N - iters_1| <= 1, dx_rel <= 1e-8"}
        log("[bench] parity vs the one-GPU solve: %s" % parity)
        if it1 < 0 or abs(int(iters) - int(it1)) > 1 or not dx <= 1e-8:
            raise RuntimeError("multi-GPU solve differs from the one-GPU solve: %s (%s)" % (parity, api.last_error()))
        ms = float(np.mean(dev_ms))
        out = {
            "metric": B.METRIC, "value": ms, "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: 3D 7-point Poisson %d^3 (%d rows, %d nnz), rhs=1, AMG-PCG tol 1e-8, "
                                   "classical RS (FASP host setup), V(1,1) L1-Jacobi; rows partitioned over %d GPUs, "
                                   "levels below %d rows replicated" % (args.n, n, A.nnz, world, args.agg_rows),
                       "levels": len(info), "iterations": int(iters), "true_relres": true_rel,
                       "l2_policy": "inputs larger than L2", "setup_s_host": t_setup, "setup_how": shared.how,
                       "upload_s": t_upload,
                       "parallelism": ("row slabs x%d, ghost push + flag barrier + all-reduce over peer-mapped memory (NVLink)"
                                       if L.fasp_cuda_comm_peer_memory() else
                                       "row slabs x%d, NCCL halo send/recv + allreduce") % world},
            "e2e": {"value": float(np.mean(e2e_ms)), "unit": B.UNIT, "h2d_bytes_per_step": int(16 * nloc),
                    "d2h_bytes_per_step": int(8 * nloc)},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": roofline, "cpu_baseline": None, "parity": parity,
        }
    barrier()
    api.unpin_host(b_loc)
    api.unpin_host(x_buf)
    solver.close()
    shared.close()
    # configs[2] proxy (27-point 256^3, the size the sequential oracle can hold) through the slab path, on the same
    # communicator: the strong-scaling line of the 27-point operator in every scaling run
    if getattr(args, "extras", 0) and getattr(args, "c3_n", 0) > 0:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
        try:
            import bench_config3 as C3
            a3 = C3.parse(["--size", str(args.c3_n), "--steps", "5", "--warmup", "3"] +
                          [x for kv in getattr(args, "opt", []) for x in ("--opt", kv)])
            res3 = C3.run(a3, own_comm=False)
            if out is not None:
                out["extra"] = {"config3_proxy": res3}
        except BaseException as e:   # the headline line is never at risk
            if out is not None:
                out["extra"] = {"config3_proxy": {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:300])}}
    L.fasp_cuda_comm_finalize()
    return out


def _quiet(fn, *a):
    import contextlib
    import io
    with contextlib.redirect_stderr(io.StringIO()):
        return fn(*a)


# ---------------------------------------------------------------------------------------
# CPU helpers for the gloo tests
# ---------------------------------------------------------------------------------------
def extract_host(A: T.CSR, nranks: int, rank: int):
    """The library's host-side slab extraction (no GPU): local ia/ja, ghost columns, send lists."""
    L = api.lib()
    n = A.shape[0]
    off = [(n * r) // nranks for r in range(nranks + 1)]
    r0, r1 = off[rank], off[rank + 1]
    nnz_loc = int(A.ia[r1] - A.ia[r0])
    ia = np.zeros(r1 - r0 + 1, dtype=np.int32)
    ja = np.zeros(max(nnz_loc, 1), dtype=np.int32)
    ghosts = np.zeros(max(nnz_loc, 1), dtype=np.int32)
    send_idx = np.zeros(max(r1 - r0, 1) * max(nranks - 1, 1), dtype=np.int32)
    send_counts = np.zeros(nranks, dtype=np.int32)
    ng = C.c_int(0)
    pi = lambda a: a.ctypes.data_as(T.PINT)
    st = L.fasp_cuda_dist_extract_host(A.ptr(), nranks, rank, pi(ia), pi(ja), pi(ghosts), ghosts.size, C.byref(ng),
                                       pi(send_idx), send_idx.size, pi(send_counts))
    api.check(st)
    val = A.val[A.ia[r0]:A.ia[r1]].copy()
    return dict(off=off, ia=ia, ja=ja[:nnz_loc], val=val, ghosts=ghosts[:ng.value],
                send_idx=send_idx[:int(send_counts.sum())], send_counts=send_counts)
