"""Distributed classical-AMG setup over row slabs — no rank ever holds a global matrix.

Why: FASP's INT is 32-bit (fasp.h:72), so the 27-point 512^3 system of BASELINE configs[2] (3.6 G nonzeros)
cannot be one dCSRmat and fasp_amg_setup_rs cannot run on it (SURVEY.md finding 4, §7 hard part 4 option a).
The hierarchy is therefore built slab by slab with FASP's OWN per-level routines — the body of the
`while` loop of fasp_amg_setup_rs (PreAMGSetupRS.c:128-243) — called on every rank's slab:

  level l, on every rank (lock step):
    1. A_loc  = the slab's diagonal block (owned rows x owned columns); the entries that couple to another
                slab are lumped onto the diagonal, so that rows next to a seam keep their row sum (the
                interpolation weights of fasp_amg_interp then still add up to one there);
    2. fasp_amg_coarsening_rs(A_loc) + fasp_amg_interp(A_loc)  -> P_loc (PreAMGCoarsenRS.c:76, PreAMGInterp.c:66);
       coarse points are numbered rank by rank (coarse offsets = prefix sum of the ranks' counts), so
       P = blockdiag(P_loc) and R = P^T (fasp_dcsr_trans) stay local;
    3. Galerkin product with the TRUE slab of A_l (all couplings across the seams): the rows of P that belong
       to the slab's ghost columns are fetched from their owners, then fasp_blas_dcsr_rap (BlaSpmvCSR.c:999)
       forms this rank's rows of A_{l+1} = R A_l P in global coarse numbering.
  Below `agg_rows` global rows the level is gathered on every rank and FASP's unmodified fasp_amg_setup_rs
  builds the replicated rest of the hierarchy.

With ONE rank every step degenerates to the reference's own call sequence on the same data, so the hierarchy
is bit-identical to fasp_amg_setup_rs (tests/test_slab_setup.py). With several ranks the coarsening near the
seams differs from the global one (slab-local C/F splitting, as in hypre's "RS0"); the hierarchy is still a
Galerkin hierarchy of the true operator. Its oracle: the slabs are assembled into global CSR matrices (small
sizes) and handed to the REFERENCE's fasp_solver_dcsr_pcg + fasp_precond_amg (`assemble_global`).

Host-side plumbing (torch.distributed, gloo) moves index lists and matrix rows between the ranks at setup;
nothing here is on the solve path.
"""
from __future__ import annotations

import ctypes as C
import pickle

import numpy as np

from . import fasp_types as T
from .fasp_types import CSR


class iCSRmat(C.Structure):
    """fasp.h:190-210"""
    _fields_ = [("row", T.INT), ("col", T.INT), ("nnz", T.INT), ("IA", T.PINT), ("JA", T.PINT), ("val", T.PINT)]


# ---------------------------------------------------------------------------------------
# host communication (setup only)
# ---------------------------------------------------------------------------------------
class HostComm:
    """All-gather / all-to-all of Python objects over torch.distributed (any CPU backend)."""

    def __init__(self, rank=0, world=1):
        self.rank, self.world = rank, world

    def allgather(self, obj):
        if self.world == 1:
            return [obj]
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out

    def alltoall(self, items):
        """items[q] goes to rank q; returns what every rank sent to me (by source rank)."""
        if self.world == 1:
            return [items[0]]
        import torch
        import torch.distributed as dist
        blobs = [pickle.dumps(it, protocol=pickle.HIGHEST_PROTOCOL) if q != self.rank else b""
                 for q, it in enumerate(items)]
        sizes = self.allgather([len(b) for b in blobs])
        reqs, recv, keep = [], {}, []
        for q in range(self.world):
            if q == self.rank:
                continue
            n_in = sizes[q][self.rank]
            if n_in:
                t = torch.empty(n_in, dtype=torch.uint8)
                recv[q] = t
                reqs.append(dist.irecv(t, src=q))
            if len(blobs[q]):
                t = torch.frombuffer(bytearray(blobs[q]), dtype=torch.uint8)
                keep.append(t)
                reqs.append(dist.isend(t, dst=q))
        for r in reqs:
            r.wait()
        out = []
        for q in range(self.world):
            if q == self.rank:
                out.append(items[q])
            else:
                out.append(pickle.loads(recv[q].numpy().tobytes()) if q in recv else None)
        return out


def _owner_split(ids, off):
    """Split ascending global ids by owner under the contiguous partition `off`."""
    cut = np.searchsorted(ids, off)
    return [ids[cut[q]:cut[q + 1]] for q in range(len(off) - 1)]


def fetch_rows(comm: HostComm, M: CSR, off, want):
    """Rows `want` (ascending global row numbers, none of them mine) of the row-partitioned matrix whose local
    slab is M (global column numbers). Collective. Returns a CSR with len(want) rows."""
    want = np.asarray(want, dtype=np.int64)
    r0 = int(off[comm.rank])
    reqs = comm.alltoall(_owner_split(want, off))
    replies = []
    for q, ids in enumerate(reqs):
        if q == comm.rank or ids is None or len(ids) == 0:
            replies.append(None)
            continue
        loc = np.asarray(ids, dtype=np.int64) - r0
        cnt = (M.ia[loc + 1] - M.ia[loc]).astype(np.int64)
        idx = _ranges(M.ia[loc].astype(np.int64), cnt)
        replies.append((cnt.astype(np.int32), M.ja[idx], M.val[idx]))
    got = comm.alltoall(replies)
    cnts, jas, vals = [], [], []
    for q in range(comm.world):
        if q != comm.rank and got[q] is not None:
            cnts.append(got[q][0]); jas.append(got[q][1]); vals.append(got[q][2])
    cnt = np.concatenate(cnts) if cnts else np.zeros(0, np.int32)
    assert cnt.size == want.size, "fetch_rows: owners returned %d of %d rows" % (cnt.size, want.size)
    ia = np.zeros(want.size + 1, dtype=np.int64)
    np.cumsum(cnt, out=ia[1:])
    return CSR(want.size, M.shape[1], ia.astype(np.int32),
               np.concatenate(jas) if jas else np.zeros(0, np.int32),
               np.concatenate(vals) if vals else np.zeros(0))


def fetch_entries(comm: HostComm, x_loc, off, want):
    """Entries `want` (ascending global indices, not mine) of a partitioned vector. Collective."""
    want = np.asarray(want, dtype=np.int64)
    r0 = int(off[comm.rank])
    reqs = comm.alltoall(_owner_split(want, off))
    got = comm.alltoall([None if (q == comm.rank or ids is None or len(ids) == 0) else x_loc[np.asarray(ids) - r0]
                         for q, ids in enumerate(reqs)])
    parts = [got[q] for q in range(comm.world) if q != comm.rank and got[q] is not None]
    return np.concatenate(parts) if parts else np.zeros(0)


def _ranges(start, count):
    """Concatenation of arange(start[i], start[i] + count[i])."""
    total = int(count.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    ends = np.cumsum(count)
    out = np.ones(total, dtype=np.int64)
    nz = count > 0
    first = (ends - count)[nz]
    s = start[nz]
    out[first] = s - np.concatenate(([0], (s + count[nz] - 1)[:-1]))
    return np.cumsum(out)


def _stack_rows(top: CSR, bottom: CSR) -> CSR:
    if bottom.shape[0] == 0:
        return top
    ia = np.concatenate((top.ia.astype(np.int64), top.ia[-1] + bottom.ia[1:].astype(np.int64)))
    return CSR(top.shape[0] + bottom.shape[0], top.shape[1], ia.astype(np.int32),
               np.concatenate((top.ja, bottom.ja)), np.concatenate((top.val, bottom.val)))


def ghost_columns(A: CSR, c0, c1):
    ja = A.ja
    g = ja[(ja < c0) | (ja >= c1)]
    return np.unique(g).astype(np.int64)


# ---------------------------------------------------------------------------------------
# FASP's per-level routines
# ---------------------------------------------------------------------------------------
class _Fasp:
    def __init__(self, hf):
        L = hf.L
        P = C.POINTER
        L.fasp_amg_coarsening_rs.argtypes = [P(T.dCSRmat), P(T.ivector), P(T.dCSRmat), P(iCSRmat), P(T.AMG_param)]
        L.fasp_amg_coarsening_rs.restype = T.SHORT
        L.fasp_amg_interp.argtypes = [P(T.dCSRmat), P(T.ivector), P(T.dCSRmat), P(iCSRmat), P(T.AMG_param)]
        L.fasp_amg_interp.restype = None
        L.fasp_dcsr_trans.argtypes = [P(T.dCSRmat), P(T.dCSRmat)]
        L.fasp_dcsr_trans.restype = T.INT
        L.fasp_blas_dcsr_rap.argtypes = [P(T.dCSRmat)] * 4
        L.fasp_blas_dcsr_rap.restype = None
        L.fasp_dcsr_free.argtypes = [P(T.dCSRmat)]
        L.fasp_dcsr_free.restype = None
        L.fasp_mem_free.argtypes = [C.c_void_p]
        L.fasp_mem_free.restype = None
        self.L = L

    def coarsen_interp(self, A_loc: CSR, param):
        """One pass of the reference's loop body (PreAMGSetupRS.c:161-199): C/F splitting, interpolation.
        Returns P_loc or None when the reference would stop coarsening here (its checks 1-3)."""
        L = self.L
        n = A_loc.shape[0]
        vert = np.zeros(max(n, 1), dtype=np.int32)
        vertices = T.ivector(n, vert.ctypes.data_as(T.PINT))
        P, S = T.dCSRmat(), iCSRmat()
        st = L.fasp_amg_coarsening_rs(A_loc.ptr(), C.byref(vertices), C.byref(P), C.byref(S), C.byref(param))
        ok = st >= 0 and P.col >= T.MIN_CDOF and not (P.row > P.col * 10.0)
        out = None
        if ok:
            L.fasp_amg_interp(A_loc.ptr(), C.byref(vertices), C.byref(P), C.byref(S), C.byref(param))
            out = CSR.from_struct(P)
        if S.IA:
            L.fasp_mem_free(C.cast(S.IA, C.c_void_p))
        if S.JA:
            L.fasp_mem_free(C.cast(S.JA, C.c_void_p))
        if st >= 0:
            L.fasp_dcsr_free(C.byref(P))
        return out

    def trans(self, P: CSR) -> CSR:
        R = T.dCSRmat()
        self.L.fasp_dcsr_trans(P.ptr(), C.byref(R))
        out = CSR.from_struct(R)
        self.L.fasp_dcsr_free(C.byref(R))
        return out

    def rap(self, R: CSR, A: CSR, P: CSR) -> CSR:
        B = T.dCSRmat()
        self.L.fasp_blas_dcsr_rap(R.ptr(), A.ptr(), P.ptr(), C.byref(B))
        out = CSR.from_struct(B)
        self.L.fasp_dcsr_free(C.byref(B))
        return out


def local_block(A: CSR, c0, c1):
    """Diagonal block of a slab (columns [c0, c1) -> 0 ..), off-slab entries lumped onto the diagonal. With one
    rank this is A itself, entry for entry."""
    n = A.shape[0]
    inside = (A.ja >= c0) & (A.ja < c1)
    if inside.all():
        return CSR(n, c1 - c0, A.ia, A.ja - np.int32(c0), A.val)
    rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(A.ia))
    out_idx = np.nonzero(~inside)[0]
    lump = np.bincount(rows[out_idx], weights=A.val[out_idx], minlength=n)
    del out_idx
    r_in = rows[inside]
    del rows
    cnt = np.bincount(r_in, minlength=n)
    ia = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=ia[1:])
    ja = A.ja[inside] - np.int32(c0)
    val = A.val[inside]
    dpos = np.nonzero(ja == r_in)[0]
    val[dpos] += lump[r_in[dpos]]
    return CSR(n, c1 - c0, ia.astype(np.int32), ja, val)


# ---------------------------------------------------------------------------------------
# the level loop
# ---------------------------------------------------------------------------------------
class SlabLevel:
    """One partitioned level on this rank: A (owned rows), P (owned rows + the ghost rows of A's slab), R (owned
    coarse rows + the ghost rows of the next level's slab); all with GLOBAL column numbers."""
    __slots__ = ("A", "P", "R", "off", "coff", "n_pext", "n_rext", "ghosts")


class SlabHierarchy:
    def __init__(self, hf, A_slab: CSR, off, amg, comm: HostComm | None = None, agg_rows=8000, log=None):
        self.hf, self.amg = hf, amg
        self.comm = comm or HostComm()
        self.F = _Fasp(hf)
        self.levels: list[SlabLevel] = []
        self.tail = None          # AMG_data array of the replicated rest (FASP's own setup)
        self.tail_A = None
        log = log or (lambda *a: None)
        comm, rank = self.comm, self.comm.rank
        if amg.AMG_type != T.CLASSIC_AMG or amg.coarsening_type != T.COARSE_RS:
            raise ValueError("slab setup: classical AMG with COARSE_RS only")
        amg.tentative_smooth = 1.0     # PreAMGSetupRS.c:83
        A = A_slab
        off = np.asarray(off, dtype=np.int64)
        max_part = int(amg.max_levels) - 2
        while int(off[-1]) >= agg_rows and len(self.levels) < max_part and int(off[-1]) > max(amg.coarse_dof, T.MIN_CDOF):
            c0, c1 = int(off[rank]), int(off[rank + 1])
            n_loc = c1 - c0
            assert A.shape[0] == n_loc
            P_loc = self.F.coarsen_interp(local_block(A, c0, c1), amg) if n_loc > 0 else None
            info = comm.allgather((P_loc is not None or n_loc == 0, 0 if P_loc is None else P_loc.shape[1]))
            if not all(ok for ok, _ in info):
                break
            ncs = np.array([nc for _, nc in info], dtype=np.int64)
            coff = np.concatenate(([0], np.cumsum(ncs)))
            NC = int(coff[-1])
            if NC >= 2 ** 31 - 1:
                raise ValueError("coarse level exceeds 32-bit column numbers")
            if P_loc is None:
                P_loc = CSR(0, 0, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0))
            nc_loc = P_loc.shape[1]
            R_loc = self.F.trans(P_loc)
            P_glob = CSR(n_loc, NC, P_loc.ia, P_loc.ja + np.int32(coff[rank]), P_loc.val)
            ghosts = ghost_columns(A, c0, c1)
            P_gh = fetch_rows(comm, P_glob, off, ghosts)
            A_next = self._galerkin(A, c0, c1, ghosts, P_loc, P_gh, R_loc, int(coff[rank]), NC)
            lv = SlabLevel()
            lv.A, lv.off, lv.coff, lv.ghosts = A, off, coff, ghosts
            lv.P = _stack_rows(P_glob, P_gh)
            lv.n_pext = int(ghosts.size)
            lv.R = CSR(nc_loc, int(off[-1]), R_loc.ia, R_loc.ja + np.int32(c0), R_loc.val)
            lv.n_rext = 0
            self.levels.append(lv)
            log("[slab setup] level %d: %d rows (%d here, %d ghosts) -> %d coarse rows" %
                (len(self.levels) - 1, int(off[-1]), n_loc, ghosts.size, NC))
            A, off = A_next, coff
        if not self.levels:
            raise ValueError("slab setup: the finest level could not be coarsened on every rank")
        # the rows of R the neighbours' A_{l+1} slabs gather (redundant ghost rows of b_{l+1}, dist.cu)
        for l in range(len(self.levels) - 1):
            lv, nx = self.levels[l], self.levels[l + 1]
            R_gh = fetch_rows(comm, lv.R, lv.coff, nx.ghosts)
            lv.R = _stack_rows(lv.R, R_gh)
            lv.n_rext = int(nx.ghosts.size)
        # gather the first replicated level everywhere and let FASP build the rest
        parts = comm.allgather((A.ia, A.ja, A.val))
        N = int(off[-1])
        ia = np.zeros(N + 1, dtype=np.int64)
        pos = 1
        for p_ia, _, _ in parts:
            m = p_ia.size - 1
            ia[pos:pos + m] = ia[pos - 1] + p_ia[1:].astype(np.int64)
            pos += m
        self.tail_A = CSR(N, N, ia.astype(np.int32), np.concatenate([p[1] for p in parts]),
                          np.concatenate([p[2] for p in parts]))
        self.tail_off = off
        self.tail_amg = type(amg).from_buffer_copy(bytes(amg))
        self.tail_amg.max_levels = int(amg.max_levels) - len(self.levels)
        self.tail = hf.amg_setup(self.tail_A, self.tail_amg)
        log("[slab setup] replicated from level %d: %d rows, %d more levels" %
            (len(self.levels), N, self.tail[0].num_levels))

    def _galerkin(self, A, c0, c1, ghosts, P_loc, P_gh, R_loc, cbase, NC):
        """This rank's rows of R A P. fasp_blas_dcsr_rap assumes square operands (its marker arrays are sized by
        the row counts, BlaSpmvCSR.c:1042-1050), so the slab is embedded: fine space = owned + ghost columns
        (ghost ROWS empty), coarse space = owned coarse points + the coarse points the ghost rows of P reach."""
        n_loc, nc_loc, ng = c1 - c0, P_loc.shape[1], int(ghosts.size)
        if ng == 0:
            return_cols = None
            A_sq = CSR(n_loc, n_loc, A.ia, A.ja - np.int32(c0), A.val)
            out = self.F.rap(R_loc, A_sq, P_loc)
            return CSR(nc_loc, NC, out.ia, out.ja + np.int32(cbase), out.val)
        ja_e = A.ja - np.int32(c0)
        outside = np.nonzero((A.ja < c0) | (A.ja >= c1))[0]
        ja_e[outside] = (n_loc + np.searchsorted(ghosts, A.ja[outside])).astype(np.int32)
        del outside
        ia_e = np.concatenate((A.ia, np.full(ng, A.ia[-1], dtype=np.int32)))
        A_sq = CSR(n_loc + ng, n_loc + ng, ia_e, ja_e, A.val)
        cg = np.unique(P_gh.ja).astype(np.int64)          # global coarse columns of the ghost rows (not mine)
        ncg = int(cg.size)
        pg_ja = (nc_loc + np.searchsorted(cg, P_gh.ja.astype(np.int64))).astype(np.int32)
        P_sq = CSR(n_loc + ng, nc_loc + ncg,
                   np.concatenate((P_loc.ia.astype(np.int64), P_loc.ia[-1] + P_gh.ia[1:].astype(np.int64))).astype(np.int32),
                   np.concatenate((P_loc.ja, pg_ja)), np.concatenate((P_loc.val, P_gh.val)))
        R_sq = CSR(nc_loc + ncg, n_loc + ng,
                   np.concatenate((R_loc.ia, np.full(ncg, R_loc.ia[-1], dtype=np.int32))), R_loc.ja, R_loc.val)
        out = self.F.rap(R_sq, A_sq, P_sq)
        nnz = int(out.ia[nc_loc])
        ja = out.ja[:nnz].astype(np.int64)
        ja_g = np.where(ja < nc_loc, ja + cbase, cg[np.clip(ja - nc_loc, 0, max(ncg - 1, 0))] if ncg else ja)
        return CSR(nc_loc, NC, out.ia[:nc_loc + 1], ja_g.astype(np.int32), out.val[:nnz])

    # -- the solver object --------------------------------------------------------------
    def create_solver(self):
        from . import api
        L = api.lib()
        n = len(self.levels)
        arr = (T.fasp_cuda_slab_level * n)()
        keep = []
        for l, lv in enumerate(self.levels):
            off32 = np.ascontiguousarray(lv.off, dtype=np.int32)
            keep.append(off32)
            arr[l].A, arr[l].P, arr[l].R = lv.A.struct, lv.P.struct, lv.R.struct
            arr[l].row_off = off32.ctypes.data_as(T.PINT)
            arr[l].n_pext, arr[l].n_rext = lv.n_pext, lv.n_rext
        toff = np.ascontiguousarray(self.tail_off, dtype=np.int32)
        h = L.fasp_cuda_dist_krylov_amg_create_slabs(n, arr, toff.ctypes.data_as(T.PINT), self.tail,
                                                     C.byref(self.tail_amg))
        if not h:
            raise api.FaspCudaError(-1, api.last_error())
        return h

    def close(self):
        if self.tail is not None:
            self.hf.amg_free(self.tail, self.tail_amg)
            self.tail = None

    # -- oracle support: the same hierarchy as global matrices (small problems only) --------
    def assemble_global(self):
        """Every rank returns the list of global (A, P, R) per partitioned level + the tail's A (CSR objects)."""
        out = []
        for lv in self.levels:
            n_loc = int(lv.off[self.comm.rank + 1] - lv.off[self.comm.rank])
            nc_loc = int(lv.coff[self.comm.rank + 1] - lv.coff[self.comm.rank])
            trip = []
            for M, rows in ((lv.A, n_loc), (lv.P, n_loc), (lv.R, nc_loc)):
                nnz = int(M.ia[rows])
                parts = self.comm.allgather((M.ia[:rows + 1], M.ja[:nnz], M.val[:nnz]))
                tot = sum(p[0].size - 1 for p in parts)
                ia = np.zeros(tot + 1, dtype=np.int64)
                pos = 1
                for p_ia, _, _ in parts:
                    m = p_ia.size - 1
                    ia[pos:pos + m] = ia[pos - 1] + p_ia[1:].astype(np.int64)
                    pos += m
                trip.append(CSR(tot, M.shape[1], ia.astype(np.int32), np.concatenate([p[1] for p in parts]),
                                np.concatenate([p[2] for p in parts])))
            out.append(tuple(trip))
        return out, self.tail_A


def assemble_mgl(hf, sh: "SlabHierarchy"):
    """The slab hierarchy as ONE global AMG_data array owned by the host FASP (small problems only): input of
    the reference's fasp_precond_amg (the oracle of the slab path) and of the one-GPU device solver."""
    L = hf.L
    glob, _ = sh.assemble_global()
    nt = sh.tail[0].num_levels
    nl = len(glob) + nt
    amg = type(sh.amg).from_buffer_copy(bytes(sh.amg))
    amg.max_levels = max(int(sh.amg.max_levels), nl)
    mgl = L.fasp_amg_data_create(amg.max_levels)

    def put(dst_owner, name, M):
        m = L.fasp_dcsr_create(M.shape[0], M.shape[1], M.nnz)
        L.fasp_dcsr_cp(M.ptr(), C.byref(m))
        setattr(dst_owner, name, m)

    for l in range(nl):
        if l < len(glob):
            A, P, R = glob[l]
        else:
            t = sh.tail[l - len(glob)]
            A = CSR.from_struct(t.A)
            P = CSR.from_struct(t.P) if l < nl - 1 else None
            R = CSR.from_struct(t.R) if l < nl - 1 else None
        put(mgl[l], "A", A)
        if P is not None:
            put(mgl[l], "P", P)
            put(mgl[l], "R", R)
        mgl[l].num_levels = nl
        mgl[l].b = L.fasp_dvec_create(A.shape[0])
        mgl[l].x = L.fasp_dvec_create(A.shape[0])
        mgl[l].cycle_type = amg.cycle_type
    mgl[0].w = L.fasp_dvec_create(glob[0][0].shape[0])
    for l in range(1, nl):
        mgl[l].w = L.fasp_dvec_create(2 * mgl[l].A.row)   # PreAMGSetupRS.c:349 (work space of the cycle)
    return mgl, amg
