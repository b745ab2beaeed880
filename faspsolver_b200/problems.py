"""Synthetic inputs of the BASELINE configs and readers for FASP's
on-disk matrix formats. numpy only; used by tests, smoke and bench (not by the product path).

All generators use natural x-fastest ordering, Dirichlet boundaries eliminated (boundary
neighbours simply omitted) and ascending column order inside a row.
"""
from __future__ import annotations

import numpy as np

from .fasp_types import BSR, CSR


def _stencil_csr(nx, ny, nz, offsets, coeffs, chunk_planes=16, zrange=None):
    """CSR of a constant-coefficient stencil on an nx*ny*nz grid.

    zrange = (z0, z1): only the rows of the planes z0 <= z < z1 (a z-slab, GLOBAL column numbers).

    offsets: list of (dx, dy, dz) sorted so that the linear offset dx + nx*(dy + ny*dz) ascends.
    coeffs : one value per offset. Built plane-chunk by plane-chunk to bound peak memory.
    """
    offsets = np.asarray(offsets, dtype=np.int64)
    coeffs = np.asarray(coeffs, dtype=np.float64)
    lin = offsets[:, 0] + nx * (offsets[:, 1] + ny * offsets[:, 2])
    assert np.all(np.diff(lin) > 0), "offsets must be sorted by linear offset"
    N = nx * ny * nz
    zlo, zhi = zrange if zrange is not None else (0, nz)
    row0, nrows = zlo * nx * ny, (zhi - zlo) * nx * ny
    counts = np.empty(nrows, dtype=np.int32)
    ja_parts, va_parts = [], []
    ix = np.arange(nx, dtype=np.int64)
    iy = np.arange(ny, dtype=np.int64)
    for z0 in range(zlo, zhi, chunk_planes):
        z1 = min(zhi, z0 + chunk_planes)
        iz = np.arange(z0, z1, dtype=np.int64)
        Z, Y, X = np.meshgrid(iz, iy, ix, indexing="ij")
        X, Y, Z = X.ravel(), Y.ravel(), Z.ravel()
        base = X + nx * (Y + ny * Z)
        m = X.size
        valid = np.empty((m, len(lin)), dtype=bool)
        for k, (dx, dy, dz) in enumerate(offsets):
            valid[:, k] = ((X + dx >= 0) & (X + dx < nx) & (Y + dy >= 0) & (Y + dy < ny)
                           & (Z + dz >= 0) & (Z + dz < nz))
        cols = (base[:, None] + lin[None, :]).astype(np.int32)
        ja_parts.append(cols[valid])
        va_parts.append(np.broadcast_to(coeffs[None, :], valid.shape)[valid])
        counts[base[0] - row0:base[0] - row0 + m] = valid.sum(axis=1, dtype=np.int32)
    ia = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(counts, out=ia[1:])
    assert ia[-1] < 2 ** 31 and N < 2 ** 31, "matrix exceeds FASP's 32-bit INT (SURVEY.md finding 4)"
    return CSR(nrows, N, ia.astype(np.int32), np.concatenate(ja_parts), np.concatenate(va_parts))


def poisson7(n, scaled=True, ny=None, nz=None, zrange=None):
    """3-D 7-point Poisson on n^3 interior nodes: diag 6 h^-2, off -h^-2, h = 1/(n+1)
    (entries as in test/src/FdmPoisson.c:518-541; ordering as the survey probes)."""
    ny = ny or n
    nz = nz or n
    s = float((n + 1) ** 2) if scaled else 1.0
    offs = [(0, 0, -1), (0, -1, 0), (-1, 0, 0), (0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
    co = [-s, -s, -s, 6 * s, -s, -s, -s]
    return _stencil_csr(n, ny, nz, offs, co, zrange=zrange)


def poisson27(n, ny=None, nz=None, zrange=None):
    """3-D 27-point M-matrix: diag 26, all 26 neighbours -1 (config 3). zrange: rows of a z-slab only."""
    ny = ny or n
    nz = nz or n
    offs = [(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    co = [26.0 if o == (0, 0, 0) else -1.0 for o in offs]
    return _stencil_csr(n, ny, nz, offs, co, zrange=zrange)


def convdiff7(n, peclet=(0.5, 0.25, 0.125)):
    """7-point upwind convection-diffusion (config 4): diag 6+cx+cy+cz, upstream (-)
    neighbours -1-c_d, downstream -1. Nonsymmetric."""
    cx, cy, cz = peclet
    offs = [(0, 0, -1), (0, -1, 0), (-1, 0, 0), (0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
    co = [-1 - cz, -1 - cy, -1 - cx, 6 + cx + cy + cz, -1.0, -1.0, -1.0]
    return _stencil_csr(n, n, n, offs, co)


def poisson5_2d(n):
    """2-D 5-point Poisson on n^2 interior nodes (diag 4, off -1)."""
    offs = [(0, -1, 0), (-1, 0, 0), (0, 0, 0), (1, 0, 0), (0, 1, 0)]
    return _stencil_csr(n, n, 1, offs, [-1.0, -1.0, 4.0, -1.0, -1.0])


# SPE01-shaped synthetic black-oil blocks
_B = np.array([[2.0, -0.5, 0.1], [-0.4, 2.0, -0.5], [0.0, -0.6, 2.0]])
_D = np.array([[0.5, 0.1, 0.0], [0.1, 0.5, 0.1], [0.0, 0.1, 0.5]])


def blockoil7(n):
    """3x3-block 7-point system on n^3 block rows: off-diagonal block s*B with s = -1 (z-,y-,
    z+,y+), -1.2 (x-), -0.8 (x+); diagonal block 6*B + D. Returns (BSR, rhs)."""
    scal = poisson7(n, scaled=False)  # pattern provider
    ia, ja = scal.ia, scal.ja
    rows = np.repeat(np.arange(n ** 3, dtype=np.int64), np.diff(ia))
    d = ja.astype(np.int64) - rows
    s = np.full(ja.size, -1.0)
    s[d == -1] = -1.2
    s[d == 1] = -0.8
    val = s[:, None, None] * _B[None, :, :]
    val[d == 0] = 6.0 * _B + _D
    A = BSR(n ** 3, n ** 3, 3, ia, ja, val)
    rhs = 1.0 + 0.01 * (np.arange(3 * n ** 3) % 7)
    return A, rhs


def rhs_ones(n_rows):
    return np.ones(n_rows)


# ---------------------------------------------------------------------------------------
# FASP file formats (BlaIO.c:164-260, :791-860)
# ---------------------------------------------------------------------------------------
def read_fasp_csr(path):
    """`n; IA[n+1]; JA[nnz]; val[nnz]`, 1-based on disk (fasp_dcsrvec_read2, BlaIO.c:164)."""
    tok = open(path).read().split()
    n = int(tok[0])
    ia = np.array(tok[1:n + 2], dtype=np.int64) - 1
    nnz = int(ia[-1])
    ja = np.array(tok[n + 2:n + 2 + nnz], dtype=np.int64) - 1
    val = np.array(tok[n + 2 + nnz:n + 2 + 2 * nnz], dtype=np.float64)
    return CSR(n, n, ia, ja, val)


def read_fasp_vec(path):
    """`n; val[n]` (fasp_dvec_read, BlaIO.c:938)."""
    tok = open(path).read().split()
    n = int(tok[0])
    return np.array(tok[1:n + 1], dtype=np.float64)


def read_fasp_vecind(path):
    """`n; (index value)[n]` (fasp_dvecind_read, BlaIO.c:887)."""
    tok = open(path).read().split()
    n = int(tok[0])
    idx = np.array(tok[1:2 * n + 1:2], dtype=np.int64)
    val = np.array(tok[2:2 * n + 2:2], dtype=np.float64)
    out = np.zeros(n)
    out[idx] = val
    return out


def read_fasp_bsr(path):
    """`ROW COL NNZ; nb; storage_manner; n IA..; n JA..; n val..` 0-based (fasp_dbsr_read)."""
    tok = open(path).read().split()
    ROW, COL, NNZ = int(tok[0]), int(tok[1]), int(tok[2])
    nb, _sm = int(tok[3]), int(tok[4])
    p = 5
    n = int(tok[p]); p += 1
    ia = np.array(tok[p:p + n], dtype=np.int64); p += n
    n = int(tok[p]); p += 1
    ja = np.array(tok[p:p + n], dtype=np.int64); p += n
    n = int(tok[p]); p += 1
    val = np.array(tok[p:p + n], dtype=np.float64)
    assert ia.size == ROW + 1 and ja.size == NNZ and val.size == NNZ * nb * nb
    return BSR(ROW, COL, nb, ia, ja, val)
