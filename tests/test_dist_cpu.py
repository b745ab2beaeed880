"""CPU tests of the multi-GPU host logic: world_size-2 gloo processes exchange halos with the plans
the library computes on the host (fasp_cuda_dist_extract_host) and must reproduce the global SpMV."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, scipy.sparse as sp, torch, torch.distributed as dist
from faspsolver_b200 import problems as PB, multigpu as MG
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
for A in (PB.poisson7(6), PB.convdiff7(5), PB.poisson27(4)):
    n = A.shape[0]
    x = np.random.default_rng(1).uniform(-1, 1, n)          # same on every rank
    p = MG.extract_host(A, world, rank)
    off = p["off"]; r0, r1 = off[rank], off[rank + 1]; nloc = r1 - r0
    # plan symmetry: what I receive from q equals what q sends to me
    owners = np.searchsorted(np.array(off), p["ghosts"], side="right") - 1
    recv_counts = np.bincount(owners, minlength=world)
    allsend = [None] * world
    dist.all_gather_object(allsend, p["send_counts"].tolist())
    for q in range(world):
        assert allsend[q][rank] == recv_counts[q], (rank, q, allsend[q][rank], recv_counts[q])
    # halo exchange with gloo: pack owned entries per peer, receive behind the owned part
    xe = np.zeros(nloc + p["ghosts"].size); xe[:nloc] = x[r0:r1]
    reqs, pos, rpos = [], 0, nloc
    bufs = []
    for q in range(world):
        c = int(p["send_counts"][q])
        if c:
            t = torch.from_numpy(xe[p["send_idx"][pos:pos + c]].copy()); bufs.append(t)
            reqs.append(dist.isend(t, dst=q)); pos += c
    recvs = []
    for q in range(world):
        c = int(recv_counts[q])
        if c:
            t = torch.empty(c, dtype=torch.float64); recvs.append((rpos, c, t))
            reqs.append(dist.irecv(t, src=q)); rpos += c
    for r in reqs: r.wait()
    for o, c, t in recvs: xe[o:o + c] = t.numpy()
    assert np.array_equal(xe[nloc:], x[p["ghosts"]])         # ghosts carry the owners' values
    y_loc = sp.csr_matrix((p["val"], p["ja"], p["ia"]), shape=(nloc, xe.size)) @ xe
    y_ref = (A.to_scipy() @ x)[r0:r1]
    assert np.allclose(y_loc, y_ref, rtol=0, atol=1e-13 * np.abs(y_ref).max()), np.abs(y_loc - y_ref).max()
dist.barrier()
print("rank", rank, "ok")
'''


@pytest.mark.parametrize("world", [2, 3])
def test_halo_plans_reproduce_global_spmv_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": str(ROOT)})
    procs = []
    port = 29600 + world + (os.getpid() % 200)
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o)
        assert "ok" in o


def test_partition_covers_rows_once():
    from faspsolver_b200 import multigpu as MG, problems as PB
    A = PB.poisson7(5)
    n = A.shape[0]
    for world in (1, 2, 4, 7):
        seen = np.zeros(n, dtype=int)
        for r in range(world):
            p = MG.extract_host(A, world, r)
            seen[p["off"][r]:p["off"][r + 1]] += 1
            nloc = p["off"][r + 1] - p["off"][r]
            assert p["ia"][-1] == p["ja"].size and p["ja"].max(initial=0) < nloc + p["ghosts"].size
            assert np.all(np.diff(p["ghosts"]) > 0)
        assert np.all(seen == 1)


SHARED_WORKER = r'''
import os, sys, hashlib
sys.path.insert(0, %(root)r)
import numpy as np, torch.distributed as dist
from faspsolver_b200 import api, problems as PB, multigpu as MG, fasp_types as T
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
hf = api.HostFasp(%(lib)r)
A = PB.poisson7(12) if rank == 0 else None            # only rank 0 holds the global matrix
amg = hf.amg_param(print_level=0, smoother=T.SMOOTHER_L1DIAG)
sh = MG.SharedHierarchy(hf, A, amg, rank, world, root=%(tmp)r)
mgl = sh.mgl
nl = mgl[0].num_levels
h = hashlib.sha256()
for l in range(nl):
    for nm in ("A", "P", "R"):
        if nm != "A" and l == nl - 1:
            continue
        m = getattr(mgl[l], nm)
        h.update(np.array([m.row, m.col, m.nnz]).tobytes())
        h.update(np.ctypeslib.as_array(m.IA, shape=(m.row + 1,)).tobytes())
        h.update(np.ctypeslib.as_array(m.JA, shape=(m.nnz,)).tobytes())
        h.update(np.ctypeslib.as_array(m.val, shape=(m.nnz,)).tobytes())
digs = [None] * world
dist.all_gather_object(digs, (h.hexdigest(), nl, float(amg.tentative_smooth), sh.how))
assert len(set(d[:3] for d in digs)) == 1, digs           # identical hierarchy and parameters on every rank
assert nl >= 3 and "rank 0" in sh.how, (nl, sh.how)
# the mapped hierarchy feeds the library's host-side slab extraction like a locally built one
A1 = T.CSR.from_struct(mgl[1].A)
p = MG.extract_host(A1, world, rank)
assert p["ia"][-1] == p["ja"].size
sh.close()
assert rank != 0 or not os.path.exists(sh.dir)
print("rank", rank, "ok")
'''


def test_shared_hierarchy_gloo(tmp_path):
    """multigpu.SharedHierarchy: FASP's setup on rank 0 only, the other ranks map the arrays read-only."""
    lib = ROOT / "oracle" / "_ref" / "libfasp_seq.so"
    if not lib.exists():
        pytest.skip("oracle/_ref/libfasp_seq.so not built")
    world = 2
    script = tmp_path / "worker.py"
    script.write_text(SHARED_WORKER % {"root": str(ROOT), "lib": str(lib), "tmp": str(tmp_path)})
    port = 29700 + (os.getpid() % 200)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o)
        assert "ok" in o
