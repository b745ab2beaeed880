#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ from the reference tree.

Run HERE (the build container, where /root/reference exists):
    python tests/golden/make_golden.py
Inputs : /root/reference/data/{csrmat_FD,rhs_FD,sol_FD,csrmat_FE,rhs_FE,sol_FE,
         bsrmat_SPE01,rhs_SPE01}.dat  (FASP's shipped test problems; data, not source)
Oracle : oracle/_ref/libfasp_seq.so (unmodified sequential FASP 2.8.7, oracle/build_ref.sh)
Outputs: tests/golden/fasp_data.npz        the matrices / vectors above as arrays
         tests/golden/oracle_answers.json  iteration counts + final relres of the reference
                                           for the recipes of SURVEY.md Appendix C, and the
                                           golden lines of test/out/reg.gcc they reproduce
         tests/golden/oracle_vectors.npz   reference outputs (SpMV, smoother sweeps, one
                                           V-cycle, solutions) on the FE problem
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from faspsolver_b200 import fasp_types as T  # noqa: E402
from faspsolver_b200 import problems as PB  # noqa: E402
from oracle.ref import RefFasp  # noqa: E402

DATA = Path("/root/reference/data")
OUT = Path(__file__).resolve().parent


def main():
    ref = RefFasp()
    FD = PB.read_fasp_csr(DATA / "csrmat_FD.dat")
    FE = PB.read_fasp_csr(DATA / "csrmat_FE.dat")
    bFD, bFE = PB.read_fasp_vec(DATA / "rhs_FD.dat"), PB.read_fasp_vec(DATA / "rhs_FE.dat")
    sFD, sFE = PB.read_fasp_vecind(DATA / "sol_FD.dat"), PB.read_fasp_vecind(DATA / "sol_FE.dat")
    SPE = PB.read_fasp_bsr(DATA / "bsrmat_SPE01.dat")
    bSPE = PB.read_fasp_vec(DATA / "rhs_SPE01.dat")
    np.savez_compressed(
        OUT / "fasp_data.npz",
        FD_ia=FD.ia, FD_ja=FD.ja, FD_val=FD.val, FD_b=bFD, FD_sol=sFD,
        FE_ia=FE.ia, FE_ja=FE.ja, FE_val=FE.val, FE_b=bFE, FE_sol=sFE,
        SPE_ia=SPE.ia, SPE_ja=SPE.ja, SPE_val=SPE.val, SPE_b=bSPE,
        SPE_dims=np.array([SPE.ROW, SPE.COL, SPE.nb]))

    answers = {"reg_gcc": {
        # lines of /root/reference/test/out/reg.gcc (sequential gcc build of the reference)
        "FE_amg_pcg_default_tol1e-10": {"iters": 6, "relres": 2.728796e-11, "line": 577},
        "FD_amg_pcg_default_tol1e-10": {"iters": 1, "relres": 4.938174e-15, "line": 255},
        "FE_amg_solver_L1DIAG_tol1e-10": {"iters": 19, "relres": 8.612004e-11, "line": 412},
        # unpreconditioned Krylov methods on the FE problem (regression.c:300-640), CSR and "BSR format" (nb = 1)
        "FE_cg_unprec_tol1e-12": {"iters": 244, "relres": 9.975280e-13, "line": 451},
        "FE_gmres_unprec_tol1e-12": {"iters": 937, "relres": 9.895899e-13, "line": 486},
        "FE_bsr_cg_unprec_tol1e-12": {"iters": 244, "relres": 9.975287e-13, "line": 535, "maxdiff": 9.9928e-08},
        "FE_bsr_gmres_unprec_tol1e-8": {"iters": 500, "relres": 6.599950e-08, "line": 549, "maxdiff": 2.8813e-06},
        "FE_bsr_vgmres_unprec_tol1e-8": {"iters": 339, "relres": 9.238968e-09, "line": 556, "maxdiff": 1.8147e-07},
        "FE_bsr_vfgmres_unprec_tol1e-8": {"iters": 339, "relres": 9.238968e-09, "line": 563, "maxdiff": 1.8147e-07},
        # unpreconditioned variable-restart GMRES and its flexible twin, tol 1e-12, restart 25
        "FE_vgmres_unprec_tol1e-12": {"iters": 493, "relres": 7.667271e-13, "line": 500},
        "FE_vfgmres_unprec_tol1e-12": {"iters": 493, "relres": 7.667271e-13, "line": 514},
    }, "recipes": []}

    def run(name, A, b, it_kw, amg_kw):
        it = ref.its_param(tol=1e-8, maxit=500, print_level=0, **it_kw)
        amg = ref.amg_param(print_level=0, **amg_kw)
        st, x = ref.krylov_amg(A, b, np.zeros_like(b), it, amg)
        r = b - A.to_scipy() @ x
        rel = float(np.linalg.norm(r) / np.linalg.norm(b))
        answers["recipes"].append({"name": name, "it": it_kw, "amg": amg_kw, "status": int(st),
                                   "true_relres": rel})
        print(name, st, rel)
        return x

    xs = {}
    run("FD_pcg_jacobi067_cdof20", FD, bFD, dict(itsolver_type=T.SOLVER_CG),
        dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67, coarse_dof=20))
    run("FD_pcg_l1_cdof20", FD, bFD, dict(itsolver_type=T.SOLVER_CG),
        dict(smoother=T.SMOOTHER_L1DIAG, coarse_dof=20))
    run("FD_pcg_l1_default", FD, bFD, dict(itsolver_type=T.SOLVER_CG), dict(smoother=T.SMOOTHER_L1DIAG))
    xs["FE_x_pcg_jacobi067"] = run("FE_pcg_jacobi067", FE, bFE, dict(itsolver_type=T.SOLVER_CG),
                                   dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67))
    xs["FE_x_pcg_l1"] = run("FE_pcg_l1", FE, bFE, dict(itsolver_type=T.SOLVER_CG),
                            dict(smoother=T.SMOOTHER_L1DIAG))
    xs["FE_x_pcg_poly3"] = run("FE_pcg_poly3", FE, bFE, dict(itsolver_type=T.SOLVER_CG),
                               dict(smoother=T.SMOOTHER_POLY, polynomial_degree=3))
    xs["FE_x_gmres_l1"] = run("FE_gmres30_l1", FE, bFE, dict(itsolver_type=T.SOLVER_GMRES, restart=30),
                              dict(smoother=T.SMOOTHER_L1DIAG))
    xs["FE_x_vgmres_poly3"] = run("FE_vgmres30_poly3", FE, bFE, dict(itsolver_type=T.SOLVER_VGMRES, restart=30),
                                  dict(smoother=T.SMOOTHER_POLY, polynomial_degree=3))
    xs["FE_x_vfgmres_l1"] = run("FE_vfgmres30_l1", FE, bFE, dict(itsolver_type=T.SOLVER_VFGMRES, restart=30),
                                dict(smoother=T.SMOOTHER_L1DIAG))
    run("FE_pcg_l1_W", FE, bFE, dict(itsolver_type=T.SOLVER_CG),
        dict(smoother=T.SMOOTHER_L1DIAG, cycle_type=T.W_CYCLE))
    run("FE_pcg_jacobi067_sa", FE, bFE, dict(itsolver_type=T.SOLVER_CG),
        dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67, AMG_type=T.SA_AMG))
    run("FE_pcg_jacobi067_ua", FE, bFE, dict(itsolver_type=T.SOLVER_CG),
        dict(smoother=T.SMOOTHER_JACOBI, relaxation=0.67, AMG_type=T.UA_AMG))
    A64 = PB.poisson7(24)
    run("p7_24_pcg_l1", A64, np.ones(A64.shape[0]), dict(itsolver_type=T.SOLVER_CG),
        dict(smoother=T.SMOOTHER_L1DIAG))
    Acd = PB.convdiff7(24)
    run("cd7_24_gmres30_poly3", Acd, np.ones(Acd.shape[0]), dict(itsolver_type=T.SOLVER_GMRES, restart=30),
        dict(smoother=T.SMOOTHER_POLY, polynomial_degree=3))

    # kernel-level reference outputs on the FE matrix
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, FE.shape[0])
    vec = {"x": x, "mxv": ref.mxv(FE, x), "aAxpy_m1": ref.aAxpy(-1.0, FE, x, bFE),
           "aAxpy_0p3": ref.aAxpy(0.3, FE, x, bFE)}
    for nm, fn in (("jacobi067", lambda u: ref.L.fasp_smoother_dcsr_jacobi(
            u.ptr(), 0, FE.shape[0] - 1, 1, FE.ptr(), T.Vec(bFE).ptr(), 2, 0.67)),
                   ("l1diag", lambda u: ref.L.fasp_smoother_dcsr_L1diag(
                       u.ptr(), 0, FE.shape[0] - 1, 1, FE.ptr(), T.Vec(bFE).ptr(), 2)),
                   ("poly3", lambda u: ref.L.fasp_smoother_dcsr_poly(
                       FE.ptr(), T.Vec(bFE).ptr(), u.ptr(), FE.shape[0], 3, 2))):
        u = T.Vec(x.copy())
        fn(u)
        vec["smooth_" + nm] = u.a.copy()
    vec.update(xs)
    np.savez_compressed(OUT / "oracle_vectors.npz", **vec)
    (OUT / "oracle_answers.json").write_text(json.dumps(answers, indent=1))


if __name__ == "__main__":
    main()
