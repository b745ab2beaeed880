"""CPU tests (no GPU) of the drop-in boundary: struct ABI, exported symbols, loud failure
without a device, host-side helpers."""
import ctypes as C
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from faspsolver_b200 import api
from faspsolver_b200 import fasp_types as T
from faspsolver_b200 import problems as PB

ROOT = Path(__file__).resolve().parent.parent
REF_INC = Path("/root/reference/base/include")

ABI_BODY = r'''
#include <stdio.h>
#include <stddef.h>
#define P(T) printf(#T " %zu\n", sizeof(T))
#define O(T,m) printf(#T "." #m " %zu\n", offsetof(T,m))
int main(void){
P(dCSRmat);P(dvector);P(ivector);P(dBSRmat);P(ITS_param);P(ILU_param);P(AMG_param);P(ILU_data);P(SWZ_data);
P(AMG_data);P(AMG_data_bsr);P(precond_data);P(precond_data_bsr);P(precond);
O(AMG_param,tol);O(AMG_param,cycle_type);O(AMG_param,smoother);O(AMG_param,relaxation);O(AMG_param,polynomial_degree);
O(AMG_param,coarse_scaling);O(AMG_param,amli_coef);O(AMG_param,strong_threshold);O(AMG_param,ILU_levels);O(AMG_param,theta);
O(AMG_data,A);O(AMG_data,R);O(AMG_data,P);O(AMG_data,b);O(AMG_data,x);O(AMG_data,cfmark);O(AMG_data,ILU_levels);
O(AMG_data,SWZ_levels);O(AMG_data,w);O(AMG_data,weight);
O(AMG_data_bsr,A);O(AMG_data_bsr,R);O(AMG_data_bsr,P);O(AMG_data_bsr,b);O(AMG_data_bsr,x);O(AMG_data_bsr,diaginv);
O(AMG_data_bsr,ILU_levels);O(AMG_data_bsr,A_nk);O(AMG_data_bsr,w);
O(precond_data,maxit);O(precond_data,mgl_data);O(precond_data,A);O(precond_data,w);
O(ITS_param,restart);O(ITS_param,tol);O(ITS_param,abstol);
return 0;}
'''


def _compile_run(tmp_path, name, includes, flags=()):
    src = tmp_path / (name + ".c")
    src.write_text("".join('#include "%s"\n' % i for i in includes) + ABI_BODY)
    exe = tmp_path / name
    subprocess.run(["gcc", *flags, "-I", str(ROOT / "include"), "-I", str(REF_INC), str(src), "-o", str(exe)], check=True)
    return subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout


def test_header_is_plain_c_and_mirrors_ctypes(tmp_path):
    """include/fasp_cuda.h compiles as C99 on its own; its mirror structs and the ctypes mirror agree."""
    out = _compile_run(tmp_path, "mine", ["fasp_cuda.h"], ["-std=c99", "-Wall", "-Werror"])
    sizes = dict(l.split() for l in out.splitlines())
    for name in ("dCSRmat", "dvector", "dBSRmat", "ITS_param", "AMG_param", "AMG_data", "AMG_data_bsr", "precond_data",
                 "precond"):
        assert int(sizes[name]) == C.sizeof(getattr(T, name)), name
    assert int(sizes["AMG_data.w"]) == T.AMG_data.w.offset
    assert int(sizes["AMG_data_bsr.diaginv"]) == T.AMG_data_bsr.diaginv.offset
    assert int(sizes["precond_data.mgl_data"]) == T.precond_data.mgl_data.offset


@pytest.mark.skipif(not REF_INC.exists(), reason="reference headers not present")
@pytest.mark.parametrize("omp", [False, True])
def test_mirror_structs_match_reference_headers(tmp_path, omp):
    """sizeof / offsetof of every struct that crosses the boundary: FASP's own fasp.h vs our mirror,
    for the sequential ABI and (FASP_CUDA_OPENMP_ABI) the OpenMP ABI."""
    ref = _compile_run(tmp_path, "ref", ["fasp.h", "fasp_block.h"], ["-fopenmp"] if omp else [])
    mine = _compile_run(tmp_path, "mine", ["fasp_cuda.h"], ["-DFASP_CUDA_OPENMP_ABI"] if omp else [])
    assert ref == mine


@pytest.mark.skipif(not REF_INC.exists(), reason="reference headers not present")
def test_header_coexists_with_fasp_h(tmp_path):
    """A FASP application includes fasp.h first, then fasp_cuda.h: no redefinitions."""
    src = tmp_path / "both.c"
    src.write_text('#include "fasp.h"\n#include "fasp_functs.h"\n#include "fasp_cuda.h"\n'
                   'int main(void){ return (int)fasp_cuda_abi_check(sizeof(dCSRmat), sizeof(AMG_data), sizeof(AMG_param)) * 0; }\n')
    subprocess.run(["gcc", "-c", "-I", str(ROOT / "include"), "-I", str(REF_INC), str(src), "-o", str(tmp_path / "both.o")],
                   check=True)


def test_library_exports_every_declared_symbol():
    L = api.lib()
    declared = api.declared_symbols()
    assert len(declared) > 60
    missing = [s for s in declared if not hasattr(L, s)]
    assert missing == [], missing
    assert L._missing == []
    # every symbol bound in api.py is declared in the header, and vice versa
    assert set(L._signatures) == set(declared)


def test_abi_check_entry_point():
    L = api.lib()
    assert L.fasp_cuda_abi_check(C.sizeof(T.dCSRmat), C.sizeof(T.AMG_data), C.sizeof(T.AMG_param)) == 0
    assert L.fasp_cuda_abi_check(C.sizeof(T.dCSRmat) + 16, C.sizeof(T.AMG_data), C.sizeof(T.AMG_param)) == T.ERROR_DATA_STRUCTURE
    assert b"ABI mismatch" in L.fasp_cuda_last_error()


def test_options_roundtrip_and_unknown_key():
    L = api.lib()
    assert L.fasp_cuda_set_option(b"strict", 1.0) == 0 and L.fasp_cuda_get_option(b"strict") == 1.0
    assert L.fasp_cuda_set_option(b"strict", 0.0) == 0
    assert L.fasp_cuda_set_option(b"no_such_option", 1.0) == T.ERROR_INPUT_PAR


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point fails loudly (no silent CPU path)."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from faspsolver_b200 import api, fasp_types as T, problems as PB
L = api.lib()
A = PB.poisson5_2d(4)
x = np.ones(A.shape[1]); y = np.zeros(A.shape[0])
st = L.fasp_cuda_blas_dcsr_mxv(A.ptr(), T.as_preal(x), T.as_preal(y))
print(st, L.fasp_cuda_last_error().decode())
assert st < 0 and not y.any()
''' % str(ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no CUDA device" in r.stdout or "CUDA error" in r.stdout


def test_generators_match_scipy_kron():
    import scipy.sparse as sp
    n = 5
    T1 = sp.diags([-1, 2, -1], [-1, 0, 1], shape=(n, n))
    I = sp.identity(n)
    K = sp.kron(sp.kron(I, I), T1) + sp.kron(sp.kron(I, T1), I) + sp.kron(sp.kron(T1, I), I)
    A = PB.poisson7(n, scaled=False).to_scipy()
    assert abs(A - K).max() == 0
    assert np.all(np.diff(PB.poisson7(n).ja.reshape(-1)[:4]) > 0)   # ascending columns
    A27 = PB.poisson27(4)
    assert A27.nnz == (3 * 4 - 2) ** 3 and np.allclose(A27.to_scipy().sum(axis=1).min(), 0.0)
    B, rhs = PB.blockoil7(3)
    assert B.NNZ == PB.poisson7(3).nnz and rhs.size == 3 * 27
    cd = PB.convdiff7(4).to_scipy()
    assert abs(cd - cd.T).max() > 0   # nonsymmetric


def test_fasp_file_readers(data):
    ref_dir = Path("/root/reference/data")
    if not ref_dir.exists():
        pytest.skip("reference data files not present")
    A = PB.read_fasp_csr(ref_dir / "csrmat_FD.dat")
    assert A.shape == (100, 100) and A.nnz == 460
    assert np.array_equal(A.val, data["FD"].val) and np.array_equal(A.ja, data["FD"].ja)
    S = PB.read_fasp_bsr(ref_dir / "bsrmat_SPE01.dat")
    assert (S.ROW, S.nb, S.NNZ) == (302, 3, 1788)


def test_multicolor_rule_matches_greedy_invariants():
    """Host colouring used by the multicolour GS smoother: every colour class is an independent set
    and the classes partition the rows (BlaSparseCSR.c:1687-1770 greedy rule)."""
    L = api.lib()
    # exercised through the oracle-free host function exported for tests
    A = PB.poisson7(6)
    n = A.shape[0]
    ic = (C.c_int * (n + 2))()
    icmap = (C.c_int * n)()
    ncol = L.fasp_cuda_multicolor_host(n, A.ia.ctypes.data_as(T.PINT), A.ja.ctypes.data_as(T.PINT), ic, icmap)
    assert ncol == 2   # 7-point stencil: red-black
    icn, mp = np.array(ic[:ncol + 1]), np.array(icmap[:])
    assert sorted(mp.tolist()) == list(range(n)) and icn[-1] == n
    S = A.to_scipy().tolil()
    for c in range(ncol):
        rows = set(mp[icn[c]:icn[c + 1]].tolist())
        for r in rows:
            assert not (set(S.rows[r]) - {r}) & rows


def test_c_application_links_and_fails_loudly_without_device(c_example_exe):
    """examples/poisson_amg_cuda.c — a FASP application in plain C99 switched to libfasp_cuda — compiles
    against include/fasp_cuda.h alone, links libfasp_cuda + the host FASP, and without a CUDA device
    stops with the library's 'no CPU fallback' message (exit status 2) instead of computing on the CPU."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([str(c_example_exe), "8", "0"], capture_output=True, text=True, env=env)
    assert r.returncode == 2, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stderr


def test_setup_shim_interposes_and_fails_loudly_without_device(tmp_path):
    """libfasp_cuda_setup.so in front of libfasp: the UNMODIFIED fasp_amg_setup_rs reaches fasp_cuda_dcsr_trans
    through FASP's own symbol name; on a machine without a GPU that is a loud FASP-style error, not a CPU fallback."""
    ref_lib = ROOT / "oracle" / "_ref" / "libfasp_seq.so"
    shim = ROOT / "faspsolver_b200" / "lib" / "libfasp_cuda_setup.so"
    if not ref_lib.exists():
        pytest.skip("oracle/_ref/libfasp_seq.so not built")
    assert shim.exists()
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from oracle.ref import RefFasp\n"
        "from faspsolver_b200 import problems as PB, fasp_types as T\n"
        "ref = RefFasp(); A = PB.poisson7(8)\n"
        "mgl = ref.amg_setup(A, ref.amg_param(print_level=0, coarse_dof=20))\n"
        "print('levels', mgl[0].num_levels)\n" % str(ROOT))
    env = dict(os.environ, LD_PRELOAD=str(shim), CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode != 0
    assert "fasp_dcsr_trans on the device failed" in r.stderr and "no CPU fallback" in r.stderr
    assert "levels" not in r.stdout
