import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def data():
    """FASP's shipped FD / FE / SPE01 problems (tests/golden/fasp_data.npz)."""
    from faspsolver_b200.fasp_types import BSR, CSR
    z = np.load(GOLDEN / "fasp_data.npz")
    d = {}
    for k in ("FD", "FE"):
        n = z[k + "_ia"].size - 1
        d[k] = CSR(n, n, z[k + "_ia"], z[k + "_ja"], z[k + "_val"])
        d[k + "_b"] = z[k + "_b"]
        d[k + "_sol"] = z[k + "_sol"]
    ROW, COL, nb = (int(v) for v in z["SPE_dims"])
    d["SPE"] = BSR(ROW, COL, nb, z["SPE_ia"], z["SPE_ja"], z["SPE_val"])
    d["SPE_b"] = z["SPE_b"]
    return d


@pytest.fixture(scope="session")
def golden_vectors():
    return np.load(GOLDEN / "oracle_vectors.npz")


@pytest.fixture(scope="session")
def golden_answers():
    return json.loads((GOLDEN / "oracle_answers.json").read_text())


@pytest.fixture(scope="session")
def ref():
    """The unmodified sequential reference (oracle/_ref/libfasp_seq.so)."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref/libfasp_seq.so not built (run oracle/build_ref.sh where /root/reference exists)")
    return R.RefFasp()


@pytest.fixture(scope="session")
def L():
    """libfasp_cuda.so through the C ABI. No fallback: a missing library is an error."""
    from faspsolver_b200 import api
    return api.lib()


@pytest.fixture(scope="session")
def gpu(L):
    st = L.fasp_cuda_init(int(os.environ.get("LOCAL_RANK", "0")))
    if st != 0:
        pytest.fail("libfasp_cuda could not initialise a CUDA device: " + L.fasp_cuda_last_error().decode())
    return L


@pytest.fixture(scope="session")
def c_example_exe(tmp_path_factory):
    """examples/poisson_amg_cuda.c built with plain gcc against include/fasp_cuda.h, libfasp_cuda and
    the host FASP (the unmodified reference build under oracle/_ref)."""
    import subprocess
    exe = tmp_path_factory.mktemp("c_example") / "poisson_amg_cuda"
    ref_dir = ROOT / "oracle" / "_ref"
    if not (ref_dir / "libfasp_seq.so").exists():
        pytest.skip("oracle/_ref/libfasp_seq.so not built")
    lib_dir = ROOT / "faspsolver_b200" / "lib"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"),
                    str(ROOT / "examples" / "poisson_amg_cuda.c"), "-L", str(lib_dir), "-lfasp_cuda",
                    "-L", str(ref_dir), "-l:libfasp_seq.so", "-lm", "-Wl,-rpath," + str(lib_dir),
                    "-Wl,-rpath," + str(ref_dir), "-o", str(exe)], check=True)
    return exe
