"""Build libfasp_cuda.so (sm_100a) in-tree with nvcc, and the oracle's C restatement with gcc.

The shared library lands in faspsolver_b200/lib/ (git-ignored, but it travels to the GPU box
with the repository snapshot). No JIT cache, no torch extension machinery: one nvcc call per
.cu file (parallel), one link.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "build"
LIB = LIBDIR / "libfasp_cuda.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-diag-suppress", "550",
    f"-I{ROOT / 'include'}", f"-I{CSRC}",
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _stamp() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h"))
                    + list((PKG / "cshim").glob("*.c"))):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(map(str, cmd)), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu for sm_100a and link libfasp_cuda.so. Returns the library path."""
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    stamp_file = LIBDIR / "libfasp_cuda.stamp"
    stamp = _stamp()
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB
    if not Path(NVCC).exists():
        if LIB.exists():
            return LIB  # GPU box without toolkit changes: keep the prebuilt library
        raise RuntimeError("nvcc not found at %s and no prebuilt %s" % (NVCC, LIB))
    srcs = _sources()
    objs = [OBJDIR / (s.stem + ".o") for s in srcs]

    def one(pair):
        s, o = pair
        extra = ["-Xptxas", "-v"] if verbose else []
        return _run([NVCC, *NVCC_FLAGS, *extra, "-c", str(s), "-o", str(o)])

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        outs = list(ex.map(one, zip(srcs, objs)))
    if verbose:
        print("\n".join(outs))
    _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB),
          *map(str, objs), "-ldl"])
    # the interposition shim (plain C): FASP's own fasp_dcsr_trans / fasp_blas_dcsr_rap, forwarded to the device
    _run(["gcc", "-O2", "-std=c99", "-Wall", "-fPIC", "-shared", f"-I{ROOT / 'include'}",
          str(PKG / "cshim" / "interpose.c"), "-o", str(LIBDIR / "libfasp_cuda_setup.so"),
          f"-L{LIBDIR}", "-lfasp_cuda", "-Wl,-rpath,$ORIGIN"])
    stamp_file.write_text(stamp)
    return LIB


def build_oracle(force: bool = False) -> Path:
    """gcc-compile oracle/fasp_oracle.c (the CPU restatement; test infrastructure only)."""
    src = ROOT / "oracle" / "fasp_oracle.c"
    out = ROOT / "oracle" / "_ref" / "libfasp_oracle.so"
    out.parent.mkdir(exist_ok=True)
    if src.exists() and (force or not out.exists() or out.stat().st_mtime < src.stat().st_mtime):
        _run(["gcc", "-O3", "-std=gnu99", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", str(out), str(src), "-lm"])
    return out


def build_reference() -> None:
    """Build the unmodified reference into oracle/_ref when /root/reference is present."""
    script = ROOT / "oracle" / "build_ref.sh"
    if script.exists():
        _run(["bash", str(script)])


if __name__ == "__main__":
    lib = build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(lib)
