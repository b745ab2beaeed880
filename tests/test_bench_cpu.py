"""CPU checks of the measurement harness: the reference arm prints the contract's JSON line (bench.py --impl
reference runs the UNMODIFIED OpenMP FASP through oracle/_ref/fasp_ref_bench; full solves, all host threads even
when a launcher exported OMP_NUM_THREADS=1), and the config-3 script parses its arguments."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "fasp_ref_bench").exists(), reason="oracle/_ref not built")
def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")          # what torchrun exports to its workers
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--n", "32", "--steps", "2",
                        "--warmup", "3"], capture_output=True, text=True, timeout=300, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["value"] > 0 and d["ms_per_step"] == d["value"] and d["steps"] == 2 and d["warmup"] == 3
    assert d["e2e"] == {"value": d["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == d["value"]
    assert cb["cores"] == (len(os.sched_getaffinity(0)) or os.cpu_count())     # not the launcher's single thread
    assert d["config"]["extrapolated"] is False and "full solves" in cb["sample"]
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--n", "32", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=str(ROOT))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_config3_script_arguments():
    sys.path.insert(0, str(ROOT / "scripts"))
    import bench_config3 as C3
    a = C3.parse(["--size", "96", "--steps", "2", "--opt", "overlap=0"])
    assert (a.n, a.steps, a.stencil, a.opt, a.lock) == (96, 2, 27, ["overlap=0"], "")
