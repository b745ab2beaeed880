#!/usr/bin/env bash
# last one-GPU call of round 2: the whole -m gpu suite and the default bench line (with extras)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r02zz
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest.log | tail -10
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.log; echo "bench rc=$?"
python scripts/show_bench.py $O/bench.json 2>/dev/null | head -24
grep "\[bench\]\|config3" $O/bench.log | tail -12
python -c "
import __graft_entry__ as g
g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
