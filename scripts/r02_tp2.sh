#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=$PWD/gpurun_out/r02tp2
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== bsr 272"; timeout 400 python scripts/bsr_sweep.py --n 272 --quick > $O/bsr272.txt 2>&1; tail -4 $O/bsr272.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.log; echo "bench rc=$?"
python scripts/show_bench.py $O/bench.json 2>/dev/null | sed -n 1,60p
timeout 600 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest.log 2>&1; tail -3 $O/pytest.log
