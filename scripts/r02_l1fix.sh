#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
O=$PWD/gpurun_out/r02l1fix
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 400 python scripts/level_sweep.py --n 256 --levels 0,1,2 --ops A,P,R --kernels 0,2,11 --reps 30 > $O/sweep.jsonl 2> $O/sweep.log; echo "rc=$?"
cut -c1-40,100-260 $O/sweep.jsonl
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extras 0 > $O/bench.json 2> $O/bench.log; echo "bench rc=$?"
python scripts/show_bench.py $O/bench.json 2>/dev/null | sed -n 1,20p
timeout 300 python -m pytest tests -m gpu -q -x --timeout 600 -k "spmv or smoother or cycle" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
