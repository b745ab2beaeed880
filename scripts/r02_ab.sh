#!/usr/bin/env bash
# A/B on one box: the round-1 library (worktree _ab/r01, commit d7c4be1) against HEAD, same bench, interleaved
set -u
cd "$(dirname "$0")/.."
O=$PWD/gpurun_out/r02ab
mkdir -p $O
export PYTHONUNBUFFERED=1
for i in 1 2; do
  (cd _ab/r01 && timeout 400 python bench.py --steps 10 --warmup 3 > $O/old_$i.json 2> $O/old_$i.log); echo "old $i rc=$?"
  python scripts/show_bench.py $O/old_$i.json | head -8
  timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extras 0 > $O/new_$i.json 2> $O/new_$i.log; echo "new $i rc=$?"
  python scripts/show_bench.py $O/new_$i.json | head -8
done
