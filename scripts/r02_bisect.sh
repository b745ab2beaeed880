#!/usr/bin/env bash
# (needs the comparison worktrees: git worktree add _ab/<commit> <commit> && build the library in each; _ab/ is git-ignored)
set -u
cd "$(dirname "$0")/.."
O=$PWD/gpurun_out/r02bisect
mkdir -p $O
export PYTHONUNBUFFERED=1
for c in ea62af1; do
  (cd _ab/$c && timeout 300 python bench.py --steps 5 --warmup 3 --extras 0 > $O/$c.json 2> $O/$c.log); echo "$c rc=$?"
  python scripts/show_bench.py $O/$c.json 2>/dev/null | sed -n 1,5p
done
