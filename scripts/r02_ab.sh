#!/usr/bin/env bash
# (needs the comparison worktrees: git worktree add _ab/<commit> <commit> && build the library in each; _ab/ is git-ignored)
# A/B on one box: stream priority on/off at HEAD, and the round-1 library (worktree _ab/r01, commit d7c4be1)
set -u
cd "$(dirname "$0")/.."
O=$PWD/gpurun_out/r02ab2
mkdir -p $O
export PYTHONUNBUFFERED=1
run() { # name, env
  env $2 timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --extras 0 > $O/$1.json 2> $O/$1.log; echo "$1 rc=$?"
  python scripts/show_bench.py $O/$1.json 2>/dev/null | sed -n 1,6p
}
run prio1_a FASP_CUDA_STREAM_PRIO=1
run prio0_a FASP_CUDA_STREAM_PRIO=0
run prio1_b FASP_CUDA_STREAM_PRIO=1
run prio0_b FASP_CUDA_STREAM_PRIO=0
