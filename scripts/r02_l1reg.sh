#!/usr/bin/env bash
# (needs the comparison worktrees: git worktree add _ab/<commit> <commit> && build the library in each; _ab/ is git-ignored)
# why did the L1-sweep kernel get slower in 25d178e with the same SASS? standalone timing + one ncu capture per build
set -u
cd "$(dirname "$0")/.."
O=$PWD/gpurun_out/r02l1
mkdir -p $O
export PYTHONUNBUFFERED=1
for c in ea62af1 HEAD; do
  D=$PWD; [ $c != HEAD ] && D=$PWD/_ab/$c
  (cd $D && timeout 300 python scripts/level_sweep.py --n 256 --levels 0,1 --ops A --kernels 0,2,11 --reps 30 > $O/sweep_$c.jsonl 2> $O/sweep_$c.log); echo "$c rc=$?"
  cut -c1-220 $O/sweep_$c.jsonl
  (cd $D && timeout 400 ncu --set full --clock-control none -k regex:csr_pipe_kernel -s 7 -c 1 -o $O/l1_level1_$c -f \
     python scripts/level_sweep.py --n 256 --levels 1 --ops A --kernels 11 --warm 5 --reps 5 > $O/ncu_$c.log 2>&1); echo "ncu $c rc=$?"
done
ls -la $O
