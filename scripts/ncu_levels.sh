#!/bin/bash
# ncu --set full capture of the level-1 / level-2 smoother kernels of the 256^3 hierarchy
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:csr_pipe_kernel|csr_vector_kernel' -c 10 \
    -o gpurun_out/r01_lvl12_256 -f python scripts/level_sweep.py --n 256 --levels 1,2 --ops A --kernels 11 \
    --reps 3 --warm 2 "$@" > gpurun_out/ncu_lvl12.log 2>&1
tail -5 gpurun_out/ncu_lvl12.log
