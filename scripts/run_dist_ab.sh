#!/bin/bash
# A/B of the ghost-exchange protocols on N GPUs: tests, then one hierarchy and several option sets
N=${1:-2}; SIZE=${2:-256}; AGG=${3:-8000}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
[ -z "$SKIP_TESTS" ] && timeout 300 python -m pytest tests/test_dist_gpu.py -q 2>&1 | tail -5
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    scripts/dist_profile.py --size $SIZE --agg-rows $AGG --opts "${4:-vec_min_avg=48}" 2>&1 | grep -v "^\*\*\|OMP_NUM\|NCCL version" | tail -40
