#!/usr/bin/env bash
# round-2 GPU call 2 (N GPUs): the 2-rank parity test, then the solve-time sweep (overlap on/off, agglomeration)
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out/r02b_N$N
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L > $O/gpus.txt 2>&1
if [ "$N" = "2" ]; then
  echo "== dist test"; date
  timeout 600 python -m pytest tests/test_dist_gpu.py -q -x > $O/pytest_dist.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_dist.log
fi
echo "== sweep"; date
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
   scripts/dist_sweep.py --size 256 --agg-list "${AGG:-8000,60000}" --opts "${OPTS:-overlap=1;overlap=0;overlap=1,overlap_min_rows=4096}" --profile \
   > $O/sweep.log 2>&1; echo "rc=$?"
grep -v "^\*\*\|OMP_NUM\|NCCL version\|^W1" $O/sweep.log | tail -60
echo "== bench"; date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
   bench.py --gpus $N --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.log; echo "rc=$?"
tail -c 1500 $O/bench.json; grep "\[bench\]" $O/bench.log | tail
date
