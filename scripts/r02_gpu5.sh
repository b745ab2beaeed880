#!/usr/bin/env bash
# round-2 multi-GPU call (N GPUs): the N-rank parity tests (N = 2), solve-time sweep on 7-pt 256^3, the bench line,
# and configs[2] through the slab path (27-pt, C3N^3).
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
C3N=${C3N:-256}
O=gpurun_out/r02e_N$N
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L > $O/gpus.txt 2>&1; nproc >> $O/gpus.txt; free -g >> $O/gpus.txt; nvidia-smi topo -m >> $O/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  echo "== dist tests"; date
  timeout 900 python -m pytest tests/test_dist_gpu.py -q > $O/pytest_dist.log 2>&1; echo "rc=$?"; tail -25 $O/pytest_dist.log
fi
if [ "${SKIP_SWEEP:-0}" != "1" ]; then
echo "== sweep"; date
timeout 900 $TR --master-port 29541 scripts/dist_sweep.py --size 256 --agg-list "${AGG:-8000,60000}" \
   --opts "${OPTS:-overlap=1;overlap=0}" --profile > $O/sweep.log 2>&1; echo "rc=$?"
grep -v "^\*\*\|OMP_NUM\|NCCL version\|^W1" $O/sweep.log | tail -40
fi
echo "== bench"; date
timeout 600 $TR --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.log; echo "rc=$?"
tail -c 1800 $O/bench.json; grep "\[bench\]" $O/bench.log | tail
echo "== config3 slab path n=$C3N"; date
timeout ${C3T:-900} $TR --master-port 29543 scripts/bench_config3.py --size $C3N --steps 5 --warmup 3 > $O/config3_$C3N.json 2> $O/config3_$C3N.log; echo "rc=$?"
cut -c1-2500 $O/config3_$C3N.json; grep "config3\|slab setup\|Error\|error" $O/config3_$C3N.log | tail -25
free -g | head -2
date
