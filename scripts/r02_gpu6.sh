#!/usr/bin/env bash
# N-GPU call: the two-rank parity tests and configs[2] through the slab path (27-pt, C3N^3)
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
C3N=${C3N:-256}
O=gpurun_out/r02f_N$N
mkdir -p $O
export PYTHONUNBUFFERED=1
nproc > $O/host.txt; free -g >> $O/host.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  echo "== dist tests"; date
  timeout 900 python -m pytest tests/test_dist_gpu.py -q > $O/pytest_dist.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_dist.log
fi
echo "== config3 slab path n=$C3N"; date
timeout ${C3T:-1200} $TR --master-port 29543 scripts/bench_config3.py --size $C3N --steps 5 --warmup 3 ${C3ARGS:-} > $O/config3_$C3N.json 2> $O/config3_$C3N.log; echo "rc=$?"
cut -c1-3000 $O/config3_$C3N.json; grep "config3\|slab setup\|Error\|error\|memory" $O/config3_$C3N.log | tail -30
free -g | head -2
if [ "${WITH_BENCH:-0}" = "1" ]; then
  echo "== bench 7-pt 256^3"; date
  timeout ${BENCHT:-600} $TR --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.log; echo "rc=$?"
  cut -c1-400 $O/bench.json; grep "\[bench\]" $O/bench.log | tail -5
fi
date
