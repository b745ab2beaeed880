#!/usr/bin/env bash
# 2-GPU validation of the final code: the two-rank parity tests, repeat-solve reproducibility, the bench line
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r02i_N2
mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_dist_gpu.py -q > $O/pytest_dist.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_dist.log
timeout 300 $TR --master-port 29544 scripts/dist_repro.py > $O/repro.txt 2> $O/repro.err; grep -v "^W1\|OMP_NUM\|^\*\*\|NCCL" $O/repro.txt | awk '{print $1, $NF}' | sort | uniq -c
FASP_BENCH_C3N=0 timeout 400 $TR --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.log; echo "bench rc=$?"
cut -c1-300 $O/bench.json; grep "parity" $O/bench.log
