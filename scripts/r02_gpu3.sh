#!/usr/bin/env bash
# round-2 GPU call 3 (one B200): full -m gpu suite, setup timing with the shim, kernel sweeps, bench
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r02c
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest"; date
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest.log | tail -30
echo "== setup timing"; date
timeout 300 python scripts/setup_timing.py 256 > $O/setup_cpu.json 2> $O/setup_cpu.log; tail -c 400 $O/setup_cpu.json
LD_PRELOAD=$PWD/faspsolver_b200/lib/libfasp_cuda_setup.so timeout 300 python scripts/setup_timing.py 256 > $O/setup_shim.json 2> $O/setup_shim.log; tail -c 400 $O/setup_shim.json; tail -3 $O/setup_shim.log
echo "== level sweeps"; date
timeout 600 python scripts/level_sweep.py --n 256 --levels 0,1,2 --ops P,R --reps 30 --reset "pipe_tpb=128,pipe_cap_mult=16,pipe_stages=2" \
   --optsets "pipe_tpb=128;pipe_tpb=256;pipe_tpb=256,pipe_cap_mult=8;pipe_tpb=128,pipe_stages=3;pipe_tpb=64" > $O/sweep_PR.txt 2>&1; tail -40 $O/sweep_PR.txt
timeout 600 python scripts/level_sweep.py --n 256 --levels 1,2,3,4,5 --ops A --kernels 11 --reps 30 --reset "vec_min_avg=48,pipe_stages=2,rowwise_max=64" \
   --optsets "vec_min_avg=48;vec_min_avg=100;vec_min_avg=250;vec_min_avg=500;vec_min_avg=250,rowwise_max=16;vec_min_avg=48,pipe_stages=3" > $O/sweep_A.txt 2>&1; tail -40 $O/sweep_A.txt
echo "== bsr 272"; date
timeout 400 python scripts/bsr_sweep.py --n 272 --quick > $O/bsr272.txt 2>&1; tail -8 $O/bsr272.txt
echo "== bench"; date
timeout 1200 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.log; echo "bench rc=$?"
python scripts/show_bench.py $O/bench.json 2>/dev/null | head -60 || tail -c 1500 $O/bench.json
date
