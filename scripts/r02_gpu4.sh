#!/usr/bin/env bash
# round-2 GPU call 4 (one B200): full -m gpu suite, bench (+extras), reference arm, setup timing with the shim,
# kernel sweeps, ncu launch list of the bench command. Every step bounded and independent.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r02d
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L > $O/gpus.txt 2>&1; nproc >> $O/gpus.txt; free -g >> $O/gpus.txt
echo "== pytest"; date
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=15 > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest.log | tail -30
echo "== bench"; date
timeout 1200 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.log; echo "bench rc=$?"
python scripts/show_bench.py $O/bench.json 2>/dev/null | head -70 || tail -c 1500 $O/bench.json
grep "\[bench\]" $O/bench.log | tail -20
echo "== reference arm"; date
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.log; echo "ref rc=$?"
cut -c1-1200 $O/bench_ref.json
echo "== setup timing"; date
timeout 300 python scripts/setup_timing.py 256 > $O/setup_cpu.json 2> $O/setup_cpu.log; tail -c 400 $O/setup_cpu.json
LD_PRELOAD=$PWD/faspsolver_b200/lib/libfasp_cuda_setup.so timeout 300 python scripts/setup_timing.py 256 > $O/setup_shim.json 2> $O/setup_shim.log; tail -c 400 $O/setup_shim.json; tail -3 $O/setup_shim.log
echo "== level sweeps"; date
timeout 400 python scripts/level_sweep.py --n 256 --levels 0,1,2 --ops P,R --reps 30 --reset "pipe_tpb=128,pipe_cap_mult=16,pipe_stages=2" \
   --optsets "pipe_tpb=128;pipe_tpb=256;pipe_tpb=256,pipe_cap_mult=8;pipe_tpb=128,pipe_stages=3;pipe_tpb=64" > $O/sweep_PR.txt 2>&1; tail -40 $O/sweep_PR.txt
timeout 400 python scripts/level_sweep.py --n 256 --levels 1,2,3,4,5 --ops A --kernels 11 --reps 30 --reset "vec_min_avg=48,pipe_stages=2,rowwise_max=64" \
   --optsets "vec_min_avg=48;vec_min_avg=100;vec_min_avg=250;vec_min_avg=500;vec_min_avg=250,rowwise_max=16;vec_min_avg=48,pipe_stages=3" > $O/sweep_A.txt 2>&1; tail -40 $O/sweep_A.txt
echo "== bsr 272"; date
timeout 400 python scripts/bsr_sweep.py --n 272 --quick > $O/bsr272.txt 2>&1; tail -8 $O/bsr272.txt
echo "== ncu launch list"; date
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv \
   python bench.py --steps 1 --warmup 3 --extras 0 --no-cpu-baseline > $O/bench_under_ncu.json 2> $O/bench_under_ncu.log; echo "ncu rc=$?"
python scripts/summarize_launches.py $O/launches.csv > $O/launches_summary.csv 2>&1; head -30 $O/launches_summary.csv
gzip -f $O/launches.csv
date
