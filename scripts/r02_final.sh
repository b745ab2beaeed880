#!/usr/bin/env bash
# round-2 final one-GPU call: the whole -m gpu suite, the default bench (with extras incl. the N = 1 configs[2] proxy),
# the reference arm, the vector-kernel depth sweep, the ncu --set full captures of the level-0 kernels (roofline.traffic)
# and the ncu launch list of the bench command. Every step bounded and independent.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r02z
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "== pytest"; date
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest.log | tail -10
echo "== bench"; date
timeout 1200 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.log; echo "bench rc=$?"
python scripts/show_bench.py $O/bench.json 2>/dev/null | head -24
grep "\[bench\]\|config3" $O/bench.log | tail -12
echo "== reference arm"; date
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.log; echo "ref rc=$?"
cut -c1-300 $O/bench_ref.json
echo "== vector kernel depth"; date
timeout 300 python scripts/level_sweep.py --n 256 --levels 3,4,5,6,7,8 --ops A,R --kernels 0,2,11 --reps 30 --optsets "vec_u=4;vec_u=8" > $O/sweep_vec_u.jsonl 2> $O/sweep_vec_u.log
python - <<'PY'
import json
t={}
for l in open("gpurun_out/r02z/sweep_vec_u.jsonl"):
    x=json.loads(l); t.setdefault((x["level"],x["op"],x["kernel"],x["nnz_per_row"]),{})[x["opts"]]=x["us"]
for k,v in t.items(): print(k,v)
PY
echo "== ncu full, level-0 kernels"; date
timeout 400 ncu --set full --clock-control none -k regex:csr_pipe_kernel -c 9 -o $O/level0_kernels -f \
   python scripts/level_sweep.py --n 256 --levels 0 --ops A --kernels 0,2,11 --warm 2 --reps 1 > $O/ncu_full.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py $O/level0_kernels.ncu-rep > $O/ncu_level0_summary.txt 2>&1; grep -c 'Kernel Name' $O/ncu_level0_summary.txt; grep 'Kernel Name\|duration\|dram__bytes' $O/ncu_level0_summary.txt | tail -12
echo "== ncu launch list"; date
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/launches.csv \
   python bench.py --steps 1 --warmup 3 --extras 0 --no-cpu-baseline > $O/bench_under_ncu.json 2> $O/bench_under_ncu.log; echo "ncu rc=$?"
python scripts/summarize_launches.py $O/launches.csv > $O/launches_summary.csv 2>&1; head -16 $O/launches_summary.csv
gzip -f $O/launches.csv
date
