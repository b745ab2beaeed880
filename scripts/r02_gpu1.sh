#!/usr/bin/env bash
# round-2 GPU call 1 (one B200): the whole -m gpu suite, the default bench (with extras), the reference arm,
# a BSR kernel-shape sweep, and the ncu launch list of the bench command. Each step is bounded and independent.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r02a
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L > $O/gpus.txt 2>&1
free -g > $O/mem.txt 2>&1; nproc >> $O/mem.txt; df -h /dev/shm /tmp >> $O/mem.txt 2>&1
echo "== pytest" ; date
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -5 $O/pytest.log
echo "== pytest (rest, no -x) if failed"
if ! grep -q " passed" $O/pytest.log || grep -q "failed" $O/pytest.log; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_all.log 2>&1; echo "pytest_all rc=$?" | tee -a $O/pytest_all.log
  grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_all.log | tail -30
fi
echo "== bench" ; date
timeout 1200 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.log; echo "bench rc=$?"
tail -c 600 $O/bench.json; tail -25 $O/bench.log
echo "== bsr sweep" ; date
timeout 600 python scripts/bsr_sweep.py --n 160 > $O/bsr_sweep.txt 2>&1; tail -20 $O/bsr_sweep.txt
echo "== reference arm" ; date
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.log; echo "ref rc=$?"
cat $O/bench_ref.json | cut -c1-900
echo "== ncu launch list" ; date
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv \
   python bench.py --steps 1 --warmup 3 --extras 0 --no-cpu-baseline > $O/bench_under_ncu.json 2> $O/bench_under_ncu.log; echo "ncu rc=$?"
python scripts/summarize_launches.py $O/launches.csv > $O/launches_summary.csv 2>&1; head -30 $O/launches_summary.csv
rm -f $O/launches.csv.tmp; gzip -f $O/launches.csv
date
