#!/usr/bin/env bash
# two-phase gather rounds per mode: stand-alone kernel times with the option off / default / on for every mode
set -u
cd "$(dirname "$0")/.."
O=$PWD/gpurun_out/r02tp
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 500 python scripts/level_sweep.py --n 256 --levels 0,1,2 --ops A,P,R --kernels 0,2,10,11 --reps 30 \
   --optsets "two_phase_mask=0;two_phase_mask=255" > $O/sweep.jsonl 2> $O/sweep.log; echo "rc=$?"
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r02tp/sweep.jsonl")]
key=lambda x:(x["level"],x["op"],x["kernel"])
t={}
for x in d: t.setdefault(key(x),{})[x["opts"]]=x["us"]
for k,v in t.items(): print(k, v)
PY
echo "== bsr 272"; timeout 400 python scripts/bsr_sweep.py --n 272 --quick > $O/bsr272.txt 2>&1; tail -4 $O/bsr272.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extras 0 > $O/bench.json 2> $O/bench.log; echo "bench rc=$?"
python scripts/show_bench.py $O/bench.json 2>/dev/null | sed -n 1,10p
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 > $O/pytest.log 2>&1; tail -3 $O/pytest.log
