#!/usr/bin/env bash
# quick 2-GPU validation of the refactored config-3 script and of the bench line with the in-process proxy
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/r02g
mkdir -p $O
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
FASP_BENCH_N=96 FASP_BENCH_C3N=64 timeout 400 $TR --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.log; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02g/bench.json"))
print(d["value"], d["parity"], list(d.get("extra",{}).keys()))
e=d.get("extra",{}).get("config3_proxy",{})
print({k:e.get(k) for k in ("value","n_gpus","unavailable","error")}, e.get("config",{}).get("iterations"))
PY
tail -4 $O/bench.log
