#!/bin/bash
# round-end check on one GPU: tests, smoke, both bench arms (outputs under gpurun_out/)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.log; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_final.json | cut -c1-1500
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.log; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref_final.json
