#!/bin/bash
# ncu launch list of the solve kernels of `python bench.py --steps 1` (durations only; compare shares)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout ${1:-240} ncu --metrics gpu__time_duration.sum --clock-control none \
    -k 'regex:csr_|k_axpby|k_copy|k_dense_gemv|k_dot|k_mul|k_pcg|k_reduce|k_scale_div|k_set|k_zero' -c 6000 --csv \
    --log-file gpurun_out/r01_launches_v2.csv python bench.py --steps 1 --warmup 1 --cpu-sample-iters 1 \
    > gpurun_out/bench_under_ncu_v2.json 2> gpurun_out/bench_under_ncu_v2.log
echo "rc=$?"; wc -l gpurun_out/r01_launches_v2.csv
