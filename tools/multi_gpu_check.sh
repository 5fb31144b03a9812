#!/bin/bash
# Multi-GPU check: bash tools/multi_gpu_check.sh N rNN  (2-rank parity tests when N >= 2, then bench lines at N)
N=${1:-2}; R=${2:-r02}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "data_parallel or two_ranks" > $O/${R}_pytest_2gpu.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" $O/${R}_pytest_2gpu.log | tail -6
bash tools/run_dp.sh $N $O/${R}_bench_${N}gpu.json --no-extra --windows 3
bash tools/run_dp.sh $N $O/${R}_bench_${N}gpu_replicated.json --no-extra --windows 3 --dp replicated
