#!/bin/bash
# 8-GPU evidence (one box): default workload at N=8 and N=4, BASELINE config 4 (synth50m) at N=8; config 5 (synth200m) with `all`.
R=${1:-r02}; O=gpurun_out; mkdir -p $O
bash tools/run_dp.sh 8 $O/${R}_bench_8gpu.json --no-extra --windows 3
bash tools/run_dp.sh 4 $O/${R}_bench_4gpu.json --no-extra --windows 3
bash tools/run_dp.sh 8 $O/${R}_bench_synth50m_8gpu.json --no-extra --windows 2 --steps 10 --warmup 3 --workload synth50m
if [ "$2" = "all" ]; then
  bash tools/run_dp.sh 8 $O/${R}_bench_synth200m_8gpu.json --no-extra --windows 2 --steps 4 --warmup 3 --workload synth200m
fi
