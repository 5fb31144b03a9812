#!/bin/bash
# Round-end evidence on one B200: ncu launch list, full-set capture of the top kernels, bench lines.
# Usage (on the GPU box, from the repo root): bash tools/profile_round.sh rNN
R=${1:-r01}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
# (1) every launch of two steps, serialised, cold cache: compare SHARES with the live event profile
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_ncu_launches.csv $B > $O/ncu1.log 2>&1
# (2) full set for the HBM-bound / serial kernels of one step
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"adam_sweep|gather_hist|scatter_hist|lstm_fwd_tc|gru_fwd_tc|lstm_bwd_tc|gru_bwd_tc|pool_bwd_kernel|h0_reduce_v4|mulrow_bwd_v4" \
  --launch-skip 19 --launch-count 19 -o $O/${R}_full_misc -f $B > $O/ncu2.log 2>&1
# (3) the tcgen05 kernels of one step: speed-of-light, memory, occupancy, warp states
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
  --section WarpStateStats --section SchedulerStats --clock-control none -k regex:"tc_gemm_kernel|tc_dw_kernel|tc_dw_group_kernel" \
  --launch-skip 42 --launch-count 42 -o $O/${R}_tc -f $B > $O/ncu3.log 2>&1   # 34 GEMM + 5 single + 3 grouped weight-gradient launches per step
for f in $O/ncu1.log $O/ncu2.log $O/ncu3.log; do tail -n 2 $f | cut -c1-200; done
ls -la $O/*.ncu-rep
