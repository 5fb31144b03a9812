#!/bin/bash
# Round-end evidence on one B200: ncu launch list, full-set capture of the top kernels, bench lines.
# Usage (on the GPU box, from the repo root): bash tools/profile_round.sh rNN
# Launch counts per step (81 in total at the end of round 2): 23 stream / serial / cooperative kernels matched by the
# first filter, TC launches (tc_gemm, tc_dw, tc_dw_group) matched by the second (count below).
R=${1:-r02}
O=gpurun_out
mkdir -p $O
export CLSR_NO_GRAPH=1   # profile the kernels as individual launches
B="python bench.py --steps 1 --warmup 1 --windows 1 --no-cpu-baseline --no-extra"
# (1) every launch of ~4 steps, serialised, cold cache: compare SHARES with the live event profile
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/${R}_ncu_launches.csv $B > $O/ncu1.log 2>&1
# (2) full set for the HBM-bound / serial kernels of one step (the second step of the run)
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"adam_sweep|gather_hist|scatter_hist|lstm_fwd_tc|gru_fwd_tc|lstm_bwd_tc|gru_bwd_tc|pool_bwd_kernel|pool_fwd_kernel|h0_reduce_v4|mulrow_bwd_v4|catmul_bwd|mark_unique|mlp_fwd_coop|mlp_bwd_coop" \
  --launch-skip 23 --launch-count 23 -o $O/${R}_full_misc -f $B > $O/ncu2.log 2>&1
# (3) the tcgen05 kernels of one step: speed of light, memory, occupancy, warp states, tensor pipe
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
  --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis \
  --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"tc_gemm_kernel|tc_dw_kernel|tc_dw_group_kernel" \
  --launch-skip 28 --launch-count 28 -o $O/${R}_tc -f $B > $O/ncu3.log 2>&1
for f in $O/ncu1.log $O/ncu2.log $O/ncu3.log; do tail -n 2 $f | cut -c1-200; done
ls -la $O/*.ncu-rep
