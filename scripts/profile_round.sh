#!/bin/bash
# Round profiling recipe (run under gpurun on ONE GPU):  scripts/profile_round.sh r01
#  1. plain bench (must exit 0), 2. ncu launch list of one steady-state step, 3. ncu --set full of the top kernels
#  (each only after the same command has exited 0 without ncu).  Reduce here with scripts/ncu_summary.py.
set -u
R=${1:-r01}
O=gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-steps 0"
$BENCH > $O/${R}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2172 -c 724 --csv --log-file $O/${R}_launches.csv \
    $BENCH > $O/${R}_launches_run.log 2>&1
tail -2 $O/${R}_plain.log | cut -c1-300
for c in attn_frame:attention_fwd attn_frame32_m:attention_fwd attn_global8x4_m:attention_fwd gemm_t_fc1:gemm_kernel \
         gemm_t_fc2:gemm_kernel gemm_t_qkvrope:gemm_kernel gemm_t_proj:gemm_kernel; do
  CASE=${c%%:*}; K=${c##*:}
  python scripts/gpu_bringup.py $CASE > $O/${R}_${CASE}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $O/${R}_${CASE} \
      python scripts/gpu_bringup.py $CASE > $O/${R}_${CASE}_ncu.log 2>&1
  tail -1 $O/${R}_${CASE}_plain.log | cut -c1-160
done
ls -la $O | grep ${R}_
