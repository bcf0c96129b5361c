#!/bin/bash
# Round profiling recipe (run under gpurun on ONE GPU):  scripts/profile_round.sh r02
#  1. plain bench (must exit 0), 2. ncu launch list of one steady-state c2 step, 3. ncu --set full of the top kernels
#  (each only after the same command has exited 0 without ncu).  Reduce here with scripts/ncu_summary.py.
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extra-workloads --profile-steps 0"
# launches per step: from a plain run's own count
$BENCH > $O/${R}_plain.log 2>&1 || { echo "plain bench failed"; tail -5 $O/${R}_plain.log; exit 1; }
L=$(python -c "import json;print(json.loads(open('$O/${R}_plain.log').read().strip().splitlines()[-1])['gpu_launches_per_step'])")
echo "launches per step: $L"
# skip the 3 warm-up steps + the first timed step, capture the second timed step
ncu --metrics gpu__time_duration.sum --clock-control none -s $((4 * L)) -c $L --csv --log-file $O/${R}_launches.csv \
    $BENCH > $O/${R}_launches_run.log 2>&1
tail -1 $O/${R}_plain.log | cut -c1-200
for c in "gemm_fc1_gelu:gemm_kernel:gpu_bringup.py gemm_t_fc1_noaux" "gemm_fc2:gemm_kernel:gpu_bringup.py gemm_t_fc2_bf16" \
         "gemm_proj:gemm_kernel:gpu_bringup.py gemm_t_proj_bf16" \
         "attn_bounded_frame8:attention_fwd_bounded:attn_case.py 8 1374 1" \
         "attn_bounded_global8x4:attention_fwd_bounded:attn_case.py 4 10992 1" \
         "attn_checked_frame8:attention_fwd_wide:attn_case.py 8 1374 0" \
         "pattn_fwd:perceiver_attn_fwd:pattn_case.py" "pattn_bwd:perceiver_attn_bwd:pattn_case.py"; do
  CASE=${c%%:*}; REST=${c#*:}; K=${REST%%:*}; CMD=${REST#*:}
  python scripts/$CMD > $O/${R}_${CASE}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $O/${R}_${CASE} \
      python scripts/$CMD > $O/${R}_${CASE}_ncu.log 2>&1
  tail -1 $O/${R}_${CASE}_plain.log | cut -c1-160
  # reduce on the box (gpurun merges at most 64 MiB back): text summary always, the report itself only for the first case
  if [ -f $O/${R}_${CASE}.ncu-rep ]; then
    python scripts/ncu_summary.py report $O/${R}_${CASE}.ncu-rep $O/${R}_${CASE}.txt > /dev/null 2>&1
    [ "$CASE" = "gemm_fc1_gelu" ] || rm -f $O/${R}_${CASE}.ncu-rep
  fi
done
python scripts/ncu_summary.py launches $O/${R}_launches.csv $O/${R}_launches_c2.txt > /dev/null 2>&1
ls -la $O | grep ${R}_
