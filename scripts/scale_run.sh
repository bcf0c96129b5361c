#!/bin/bash
# Weak-scaling run on ONE box: N = 1, 2, 4, 8 (as many as are visible), c2/c5 training step.  usage: gpurun --gpus 8 -- 'bash scripts/scale_run.sh r02'
R=${1:-r02}
mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
for N in 1 2 4 8; do
  [ $N -gt $n ] && break
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-extra-workloads \
      > gpurun_out/${R}_bench_c2_n1.json 2> gpurun_out/${R}_bench_c2_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${R}_bench_c2_n$N.json 2> gpurun_out/${R}_bench_c2_n$N.err
  fi
  echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${R}_bench_c2_n$N.json") if l.startswith("{")][-1])
    print("N=$N", d["value"], "f/s", d["ms_per_step"], "ms/step  e2e", d["e2e"]["value"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("N=$N: no line", e)
PY
done
