#!/bin/bash
# A/B of two builds of the library on the same box: default vs $1 (path of the variant .so), c2 + c3, interleaved twice
mkdir -p gpurun_out
for rep in 1 2; do
  for v in default variant; do
    if [ $v = variant ]; then export VQ_B200_LIB=$PWD/$1; else unset VQ_B200_LIB; fi
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out gpurun_out/ab_${v}_${rep}.txt > gpurun_out/ab_${v}_${rep}.json 2>/dev/null
    python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${v}_${rep}.json").read().strip().splitlines()[-1])
print("$v $rep c2", d["value"], d["ms_per_step"], " ".join(f"{k} {w['value']} {w['ms_per_step']}" for k,w in d["workloads"].items()))
PY
    grep -E "gemm_gelu 10992" gpurun_out/ab_${v}_${rep}.txt
  done
done
