#!/bin/bash
# One gpurun call: the GPU test suite, a short N = 1 bench (all legs) and, when 2+ GPUs are visible, the N = 2 bench.
# usage: gpurun [--gpus 2] -- 'bash scripts/gpu_check.sh <tag> [pytest-args]'
tag=${1:-chk}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 --durations=15 -p no:cacheprovider "$@" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/${tag}_kernels_c2.txt \
  > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/${tag}_bench_c2.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_c2.json").read().strip().splitlines()[-1])
    print("c2", d["value"], "f/s", d["ms_per_step"], "ms e2e", d["e2e"]["value"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"])
    for k, v in d.get("workloads", {}).items():
        print(k, v["value"], "f/s", v["ms_per_step"], "ms", v["frac_of_bf16_peak"], "e2e", v["e2e"]["value"])
    print("eager", json.dumps(d.get("gpu_eager_bf16")))
except Exception as e:
    print("no bench line:", e)
PY
n=$(nvidia-smi -L | wc -l)
if [ "$n" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err
  echo "bench n2 rc=$?"; tail -c 800 gpurun_out/${tag}_bench_n2.err; cut -c 1-400 gpurun_out/${tag}_bench_n2.json
fi
