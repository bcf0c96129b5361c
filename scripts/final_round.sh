#!/bin/bash
# Final validation of a round on ONE GPU: full GPU test suite, smoke(), the reference arm and the default bench line
# exactly as the driver runs them, plus the per-workload kernel tables.  usage: gpurun -- 'bash scripts/final_round.sh r02'
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > $O/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${R}_pytest_gpu.log
tail -3 $O/${R}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.log 2>&1; tail -1 $O/${R}_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/${R}_bench_reference_arm.json 2> $O/${R}_bench_reference_arm.err; echo "reference rc=$?"
timeout 1200 python bench.py --steps 20 --warmup 5 --profile-out $O/${R}_kernels_c2_events.txt > $O/${R}_bench_c2.json 2> $O/${R}_bench_c2.err; echo "bench rc=$?"
for w in c3 c4; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline \
    --profile-out $O/${R}_kernels_${w}_events.txt > $O/${R}_bench_$w.json 2> $O/${R}_bench_$w.err
done
python - <<PY
import json
for n in ("reference_arm", "c2", "c3", "c4"):
    try:
        d = json.loads([l for l in open("$O/${R}_bench_%s.json" % n) if l.startswith("{")][-1])
        print(n, d["value"], d["unit"], d["ms_per_step"], "ms", "e2e", d["e2e"]["value"], d.get("frac_of_bf16_peak"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
        if n == "c2":
            print(" eager", {k: v.get("value") for k, v in d.get("gpu_eager_bf16", {}).items()}, "cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("cores"))
            for k, v in d.get("workloads", {}).items():
                print(" ", k, v["value"], v["ms_per_step"], v["frac_of_bf16_peak"])
    except Exception as e:
        print(n, "no line:", e)
PY
