#!/bin/bash
# In-step A/B of launch heuristics on ONE box (two interleaved rounds through bench.py; isolated kernel timings mislead at
# the power cap).  usage: WL=c2 bash scripts/knob_sweep.sh tag1:ENV=VAL tag2:ENV=VAL ...   (tag:VQ_X=0 = defaults)
O=gpurun_out; mkdir -p $O
WL=${WL:-c2}
STEPS=20; [ "$WL" != "c2" ] && STEPS=10
for round in 1 2; do
  for v in "$@"; do
    tag=${v%%:*}; asg=${v#*:}
    env $asg python bench.py --workload $WL --steps $STEPS --warmup 5 --no-cpu-baseline --no-eager-baseline --no-extra-workloads \
      --profile-steps 0 > $O/knob_${WL}_${tag}_r$round.json 2> $O/knob_${WL}_${tag}_r$round.err
    python - $WL $tag $round <<'P'
import json, sys
wl, tag, r = sys.argv[1:4]
try:
    d = json.loads(open(f"gpurun_out/knob_{wl}_{tag}_r{r}.json").read().strip().splitlines()[-1])
    print(f"{wl} {tag}_r{r}: {d['value']} frames/s  {d['ms_per_step']} ms  sm_mhz {d['clocks']['sm_mhz']}")
except Exception as e:
    print(f"{wl} {tag}_r{r}: no line ({e})")
P
  done
done
