#!/bin/bash
# In-step sweep of the FMA-pipe share of the bounded attention kernel on ONE box (the power cap makes isolated kernel
# timings a poor predictor): c3 frames/s for VQ_ATTN_BOUNDED_EMU = 0..4, two interleaved rounds.
O=gpurun_out; mkdir -p $O
for round in 1 2; do
  for e in 0 1 2 3 4; do
    VQ_ATTN_BOUNDED_EMU=$e python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline \
      --profile-steps 0 > $O/emu${e}_r$round.json 2> $O/emu${e}_r$round.err
    python - $e $round <<'P'
import json, sys
e, r = sys.argv[1:3]
d = json.loads(open(f"gpurun_out/emu{e}_r{r}.json").read().strip().splitlines()[-1])
print(f"emu {e} round {r}: {d['value']} frames/s  {d['ms_per_step']} ms  sm_mhz {d['clocks']['sm_mhz']}  power {d['clocks']['power_w_max']}")
P
  done
done
