#!/bin/bash
# c2 (bench.py) and c4 throughput as a function of the hybrid split width (tensor-core prefilter for W >= split)
cd "$(dirname "$0")/.."
for s in "$@"; do
  echo -n "split $s: "
  SMEAGOL_B200_TC_MIN_WIDTH=$s python bench.py --steps 3 --warmup 3 --cpu-sample 4 2>&1 | tail -1 | \
    python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("c2", round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1))'
done
for s in "$@"; do
  for c in c3 c4; do
    echo -n "split $s: "
    SMEAGOL_B200_TC_MIN_WIDTH=$s python bench_configs.py --config $c --steps 3 2>/dev/null | tail -1 | \
      python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("'$c'", round(d["value"],1), round(d["ms_per_step"],2))'
  done
done
