#!/bin/bash
# c2 / c4 throughput for different blocks-per-warp tables of the tensor-core prefilter (SMEAGOL_B200_HMMA_NB)
cd "$(dirname "$0")/.."
for nb in "$@"; do
  export SMEAGOL_B200_HMMA_NB=$nb
  echo -n "nb $nb: "
  python bench.py --steps 3 --warmup 3 --cpu-sample 4 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("c2", round(d["value"],1), end="  ")'
  python bench_configs.py --config c4 --steps 3 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("c4", round(d["value"],1))'
done
