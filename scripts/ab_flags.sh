#!/bin/bash
# A/B of DSC_TC5_FLAGS on the dominant shapes: bash scripts/ab_flags.sh "0 1 2 3" "4096x40,1024x80"
FLAGS=${1:-"0 1 2 3"}; SHAPES=${2:-"4096x40"}
for f in $FLAGS; do
  echo "== DSC_TC5_FLAGS=$f"
  DSC_XATTN_IMPL=tc5 DSC_TC5_FLAGS=$f timeout 120 python scripts/tc5_debug.py fwd 16 4096 40 2>&1 | tail -1
  DSC_XATTN_IMPL=tc5 DSC_TC5_FLAGS=$f timeout 120 python scripts/tc5_debug.py stats 16 1024 80 2>&1 | tail -1
  DSC_XATTN_IMPL=tc5 DSC_TC5_FLAGS=$f timeout 300 python scripts/microbench.py --quick --no-ref --shapes $SHAPES --out gpurun_out/ab_$f.jsonl 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k in ('L','D','ms_stats','ms_forward','ms_both','ms_call_sustained','frac')})
"
done
