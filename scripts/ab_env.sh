#!/bin/bash
# dev: bench.py (no CPU baseline) under a list of environment settings: scripts/ab_env.sh "A=1" "A=0 B=2" ...
for e in "$@"; do
  echo "== $e"
  env $e python bench.py --cpu-sample 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value']), d['stage_ms'], round(d['hsddp_iteration_ms_max'],2))"
done
