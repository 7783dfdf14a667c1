#!/bin/bash
# A/B of an environment switch on the full step: tools/ab_env.sh VAR A B
for v in "$2" "$3" "$2" "$3"; do
  echo -n "$1=$v: "; env $1=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2),'ms/step')"
done
