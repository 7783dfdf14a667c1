#!/bin/bash
# A/B of two builds of the library on the full step, alternating on the same box:
#   tools/ab_step.sh tools/ab/libubssfp_old.so [reps]
old=$1; reps=${2:-3}
for i in $(seq $reps); do
  for lib in "$old" ""; do
    echo -n "${lib:-in-tree}: "
    UB_LIB_PATH=${lib:+$PWD/$lib} python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e 2>/dev/null \
      | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2),'ms/step', d['clocks']['sm_mhz'], 'MHz')"
  done
done
