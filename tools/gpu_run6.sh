#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/phase_times.py 6 > $O/phase_times.txt 2>&1
cat $O/phase_times.txt
python tools/profile_step.py --out $O/r02e_step_profile.txt > /dev/null 2>&1
head -22 $O/r02e_step_profile.txt
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e --no-roofline"
for i in 1 2; do
  UB_WGRAD_STREAM=0 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('one stream', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['power_w'])"
  $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('side stream', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['power_w'])"
done > $O/ab6.txt 2>&1
cat $O/ab6.txt
nvidia-smi -q -d POWER | grep -iE "power limit|default|enforced|max power|min power" | head -12
