#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/profile_step.py --out $O/r02g_step_profile.txt > /dev/null 2>&1
head -60 $O/r02g_step_profile.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-roofline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks'])"
