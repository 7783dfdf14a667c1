#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/pytest5.log 2>&1; echo "pytest rc=$?" >> $O/pytest5.log
tail -4 $O/pytest5.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e --no-roofline"
for i in 1 2 3; do
  UB_NAB_GENERIC=1 UB_HEAD_FUSE=0 UB_HEAD_CUDA_CORES=1 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('old', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['power_w'])"
  $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks']['power_w'])"
done > $O/ab5.txt 2>&1
cat $O/ab5.txt
