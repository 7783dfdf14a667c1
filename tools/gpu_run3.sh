#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_pointwise_gpu.py tests/test_deferred_gpu.py -x -q -m gpu > $O/pytest3.log 2>&1; echo "pytest rc=$?" >> $O/pytest3.log
tail -3 $O/pytest3.log
for k in head_fwd head_bwd head_bwd_fused nab poolbwd pool_fused; do python tools/prof_mem.py $k 3; done > $O/prof_mem3.txt 2>&1
UB_HEAD_CUDA_CORES=1 python tools/prof_mem.py head_fwd 3 >> $O/prof_mem3.txt 2>&1
cat $O/prof_mem3.txt
python -m pytest tests/test_model_gpu.py tests/test_reference_golden_gpu.py -x -q -m gpu > $O/pytest3b.log 2>&1; echo "pytest rc=$?" >> $O/pytest3b.log
tail -3 $O/pytest3b.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e --no-roofline"
for i in 1 2; do
  UB_NAB_GENERIC=1 UB_POOL_FUSE=0 UB_HEAD_FUSE=0 UB_HEAD_CUDA_CORES=1 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('old', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
  $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done > $O/ab3.txt 2>&1
cat $O/ab3.txt
python tools/profile_step.py --out $O/r02c_step_profile.txt > /dev/null 2>&1
head -24 $O/r02c_step_profile.txt
