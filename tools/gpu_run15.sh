#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_pointwise_gpu.py -x -q -m gpu > $O/pytest15.log 2>&1; echo "rc=$?" >> $O/pytest15.log; tail -4 $O/pytest15.log
for sh in "32 0 64 8 64 64 64" "64 0 128 8 32 32 32" "128 0 256 8 16 16 16" "256 0 512 8 8 8 8"; do
  for k in 2 4; do for w in fwd wgrad; do timeout 120 python tools/prof_conv.py $w $k $sh 4 | tail -1; done; done
done > $O/d_s2d.txt 2>&1
cat $O/d_s2d.txt
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_reference_golden_gpu.py tests/test_fp32_mode_gpu.py tests/test_ddp_gpu.py -x -q -m gpu > $O/pytest15b.log 2>&1; echo "rc=$?" >> $O/pytest15b.log; tail -4 $O/pytest15b.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e --no-roofline"
for i in 1 2; do
  UB_D_S2D=0 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('strided', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
  $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('s2d copies', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done > $O/ab15.txt 2>&1
cat $O/ab15.txt
