#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_reference_golden_gpu.py -x -q -m gpu > $O/pytest31.log 2>&1; echo "rc=$?" >> $O/pytest31.log; tail -5 $O/pytest31.log
{
for a in "dgrad 0 32 64 32 8 128 128 128" "fwd 0 64 0 64 8 64 64 64" "fwd 0 64 64 64 8 64 64 64" "dgrad 0 64 0 64 8 64 64 64" "dgrad 0 64 64 64 8 64 64 64" "fwd 0 128 0 128 8 32 32 32" "fwd 0 128 128 128 8 32 32 32" "dgrad 0 128 0 128 8 32 32 32" "fwd 0 32 0 64 8 64 64 64" "fwd 0 64 0 128 8 32 32 32"; do
  echo -n "one box per depth tap: "; UB_WBOX_MERGE=0 timeout 120 python tools/prof_conv.py $a 6 | tail -1
  echo -n "merged weight boxes  : "; timeout 120 python tools/prof_conv.py $a 6 | tail -1
done
} > $O/r02i_wbox_ab.txt 2>&1
cat $O/r02i_wbox_ab.txt
bash tools/ab_step.sh tools/ab/libubssfp_prev.so 2 2>&1 | tee $O/r02i_step_ab7.txt
