#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_reference_golden_gpu.py tests/test_deferred_gpu.py -x -q -m gpu > $O/pytest37.log 2>&1; echo "rc=$?" >> $O/pytest37.log; tail -3 $O/pytest37.log
{
for a in "wgrad 0 32 0 32 8 128 128 128" "wgrad 0 32 64 32 8 128 128 128" "wgrad 0 24 0 32 8 128 128 128" "wgrad 0 64 0 64 8 64 64 64" "wgrad 0 64 64 64 8 64 64 64" "wgrad 0 128 0 128 8 32 32 32" "wgrad 4 30 0 32 8 128 128 128" "wgrad 4 32 0 64 8 64 64 64"; do
  echo -n "one issuer : "; UB_WGRAD_MMA2=0 timeout 60 python tools/prof_conv.py $a 6 | tail -1
  echo -n "two issuers: "; timeout 60 python tools/prof_conv.py $a 6 | tail -1
done
} > $O/r02j_wgrad_mma2.txt 2>&1
cat $O/r02j_wgrad_mma2.txt
bash tools/ab_step.sh tools/ab/libubssfp_prev.so 2 2>&1 | tee $O/r02j_step_ab.txt
