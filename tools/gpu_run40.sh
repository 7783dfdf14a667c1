#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_reference_golden_gpu.py tests/test_deferred_gpu.py -x -q -m gpu > $O/pytest40.log 2>&1; echo "rc=$?" >> $O/pytest40.log; tail -3 $O/pytest40.log
{
for a in "wgrad 0 32 0 32 8 128 128 128" "wgrad 0 32 64 32 8 128 128 128" "wgrad 0 64 0 64 8 64 64 64" "wgrad 0 64 64 64 8 64 64 64" "wgrad 0 128 0 128 8 32 32 32"; do
  echo -n "one issuer               : "; UB_WGRAD_MMA2=0 timeout 60 python tools/prof_conv.py $a 6 | tail -1
  echo -n "two issuers, 12 : 12     : "; timeout 60 python tools/prof_conv.py $a 6 | tail -1
done
} > $O/r02k_wgrad_mma2_balanced.txt 2>&1
cat $O/r02k_wgrad_mma2_balanced.txt
