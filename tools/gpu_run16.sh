#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q -m gpu > $O/pytest16.log 2>&1; echo "rc=$?" >> $O/pytest16.log; tail -4 $O/pytest16.log
for sh in "4 30 0 32 8 128 128 128" "2 32 0 64 8 64 64 64" "2 64 0 128 8 32 32 32" "2 128 0 256 8 16 16 16"; do
  UB_K4_DGRAD_FOLD=0 timeout 120 python tools/prof_conv.py dgrad $sh 4 | tail -1
  timeout 120 python tools/prof_conv.py dgrad $sh 4 | tail -1
done > $O/k4_dgrad_fold.txt 2>&1
cat $O/k4_dgrad_fold.txt
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_reference_golden_gpu.py -x -q -m gpu > $O/pytest16b.log 2>&1; echo "rc=$?" >> $O/pytest16b.log; tail -3 $O/pytest16b.log
