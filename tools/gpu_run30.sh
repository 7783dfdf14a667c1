#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_reference_golden_gpu.py tests/test_deferred_gpu.py -x -q -m gpu > $O/pytest30.log 2>&1; echo "rc=$?" >> $O/pytest30.log; tail -5 $O/pytest30.log
{
for a in "wgrad 4 30 0 32 8 128 128 128" "wgrad 4 32 0 64 8 64 64 64" "wgrad 0 32 0 32 8 128 128 128" "wgrad 0 32 64 32 8 128 128 128" "wgrad 0 64 0 64 8 64 64 64" "wgrad 0 128 0 128 8 32 32 32" "wgrad 3 64 0 64 8 64 64 64" "wgrad 3 128 0 64 8 32 32 32" "wgrad 4 64 0 128 8 32 32 32" "wgrad 2 128 0 256 8 16 16 16" "wgrad 2 256 0 512 8 8 8 8" "wgrad 1 24 0 24 8 128 128 128" "wgrad 0 512 0 512 8 8 8 8"; do
  echo -n "one producer thread : "; UB_LIB_PATH=$PWD/tools/ab/libubssfp_prev.so timeout 120 python tools/prof_conv.py $a 6 | tail -1
  echo -n "two producer threads: "; timeout 120 python tools/prof_conv.py $a 6 | tail -1
done
} > $O/r02i_wgrad_producer_ab.txt 2>&1
cat $O/r02i_wgrad_producer_ab.txt
bash tools/ab_step.sh tools/ab/libubssfp_prev.so 3 2>&1 | tee $O/r02i_step_ab6.txt
