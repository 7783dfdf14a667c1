#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_reference_golden_gpu.py tests/test_guard_bands_gpu.py -x -q -m gpu > $O/pytest29.log 2>&1; echo "rc=$?" >> $O/pytest29.log; tail -5 $O/pytest29.log
{
for a in "fwd 4 30 0 32 8 128 128 128" "dgrad 4 30 0 32 8 128 128 128" "fwd 4 32 0 64 8 64 64 64" "dgrad 4 32 0 64 8 64 64 64" "dgrad 0 32 64 32 8 128 128 128" "fwd 0 64 0 64 8 64 64 64" "fwd 0 64 64 64 8 64 64 64" "dgrad 0 64 0 64 8 64 64 64" "fwd 0 128 0 128 8 32 32 32" "fwd 0 128 128 128 8 32 32 32" "fwd 0 32 0 64 8 64 64 64" "fwd 1 24 0 24 8 128 128 128" "fwd 0 256 0 256 8 16 16 16" "fwd 0 512 0 512 8 8 8 8" "fwd 2 256 0 512 8 8 8 8" "fwd 2 128 0 256 8 16 16 16" "fwd 3 64 0 64 8 64 64 64" "dgrad 3 64 0 64 8 64 64 64" "fwd 3 128 0 64 8 32 32 32"; do
  echo -n "one producer (round-start build) : "; UB_LIB_PATH=$PWD/tools/ab/libubssfp_old.so timeout 120 python tools/prof_conv.py $a 6 | tail -1
  echo -n "two producers, one box per plane : "; UB_PLANE_BOX=0 timeout 120 python tools/prof_conv.py $a 6 | tail -1
  echo -n "two producers, one box per tile  : "; timeout 120 python tools/prof_conv.py $a 6 | tail -1
done
} > $O/r02i_producer_ab.txt 2>&1
cat $O/r02i_producer_ab.txt
bash tools/ab_step.sh tools/ab/libubssfp_old.so 2 2>&1 | tee $O/r02i_step_ab5.txt
