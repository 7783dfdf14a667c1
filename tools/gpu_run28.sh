#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_reference_golden_gpu.py -x -q -m gpu > $O/pytest28.log 2>&1; echo "rc=$?" >> $O/pytest28.log; tail -5 $O/pytest28.log
{
for a in "wgrad 3 64 0 64 8 64 64 64" "wgrad 3 128 0 64 8 32 32 32" "wgrad 3 256 0 128 8 16 16 16" "wgrad 3 512 0 256 8 8 8 8"; do
  echo -n "per-chunk, per-sub-position CTAs: "; UB_DC_WGRAD_MODE=0 timeout 120 python tools/prof_conv.py $a 6 | tail -1
  echo -n "chunks on M, sub-positions in N : "; timeout 120 python tools/prof_conv.py $a 6 | tail -1
done
} > $O/r02i_dc_wgrad_ab.txt 2>&1
cat $O/r02i_dc_wgrad_ab.txt
bash tools/ab_step.sh tools/ab/libubssfp_old.so 3 2>&1 | tee $O/r02i_step_ab4.txt
