#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q -m gpu > $O/pytest19.log 2>&1; echo "rc=$?" >> $O/pytest19.log; tail -3 $O/pytest19.log
for a in "fwd 3 64 0 64 8 64 64 64" "dgrad 3 64 0 64 8 64 64 64" "dgrad 4 30 0 32 8 128 128 128" "fwd 4 30 0 32 8 128 128 128" "dgrad 0 32 64 32 8 128 128 128" "dgrad 0 64 0 64 8 64 64 64" "dgrad 0 64 64 64 8 64 64 64" "dgrad 2 32 0 64 8 64 64 64" "dgrad 0 128 0 128 8 32 32 32" "fwd 3 128 0 64 8 32 32 32"; do
  echo -n " 8 epilogue warps: "; UB_EPI16=0 timeout 120 python tools/prof_conv.py $a 4 | tail -1
  echo -n "16 epilogue warps: "; timeout 120 python tools/prof_conv.py $a 4 | tail -1
done > $O/epi16.txt 2>&1
cat $O/epi16.txt
