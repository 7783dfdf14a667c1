#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
UB_MMA2=1 timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu > $O/pytest22.log 2>&1; echo "rc=$?" >> $O/pytest22.log; tail -5 $O/pytest22.log
for a in "fwd 4 30 0 32 8 128 128 128" "dgrad 4 30 0 32 8 128 128 128" "dgrad 0 32 64 32 8 128 128 128" "fwd 0 64 0 64 8 64 64 64" "dgrad 0 64 0 64 8 64 64 64" "fwd 0 64 64 64 8 64 64 64" "fwd 0 128 0 128 8 32 32 32" "fwd 0 32 0 64 8 64 64 64" "fwd 4 32 0 64 8 64 64 64" "dgrad 2 32 0 64 8 64 64 64" "fwd 1 24 0 24 8 128 128 128" "fwd 0 256 0 256 8 16 16 16" "fwd 2 256 0 512 8 8 8 8"; do
  echo -n "1 issuer : "; timeout 120 python tools/prof_conv.py $a 4 | tail -1
  echo -n "2 issuers: "; UB_MMA2=1 timeout 120 python tools/prof_conv.py $a 4 | tail -1
done > $O/mma2.txt 2>&1
cat $O/mma2.txt
