#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q -m gpu > $O/pytest17.log 2>&1; echo "rc=$?" >> $O/pytest17.log; tail -3 $O/pytest17.log
for a in "fwd 1 24 0 24 8 128 128 128" "fwd 4 30 0 32 8 128 128 128" "dgrad 4 30 0 32 8 128 128 128" "dgrad 2 32 0 64 8 64 64 64" "fwd 4 32 0 64 8 64 64 64" "fwd 1 32 0 6 8 128 128 128"; do
  for m in 2 4; do echo -n "nacc<=$m: "; UB_NACC_MAX=$m timeout 120 python tools/prof_conv.py $a 4 | tail -1; done
done > $O/nacc.txt 2>&1
cat $O/nacc.txt
