#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
UB_MARCH_PAIR_MMA2=1 timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -x -q -m gpu > $O/pytest38.log 2>&1; echo "rc=$?" >> $O/pytest38.log; tail -3 $O/pytest38.log
{
for a in "fwd 0 32 64 32 8 128 128 128" "fwd 0 32 32 32 8 128 128 128"; do
  echo -n "CTA pairs, one issuer : "; timeout 60 python tools/prof_conv.py $a 6 | tail -1
  echo -n "CTA pairs, two issuers: "; UB_MARCH_PAIR_MMA2=1 timeout 60 python tools/prof_conv.py $a 6 | tail -1
done
} > $O/r02j_march_pair_mma2.txt 2>&1
cat $O/r02j_march_pair_mma2.txt
