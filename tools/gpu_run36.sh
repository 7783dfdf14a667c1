#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -x -q -m gpu > $O/pytest36a.log 2>&1; echo "rc=$?" >> $O/pytest36a.log; tail -3 $O/pytest36a.log
UB_MARCH_MMA2=1 timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -x -q -m gpu > $O/pytest36b.log 2>&1; echo "rc=$?" >> $O/pytest36b.log; tail -3 $O/pytest36b.log
{
for a in "fwd 0 32 0 32 8 128 128 128" "dgrad 0 32 0 32 8 128 128 128" "fwd 0 24 0 32 8 128 128 128"; do
  echo -n "one issuer : "; timeout 60 python tools/prof_conv.py $a 6 | tail -1
  echo -n "two issuers: "; UB_MARCH_MMA2=1 timeout 60 python tools/prof_conv.py $a 6 | tail -1
done
for a in "fwd 0 32 64 32 8 128 128 128"; do
  echo -n "CTA pairs              : "; timeout 60 python tools/prof_conv.py $a 6 | tail -1
  echo -n "single CTA, one issuer : "; UB_MARCH_PAIR=0 timeout 60 python tools/prof_conv.py $a 6 | tail -1
  echo -n "single CTA, two issuers: "; UB_MARCH_PAIR=0 UB_MARCH_MMA2=3 timeout 60 python tools/prof_conv.py $a 6 | tail -1
done
} > $O/r02j_march_mma2.txt 2>&1
cat $O/r02j_march_mma2.txt
