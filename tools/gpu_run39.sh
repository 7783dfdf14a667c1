#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
UB_MMA_PLANES=1 timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -x -q -m gpu > $O/pytest39.log 2>&1; echo "rc=$?" >> $O/pytest39.log; tail -3 $O/pytest39.log
{
for a in "fwd 0 64 0 64 8 64 64 64" "fwd 0 64 64 64 8 64 64 64" "dgrad 0 64 0 64 8 64 64 64" "dgrad 0 32 64 32 8 128 128 128" "fwd 0 128 0 128 8 32 32 32" "fwd 0 128 128 128 8 32 32 32" "fwd 0 32 0 64 8 64 64 64" "fwd 4 30 0 32 8 128 128 128" "dgrad 4 30 0 32 8 128 128 128" "fwd 0 256 0 256 8 16 16 16" "fwd 0 512 0 512 8 8 8 8" "fwd 3 64 0 64 8 64 64 64" "fwd 1 24 0 24 8 128 128 128"; do
  echo -n "one issuer                 : "; timeout 60 python tools/prof_conv.py $a 6 | tail -1
  echo -n "two issuers, planes split  : "; UB_MMA_PLANES=1 timeout 60 python tools/prof_conv.py $a 6 | tail -1
done
} > $O/r02j_fwd_mma_planes.txt 2>&1
cat $O/r02j_fwd_mma_planes.txt
