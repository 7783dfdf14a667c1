#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_reference_golden_gpu.py -x -q -m gpu > $O/pytest34.log 2>&1; echo "rc=$?" >> $O/pytest34.log; tail -5 $O/pytest34.log
{
for a in "fwd 3 64 0 64 8 64 64 64" "dgrad 3 64 0 64 8 64 64 64" "fwd 3 128 0 64 8 32 32 32" "dgrad 3 128 0 64 8 32 32 32" "fwd 3 256 0 128 8 16 16 16" "dgrad 3 256 0 128 8 16 16 16" "fwd 3 512 0 256 8 8 8 8" "dgrad 3 512 0 256 8 8 8 8"; do
  echo -n "32-channel chunks (64-byte rows) : "; UB_KC64=0 timeout 120 python tools/prof_conv.py $a 6 | tail -1
  echo -n "64-channel chunks (128-byte rows): "; timeout 120 python tools/prof_conv.py $a 6 | tail -1
done
} > $O/r02j_kc64_ab.txt 2>&1
cat $O/r02j_kc64_ab.txt
