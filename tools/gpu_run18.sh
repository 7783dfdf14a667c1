#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q -m gpu > $O/pytest18.log 2>&1; echo "rc=$?" >> $O/pytest18.log; tail -3 $O/pytest18.log
for a in "fwd 3 64 0 64 8 64 64 64" "fwd 3 128 0 64 8 32 32 32"; do
  echo -n "8 N tiles: "; UB_DECONV_PAIR=0 timeout 120 python tools/prof_conv.py $a 4 | tail -1
  echo -n "paired:    "; timeout 120 python tools/prof_conv.py $a 4 | tail -1
done > $O/deconv_pair.txt 2>&1
cat $O/deconv_pair.txt
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_inference_gpu.py -x -q -m gpu > $O/pytest18b.log 2>&1; echo "rc=$?" >> $O/pytest18b.log; tail -3 $O/pytest18b.log
