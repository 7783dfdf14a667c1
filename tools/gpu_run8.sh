#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_pointwise_gpu.py tests/test_model_gpu.py -x -q -m gpu -k "pack or bf16_input or transport" > $O/pytest8.log 2>&1; tail -2 $O/pytest8.log
UB_WIDEN_BF16=0 python tools/bench_bf16_input.py > $O/bf16_in_direct2.txt 2>&1
cat $O/bf16_in_direct2.txt
python tools/bench_bf16_input.py > $O/bf16_in_widen2.txt 2>&1
cat $O/bf16_in_widen2.txt
