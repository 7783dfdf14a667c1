#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/phase_times.py 5 > $O/phase_times2.txt 2>&1
tail -3 $O/phase_times2.txt
python tools/bench_bf16_input.py > $O/bf16_in_widen.txt 2>&1
UB_WIDEN_BF16=0 python tools/bench_bf16_input.py > $O/bf16_in_direct.txt 2>&1
cat $O/bf16_in_widen.txt $O/bf16_in_direct.txt
