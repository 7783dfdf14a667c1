#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/step_jitter.py 40 > $O/step_jitter.txt 2>&1
cat $O/step_jitter.txt | awk 'NR<=8 || $3>70' | head -30
python tools/step_jitter.py 30 > $O/step_jitter2.txt 2>&1
cat $O/step_jitter2.txt | awk 'NR<=6 || $3>70' | head -30
nvidia-smi --query-gpu=temperature.gpu,clocks.sm,power.draw,clocks_event_reasons.active --format=csv
