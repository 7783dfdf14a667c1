#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/step_jitter.py 30 > $O/step_jitter3.txt 2>&1
awk '{print $2, $3, $7, $13}' $O/step_jitter3.txt | tr '\n' ';'; echo
python -m pytest tests -x -q -m gpu > $O/pytest10.log 2>&1; echo "pytest rc=$?" >> $O/pytest10.log
tail -3 $O/pytest10.log
