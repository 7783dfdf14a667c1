#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest24.log 2>&1; echo "pytest rc=$?" >> $O/pytest24.log
tail -4 $O/pytest24.log
