#!/bin/bash
cd $GRAFT_REPO_ROOT
python tools/parity_values.py > gpurun_out/parity_values.txt 2>&1
cat gpurun_out/parity_values.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
