#!/bin/bash
cd $GRAFT_REPO_ROOT
python tools/grad_yardstick_probe.py > gpurun_out/grad_yardstick.txt 2>&1
cat gpurun_out/grad_yardstick.txt | tail -40
python -m pytest tests/test_model_gpu.py tests/test_reference_golden_gpu.py -x -q -m gpu 2>&1 | tail -3
