#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_pointwise_gpu.py -x -q -m gpu > $O/pytest20.log 2>&1; echo "rc=$?" >> $O/pytest20.log; tail -3 $O/pytest20.log
timeout 120 python tools/prof_conv.py fwd 1 24 0 24 8 128 128 128 4 | tail -1
timeout 120 python tools/prof_conv.py fwd 0 64 0 64 8 64 64 64 4 | tail -1
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_reference_golden_gpu.py tests/test_deferred_gpu.py -x -q -m gpu > $O/pytest20b.log 2>&1; echo "rc=$?" >> $O/pytest20b.log; tail -3 $O/pytest20b.log
