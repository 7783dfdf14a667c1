#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_reference_golden_gpu.py tests/test_ddp_gpu.py -x -q -m gpu > $O/pytest27.log 2>&1; echo "rc=$?" >> $O/pytest27.log; tail -5 $O/pytest27.log
bash tools/ab_step.sh tools/ab/libubssfp_old.so 2 2>&1 | tee $O/r02i_step_ab3.txt
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-roofline --no-e2e > $O/b27.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02i_launches_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-roofline --no-e2e > $O/ncu27.log 2>&1
python tools/launch_summary.py $O/r02i_launches_step.csv 520 | head -50
