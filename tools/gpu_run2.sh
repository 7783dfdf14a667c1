#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_pointwise_gpu.py tests/test_conv_gpu.py -x -q -m gpu > $O/pytest2.log 2>&1; echo "pytest rc=$?" >> $O/pytest2.log
for k in nab poolbwd; do python tools/prof_mem.py $k 3; UB_NAB_GENERIC=1 python tools/prof_mem.py $k 3; done > $O/prof_mem2.txt 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > $O/bench2.json 2> $O/bench2.err
UB_NAB_GENERIC=1 UB_POOL_FUSE=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e --no-roofline > $O/bench2_old.json 2>> $O/bench2.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e --no-roofline > $O/bench2_b.json 2>> $O/bench2.err
tail -3 $O/pytest2.log
