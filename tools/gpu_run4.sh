#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests/test_pointwise_gpu.py tests/test_deferred_gpu.py -x -q -m gpu > $O/pytest4.log 2>&1; echo "pytest rc=$?" >> $O/pytest4.log
tail -3 $O/pytest4.log
for k in head_bwd_fused poolbwd pool_fused; do python tools/prof_mem.py $k 3; done > $O/prof_mem4.txt 2>&1
cat $O/prof_mem4.txt
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:head_bwd_sums_mma -s 1 -c 1 -o $O/r02d_head_k1 python tools/prof_mem.py head_bwd_fused 2 > $O/ncu41.log 2>&1
$NCU -k regex:head_bwd_apply_mma -s 1 -c 1 -o $O/r02d_head_k2 python tools/prof_mem.py head_bwd_fused 2 > $O/ncu42.log 2>&1
$NCU -k regex:maxpool_bwd_sums -s 1 -c 1 -o $O/r02d_pool_fused python tools/prof_mem.py pool_fused 2 > $O/ncu43.log 2>&1
$NCU -k regex:head_fwd_mma -s 1 -c 1 -o $O/r02d_head_fwd python tools/prof_mem.py head_fwd 2 > $O/ncu44.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e --no-roofline > $O/ncu45.log 2>&1
ls -la $O | tail -8
