#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-roofline --no-e2e > $O/b33.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02j_launches_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-roofline --no-e2e > $O/ncu33.log 2>&1
python tools/launch_summary.py $O/r02j_launches_step.csv 516 | head -30
python tools/prof_conv.py fwd 4 30 0 32 8 128 128 128 2 | tail -1 && $NCU -k regex:igemm_fwd_kernel -s 1 -c 1 -o $O/r02j_stem_fwd python tools/prof_conv.py fwd 4 30 0 32 8 128 128 128 2 > $O/ncu331.log 2>&1
python tools/prof_conv.py wgrad 3 64 0 64 8 64 64 64 2 | tail -1 && $NCU -k regex:igemm_wgrad_kernel -s 1 -c 1 -o $O/r02j_deconv_wgrad python tools/prof_conv.py wgrad 3 64 0 64 8 64 64 64 2 > $O/ncu332.log 2>&1
python tools/prof_conv.py dgrad 0 32 64 32 8 128 128 128 2 | tail -1 && $NCU -k regex:igemm_fwd_kernel -s 1 -c 1 -o $O/r02j_dgrad96 python tools/prof_conv.py dgrad 0 32 64 32 8 128 128 128 2 > $O/ncu333.log 2>&1
ls -la $O/r02j*.ncu-rep
