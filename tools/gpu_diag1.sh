#!/bin/bash
# round-2 diagnostics: per-shape step profile, e2e probe, ncu captures of the slow memory-bound kernels
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
python tools/profile_step.py --out $O/r02b_step_profile.txt > /dev/null 2>$O/prof_step.err
python tools/e2e_probe.py 8 > $O/e2e_probe.txt 2>&1
for k in head_fwd head_bwd nab naf pool poolbwd; do python tools/prof_mem.py $k 3; done > $O/prof_mem.txt 2>&1
python tools/prof_conv.py fwd 3 64 0 64 8 64 64 64 3 >> $O/prof_mem.txt 2>&1
python tools/prof_conv.py fwd 1 24 0 24 8 128 128 128 3 >> $O/prof_mem.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:conv1x1_to_ncdhw -s 1 -c 1 -o $O/r02b_head_fwd python tools/prof_mem.py head_fwd 2 > $O/ncu1.log 2>&1
$NCU -k regex:conv1x1_from_ncdhw_bwd_kernel -s 1 -c 1 -o $O/r02b_head_bwd python tools/prof_mem.py head_bwd 2 > $O/ncu2.log 2>&1
$NCU -k regex:norm_act_bwd_apply -s 1 -c 1 -o $O/r02b_nab_apply python tools/prof_mem.py nab 2 > $O/ncu3.log 2>&1
$NCU -k regex:igemm_fwd_kernel -s 1 -c 1 -o $O/r02b_deconv64 python tools/prof_conv.py fwd 3 64 0 64 8 64 64 64 2 > $O/ncu4.log 2>&1
$NCU -k regex:igemm_fwd_kernel -s 1 -c 1 -o $O/r02b_k1_24 python tools/prof_conv.py fwd 1 24 0 24 8 128 128 128 2 > $O/ncu5.log 2>&1
ls -la $O | tail -12
