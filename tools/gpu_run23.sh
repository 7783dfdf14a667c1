#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
python tools/prof_conv.py wgrad 3 64 0 64 8 64 64 64 3 | tail -1
python tools/prof_conv.py wgrad 1 24 0 24 8 128 128 128 3 | tail -1
python tools/prof_conv.py dgrad 3 64 0 64 8 64 64 64 3 | tail -1
$NCU -k regex:igemm_wgrad_kernel -s 1 -c 1 -o $O/r02h_deconv_wgrad python tools/prof_conv.py wgrad 3 64 0 64 8 64 64 64 2 > $O/ncu231.log 2>&1
$NCU -k regex:igemm_fwd_kernel -s 1 -c 1 -o $O/r02h_deconv_dgrad python tools/prof_conv.py dgrad 3 64 0 64 8 64 64 64 2 > $O/ncu232.log 2>&1
ls -la $O/r02h*.ncu-rep
