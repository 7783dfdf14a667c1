#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
python tools/prof_conv.py fwd 4 30 0 32 8 128 128 128 3 > $O/stem.txt 2>&1
python tools/prof_conv.py wgrad 4 30 0 32 8 128 128 128 3 >> $O/stem.txt 2>&1
python tools/prof_conv.py dgrad 4 30 0 32 8 128 128 128 3 >> $O/stem.txt 2>&1
python tools/prof_conv.py fwd 2 32 0 64 8 64 64 64 3 >> $O/stem.txt 2>&1
cat $O/stem.txt
$NCU -k regex:igemm_fwd_kernel -s 1 -c 1 -o $O/r02f_stem_fwd python tools/prof_conv.py fwd 4 30 0 32 8 128 128 128 2 > $O/ncu111.log 2>&1
$NCU -k regex:igemm_wgrad_kernel -s 1 -c 1 -o $O/r02f_stem_wgrad python tools/prof_conv.py wgrad 4 30 0 32 8 128 128 128 2 > $O/ncu112.log 2>&1
$NCU -k regex:igemm_fwd_kernel -s 1 -c 1 -o $O/r02f_d2_fwd python tools/prof_conv.py fwd 2 32 0 64 8 64 64 64 2 > $O/ncu113.log 2>&1
ls -la $O/*.ncu-rep | tail -4
