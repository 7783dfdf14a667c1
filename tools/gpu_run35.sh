#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
{
echo "TEMPORARY experiment (debug field, not in the tree): the marching kernels with kh groups of UMMAs left out -- same TMA loads,"
echo "same barriers and commits, same epilogue; 3 = the real kernel, 2 / 1 / 0 = two / one / none of the three kh groups issued"
for a in "wgrad 0 32 0 32 8 128 128 128" "wgrad 0 32 64 32 8 128 128 128" "wgrad 0 64 0 64 8 64 64 64" "fwd 0 32 0 32 8 128 128 128" "fwd 0 32 64 32 8 128 128 128" "dgrad 0 32 0 32 8 128 128 128"; do
  for k in 3 2 1 0; do
    echo -n "kh groups $k: "; UB_DBG_KH=$k timeout 120 python tools/prof_conv.py $a 5 | tail -1
  done
done
} > $O/r02j_march_mma_ablation.txt 2>&1
cat $O/r02j_march_mma_ablation.txt
