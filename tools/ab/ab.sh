#!/bin/bash
# A/B timing: the same prof_conv invocations against the previous build (tools/ab/libubssfp_old.so) and the in-tree build
for args in "fwd 0 64 0 64 8 64 64 64" "fwd 0 64 64 64 8 64 64 64" "dgrad 0 64 0 64 8 64 64 64" "dgrad 0 32 64 32 8 128 128 128" "fwd 0 128 0 128 8 32 32 32" "fwd 0 32 0 64 8 64 64 64"; do
  for rep in 1 2; do
    echo -n "old: "; UB_LIB_PATH=$PWD/tools/ab/libubssfp_old.so python tools/prof_conv.py $args 5 | tail -1
    echo -n "new: "; python tools/prof_conv.py $args 5 | tail -1
  done
done
