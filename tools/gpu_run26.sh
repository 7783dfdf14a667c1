#!/bin/bash
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_reference_golden_gpu.py -x -q -m gpu > $O/pytest26.log 2>&1; echo "rc=$?" >> $O/pytest26.log; tail -5 $O/pytest26.log
{
for a in "wgrad 3 64 0 64 8 64 64 64" "wgrad 3 128 0 64 8 32 32 32" "wgrad 3 256 0 128 8 16 16 16" "wgrad 3 512 0 256 8 8 8 8" "wgrad 4 32 0 64 8 64 64 64" "wgrad 4 64 0 128 8 32 32 32" "wgrad 2 128 0 256 8 16 16 16" "wgrad 2 256 0 512 8 8 8 8" "wgrad 1 24 0 24 8 128 128 128" "wgrad 0 512 0 512 8 8 8 8" "wgrad 0 256 0 512 8 8 8 8"; do
  echo -n "old split: "; UB_WGRAD_SPLIT_OLD=1 timeout 120 python tools/prof_conv.py $a 5 | tail -1
  echo -n "new split: "; timeout 120 python tools/prof_conv.py $a 5 | tail -1
done
for a in "dgrad 3 64 0 64 8 64 64 64" "dgrad 3 128 0 64 8 32 32 32" "dgrad 3 256 0 128 8 16 16 16" "dgrad 3 512 0 256 8 8 8 8"; do
  echo -n "td2: "; UB_DC_DGRAD_TD=2 timeout 120 python tools/prof_conv.py $a 5 | tail -1
  echo -n "td1: "; timeout 120 python tools/prof_conv.py $a 5 | tail -1
done
} > $O/r02i_wgrad_split_ab.txt 2>&1
cat $O/r02i_wgrad_split_ab.txt
python - <<'PY' 2>&1 | tee $O/r02i_pack_ab.txt
import os, torch, time
import unet_bssfp_b200 as ub
from unet_bssfp_b200 import modules as M
g = ub.Generator("bssfp").cuda(); d = ub.Discriminator("bssfp").cuda()
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x = torch.rand(1, 24, 32, 32, 32, device="cuda")
g(x)
net = g._graph
for name, env in (("gather", "1"), ("tiled", "0")):
    os.environ["UB_PACK_OLD"] = env
    print(name, "G fwd pack ms", round(t(lambda: net.cache.refresh(net.fwd_blocks(), 0)), 4),
          "G dgrad pack ms", round(t(lambda: net.cache.refresh(net.fwd_blocks(), 1)), 4))
    d(x, torch.rand(1, 6, 32, 32, 32, device="cuda"))
    ch = d._chain
    print(name, "D fwd pack ms", round(t(lambda: ch.cache.refresh(ch.blocks, 0)), 4), "D dgrad pack ms", round(t(lambda: ch.cache.refresh(ch.blocks, 1)), 4))
PY
bash tools/ab_step.sh tools/ab/libubssfp_old.so 2 2>&1 | tee $O/r02i_step_ab2.txt
