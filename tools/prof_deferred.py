"""One launch each of the full-resolution 32 -> 32 forward conv in its three source-0 forms (for ncu):
materialised fp16 operand copy (the default), deferred activation rewritten in shared memory in the fp16 form and
in the bf16 form (UB_DEFER_CONV=1).   python tools/prof_deferred.py [n size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import unet_bssfp_b200 as ub  # noqa: E402

ops = ub.ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
y = (torch.randn((N, S, S, S, 32), device=dev, generator=g) * 1.5).to(torch.float16)
scale = torch.rand((N, 32), device=dev, generator=g) + 0.5
shift = torch.randn((N, 32), device=dev, generator=g) * 0.3
spec = ops.ConvSpec(0, 32, 32)
wt = torch.randn((32, 32, 3, 3, 3), device=dev, generator=g) * 0.05
b = torch.zeros(32, device=dev)
w16 = ops.pack_conv_weights(spec, wt, ops.UB_PACK_F16_SRC0)
wbf = ops.pack_conv_weights(spec, wt, 0)
_, _, a16 = ops.norm_act_fwd(y, scale, shift, 0.1, 0.05, 1234, f16_copy=True, materialize=False)
for _ in range(2):
    ops.conv_fwd(spec, a16, None, w16, b, want_stats=True)
    ops.conv_fwd(spec, ops.DeferredAct(y, scale, shift, 0.1, 0.05, 1234, f16_operand=True), None, w16, b, want_stats=True)
    ops.conv_fwd(spec, ops.DeferredAct(y, scale, shift, 0.1, 0.05, 1234), None, wbf, b, want_stats=True)
torch.cuda.synchronize()
print("done")
