import sys, torch, time
sys.path.insert(0, '/root/repo')
import unet_bssfp_b200 as ub
from unet_bssfp_b200.train_step import GanTrainer
dev='cuda'
def timed(fn, iters=5):
    fn(); fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/iters
x=torch.rand(8,24,128,128,128,device=dev); y=torch.rand(8,6,128,128,128,device=dev); x16=x.to(torch.bfloat16)
for nm,xx in (("fp32",x),("bf16",x16)):
    print(nm, "pack plain %.3f ms"%timed(lambda: ub.ops.pack_ncdhw(xx)), "pack s2d %.3f ms"%timed(lambda: ub.ops.pack_ncdhw(xx,y,s2d=True)))
torch.manual_seed(0)
g,d=ub.Generator('bssfp').to(dev),ub.Discriminator('bssfp').to(dev)
tr=GanTrainer(g,d)
xs=[(x,y),(x.flip(0).contiguous(),y.flip(0).contiguous())]
xs16=[(a.to(torch.bfloat16),b) for a,b in xs]
for nm,bs in (("fp32",xs),("bf16",xs16),("fp32",xs),("bf16",xs16)):
    i=[0]
    def step():
        tr.step(*bs[i[0]%2]); i[0]+=1
    print(nm, "step %.2f ms"%timed(step, 6))
