import math, sys, os
sys.path.insert(0, os.getcwd())
import torch
from video_diffusion_nnx_b200 import ops
torch.manual_seed(0)
DEV="cuda"
def run(B,Fr,side):
    HW=side*side; P=B*Fr*HW
    qkv=(torch.randn(P,768,device=DEV)).to(torch.bfloat16)
    qf=qkv.float().requires_grad_(True)
    t=qf.reshape(B,Fr,HW,3,8,32).permute(0,2,1,3,4,5)
    q,k,v=t[...,0,:,:],t[...,1,:,:],t[...,2,:,:]
    att=torch.einsum("...ihd,...jhd->...hij",q/math.sqrt(32),k).softmax(-1)
    o=torch.einsum("...hij,...jhd->...ihd",att,v).permute(0,2,1,3,4).reshape(P,256)
    do=torch.randn(P,256,device=DEV).to(torch.bfloat16)
    o.backward(do.float())
    out=torch.empty(P,256,dtype=torch.bfloat16,device=DEV); lse=torch.empty(P,8,device=DEV)
    ops.mha_core_fwd(qkv,out,lse,0,B,Fr,HW); torch.cuda.synchronize(); print("fwd ok")
    for name,fn in (("smem_bwd", lambda d: ops.mha_temporal_bwd(qkv,out,do,lse,d,B,Fr,side,side)),
                    ("tc_bwd", lambda d: ops.mha_temporal_tc_bwd(qkv,do,lse,d,B,Fr,side,side))):
        d=torch.zeros_like(qkv)
        try:
            fn(d); torch.cuda.synchronize()
            for part,nm in enumerate("qkv"):
                a=d[:,part*256:(part+1)*256].float(); b=qf.grad[:,part*256:(part+1)*256]
                print(name,B,Fr,side,nm,((a-b).abs().max()/b.abs().max()).item())
        except Exception as e:
            print(name,"FAILED",str(e)[:200]); return
for cfg in [(2,10,8),(1,16,16),(1,10,64),(4,10,64)]:
    run(*cfg)
