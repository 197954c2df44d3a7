import sys, torch
sys.path.insert(0, "/root/repo")
from video_diffusion_nnx_b200 import ops
from video_diffusion_nnx_b200._lib import debug_switches
dev="cuda"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e3
for (B,Fr,HW) in ((4,10,64),(4,16,256)):
    P=B*Fr*HW
    qkv=torch.randn(P,768,device=dev).to(torch.bfloat16); do=torch.randn(P,256,device=dev).to(torch.bfloat16)
    out=torch.empty(P,256,dtype=torch.bfloat16,device=dev); lse=torch.empty(P,8,device=dev); D=torch.empty(P,8,device=dev); dq=torch.empty_like(qkv)
    for sw in ({"VDN_MHA_SPATIAL_SCALAR":1},{}):
        with debug_switches(**sw):
            f=timeit(lambda: ops.mha_core_fwd(qkv,out,lse,1,B,Fr,HW)); b=timeit(lambda: ops.mha_core_bwd(qkv,out,do,lse,D,dq,1,B,Fr,HW))
        print(B,Fr,HW,sw,"fwd %.1f us bwd %.1f us"%(f,b))
