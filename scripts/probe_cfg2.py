import time, torch, sys
sys.path.insert(0, '.')
import recbole_gnn_b200 as rg
from recbole_gnn_b200 import functional as F_
dev='cuda:0'
U=I=1_000_000; E=100_000_000; D=64; L=3
gen=torch.Generator(device=dev).manual_seed(0)
t0=time.time()
uid=torch.randint(1,U,(E,),generator=gen,device=dev); iid=torch.randint(1,I,(E,),generator=gen,device=dev)
h=rg.GraphHandle.from_interactions(uid,iid,U,I).gcn_norm().to(dev)
torch.cuda.synchronize(); print('build s',time.time()-t0, 'hubs',h._n_hubs)
del uid,iid
xu=(torch.rand(U,D,device=dev)*2-1)*0.0024; xi=(torch.rand(I,D,device=dev)*2-1)*0.0024
for name,fn in [('fused3',lambda: F_.lightgcn_propagate(h,xu,xi,3))]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): fn()
    e.record(); torch.cuda.synchronize()
    ms=s.elapsed_time(e)/10
    print(name,'ms/step',ms,'Gedges/s',200e6*3/ms/1e6, 'GB/s algo', (200e6*264+2e6*260)*3/ms/1e6)
x=torch.cat([xu,xi]); y=torch.empty_like(x)
for _ in range(3): F_.spmm_raw(h,x,y=y)
s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): F_.spmm_raw(h,x,y=y)
e.record(); torch.cuda.synchronize()
ms=s.elapsed_time(e)/10
print('single layer ms',ms,'GB/s algo',(200e6*264+2e6*260)/ms/1e6)
print(torch.cuda.max_memory_allocated()/1e9,'GB peak')
