import torch, numpy as np
dev=torch.device('cuda',0)
flush=torch.empty(256<<20,dtype=torch.uint8,device=dev)
for mb in (103, 206, 412, 2048):
    x=torch.randn(mb*1024*1024//4,device=dev)
    for name,fn in (('sum',lambda: x.sum()),('copy',lambda: x.clone())):
        ts=[]
        for k in range(8):
            flush.fill_(k)
            a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
            a.record(); y=fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        t=np.median(ts[2:])
        nbytes=x.numel()*4*(1 if name=='sum' else 2)
        print(mb,'MB',name,'%.1f us'%(t*1e3),'%.0f GB/s'%(nbytes/t/1e6))
