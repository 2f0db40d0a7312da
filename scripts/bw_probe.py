import torch, time
x = torch.empty(1<<30, dtype=torch.float32, device='cuda').normal_()   # 4 GiB
y = torch.empty_like(x)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best=1e9
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best=min(best,a.elapsed_time(b))
    return best
ms=t(lambda: x.sum()); print('read-only sum   %.1f GB/s'%(x.numel()*4/ms/1e6))
ms=t(lambda: x.max()); print('read-only max   %.1f GB/s'%(x.numel()*4/ms/1e6))
ms=t(lambda: y.copy_(x)); print('copy (r+w)      %.1f GB/s'%(2*x.numel()*4/ms/1e6))
ms=t(lambda: y.fill_(1.0)); print('write-only fill %.1f GB/s'%(x.numel()*4/ms/1e6))
