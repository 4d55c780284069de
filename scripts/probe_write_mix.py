import torch, sys
N,H,W=4,3000,4000
raw=torch.rand(N,1,H,W,device='cuda'); y=torch.empty(N,3,H,W,device='cuda'); y2=torch.empty(N,3,H,W,device='cuda')
def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
px=N*H*W
t=timeit(lambda: y.copy_(raw.expand(N,3,H,W))); print('expand-copy 4r+12w: %.4f ms  %.0f GB/s (16B/px)'%(t,16*px/t/1e6))
t=timeit(lambda: y.fill_(1.0)); print('fill 12w: %.4f ms %.0f GB/s'%(t,12*px/t/1e6))
t=timeit(lambda: y2.copy_(y)); print('copy 12r+12w: %.4f ms %.0f GB/s'%(t,24*px/t/1e6))
t=timeit(lambda: torch.sum(y)); print('sum 12r: %.4f ms %.0f GB/s'%(t,12*px/t/1e6))
t=timeit(lambda: torch.add(y[:,0],y[:,1],out=y2[:,0])); print('add 8r+4w: %.4f ms %.0f GB/s'%(t,12*px/t/1e6))
