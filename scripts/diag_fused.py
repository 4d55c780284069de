import sys, os, torch
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(),'tests'))
from oracle import isp_oracle as O
import reconfigisp_b200.ops as ops
from test_fused_gpu import SIGS, DM, stage_params, oracle_chain
kind, sig, shape = 'malvar', 'A', (3, 50, 520)
N,H,W = shape
g = torch.Generator().manual_seed(hash((kind, sig, shape)) % 1000)
raw = torch.rand(N, 1, H, W, generator=g) * 1.05
raw.view(-1)[:3] = torch.tensor([0., 1., 0.5])
gt = torch.rand(N, 3, H, W, generator=g)
st = SIGS[sig]; chain = ops.Chain(st); params = stage_params(st, g)
yo = oracle_chain(DM[kind](raw), st, params.expand(N,-1))
yd = oracle_chain(DM[kind](raw.double()), st, params.double().expand(N,-1))
yg = ops.pipeline_fwd(raw.cuda(), kind, chain, params.cuda()).cpu()
os.environ['X']='1'
print('gpu vs f32 oracle', float((yg-yo).abs().max()), 'gpu vs f64', float((yg.double()-yd).abs().max()), 'f32 oracle vs f64', float((yo.double()-yd).abs().max()))
i = (yg.double()-yd).abs().argmax(); idx = torch.unravel_index(i, yg.shape); print(idx, yg[idx], yo[idx], yd[idx])
# intermediates at that pixel
dm = DM[kind](raw.double()); n,c,y,x = [int(v) for v in idx]
a = O.wb_manual(dm, params.double()[:,0:3].expand(N,3)); b = O.wb_quadratic(a, ((params.double()[:,3:33]+5)/10).expand(N,30)); cc = O.gamma_manual(b, params.double()[:,33:34].expand(N,1))
print('dm', dm[n,:,y,x], 'poly', b[n,:,y,x], 'gamma', cc[n,:,y,x])
