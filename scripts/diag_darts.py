import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
from test_search_gpu import _opt
from reconfigisp_b200.search import DartsModel
from reconfigisp_b200 import ops
from oracle import pipeline_oracle as PO
g = np.load('tests/golden/darts_model.npz'); T = torch.from_numpy
m = DartsModel(_opt(3)); net = m.netG
oG = PO.Supernet(3, 0.2, 10)
with torch.no_grad():
    for i, a in enumerate(oG.alphas): a.copy_(T(g['it0_alpha_%d' % i]))
    for a, b in zip(net.alphas, oG.alphas): a.copy_(b.cuda())
img, gt = T(g['img']), T(g['gt'])
y, inter = oG.forward(img)
go = torch.autograd.grad(((y - gt) ** 2).mean(), inter)
yg = net(img.cuda())
gg = torch.autograd.grad(ops.mse_loss(yg, gt.cuda()), net.middle_results)
k = 2
err = (gg[k].cpu() - go[k]).abs(); s = float(go[k].abs().max())
bad = (err > 2e-3 * s).nonzero()
print('outlier elements', bad.shape[0], 'of', err.numel())
xo, xg = inter[k].detach(), net.middle_results[k].detach().cpu()
for n in range(xo.shape[0]):
    for c in range(3):
        po, pg = xo[n, c], xg[n, c]
        print('n%d c%d' % (n, c), 'argmin cpu', int(po.argmin()), 'gpu', int(pg.argmin()), 'argmax cpu', int(po.argmax()), 'gpu', int(pg.argmax()),
              'min gap cpu %.2e' % float(po.flatten().sort().values[1] - po.min()), 'max gap %.2e' % float(po.max() - po.flatten().sort().values[-2]),
              'n(==min) %d n(==max) %d' % (int((po == po.min()).sum()), int((po == po.max()).sum())))
print(bad[:12].tolist())
