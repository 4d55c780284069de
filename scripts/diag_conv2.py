import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from reconfigisp_b200.modules import tools_proxy as P
g = np.load('tests/golden/cnn_candidates.npz')
T = torch.from_numpy
torch.backends.cudnn.allow_tf32 = False
net = P.ProxyNet(3, None); net.load_state_dict(P.seeded_state_dict(net, 10))
x = T(g['x3']); par = T(g['srcnn_res3_par'])
N, _, H, W = x.shape
feat = torch.cat([x.amin(dim=(2, 3)), x.mean(dim=3).mean(dim=2), x.amax(dim=(2, 3)), par], dim=1)
fin = torch.cat([x, feat.view(N, -1, 1, 1).expand(N, feat.shape[1], H, W)], dim=1)
c1, c2, c3 = net.srcnn[0], net.srcnn[2], net.srcnn[4]
h1 = torch.relu(F.conv2d(fin, c1.weight, c1.bias, padding=4))
h2 = torch.relu(F.conv2d(h1, c2.weight, c2.bias, padding=2))
h3 = F.conv2d(h2, c3.weight, c3.bias, padding=2)
print('cpu module vs golden', ((x + h3) - T(g['srcnn_res3_y'])).abs().max().item())
for name, en in (('cudnn', True), ('native', False)):
    torch.backends.cudnn.enabled = en
    d = lambda t: t.detach().cuda()
    g1 = torch.relu(F.conv2d(d(fin), d(c1.weight), d(c1.bias), padding=4))
    print(name, 'conv1 err', (g1.cpu() - h1).abs().max().item(), 'scale', h1.abs().max().item())
    g2 = torch.relu(F.conv2d(d(h1), d(c2.weight), d(c2.bias), padding=2))
    print(name, 'conv2 err', (g2.cpu() - h2).abs().max().item(), 'scale', h2.abs().max().item())
    g3 = F.conv2d(d(h2), d(c3.weight), d(c3.bias), padding=2)
    print(name, 'conv3 err', (g3.cpu() - h3).abs().max().item(), 'scale', h3.abs().max().item())
