import torch, torch.nn.functional as F
torch.manual_seed(0)
x = torch.rand(2, 15, 20, 24); w = torch.randn(64, 15, 9, 9) * 0.05; b = torch.randn(64) * 0.05
ref = F.conv2d(x.double(), w.double(), b.double(), padding=4).float()
cpu = F.conv2d(x, w, b, padding=4)
print('cpu fp32 vs fp64', (cpu - ref).abs().max().item())
for name, setter in (('default', lambda: None),
                     ('allow_tf32=False', lambda: setattr(torch.backends.cudnn, 'allow_tf32', False)),
                     ('matmul_prec highest', lambda: torch.set_float32_matmul_precision('highest')),
                     ):
    setter()
    y = F.conv2d(x.cuda(), w.cuda(), b.cuda(), padding=4).cpu()
    print(name, (y - ref).abs().max().item())
try:
    print('conv.fp32_precision =', torch.backends.cudnn.conv.fp32_precision)
    torch.backends.cudnn.conv.fp32_precision = 'ieee'
    y = F.conv2d(x.cuda(), w.cuda(), b.cuda(), padding=4).cpu()
    print('ieee', (y - ref).abs().max().item())
except Exception as e:
    print('no fp32_precision api', e)
torch.backends.cudnn.enabled = False
y = F.conv2d(x.cuda(), w.cuda(), b.cuda(), padding=4).cpu()
print('cudnn disabled', (y - ref).abs().max().item())
