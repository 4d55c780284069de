"""`DartsFtModel` -- DARTS search with ONLINE PROXY FINE-TUNING (codes/models/darts_ft_model.py:20-368,
driven by codes/train_ft.py): the search loop of `DartsModel`, plus a FIFO memory of the sRGB intermediates
the supernet produced, and every `ft_interval` iterations a few Adam steps that pull each flagged proxy CNN
(crysisengine / whiteworld / bilateral / median / fastnlm) back towards its classical original on those intermediates
with random parameters; the tuned weights are then copied into every sRGB step of the supernet.

What it runs on: proxy forward / data gradient = tcgen05 convolution, weight gradient = `risp_conv2d_bwd_weight`,
targets = the classical CUDA ops (`Origin*`).  B200 changes: the FIFO stays in HBM (the reference moves every
intermediate to the host and back, :174-175/:199); proxy gradients of all ranks are averaged in one flattened
all-reduce (the reference wraps each proxy in DDP, :83-85).
"""
import logging
import random

import torch

from . import dist as D
from .modules import tools_origin as TO
from .search import DartsModel

logger = logging.getLogger('base')

PARAM_NUM = {'reinhard': 2, 'crysisengine': 1, 'filmic': 2, 'whiteworld': 1, 'bilateral': 3, 'median': 1, 'fastnlm': 3}


class DartsFtModel(DartsModel):
    def __init__(self, opt):
        super().__init__(opt)
        ft = opt['proxy_ft_params']
        self.memory_size = ft['memory_size']
        self.memory_bytes = int(ft.get('memory_bytes', 16 << 30))      # device FIFO cap (default 16 GiB of the 180 GB)
        self.ft_steps = ft['ft_steps']
        self.ft_data = []
        t = opt['train']
        targets = {'reinhard': TO.OriginToneReinhard, 'crysisengine': TO.OriginToneCrysis, 'filmic': TO.OriginToneFilmic,
                   'whiteworld': TO.OriginWbWhiteworld, 'bilateral': TO.OriginNoiseBilateral, 'median': TO.OriginNoiseMedian,
                   'fastnlm': TO.OriginNoiseFastnlm}
        self.ft_nets = []          # [name, proxy, target, optimizer]; the proxy is the LAST step's instance (:79)
        for (name, flag), proxy in zip(self.netG.proxy_ft_flag, self.netG.all_modules[-1]):
            if not flag:
                continue
            if name not in targets:
                logger.warning('proxy fine-tuning of %s skipped: its original is outside the rebuilt path', name)
                continue
            optimizer = torch.optim.Adam(proxy.parameters(), lr=t['lr_G'], betas=(t['beta1'], t['beta2']))
            self.ft_nets.append([name, proxy, targets[name](), optimizer])

    def optimize_parameters(self):
        """pass #5 + the FIFO update (:159-183).  Only the 3-plane (sRGB) intermediates are remembered."""
        super().optimize_parameters()
        self.ft_data.extend(t.detach().clone() for t in self.netG.intermediate_results if t.shape[1] == 3)
        if len(self.ft_data) > self.memory_size:
            self.ft_data = self.ft_data[len(self.ft_data) - self.memory_size:]
        # the reference keeps this FIFO on the host; here it lives in HBM, so it is also capped by bytes (oldest first)
        cap = self.memory_bytes
        total = sum(t.numel() * 4 for t in self.ft_data)
        while len(self.ft_data) > 1 and total > cap:
            total -= self.ft_data.pop(0).numel() * 4

    def _module_grads(self, loss, params):
        return D.allreduce_mean_flat(torch.autograd.grad(loss, params, allow_unused=True))

    def finetune_proxies(self):
        """:185-231.  Same host RNG calls as the reference (random.random() for the memory index, torch.rand on
        the CPU generator for the parameters), so a seeded run draws the same samples."""
        name_net = {}
        last_losses = {}
        for name, proxy, target, optimizer in self.ft_nets:
            n = len(self.ft_data)
            if n == 0:
                logger.warning('[Warning] Data is not ready for proxy fine-tuning!')
                continue
            params = [p for p in proxy.parameters()]
            # weights are trainable only inside this loop: the search passes must not pay for weight gradients
            for p in params:
                p.requires_grad_(True)
            for _ in range(self.ft_steps):
                data = self.ft_data[int(random.random() * n)]
                param = torch.rand(1, PARAM_NUM[name]).repeat(data.shape[0], 1).to(self.device)
                out = proxy(data, param)
                with torch.no_grad():
                    gt = target(data, param)
                loss = self._loss(out, gt)
                optimizer.zero_grad()
                for p, g in zip(params, self._module_grads(loss, params)):
                    p.grad = g
                optimizer.step()
                last_losses[name] = loss.detach()
            for p in params:
                p.requires_grad_(False)
            name_net[name] = proxy
        if name_net:
            self._load_proxy_nets(name_net)
        self.log_dict.update(('ft_' + k, v) for k, v in last_losses.items())

    def _load_proxy_nets(self, name_net):
        """`load_proxy_nets` (:194-209 of the supernet file) for the proxies that were tuned; the skipped ones
        keep their weights."""
        net = self.netG
        for idx, (name, flag) in enumerate(net.proxy_ft_flag):
            if flag and name in name_net:
                state = name_net[name].state_dict()
                for k in range(net.n_step):
                    net.all_modules[-1 - k][idx].load_state_dict(state)

    def save(self, path_fmt):
        """:154-158 -- the supernet and every tuned proxy; `path_fmt` has one `{}` for the name."""
        self.save_network(path_fmt.format('G'))
        for name, proxy, _, _ in self.ft_nets:
            torch.save({k: v.cpu() for k, v in proxy.state_dict().items()}, path_fmt.format(name))
