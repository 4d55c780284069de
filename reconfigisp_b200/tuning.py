"""`IspModel` -- proxy tuning / inference driver of a fixed pipeline (codes/models/isp_model.py:15-151):
same public methods (`feed_data`, `optimize_parameters`, `test`, `get_current_log`, `save`, `load`) and
option keys.  When the pipeline is a classical demosaic followed by differentiable per-pixel stages,
`optimize_parameters` runs the WHOLE step -- forward, MSE, gradients of every stage parameter -- as one
pass over the frame (`ops.pipeline_mse`, 16 B/px); otherwise it runs the planned segments with autograd."""
from collections import OrderedDict

import torch
import torch.nn as nn

from . import dist as D
from . import ops
from .networks import define_G


class IspModel:
    def __init__(self, opt):
        self.opt = opt
        self.device = torch.device('cuda')
        if not torch.cuda.is_available():
            raise RuntimeError('IspModel needs a CUDA device (no CPU fallback)')
        self.is_train = opt.get('is_train', True)
        self.netG = define_G(opt).to(self.device)
        self.netG_attr = self.netG
        self.img = self.gt = self.val_img = self.val_gt = self.meta = None
        self._output = None
        self.log_dict = OrderedDict()
        self.optimizers, self.schedulers = [], []
        if self.is_train:
            t = opt['train']
            self.loss_type = t['pixel_criterion']
            if self.loss_type not in ('l1', 'l2'):
                raise NotImplementedError('pixel_criterion %r' % self.loss_type)
            self.optimizer_G = torch.optim.Adam([p for p in self.netG.trainable_parameters], t['lr_G'], (t['beta1'], t['beta2']))
            self.optimizers.append(self.optimizer_G)
            if t.get('lr_scheme', 'MultiStepLR') == 'MultiStepLR':
                self.schedulers.append(torch.optim.lr_scheduler.MultiStepLR(self.optimizer_G, t['lr_steps'], t['lr_gamma']))
            else:
                raise NotImplementedError('MultiStepLR learning rate scheme is enough.')
        self.load()

    # -- data ------------------------------------------------------------------------------------------------
    def feed_data(self, data):
        """(img, gt[, val_img, val_gt]) host or device tensors; pinned host tensors copy asynchronously."""
        if len(data) == 2:
            img, gt = data
        elif len(data) == 4:
            img, gt, val_img, val_gt = data
            self.val_img = val_img.to(self.device, non_blocking=True)
            self.val_gt = val_gt.to(self.device, non_blocking=True)
        else:
            raise ValueError('Invalid data format.')
        self.img = self._to_device(img, self.opt.get('raw_white_level', 1023.))
        self.gt = self._to_device(gt, 255.)
        self._output = None

    def _to_device(self, t, denom):
        """fp32 tensors are copied as they are (the reference's contract); integer sensor / display codes
        (uint8, uint16/int16) are copied as codes and normalised on the device (`ops.decode_codes`)."""
        d = t.to(self.device, non_blocking=True)
        return d if d.dtype.is_floating_point else ops.decode_codes(d, denom)

    # -- training step ---------------------------------------------------------------------------------------
    def _fused_loss(self):
        plan = self.netG.fused_mse_step_plan() if (self.netG.fuse and self.loss_type == 'l2') else None
        H, W = self.img.shape[2:]
        if plan is None or H % 2 or W % 4:
            return None
        dm_kind, chain, keep = plan
        table = self.netG._segment_table(keep, self.img.shape[0])
        return ops.pipeline_mse(table, self.img, self.gt, dm_kind, chain)

    def optimize_parameters(self):
        l_pix = self._fused_loss()
        if l_pix is None:
            self._output = self.netG(self.img)
            l_pix = (ops.l1_loss if self.loss_type == 'l1' else ops.mse_loss)(self._output, self.gt)
        self.optimizer_G.zero_grad()
        l_pix.backward()
        if D.is_dist():
            ps = [p for p in self.netG.trainable_parameters if p.grad is not None]
            for p, g in zip(ps, D.allreduce_mean_flat([p.grad for p in ps])):
                p.grad.copy_(g)
        self.optimizer_G.step()
        self.l_pix = l_pix.detach()
        self.log_dict['loss'] = self.l_pix             # device scalar; `.item()` it when logging

    @property
    def output(self):
        """The output image of the last step (computed on demand when the fused step skipped it)."""
        if self._output is None and self.img is not None:
            with torch.no_grad():
                self._output = self.netG(self.img)
        return self._output

    def test(self):
        self._output = self.netG(self.img)
        return self._output, self.netG.intermediate_results

    def get_current_log(self):
        return self.log_dict

    # -- checkpoints (base_model.py:77-97): {iter}_G.pth = CPU state_dict, 'module.' prefixes stripped on load --
    def save_network(self, path):
        torch.save(OrderedDict((k, v.cpu()) for k, v in self.netG.state_dict().items()), path)

    def load_network(self, path, strict=True):
        state = torch.load(path, map_location='cpu')
        clean = OrderedDict((k[7:] if k.startswith('module.') else k, v) for k, v in state.items())
        self.netG.load_state_dict(clean, strict=strict)

    def load(self):
        path = (self.opt.get('path') or {}).get('pretrain_model_G')
        if path is not None:
            self.load_network(path, (self.opt.get('path') or {}).get('strict_load', True))

    def save(self, path):
        self.save_network(path)
