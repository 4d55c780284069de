"""`DartsModel` -- the second-order DARTS search driver (codes/models/darts_model.py:19-330): netG + netV
(two independent supernets), SGD-momentum on the module parameters, Adam on the architecture weights, and
per iteration  optimize_alphas() [virtual step -> unrolled validation loss -> finite-difference Hessian]
then optimize_parameters(): five supernet forward+backward passes.

Kept from the reference, because they change the numbers: the virtual step  p' = p - xi*(mu*buf + g)
(:212-218), eps = 0.01/||dp|| with the 1e-6 guard (:278-283), the hessian expression `(pos-neg)/2.*eps`
(:323 -- it MULTIPLIES by eps), the NaN guard (:260-263), alpha.grad = dalpha - xi*h (:265).

Changed for B200: every gradient list of a pass (module-parameter and alpha gradients alike) is averaged
across ranks as ONE flattened buffer in one NCCL all-reduce (`dist.allreduce_mean_flat`); the reference
reduces only pass #5 through DDP and leaves the alpha gradients rank-local."""
from collections import OrderedDict

import torch

from . import dist as D
from . import ops
from .networks import define_G


class DartsModel:
    def __init__(self, opt):
        self.opt = opt
        if not torch.cuda.is_available():
            raise RuntimeError('DartsModel needs a CUDA device (no CPU fallback)')
        self.device = torch.device('cuda')
        self.netG = define_G(opt).to(self.device)
        self.netV = define_G(opt).to(self.device)          # "build another net instead of deep copy" (:28-29)
        self.netG_attr = self.netG
        t = opt['train']
        self.loss_type = t['pixel_criterion']
        if self.loss_type not in ('l1', 'l2'):
            raise NotImplementedError('pixel_criterion %r (latency / local_global losses have no shipped producer)' % self.loss_type)
        self.momentum_G = t['momentum_G']
        self.lr_meta = t['lr_meta']
        self.optimizer_G = torch.optim.SGD(self.netG.trainable_parameters, t['lr_G'], momentum=self.momentum_G)
        self.optimizer_alpha = torch.optim.Adam(self.netG.alphas, lr=t['lr_G'], betas=(t['beta1'], t['beta2']))
        self.optimizers = [self.optimizer_G, self.optimizer_alpha]
        self.img = self.gt = self.val_img = self.val_gt = None
        self.output = None
        self.val_loss = None
        self.log_dict = OrderedDict()

    def _loss(self, y, gt):
        return ops.l1_loss(y, gt) if self.loss_type == 'l1' else ops.mse_loss(y, gt)

    def feed_data(self, data):
        if len(data) == 2:
            img, gt = data
        elif len(data) == 4:
            img, gt, val_img, val_gt = data
            self.val_img = val_img.to(self.device, non_blocking=True)
            self.val_gt = val_gt.to(self.device, non_blocking=True)
        else:
            raise ValueError('Invalid data format.')
        self.img = img.to(self.device, non_blocking=True)
        self.gt = gt.to(self.device, non_blocking=True)

    # ---------------------------------------------------------------------------------------------------------
    def _grads(self, loss, wrt):
        return D.allreduce_mean_flat(torch.autograd.grad(loss, wrt, allow_unused=True))

    def optimize_parameters(self):
        """pass #5: plain training step on the module parameters (:159-180)."""
        self.output = self.netG(self.img)
        l_pix = self._loss(self.output, self.gt)
        self.optimizer_G.zero_grad()
        ps = [p for p in self.netG.trainable_parameters if p.nelement() > 0]
        for p, g in zip(ps, self._grads(l_pix, ps)):
            p.grad = g
        self.optimizer_G.step()
        self.log_dict['loss'] = l_pix.detach()

    def virtual_step(self):
        """pass #1: p' = p - lr_meta * (momentum * buf + g), written into netV; alphas copied (:182-222).
        All parameter tensors move through multi-tensor (`torch._foreach_*`) launches: same arithmetic per element."""
        loss = self._loss(self.netG(self.img), self.gt)
        P = self.netG.trainable_parameters
        nz = [p for p in P if p.nelement() > 0]
        grads = iter(self._grads(loss, nz))
        with torch.no_grad():
            ps, vps, gs, bufs, plain = [], [], [], [], []
            for p, vp in zip(P, self.netV.trainable_parameters):
                if p.nelement() == 0:
                    continue
                g = next(grads)
                if g is None:
                    plain.append((p, vp))
                    continue
                ps.append(p)
                vps.append(vp)
                gs.append(g)
                bufs.append(self.optimizer_G.state[p].get('momentum_buffer'))
            if ps:
                if all(b is not None for b in bufs):
                    t = torch._foreach_mul(bufs, self.momentum_G)
                    torch._foreach_add_(t, gs)                       # momentum + g
                else:                                              # before the first optimizer step: buffers are 0.
                    t = [g + (0. if b is None else b * self.momentum_G) for g, b in zip(gs, bufs)]
                torch._foreach_mul_(t, self.lr_meta)
                torch._foreach_copy_(vps, torch._foreach_sub(ps, t))
            if plain:
                torch._foreach_copy_([vp for _, vp in plain], [p for p, _ in plain])
            torch._foreach_copy_(list(self.netV.alphas), list(self.netG.alphas))

    def optimize_alphas(self):
        """passes #1-#4 (:224-268)."""
        self.optimizer_alpha.zero_grad()
        self.virtual_step()
        loss = self._loss(self.netV(self.val_img), self.val_gt)                      # pass #2
        self.val_loss = loss.detach()
        v_alphas = tuple(self.netV.alphas)
        v_params = tuple(p for p in self.netV.trainable_parameters if p.nelement() > 0)
        v_grads = self._grads(loss, v_alphas + v_params)
        dalpha, dp = v_grads[:len(v_alphas)], v_grads[len(v_alphas):]
        hessian = self.compute_hessian(dp)
        with torch.no_grad():
            live = [i for i, (da, h) in enumerate(zip(dalpha, hessian)) if da is not None and h is not None]
            if live:
                hs = [hessian[i] for i in live]
                gs = torch._foreach_sub([dalpha[i] for i in live], torch._foreach_mul(hs, self.lr_meta))    # da - lr_meta * h
            for idx, alpha in enumerate(self.netG.alphas):
                if idx not in live:
                    alpha.grad = torch.zeros_like(alpha)
                else:
                    # NaN guard (:260-263) without the host sync: where() on the device
                    k = live.index(idx)
                    alpha.grad = torch.where(torch.isnan(hs[k]).any(), 0., gs[k])
        self.optimizer_alpha.step()

    def compute_hessian(self, dp):
        """passes #3/#4: finite difference of d L_trn / d alpha along dp (:270-324)."""
        nz = [p for p in self.netG.trainable_params if p.nelement() > 0]
        norm = torch.cat([w.reshape(-1) for w in dp if w is not None]).norm()
        eps = torch.where(norm < 1e-6, torch.zeros_like(norm), 0.01 / norm)          # device scalar, no sync
        ps = [p for p, d in zip(nz, dp) if d is not None]
        ds = [d for d in dp if d is not None]
        with torch.no_grad():
            step = torch._foreach_mul(ds, eps)                  # eps * d
            step2 = torch._foreach_mul(ds, 2. * eps)            # (2. * eps) * d
            torch._foreach_add_(ps, step)
        dalpha_pos = self._grads(self._loss(self.netG(self.img), self.gt), self.netG.alphas)
        with torch.no_grad():
            torch._foreach_sub_(ps, step2)
        dalpha_neg = self._grads(self._loss(self.netG(self.img), self.gt), self.netG.alphas)
        with torch.no_grad():
            torch._foreach_add_(ps, step)
            live = [i for i, (p, n) in enumerate(zip(dalpha_pos, dalpha_neg)) if p is not None and n is not None]
            h = torch._foreach_sub([dalpha_pos[i] for i in live], [dalpha_neg[i] for i in live])
            torch._foreach_div_(h, 2.)
            torch._foreach_mul_(h, eps)                         # (pos - neg) / 2. * eps
        out = [None] * len(dalpha_pos)
        for k, i in enumerate(live):
            out[i] = h[k]
        return out

    def test(self):
        self.output = self.netG(self.img)
        return self.output, self.netG.intermediate_results

    def get_current_log(self):
        return self.log_dict

    def save_network(self, path):
        torch.save(OrderedDict((k, v.cpu()) for k, v in self.netG.state_dict().items()), path)
