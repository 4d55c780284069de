"""`IspModel` -- proxy tuning / inference driver of a fixed pipeline (codes/models/isp_model.py:15-151):
same public methods (`feed_data`, `optimize_parameters`, `test`, `get_current_log`, `save`, `load`) and
option keys.  When the pipeline is a classical demosaic followed by differentiable per-pixel stages,
`optimize_parameters` runs the WHOLE step -- forward, MSE, gradients of every stage parameter -- as one
pass over the frame (`ops.pipeline_mse`, 16 B/px); otherwise it runs the planned segments with autograd.

B200 additions: `feed_data` copies into persistent device buffers (and decodes integer sensor codes on the device),
and the fused step -- parameter table, the pipeline kernel, the gradient all-reduce, Adam -- is captured once into
a CUDA graph and replayed (`opt['cuda_graph']`, default on): the ~25 tiny launches around the one big kernel are
launch-latency-bound and cost ~10 % of a 48 MP step when issued one by one."""
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
            self.use_graph = bool(opt.get('cuda_graph', True))
            # capturable: the step counter lives on the device, so the update can be replayed from a CUDA graph;
            # fused: one multi-tensor kernel for the whole update instead of ~7 foreach launches
            self.optimizer_G = torch.optim.Adam([p for p in self.netG.trainable_parameters], t['lr_G'], (t['beta1'], t['beta2']),
                                                capturable=self.use_graph, fused=self.use_graph)
            self._graph, self._graph_key, self._eager_steps = None, None, 0
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
        self.img = self._to_device('img', img, self.opt.get('raw_white_level', 1023.))
        self.gt = self._to_device('gt', gt, 255.)
        self._output = None

    def _to_device(self, slot, t, denom):
        """fp32 tensors are copied as they are (the reference's contract); integer sensor / display codes
        (uint8, uint16/int16) are copied as codes and normalised on the device (`ops.decode_codes`).  The device
        buffers persist across calls of the same shape (stable addresses: the captured step reads them)."""
        bufs = self.__dict__.setdefault('_bufs', {})
        if t.is_cuda and t.dtype == torch.float32:
            return t                                        # caller-owned device tensor: used in place
        dst = bufs.get(slot)
        if dst is None or dst.shape != t.shape:
            dst = bufs[slot] = torch.empty(t.shape, device=self.device, dtype=torch.float32)
        if t.dtype.is_floating_point:
            dst.copy_(t, non_blocking=True)
            return dst
        key = slot + '_codes'
        codes = bufs.get(key)
        if codes is None or codes.shape != t.shape or codes.dtype != t.dtype:
            codes = bufs[key] = torch.empty(t.shape, device=self.device, dtype=t.dtype)
        codes.copy_(t, non_blocking=True)
        return ops.decode_codes(codes, denom, out=dst)

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
        """isp_model.py:132-141.  Eager for the first two calls (allocator / optimizer-state warm-up), then the fused
        step is captured into a CUDA graph and replayed for as long as buffers, shapes and learning rates stay put."""
        if self.use_graph and self._graph_replay():
            return
        self._step_eager()

    def _graph_signature(self):
        return (self.img.data_ptr(), self.gt.data_ptr(), tuple(self.img.shape), tuple(self.gt.shape),
                tuple(float(g['lr']) for g in self.optimizer_G.param_groups))

    def _graph_replay(self):
        key = self._graph_signature()
        if self._graph is not None and key == self._graph_key:
            g_loss, g_update = self._graph
            g_loss.replay()
            if g_update is not None:                       # data parallel: the collective stays outside the graphs
                self._allreduce_grads()
                g_update.replay()
            self.log_dict['loss'] = self.l_pix
            return True
        self._graph = None
        stable, self._last_key = key == getattr(self, '_last_key', None), key
        # capture only once the same buffers / shapes / learning rates have been seen twice in a row (a loop that
        # feeds fresh device tensors every step would otherwise re-capture every step)
        if not stable or self._eager_steps < 2 or not (self.netG.fuse and self.loss_type == 'l2') or \
                self.netG.fused_mse_step_plan() is None:
            return False
        try:
            torch.cuda.synchronize()
            g_loss, g_update = torch.cuda.CUDAGraph(), None
            if not D.is_dist():
                with torch.cuda.graph(g_loss):
                    self._step_eager(count=False)
            else:
                # NCCL kernels are kept out of the graphs (a captured collective ties the communicator's lifetime to
                # the graph's): [table, pipeline kernel, finaliser, backward] | all-reduce | [Adam]
                with torch.cuda.graph(g_loss):
                    self._loss_and_grads()
                g_update = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_update, pool=g_loss.pool()):
                    self.optimizer_G.step()
            self._graph, self._graph_key = (g_loss, g_update), key
            return self._graph_replay()                    # capturing does not execute: run the step now
        except Exception as e:                             # capture is an optimisation: fall back to eager launches
            import logging
            logging.getLogger('base').warning('CUDA-graph capture of the tuning step failed (%s); running eagerly', e)
            self.use_graph, self._graph = False, None
            torch.cuda.synchronize()
            return False

    def _loss_and_grads(self):
        l_pix = self._fused_loss()
        if l_pix is None:
            self._output = self.netG(self.img)
            l_pix = (ops.l1_loss if self.loss_type == 'l1' else ops.mse_loss)(self._output, self.gt)
        self.optimizer_G.zero_grad()
        l_pix.backward()
        self.l_pix = l_pix.detach()
        self.log_dict['loss'] = self.l_pix             # device scalar; `.item()` it when logging

    def _allreduce_grads(self):
        if D.is_dist():
            ps = [p for p in self.netG.trainable_parameters if p.grad is not None]
            for p, g in zip(ps, D.allreduce_mean_flat([p.grad for p in ps])):
                p.grad.copy_(g)

    def _step_eager(self, count=True):
        if count:
            self._eager_steps += 1
        self._loss_and_grads()
        self._allreduce_grads()
        self.optimizer_G.step()

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
