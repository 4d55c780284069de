"""DARTS supernet with online pruning -- `SuperPruneFifteenDemosFourBayerTwo(n_step, threshold, module_path)`
(codes/models/modules/super_prune_fifteen_demos_four_bayer_two.py:13-230): a Bayer step (2 candidates), a
demosaic step (4) and `n_step` sRGB steps of 15 candidates each.  Same attributes (`all_modules`,
`all_params`, `all_alphas`, `trainable_params`, `pruned_paths`, ...), properties and registered parameter
names (`alpha_bayer`, `alpha_demosaic`, `alpha_step{k}`, `param_step{k}_{name}`).

Mixed-op evaluation (reference :175-214) re-designed for the device:
  * softmax / prune / renormalise of every step's alpha is a one-warp kernel; ONE device->host copy of
    all post-prune weights per forward replaces the per-step `.item()` (:193) and the per-candidate
    `if prob < 1e-9` syncs (:197) -- it is only needed to skip the CNN work of pruned candidates;
  * the classical candidates of an sRGB step (gamma, grayworld, skip, wbmanual, wbquadratic, gtmmanual) are
    evaluated in registers inside `ops.mixed_op` from a single read of x; only CNN candidates are
    materialised; the weighted sum, all K alpha dot-products and the parameter gradients are one kernel each way;
  * quirks kept: mask from detached probs and detached normaliser (:188-192), zero (not None) gradients for
    the parameters of pruned candidates (:199-201), Skip aliasing, GtmManual using batch row 0.
"""
import os

import torch
import torch.nn as nn

from .. import ops
from . import registry as R
from . import tools_origin as T
from . import tools_proxy as P

SRGB_NAMES = R.NAMES['sRGB'][:R.N_SRGB_ORIGIN]
# classical sRGB candidates evaluated in registers by the mixed-op kernel: (pool index, chain op)
SRGB_CLASSICAL = [(0, ('gamma', 0)), (4, ('gain_clip', 0)), (9, ('skip', 0)), (10, ('gain', 0)), (12, ('poly10', 0)),
                  (13, ('gtm', 4))]


class _PlaneMeanFn(torch.autograd.Function):
    """Per-plane mean with the broadcast backward (gray-world gains depend on the image)."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = x.shape
        return ops.plane_stats(x)[..., 1].contiguous()

    @staticmethod
    def backward(ctx, dm):
        N, C, H, W = ctx.shape
        return (dm / float(H * W)).view(N, C, 1, 1).expand(N, C, H, W)


def grayworld_gains(x):
    m = _PlaneMeanFn.apply(x)
    return m.mean(dim=1, keepdim=True) / torch.clamp(m, min=1e-6)


class SuperPruneFifteenDemosFourBayerTwo(nn.Module):
    def __init__(self, n_step, threshold, module_path, weight_seed=None):
        super().__init__()
        self.threshold = threshold
        self.n_step = n_step
        self.middle_results = None
        self.pruned_paths = [0] * (n_step + 2)
        seed = [weight_seed]

        def net(name):
            if seed[0] is None:
                return R.build_net(name, module_path + R.PROXY_CKPT[name][1])
            m = R.build_net(name, None, seed[0])
            seed[0] += 1
            return m

        empty = lambda: nn.Parameter(torch.Tensor([]))
        self.all_modules, self.all_params, self.all_alphas = [], [], []
        self.trainable_params, self.param_and_alpha = [], []

        # Bayer step (:57-74)
        self.all_modules.append(nn.ModuleList([net('path_bayer'), T.Skip()]))
        self.all_params.append(nn.ParameterList([empty(), empty()]))
        self.alpha_bayer = nn.Parameter(torch.zeros((2,)))
        self.all_alphas.append(self.alpha_bayer)
        # demosaic step (:77-98)
        self.all_modules.append(nn.ModuleList([T.DemosaicNearest(), net('bilinear'), net('laplacian'), T.DemosaicNet()]))
        self.all_params.append(nn.ParameterList([empty() for _ in range(4)]))
        self.alpha_demosaic = nn.Parameter(torch.zeros((4,)))
        self.all_alphas.append(self.alpha_demosaic)
        # sRGB steps (:101-171)
        for k in range(n_step):
            mods, pars = [], []
            for name in SRGB_NAMES:
                if name in R.CLASSICAL:
                    mods.append(R.CLASSICAL[name]())
                elif name == 'gtmmanual':
                    mods.append(T.GtmManual(4))
                else:
                    mods.append(net(name))
                logits = R.DEFAULT_LOGITS[name]
                if len(logits):
                    key = 'param_step{}_{}'.format(k + 1, name)
                    setattr(self, key, nn.Parameter(torch.Tensor(logits)))
                    pars.append(getattr(self, key))
                else:
                    pars.append(empty())
            setattr(self, 'alpha_step{}'.format(k + 1), nn.Parameter(torch.zeros((15,))))
            self.trainable_params += pars
            self.all_modules.append(nn.ModuleList(mods))
            self.all_params.append(nn.ParameterList(pars))
            self.all_alphas.append(getattr(self, 'alpha_step{}'.format(k + 1)))
        self.param_and_alpha = self.trainable_params + self.all_alphas
        # like the reference, the per-step containers are plain Python lists: candidate-net weights stay out
        # of named_parameters()/state_dict()/DDP
        self._chain = ops.Chain([op for _, op in SRGB_CLASSICAL])
        self.n_streams = int(os.environ.get('RISP_SEARCH_STREAMS', '4'))
        self.use_bank = os.environ.get('RISP_SRCNN_BANK', '1') != '0'
        self._streams = []

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        for ml in self.all_modules:
            ml._apply(fn, *a, **k)
        return self

    # ------------------------------------------------------------------------------------------------------
    def _post_probs(self):
        """post-prune weights of every step (device, differentiable) + one host copy for control flow."""
        posts = [ops.alpha_prune(a, self.threshold) for a in self.all_alphas]
        host = torch.cat([p.detach() for p in posts]).cpu().tolist()      # plain floats: the control flow below is pure Python
        out, o = [], 0
        for i, p in enumerate(posts):
            h = host[o:o + p.numel()]
            o += p.numel()
            self.pruned_paths[i] = sum(1 for v in h if v == 0)
            out.append((p, h))
        return out

    @staticmethod
    def _dummy(par_list, idxs):
        """0 * sum(params) of pruned candidates: zero instead of missing gradients (:199-201)."""
        live = [par_list[i].view(-1) for i in idxs if par_list[i].nelement() > 0]
        if not live:
            return None
        return (live[0] if len(live) == 1 else torch.cat(live)).sum() * 0.0

    # ------------------------------------------------------------------------------------------------------
    def _fan_out(self, jobs):
        """Run the materialised (CNN) candidates of one step: jobs = [callable -> tensor].  With `self.n_streams` > 1 they are
        spread over side streams -- the candidates of a step are independent (reference :195-210 evaluates them one after the
        other), and at the reference's 4 x 256^2 batch a single candidate's grid leaves a partial last wave on 148 SMs, which
        the other candidates' CTAs fill.  Autograd replays each candidate's backward on the stream its forward ran on."""
        n = min(self.n_streams, len(jobs)) if self._small_batch else 1      # big batches fill the SMs on their own (and
        if n <= 1 or not torch.cuda.is_available():                          # per-stream allocator pools would cost HBM)
            return [j() for j in jobs]
        main = torch.cuda.current_stream()
        if len(self._streams) < n:
            self._streams += [torch.cuda.Stream() for _ in range(n - len(self._streams))]
        fork = torch.cuda.Event()
        fork.record(main)
        outs = []
        for k, job in enumerate(jobs):
            s = self._streams[k % n]
            s.wait_event(fork)
            with torch.cuda.stream(s):
                outs.append(job())
        for s in self._streams[:n]:
            main.wait_stream(s)
        for o in outs:
            for t in (o if isinstance(o, (list, tuple)) else (o,)):
                t.record_stream(main)
        return outs

    def _srcnn_bank(self, s, mods):
        banks = self.__dict__.setdefault('_banks', {})
        if s not in banks:
            index = [i for i in range(15) if isinstance(mods[i], P.SRCNNRes)]
            banks[s] = P.SRCNNResBank([mods[i] for i in index])
            banks[s].index = index
        return banks[s]

    def _index(self, idx, device):
        """device index tensor of a candidate list (cached: a Python list index costs a host->device copy every time)."""
        cache = self.__dict__.setdefault('_idx_cache', {})
        k = (tuple(idx), str(device))
        if k not in cache:
            cache[k] = torch.tensor(list(idx), dtype=torch.int64, device=device)
        return cache[k]

    def _table_affine(self, device):
        """[0,1] -> kernel range of the classical candidates' parameters as one affine map over the 37 table entries."""
        t = self.__dict__.get('_affine')
        if t is None or t[0].device != device:
            scale = torch.tensor([1.0] + [5.0] * 3 + [10.0] * 30 + [1.0] * 3, device=device).view(1, -1)
            shift = torch.tensor([0.0] + [0.0] * 3 + [-5.0] * 30 + [0.0] * 3, device=device).view(1, -1)
            t = self._affine = (scale, shift)
        return t

    def forward(self, x):
        """x: (N,1,H,W) RGGB -> (N,3,H,W) BGR."""
        N = x.size(0)
        self._small_batch = x.numel() <= (1 << 20)
        self.middle_results = []
        posts = self._post_probs()

        # ---- Bayer step: [path_bayer, skip] -------------------------------------------------------------------
        post, host = posts[0]
        mods = self.all_modules[0]
        ext, widx = [], [1]                                   # kernel order: [skip, materialised...]
        if not host[0] < 1e-9:
            ext.append(mods[0](x, None))
            widx.append(0)
        if host[1] < 1e-9:                                    # skip pruned: zero weight keeps the kernel generic
            pass
        y = ops.mixed_op(x, ops.Chain(['skip']), None, post.index_select(0, self._index(widx, post.device)), ext)
        self.middle_results.append(y)
        x = y

        # ---- demosaic step: every candidate maps 1 -> 3 planes, all are materialised ---------------------------
        post, host = posts[1]
        mods = self.all_modules[1]
        jobs, widx = [], []
        for k in range(4):
            if host[k] < 1e-9:
                continue
            jobs.append(lambda k=k, x=x: mods[k](x, None))
            widx.append(k)
        ext = self._fan_out(jobs)
        y = ops.mixed_op(ext[0].detach(), ops.Chain([]), None, post.index_select(0, self._index(widx, post.device)), ext)
        self.middle_results.append(y)
        x = y

        # ---- sRGB steps ----------------------------------------------------------------------------------------
        cls_idx = [i for i, _ in SRGB_CLASSICAL]
        for s in range(self.n_step):
            post, host = posts[2 + s]
            mods, pars = self.all_modules[2 + s], self.all_params[2 + s]
            # sigmoid of ALL parameter logits of the step in one launch (one cat, one sigmoid; the per-candidate values are views)
            sizes = [p.nelement() for p in pars]
            sg_all = torch.sigmoid(torch.cat([p.view(-1) for p in pars if p.nelement()]))
            # one split instead of a slice per candidate: its backward is ONE cat of the candidates' gradients (a slice each
            # costs a zero-fill, a copy and an accumulation add per candidate and pass)
            parts = torch.split(sg_all, [n_ for n_ in sizes if n_])
            slot, o = {}, 0
            for i_, n_ in enumerate(sizes):
                if n_:
                    slot[i_] = o
                    o += 1
            sig = lambda i: parts[slot[i]].view(1, -1)
            # kernel-level parameter table of the classical candidates, one row per image:
            # [gamma | grayworld gains | wbmanual p*5 | wbquadratic p*10-5 | gtm knots]  (tools_origin.py:214, :326)
            gw = grayworld_gains(x) if not host[4] < 1e-9 else torch.ones((N, 3), device=x.device)
            scale, shift = self._table_affine(x.device)
            row = torch.addcmul(shift, torch.cat([sig(0), sig(10), sig(12), sig(13)], dim=1), scale)       # (1, 37)
            table = torch.cat([row[:, :1].expand(N, 1), gw, row[:, 1:].expand(N, 36)], dim=1)
            jobs, widx = [], list(cls_idx)
            pruned = [i for i in range(15) if host[i] < 1e-9]
            # the SRCNNRes proxies of the step (same layer shapes once their constant channels are folded): one grouped job
            bank = self._srcnn_bank(s, mods)
            banked = [i for i in bank.index if not host[i] < 1e-9] if (self.use_bank and bank.usable(x)) else []
            if len(banked) < 2:
                banked = []
            slot_of = {i: k for k, i in enumerate(bank.index)}
            where = {}                                     # candidate index -> (job, position in the job's result list)
            if banked:
                where.update({i: (0, k) for k, i in enumerate(banked)})
                jobs.append(lambda x=x, b=banked: bank(x, [sig(i) for i in b], [slot_of[i] for i in b]))
            for i in range(15):
                if i in cls_idx or host[i] < 1e-9:
                    continue
                widx.append(i)
                if i in where:
                    continue
                par = pars[i]
                par_tensor = None if par.nelement() == 0 else sig(i).expand(N, -1)
                where[i] = (len(jobs), None)
                jobs.append(lambda i=i, x=x, p=par_tensor: mods[i](x, p))
            if jobs:
                P.prepare_shared(x)          # blocked copy + statistics of the stage input, once, on the main stream
            res = self._fan_out(jobs)
            P.release_shared(x)
            ext = []
            for i in widx[len(cls_idx):]:
                j, k = where[i]
                ext.append(res[j] if k is None else res[j][k])
            w = post.index_select(0, self._index(widx, post.device))
            dummy = self._dummy(pars, pruned)
            if dummy is not None:
                w = w + dummy
            y = ops.mixed_op(x, self._chain, table, w, ext)
            self.middle_results.append(y)
            x = y
        return x

    @property
    def trainable_parameters(self):
        return self.trainable_params

    @property
    def parameters_and_alpha(self):
        return self.param_and_alpha

    @property
    def alphas(self):
        return self.all_alphas

    @property
    def intermediate_results(self):
        return self.middle_results


def mixed_op_probe(timed, N, H, W, hbm_peak_gbs, K_ext=9):
    """Time the fused mixed-op of ONE sRGB step in isolation at a size where the HBM roofline means something
    (bench.py --workload search): the 6 classical candidates are evaluated in registers from one read of x, the K_ext = 9
    CNN candidates are materialised inputs.  Algorithmic bytes: forward 12 (x) + 12*K_ext + 12 (y) = 132 B/px,
    backward 12 (dy) + 12 (x) + 12*K_ext (dot products with the CNN outputs) + 12 (dx) + 12*K_ext (the upstream gradients
    w_i*dy of the CNN candidates, written by the same kernel) = 252 B/px."""
    dev = torch.device('cuda')
    chain = ops.Chain([op for _, op in SRGB_CLASSICAL])
    x = torch.rand(N, 3, H, W, device=dev).requires_grad_()
    ext = [torch.rand(N, 3, H, W, device=dev).requires_grad_() for _ in range(K_ext)]
    ident = [0.0] * 30
    ident[6] = ident[17] = ident[28] = 1.0
    table = torch.tensor([[0.5, 1.0, 1.0, 1.0, 1.05, 1.0, 0.95] + ident + [0.25, 0.5, 0.75]], device=dev).repeat(N, 1).requires_grad_()
    w = torch.full((len(SRGB_CLASSICAL) + K_ext,), 1.0 / 15, device=dev).requires_grad_()
    dy = torch.randn(N, 3, H, W, device=dev)
    px = N * H * W
    with torch.no_grad():
        ms_f = timed(lambda: ops.mixed_op(x, chain, table, w, ext), 5, 2) / 5
    y = ops.mixed_op(x, chain, table, w, ext)
    ms_b = timed(lambda: torch.autograd.grad(y, [x, table, w] + ext, dy, retain_graph=True), 5, 2) / 5
    bf, bb = 12 + 12 * K_ext + 12, 12 + 12 + 12 * K_ext + 12 + 12 * K_ext
    return {'shape': [N, 3, H, W], 'K_classical': len(SRGB_CLASSICAL), 'K_ext': K_ext,
            'fwd': {'ms': round(ms_f, 4), 'bytes_per_px': bf, 'GBps': round(bf * px / ms_f / 1e6, 1), 'frac': round(bf * px / ms_f / 1e6 / hbm_peak_gbs, 4)},
            'bwd': {'ms': round(ms_b, 4), 'bytes_per_px': bb, 'GBps': round(bb * px / ms_b / 1e6, 1), 'frac': round(bb * px / ms_b / 1e6 / hbm_peak_gbs, 4)},
            'kernel': 'risp::mixed_fwd_kernel / mixed_bwd_kernel'}
