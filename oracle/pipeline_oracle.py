"""CPU oracle for the pipeline containers and the supernet  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Restates, on top of `oracle/isp_oracle.py`:
  * the candidate-module wrappers' parameter mappings (`tools_origin.py`, cited per function),
  * the stage registry / architecture-string grammar (`isp_universal.py:62-208`,
    `origin_universal.py:36-141`),
  * the sequential container forward (`isp_universal.py:210-232`, `origin_universal.py:143-161`),
  * the DARTS supernet forward (`super_prune_fifteen_demos_four_bayer_two.py:57-214`),
  * the DARTS second-order step (`darts_model.py:159-324`).
Pinned by `tests/golden/*.npz` (generated from the reference by `oracle/gen_golden.py`).
Also the thing `bench.py --impl reference` / `cpu_baseline` times (kind "port").
"""
import numpy as np
import torch

from . import isp_oracle as O

# ---- weight tables (state-dict key order of the reference architectures) -----------------
def _srcnn_res_shapes(P):
    return (('srcnn.0.weight', (64, 12 + P, 9, 9)), ('srcnn.0.bias', (64,)),
            ('srcnn.2.weight', (32, 64, 5, 5)), ('srcnn.2.bias', (32,)),
            ('srcnn.4.weight', (3, 32, 5, 5)), ('srcnn.4.bias', (3,)))


def _path14l_shapes(cin, cout):
    s = [('path_restore_14l.0.weight', (64, cin, 3, 3)), ('path_restore_14l.0.bias', (64,))]
    for i in range(6):
        for j in (1, 3):
            s += [('path_restore_14l.1.%d.basic.%d.weight' % (i, j), (64, 64, 3, 3)),
                  ('path_restore_14l.1.%d.basic.%d.bias' % (i, j), (64,))]
    s += [('path_restore_14l.3.weight', (cout, 64, 3, 3)), ('path_restore_14l.3.bias', (cout,))]
    return tuple(s)


def seeded_weights(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(shp, generator=g) * 0.05 for k, shp in shapes}


ARCH_SHAPES = {
    'srcnn_res': _srcnn_res_shapes,
    'srcnn_demosaic': lambda P=0: O.SRCNN_DEMOSAIC_SHAPES,
    'path14l_bayer': lambda P=0: _path14l_shapes(4, 4),
    'path14l_bgr': lambda P=0: _path14l_shapes(3, 3),
}

PROXY_PARAMS = {'reinhard': 2, 'crysisengine': 1, 'filmic': 2, 'whiteworld': 1, 'bilateral': 3,
                'median': 1, 'fastnlm': 3, 'bm3d': 5}

# ---- module wrappers: (img NCHW, params (N,P) in [0,1]) -> img ---------------------------
def m_gamma(x, p): return O.gamma_manual(x, p)                                   # tools_origin.py:48-73
def m_grayworld(x, p=None): return O.wb_grayworld(x)                             # :22-45
def m_wbmanual(x, p): return O.wb_manual(x, p * 5)                               # :200-225
def m_skip(x, p=None): return x                                                  # :256-262
def m_wbquadratic(x, p): return O.wb_quadratic(x, p)                             # :313-359
def m_gtmmanual(x, p): return O.gtm_manual(x, p, 4)                              # :409-440
def m_nearest(x, p=None): return O.demosaic_nearest(x)                           # :265-286
def m_bilinear(x, p=None): return O.demosaic_bilinear(x * 255.) / 255.           # :445-475
def m_laplacian(x, p=None): return O.demosaic_laplacian(x * 255., 255.) / 255.   # :479-509
def m_demosaicnet(x, p=None): return O.srcnn_demosaic(x, O.demosaicnet_standin_state())


def m_reinhard(x, p):                                                            # :513-550
    p = p.detach()
    return O.tone_reinhard(x * 255., p[:, 0], p[:, 1]).float() / 255.


def m_crysis(x, p):                                                              # :554-588
    return O.tone_crysis(x * 255., p.detach()[:, 0]).float() / 255.


def m_filmic(x, p):                                                              # :592-630
    p = p.detach()
    return O.tone_filmic(x * 255., p[:, 0], p[:, 1] * 9. + 1.).float() / 255.


def m_whiteworld(x, p):                                                          # :634-669
    return O.wb_whiteworld(x * 255., p.detach()[:, 0].numpy()).float() / 255.


def m_bilateral(x, p):                                                           # :673-717
    p = p.detach()
    win = O.bilateral_window_from_param(p[:, 0])
    return O.denoise_bilateral(x * 255., win, p[:, 1] * 99 + 1, p[:, 2] * 99 + 1).float() / 255.


def m_median(x, p):                                                              # :721-758
    k = O.median_size_from_param(p.detach()[0, 0].item())
    return O.denoise_median(x * 255., k).float() / 255.


def m_fastnlm(x, p):                                                             # :762-804
    p = p.detach()
    blk, srch = O.bilateral_window_from_param(p[:, 0]), O.bilateral_window_from_param(p[:, 1])
    return O.denoise_fastnlm(x * 255., blk, srch, p[:, 2] * 99 + 1).float() / 255.


class Net:
    """A CNN candidate with its weights (tools_proxy.py)."""
    def __init__(self, arch, P, seed):
        self.arch, self.P = arch, P
        self.sd = seeded_weights(ARCH_SHAPES[arch](P), seed)

    def __call__(self, x, p=None):
        if self.arch == 'srcnn_res':
            return O.srcnn_res(x, p, self.sd)
        return getattr(O, self.arch)(x, self.sd)


# ---- registry (isp_universal.py:62-127 / origin_universal.py:36-83) ----------------------
NAMES = {'Bayer': ['path_bayer', 'skip'],
         'Demosaic': ['nearest', 'bilinear', 'laplacian', 'demosaicnet'],
         'sRGB': ['gamma', 'reinhard', 'crysisengine', 'filmic', 'grayworld', 'whiteworld', 'bilateral',
                  'median', 'fastnlm', 'skip', 'wbmanual', 'path_bgr', 'wbquadratic', 'gtmmanual', 'bm3d']}
WBQ_INIT = [0, 0, 0, 0, 0, 0, 0.406, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.406, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.406, 0]
DEFAULT_LOGITS = {'gamma': [0.], 'reinhard': [0., 0.], 'crysisengine': [0.], 'filmic': [0., 0.],
                  'whiteworld': [0.], 'bilateral': [0., 0., 0.], 'median': [0.], 'fastnlm': [0., 0., 0.],
                  'wbmanual': [-1.38] * 3, 'wbquadratic': WBQ_INIT, 'gtmmanual': [-1.099, 0, 1.099],
                  'bm3d': [-1.946, 1.099, -1.099, -1.099, 2.708]}
CLASSICAL = {'gamma': m_gamma, 'grayworld': m_grayworld, 'skip': m_skip, 'wbmanual': m_wbmanual,
             'wbquadratic': m_wbquadratic, 'gtmmanual': m_gtmmanual, 'nearest': m_nearest,
             'demosaicnet': m_demosaicnet}
ORIGIN = {'reinhard': m_reinhard, 'crysisengine': m_crysis, 'filmic': m_filmic, 'whiteworld': m_whiteworld,
          'bilateral': m_bilateral, 'median': m_median, 'fastnlm': m_fastnlm, 'bilinear': m_bilinear, 'laplacian': m_laplacian}


def parse_architecture(architecture):
    """-> [(step, domain, name)]; grammar of isp_universal.py:131-164."""
    domain, step, out = None, 0, []
    for tok in architecture.split('_'):
        if tok in NAMES:
            domain = tok
            continue
        if domain is None:
            raise ValueError('Domain (Bayer, Demosaic, sRGB) is not specified in ISP architecture!')
        step += 1
        idx = int(tok)
        assert 1 <= idx <= len(NAMES[domain])
        out.append((step, domain, NAMES[domain][idx - 1]))
    return out


class FixedPipeline:
    """OriginUniversal (variant='origin') / IspUniversal (variant='isp') on CPU."""

    def __init__(self, architecture, variant='origin', weight_seed=10):
        self.stages, self.logits, self.keys = [], [], []
        n_nets = 0
        for step, domain, name in parse_architecture(architecture):
            if name == 'path_bayer':
                fn = Net('path14l_bayer', 0, weight_seed + n_nets); n_nets += 1
            elif name == 'path_bgr':
                fn = Net('path14l_bgr', 0, weight_seed + n_nets); n_nets += 1
            elif name == 'bm3d' or (variant == 'isp' and name in PROXY_PARAMS):
                fn = Net('srcnn_res', PROXY_PARAMS[name], weight_seed + n_nets); n_nets += 1
            elif variant == 'isp' and name in ('bilinear', 'laplacian'):
                fn = Net('srcnn_demosaic', 0, weight_seed + n_nets); n_nets += 1
            elif name in CLASSICAL:
                fn = CLASSICAL[name]
            elif name in ORIGIN:
                fn = ORIGIN[name]
            else:
                raise NotImplementedError(name)
            self.stages.append(fn)
            self.keys.append('param_step{}_{}'.format(step, name))
            self.logits.append(torch.tensor(DEFAULT_LOGITS.get(name, []), dtype=torch.float32, requires_grad=True))

    def forward(self, x, logits=None):
        logits = self.logits if logits is None else logits
        N = x.shape[0]
        inter = []
        for fn, lg in zip(self.stages, logits):
            par = None if lg.numel() == 0 else torch.sigmoid(lg).repeat(N, 1)
            x = fn(x, par)
            inter.append(x)
        return x, inter


SRGB_STEP = [('gamma', None), ('reinhard', 'srcnn_res'), ('crysisengine', 'srcnn_res'), ('filmic', 'srcnn_res'),
             ('grayworld', None), ('whiteworld', 'srcnn_res'), ('bilateral', 'srcnn_res'), ('median', 'srcnn_res'),
             ('fastnlm', 'srcnn_res'), ('skip', None), ('wbmanual', None), ('path_bgr', 'path14l_bgr'),
             ('wbquadratic', None), ('gtmmanual', None), ('bm3d', 'srcnn_res')]


class Supernet:
    """SuperPruneFifteenDemosFourBayerTwo on CPU (super_prune…:14-214).  Network weights
    follow the construction order of the reference (one seed per network instance)."""

    def __init__(self, n_step, threshold, weight_seed=10):
        self.threshold = threshold
        s = [weight_seed]

        def net(arch, P=0):
            n = Net(arch, P, s[0]); s[0] += 1
            return n
        self.steps = [[net('path14l_bayer'), m_skip],
                      [m_nearest, net('srcnn_demosaic'), net('srcnn_demosaic'), m_demosaicnet]]
        self.logits = [[torch.zeros(0)] * 2, [torch.zeros(0)] * 4]
        self.keys = ['alpha_bayer', 'alpha_demosaic']
        self.alphas = [torch.zeros(2, requires_grad=True), torch.zeros(4, requires_grad=True)]
        for k in range(n_step):
            mods, lgs = [], []
            for name, arch in SRGB_STEP:
                if arch == 'srcnn_res':
                    mods.append(net(arch, PROXY_PARAMS[name]))
                elif arch is not None:
                    mods.append(net(arch))
                else:
                    mods.append(CLASSICAL[name])
                lgs.append(torch.tensor(DEFAULT_LOGITS.get(name, []), dtype=torch.float32, requires_grad=True))
            self.steps.append(mods); self.logits.append(lgs)
            self.alphas.append(torch.zeros(15, requires_grad=True))
            self.keys.append('alpha_step%d' % (k + 1))
        self.pruned_paths = [0] * (n_step + 2)

    @property
    def trainable(self):
        return [l for lg in self.logits[2:] for l in lg]

    def forward(self, x):
        N = x.shape[0]
        inter = []
        for i, (mods, lgs, alpha) in enumerate(zip(self.steps, self.logits, self.alphas)):
            post, self.pruned_paths[i] = O.prune_probs(alpha, self.threshold)
            y = 0
            for mod, lg, pr in zip(mods, lgs, post):
                if pr < 1e-9:
                    if lg.numel() > 0:
                        y = y + torch.zeros(x.shape) * lg.sum()
                    continue
                par = None if lg.numel() == 0 else torch.sigmoid(lg).repeat(N, 1)
                y = y + mod(x, par) * pr
            inter.append(y)
            x = y
        return x, inter


def darts_step(netG, netV, img, gt, val_img, val_gt, lr_G, momentum_G, lr_meta, momentum_buf=None,
               adam_state=None, beta1=0.9, beta2=0.999):
    """One `optimize_alphas` + `optimize_parameters` (darts_model.py:159-324) with MSE loss, plain
    SGD-momentum for params and Adam for alphas, on two `Supernet` instances.  Returns a dict of
    everything a parity test wants to look at."""
    mse = lambda a, b: ((a - b) ** 2).mean()
    P, A = netG.trainable, netG.alphas
    PV, AV = netV.trainable, netV.alphas
    nz = [p for p in P if p.numel() > 0]
    # virtual step (:182-222)
    loss = mse(netG.forward(img)[0], gt)
    g = torch.autograd.grad(loss, nz, allow_unused=True)
    gi = iter(g)
    with torch.no_grad():
        for i, (p, vp) in enumerate(zip(P, PV)):
            if p.numel() == 0:
                continue
            gr = next(gi)
            mom = (momentum_buf[i] if momentum_buf is not None else 0.) * momentum_G
            vp.copy_(p if gr is None else p - lr_meta * (mom + gr))
        for a, va in zip(A, AV):
            va.copy_(a)
    # unrolled loss (:236-250)
    vloss = mse(netV.forward(val_img)[0], val_gt)
    nzv = [p for p in PV if p.numel() > 0]
    vg = torch.autograd.grad(vloss, list(AV) + nzv, allow_unused=True)
    dalpha, dp = vg[:len(AV)], vg[len(AV):]
    # hessian (:270-324)
    norm = torch.cat([w.reshape(-1) for w in dp if w is not None]).norm()
    eps = 0. if norm < 1e-6 else 0.01 / norm
    with torch.no_grad():
        for p, d in zip(nz, dp):
            if d is not None:
                p += eps * d
    pos = torch.autograd.grad(mse(netG.forward(img)[0], gt), A)
    with torch.no_grad():
        for p, d in zip(nz, dp):
            if d is not None:
                p -= 2. * eps * d
    neg = torch.autograd.grad(mse(netG.forward(img)[0], gt), A)
    with torch.no_grad():
        for p, d in zip(nz, dp):
            if d is not None:
                p += eps * d
    hess = O.darts_hessian(pos, neg, eps)
    agrad = [da - lr_meta * h for da, h in zip(dalpha, hess)]
    # optimize_parameters (:159-180): plain loss/grad on train data
    loss2 = mse(netG.forward(img)[0], gt)
    pgrad = torch.autograd.grad(loss2, nz, allow_unused=True)
    return dict(loss=loss, val_loss=vloss, dalpha=dalpha, dp=dp, eps=eps, hessian=hess, alpha_grad=agrad,
                loss2=loss2, param_grad=pgrad)
