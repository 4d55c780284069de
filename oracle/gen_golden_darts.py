"""Record `tests/golden/darts_model.npz` by RUNNING THE REFERENCE's own `models.darts_model.DartsModel`
(codes/models/darts_model.py:19-330: optimize_alphas + optimize_parameters, SGD-momentum on the module parameters, Adam on
the architecture weights) and `models.darts_ft_model.DartsFtModel.finetune_proxies` on CPU.  TEST INFRASTRUCTURE.

    python -m oracle.gen_golden_darts          (only in the build container: needs /root/reference)

Set-up (SURVEY.md Appendix B through oracle/ref_loader.py): the five un-shipped kernel modules are bound to
oracle/isp_oracle.py, CUDA placement is neutralised, every candidate network gets deterministic seeded weights, and netV
gets the same weights as netG (the reference loads the same checkpoints into both).  n_step = 3, threshold 0.2, two
iterations on 2+2 patches of 16x16 with alphas that put two candidates under the pruning threshold.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as RL           # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
N_STEP, ITERS = 3, 2
HP = dict(lr_G=0.01, momentum_G=0.9, lr_meta=0.01, beta1=0.9, beta2=0.999)


def ref_opt(which='SuperPruneFifteenDemosFourBayerTwo', model='darts'):
    return {'dist': False, 'gpu_ids': None, 'is_train': True, 'model': model,
            'network_G': {'which_model_G': which, 'n_step': N_STEP, 'n_modules': 15, 'prune_threshold': 0.2},
            'train': dict(HP, pixel_criterion='l2', lr_scheme='MultiStepLR', lr_steps=[100000], restarts=None,
                          restart_weights=None, lr_gamma=0.5, clear_state=None),
            'path': {'pretrain_model_G': None, 'strict_load': True}}


def inputs(seed=10):
    g = torch.Generator().manual_seed(seed)
    N, H, W = 2, 16, 16
    img, vimg = torch.rand(N, 1, H, W, generator=g) * 0.8 + 0.1, torch.rand(N, 1, H, W, generator=g) * 0.8 + 0.1
    gt, vgt = torch.rand(N, 3, H, W, generator=g), torch.rand(N, 3, H, W, generator=g)
    alphas0 = [torch.randn(k, generator=g) * 0.5 for k in (2, 4) + (15,) * N_STEP]
    alphas0[2][3] = -3.0      # two candidates under the threshold: pruning is active in the recorded run
    alphas0[3][7] = -3.0
    return img, gt, vimg, vgt, alphas0


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = RL.load_reference(weight_seed=10)
    import models.networks as networks
    import models.darts_model as dm
    orig_define = networks.define_G

    def define_G(opt):                      # netG and netV load the same candidate-network weights
        ref.counter['n'] = 0
        return orig_define(opt)
    networks.define_G = define_G
    dm.networks.define_G = define_G
    # x^gamma has an unbounded derivative at 0 and the clamps have kinks: a recorded run in which some intermediate pixel
    # sits within 1e-4 of 0 cannot be reproduced to a tolerance by ANY other fp32 implementation (a 1e-6 difference in that
    # pixel moves a parameter gradient by 10 %).  Pick the first seed whose intermediates stay clear of the singularity.
    from oracle import pipeline_oracle as PO
    for seed in range(10, 60):
        img, gt, vimg, vgt, alphas0 = inputs(seed)
        probe = PO.Supernet(N_STEP, 0.2, 10)
        with torch.no_grad():
            for a, v in zip(probe.alphas, alphas0):
                a.copy_(v)
        worst = 1.0
        for x in (img, vimg):
            for t in probe.forward(x)[1][1:]:
                t = t.detach()
                near0 = t[(t > -1e-4) & (t < 1e-4)]
                if near0.numel():
                    worst = 0.0
        if worst > 0:
            break
    assert worst > 0, 'no well-conditioned seed found'
    print('input seed', seed)
    rec = {'img': img, 'gt': gt, 'vimg': vimg, 'vgt': vgt, 'input_seed': torch.tensor(seed)}
    for i, a in enumerate(alphas0):
        rec['alpha0_%d' % i] = a
    with RL.cpu_only():
        m = dm.DartsModel(ref_opt())
        with torch.no_grad():
            for a, v in zip(m.netG_attr.alphas, alphas0):
                a.copy_(v)
        m.feed_data((img, gt, vimg, vgt))
        for it in range(ITERS):
            m.optimize_alphas()
            rec['it%d_val_loss' % it] = m.val_loss.detach()
            for i, a in enumerate(m.netG_attr.alphas):
                rec['it%d_alpha_grad_%d' % (it, i)] = a.grad.clone()
                rec['it%d_alpha_%d' % (it, i)] = a.detach().clone()
            m.optimize_parameters()
            rec['it%d_loss' % it] = torch.tensor(m.log_dict['loss'])
            rec['it%d_pruned' % it] = torch.tensor(m.netG_attr.pruned_paths)
            nz = [p for p in m.netG_attr.trainable_parameters if p.nelement() > 0]
            for i, p in enumerate(nz):
                rec['it%d_param_%d' % (it, i)] = p.detach().clone()
                rec['it%d_param_grad_%d' % (it, i)] = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
        rec['n_params'] = torch.tensor(len(nz))
    np.savez_compressed(os.path.join(OUT, 'darts_model.npz'), **{k: v.detach().cpu().numpy() for k, v in rec.items()})
    print('wrote darts_model.npz with', len(rec), 'arrays; losses', [float(rec['it%d_loss' % i]) for i in range(ITERS)],
          'pruned', rec['it0_pruned'].tolist())


if __name__ == '__main__':
    main()
