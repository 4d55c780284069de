"""Record `tests/golden/darts_ft.npz` by RUNNING THE REFERENCE's own `models.darts_ft_model.DartsFtModel`
(codes/models/darts_ft_model.py:20-368) on CPU: one `optimize_parameters()` (fills the FIFO of sRGB intermediates, :194-201)
followed by one `finetune_proxies()` (:206-246) with `ft_steps = 2`.  TEST INFRASTRUCTURE.

    python -m oracle.gen_golden_darts_ft          (only in the build container: needs /root/reference)

The reference does not expose the fine-tuning losses or gradients, so its own objects are instrumented, not re-implemented:
`cri_pix` (the model's own nn.MSELoss instance) is wrapped to log every loss it returns, and every proxy optimiser's
`step` logs the gradients the reference's `loss.backward()` left in `.grad` (compact: first 24 values, sum and abs-sum of
every tensor).  The host RNG (python `random` for the memory index :219, torch CPU generator for the parameters :225) is
seeded right before `finetune_proxies()`; the drawn indices / parameters are recorded too.  The `Origin*` targets are the
reference's wrappers over the oracle (the kernels they call are not shipped, oracle/SPEC.md)."""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as RL           # noqa: E402
from oracle.gen_golden_darts import ref_opt, inputs, OUT, N_STEP   # noqa: E402

FT_STEPS, FT_SEED, LR = 2, 1234, 1e-4


def summary(t):
    f = t.detach().reshape(-1).double()
    return torch.cat([f[:24].float() if f.numel() >= 24 else torch.nn.functional.pad(f.float(), (0, 24 - f.numel())),
                      torch.tensor([float(f.sum()), float(f.abs().sum())])])


def main():
    ref = RL.load_reference(weight_seed=10)
    import models.networks as networks
    import models.darts_ft_model as dfm
    orig_define = networks.define_G

    def define_G(opt):
        ref.counter['n'] = 0
        return orig_define(opt)
    networks.define_G = define_G
    dfm.networks.define_G = define_G
    img, gt, vimg, vgt, alphas0 = inputs(10)
    rec = {'img': img, 'gt': gt, 'vimg': vimg, 'vgt': vgt}
    opt = ref_opt('SuperPruneFifteenDemosFourBayerTwoFt', 'darts_ft')
    opt['proxy_ft_params'] = {'memory_size': 8, 'ft_steps': FT_STEPS}
    opt['train']['lr_G'] = LR          # Adam moves every weight by ~lr per step whatever its gradient: keep the second step comparable
    with RL.cpu_only():
        m = dfm.DartsFtModel(opt)
        m.feed_data((img, gt, vimg, vgt))
        m.optimize_parameters()
        rec['loss_G'] = torch.tensor(m.log_dict['loss'])
        rec['n_ft_data'] = torch.tensor(len(m.ft_data))
        names = [e[0] for e in m.ft_nets]
        rec['names'] = np.array(names)
        losses = []
        cri = m.cri_pix

        class Logged(torch.nn.Module):
            def forward(self, a, b):
                l = cri(a, b)
                losses.append(float(l.detach()))
                return l
        m.cri_pix = Logged()
        grads = {n: [] for n in names}
        for name, _, attr, _, optim in m.ft_nets:
            orig_step = optim.step

            def step(orig_step=orig_step, attr=attr, name=name):
                grads[name].append(torch.stack([summary(p.grad) for p in attr.parameters()]))
                return orig_step()
            optim.step = step
        random.seed(FT_SEED)
        torch.manual_seed(FT_SEED)
        m.finetune_proxies()
        rec['losses'] = torch.tensor(losses).view(len(names), FT_STEPS)
        for n in names:
            rec['grads_' + n] = torch.stack(grads[n])                 # (FT_STEPS, n_tensors, 26)
    rec['ft_seed'] = torch.tensor(FT_SEED)
    rec['lr'] = torch.tensor(LR)
    np.savez_compressed(os.path.join(OUT, 'darts_ft.npz'), **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in rec.items()})
    print('wrote darts_ft.npz:', names, 'losses', rec['losses'].tolist(), 'ft_data', int(rec['n_ft_data']))


if __name__ == '__main__':
    main()
