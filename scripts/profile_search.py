"""Kernel-level breakdown of one DARTS search iteration with torch.profiler (CUPTI; low overhead compared with ncu):
    python scripts/profile_search.py [--batch 4] [--out gpurun_out/search_kernels.json]
Aggregates the device time per kernel name over ONE timed iteration and prints the top entries."""
import argparse, collections, json, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from reconfigisp_b200.search import DartsModel
from reconfigisp_b200.synthetic import synthetic_frames


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--out', default='gpurun_out/search_kernels.json')
    a = ap.parse_args()
    opt = {'model': 'darts', 'network_G': {'which_model_G': 'SuperPruneFifteenDemosFourBayerTwo', 'n_step': 3, 'n_modules': 15,
                                           'prune_threshold': 0.2, 'weight_seed': 10},
           'train': {'lr_G': 1e-3, 'momentum_G': 0.9, 'lr_meta': 1e-3, 'beta1': 0.9, 'beta2': 0.999, 'pixel_criterion': 'l2'}}
    m = DartsModel(opt)
    raw, gt = synthetic_frames(2 * a.batch, a.size, a.size, seed=10, pin=True)
    m.feed_data((raw[:a.batch], gt[:a.batch], raw[a.batch:], gt[a.batch:]))

    def it():
        m.optimize_alphas()
        m.optimize_parameters()
    for _ in range(2):
        it()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
        it()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    t_min, t_max = None, None
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = re.sub(r'\(.*', '', e.name)[:110]
            agg[k][0] += 1
            agg[k][1] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
            s, t = e.time_range.start, e.time_range.end
            t_min = s if t_min is None else min(t_min, s)
            t_max = t if t_max is None else max(t_max, t)
    tot = sum(v[1] for v in agg.values())
    rows = sorted(([k, v[0], round(v[1], 1)] for k, v in agg.items()), key=lambda r: -r[2])
    own = sum(r[2] for r in rows if 'risp' in r[0])
    out = {'batch': a.batch, 'launches': sum(r[1] for r in rows), 'own_launches': sum(r[1] for r in rows if 'risp' in r[0]),
           'sum_kernel_us': round(tot, 1), 'span_us': None if t_min is None else round(t_max - t_min, 1),
           'own_share': round(own / tot, 4), 'kernels': rows}
    os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
    json.dump(out, open(a.out, 'w'), indent=1)
    print(json.dumps({k: out[k] for k in out if k != 'kernels'}))
    for r in rows[:32]:
        print('%6d %10.1f %5.1f%%  %s' % (r[1], r[2], 100 * r[2] / tot, r[0]))
    # which ATen ops (by input shape) the glue comes from
    ops_ = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith('aten::') and e.device_time_total > 300]
    for e in sorted(ops_, key=lambda e: -e.device_time_total)[:24]:
        print('%-28s %5d %9.1f us  %s' % (e.key, e.count, e.device_time_total, str(e.input_shapes)[:150]))


if __name__ == '__main__':
    main()
