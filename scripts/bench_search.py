"""Secondary benchmark (BASELINE.json configs[2]/[4]): one full DARTS search iteration of the supernet
(n_step=3, threshold 0.2, alpha=0 -> nothing pruned: 2+4+45 candidates) = optimize_alphas() +
optimize_parameters() = 5 supernet forward+backward passes on 256x256 raw patches.

    python scripts/bench_search.py [--batch 4] [--iters 5]        (torchrun for N>1: per-GPU batch fixed)

Prints one JSON line (rank 0): iterations/s, MP/s (raw pixels through ONE fwd+bwd pass x 5 passes), ms/iter."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from reconfigisp_b200 import dist as D
from reconfigisp_b200.search import DartsModel
from reconfigisp_b200.synthetic import synthetic_frames


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    ap.add_argument('--n_step', type=int, default=3)
    a = ap.parse_args()
    rank, world, local = D.init_from_env('nccl')
    torch.cuda.set_device(local)
    opt = {'model': 'darts', 'network_G': {'which_model_G': 'SuperPruneFifteenDemosFourBayerTwo', 'n_step': a.n_step, 'n_modules': 15,
                                           'prune_threshold': 0.2, 'weight_seed': 10},
           'train': {'lr_G': 1e-3, 'momentum_G': 0.9, 'lr_meta': 1e-3, 'beta1': 0.9, 'beta2': 0.999, 'pixel_criterion': 'l2'}}
    m = DartsModel(opt)
    raw, gt = synthetic_frames(2 * a.batch, a.size, a.size, seed=10 + rank, pin=True)
    data = (raw[:a.batch], gt[:a.batch], raw[a.batch:], gt[a.batch:])       # train half / validation half
    m.feed_data(data)

    def it():
        m.optimize_alphas()
        m.optimize_parameters()
    for _ in range(a.warmup):
        it()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        it()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.iters], device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    if rank == 0:
        px = a.batch * a.size * a.size
        print(json.dumps({'metric': 'DARTS supernet search iteration (5 fwd+bwd passes, 51 candidates)', 'n_gpus': world,
                          'batch_per_gpu': a.batch, 'patch': a.size, 'n_step': a.n_step, 'ms_per_iter': round(ms, 2),
                          'iters_per_s': round(1e3 / ms, 3), 'MP_per_s_per_pass': round(world * 5 * px / 1e6 / (ms / 1e3), 2),
                          'loss': float(m.log_dict['loss']), 'pruned_paths': m.netG.pruned_paths}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
