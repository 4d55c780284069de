"""Host-side cost of one DARTS search iteration: enqueue time (no synchronisation) against device time, and a cProfile of
the Python side (which calls dominate the launch path)."""
import cProfile, os, pstats, sys, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reconfigisp_b200.search import DartsModel
from reconfigisp_b200.synthetic import synthetic_frames

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
opt = {'model': 'darts', 'network_G': {'which_model_G': 'SuperPruneFifteenDemosFourBayerTwo', 'n_step': 3, 'n_modules': 15,
                                       'prune_threshold': 0.2, 'weight_seed': 10},
       'train': {'lr_G': 1e-3, 'momentum_G': 0.9, 'lr_meta': 1e-3, 'beta1': 0.9, 'beta2': 0.999, 'pixel_criterion': 'l2'}}
m = DartsModel(opt)
raw, gt = synthetic_frames(2 * B, 256, 256, seed=10, pin=True)
m.feed_data((raw[:B], gt[:B], raw[B:], gt[B:]))


def it():
    m.optimize_alphas()
    m.optimize_parameters()


for _ in range(3):
    it()
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    it()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print('enqueue %.1f ms, until the device is idle %.1f ms' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3), flush=True)
pr = cProfile.Profile()
pr.enable()
it()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40)
print(s.getvalue()[:9000])
