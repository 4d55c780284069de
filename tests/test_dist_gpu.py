"""Multi-GPU (NCCL) checks; skipped unless >= 2 GPUs are visible.  The N>1 host logic itself is covered on
CPU by tests/test_dist_cpu.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ['RISP_ROOT'])
import torch.distributed as dist
from reconfigisp_b200 import dist as D
from reconfigisp_b200.tuning import IspModel
from reconfigisp_b200.synthetic import synthetic_frames
rank, world, local = D.init_from_env('nccl')
torch.cuda.set_device(local)
opt = {'model': 'isp', 'is_train': True, 'network_G': {'which_model_G': 'OriginUniversal', 'architecture': 'Bayer_02_Demosaic_02_sRGB_11_13_01_14', 'weight_seed': 10},
       'train': {'lr_G': 1e-2, 'beta1': 0.9, 'beta2': 0.99, 'pixel_criterion': 'l2', 'lr_scheme': 'MultiStepLR', 'lr_steps': [100], 'lr_gamma': 0.5}}
m = IspModel(opt)
raw, gt = synthetic_frames(2 * world, 64, 96, seed=10)            # one global batch, sharded by rank
sl = D.shard_batch(2 * world, rank, world)
for _ in range(3):
    m.feed_data((raw[sl], gt[sl]))
    m.optimize_parameters()
flat = torch.cat([p.detach().reshape(-1) for p in m.netG.trainable_parameters if p.nelement()])
both = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(both, flat)
assert all(torch.equal(both[0], b) for b in both), 'ranks diverged'
# the gradient exchange ran over NVLink peer memory (csrc/risp_p2p.cu), not NCCL: hold it to NCCL on random vectors,
# many epochs in a row (parity double-buffering), different lengths, bit-identical across ranks
p2p = m.__dict__.get('_p2p')
assert p2p is not None, 'peer-memory all-reduce was not set up'
g = torch.Generator(device='cuda').manual_seed(100 + rank)
for it in range(64):
    n = 1 + (it * 37) % 1000
    x = torch.randn(n, device='cuda', generator=g)
    ref = x.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    ref /= world
    p2p(x)
    assert float((x - ref).abs().max()) <= 1e-6, (it, float((x - ref).abs().max()))
    allx = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(allx, x)
    assert all(torch.equal(allx[0], t) for t in allx), 'p2p averages differ between ranks'
assert p2p.timeouts() == 0
if rank == 0:
    # single-process reference on the full batch: averaging per-rank mean-gradients == full-batch gradient
    torch.save(flat.cpu(), os.environ['RISP_OUT'])
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
def test_data_parallel_tuning_matches_single_gpu(tmp_path):
    out = str(tmp_path / 'dp.pt')
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    env = dict(os.environ, RISP_ROOT=ROOT, RISP_OUT=out)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29631', str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    dp = torch.load(out)
    # the same 3 steps on one GPU with the full batch
    sys.path.insert(0, ROOT)
    from reconfigisp_b200.tuning import IspModel
    from reconfigisp_b200.synthetic import synthetic_frames
    opt = {'model': 'isp', 'is_train': True, 'network_G': {'which_model_G': 'OriginUniversal', 'architecture': 'Bayer_02_Demosaic_02_sRGB_11_13_01_14', 'weight_seed': 10},
           'train': {'lr_G': 1e-2, 'beta1': 0.9, 'beta2': 0.99, 'pixel_criterion': 'l2', 'lr_scheme': 'MultiStepLR', 'lr_steps': [100], 'lr_gamma': 0.5}}
    m = IspModel(opt)
    raw, gt = synthetic_frames(4, 64, 96, seed=10)
    for _ in range(3):
        m.feed_data((raw, gt))
        m.optimize_parameters()
    flat = torch.cat([p.detach().reshape(-1) for p in m.netG.trainable_parameters if p.nelement()]).cpu()
    assert float((flat - dp).abs().max()) <= 2e-5
