#!/usr/bin/env python
"""Benchmark of the ReconfigISP hot path on B200 (driver contract: one JSON line on stdout from rank 0).

Workload = BASELINE.json configs[1]: one proxy-tuning step (forward + MSE + gradients of every stage
parameter + Adam update) of the fixed all-classical pipeline  Bayer_02_Demosaic_02_sRGB_11_13_01_14
(skip | bilinear demosaic | wb-manual | wb-quadratic | gamma | gtm-manual) on a batch of
synthetic 12 MP RGGB raws per GPU.  Metric: ISP-stack MP/s, fwd+bwd (MP = N*H*W of the raw frame).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...          (one rank per GPU, weak scaling)

  value        steps with raw/GT already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e          the same step through the public model API with PINNED HOST buffers: feed_data() (H2D of raw + GT)
               -> optimize_parameters() -> loss.item() (D2H) every step
  roofline     the dominant kernel (fused pipeline step) alone: algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline the oracle port of the same pipeline on the host cores, bounded sample (rank 0, N=1 only)
  --impl reference   times that CPU port as its own arm (rank 0 only under torchrun)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARCH = 'Bayer_02_Demosaic_02_sRGB_11_13_01_14'
H, W = 3000, 4000
FRAMES_PER_GPU = 4
ALGO_BYTES_PER_PX = 16          # read raw 4 B + GT 12 B; nothing but ~100 floats is written
METRIC = 'ISP-stack MP/s fwd+bwd (fixed pipeline proxy-tuning step, synthetic 12MP RGGB raws)'


def measured_peak_gbs():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(self.samples)}


def opt_for_bench():
    return {'model': 'isp', 'is_train': True,
            'network_G': {'which_model_G': 'OriginUniversal', 'architecture': ARCH, 'weight_seed': 10},
            'train': {'lr_G': 1e-3, 'beta1': 0.9, 'beta2': 0.99, 'pixel_criterion': 'l2', 'lr_scheme': 'MultiStepLR',
                      'lr_steps': [20000, 40000, 60000], 'lr_gamma': 0.5},
            'path': {'pretrain_model_G': None}}


# ------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(sample_hw, reps, threads=None):
    """The oracle port of the pipeline (reference code path restated on CPU torch): fwd + MSE + backward of the
    stage parameters on one frame of `sample_hw`.  Returns (MP/s, cores, seconds per rep)."""
    import torch
    from oracle import pipeline_oracle as PO
    from reconfigisp_b200.synthetic import synthetic_frames
    cores = threads or os.cpu_count()
    torch.set_num_threads(cores)
    h, w = sample_hw
    raw, gt = synthetic_frames(1, h, w, seed=10)
    pipe = PO.FixedPipeline(ARCH, 'origin', 10)

    def step():
        y, _ = pipe.forward(raw)
        loss = ((y - gt) ** 2).mean()
        nz = [l for l in pipe.logits if l.numel()]
        torch.autograd.grad(loss, nz)
    step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    return h * w / 1e6 / dt, torch.get_num_threads(), dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    # a "step" = one fwd+bwd of the CPU port on a bounded sample of the workload (a crop of one 12 MP frame), sized
    # from a probe so that the whole --steps/--warmup run stays within ~2 minutes
    _, _, probe_dt = cpu_reference_rate((752, 1000), 1)
    n = max(1, args.steps + args.warmup)
    scale = min(4.0, max(0.25, (120.0 / n) / max(probe_dt, 1e-3)))         # pixels relative to the probe crop
    sample = (int(752 * scale ** 0.5) // 2 * 2, int(1000 * scale ** 0.5) // 4 * 4)
    rate, cores, dt = cpu_reference_rate(sample, max(1, args.steps + args.warmup - 1))
    line = {'impl': 'reference', 'metric': METRIC, 'value': round(rate, 3), 'unit': 'MP/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(dt * 1e3, 2), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'configs[1]: fixed pipeline %s fwd+bwd (proxy tuning), 12MP RGGB raws' % ARCH,
                       'note': 'CPU port of the reference path; each step = one bounded %dx%d sample of the workload' % sample},
            'cpu_baseline': {'value': round(rate, 3), 'unit': 'MP/s', 'cores': cores, 'kind': 'port',
                             'sample': '1 frame %dx%d per step (oracle/pipeline_oracle.py FixedPipeline, torch CPU)' % sample},
            'e2e': {'value': round(rate, 3), 'unit': 'MP/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from reconfigisp_b200 import dist as D
    from reconfigisp_b200 import ops
    from reconfigisp_b200.synthetic import synthetic_frames
    from reconfigisp_b200.tuning import IspModel

    rank, world, local = D.init_from_env('nccl')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    B = args.frames
    raw_h, gt_h = synthetic_frames(B, H, W, seed=10 + rank, pin=True)
    model = IspModel(opt_for_bench())
    px_per_step_rank = B * H * W

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- value: inputs resident in HBM ------------------------------------------------------------------------
    model.feed_data((raw_h, gt_h))
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total = timed(model.optimize_parameters, args.steps, args.warmup)
    ms_step = ms_total / args.steps
    value = world * px_per_step_rank / 1e6 / (ms_step / 1e3)

    # ---- e2e: host buffers through the public API, H2D + step + D2H of the loss every step --------------------
    def e2e_step():
        model.feed_data((raw_h, gt_h))
        model.optimize_parameters()
        return float(model.log_dict['loss'].item())
    n_e2e = max(2, min(args.steps, 40))
    ms_e2e = timed(e2e_step, n_e2e, 3) / n_e2e
    e2e = world * px_per_step_rank / 1e6 / (ms_e2e / 1e3)

    # ---- e2e with the device-side codec: the loaders' integer codes cross PCIe, /1023 and /255 happen on the GPU --
    raw_c = torch.round(raw_h * 1023).to(torch.int16).pin_memory()
    gt_c = torch.round(gt_h * 255).to(torch.uint8).pin_memory()

    def e2e_codec_step():
        model.feed_data((raw_c, gt_c))
        model.optimize_parameters()
        return float(model.log_dict['loss'].item())
    ms_codec = timed(e2e_codec_step, n_e2e, 3) / n_e2e
    e2e_codec = world * px_per_step_rank / 1e6 / (ms_codec / 1e3)

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------
    dm_kind, chain, keep = model.netG.fused_mse_step_plan()
    step = ops.PipelineStep(B, H, W, dm_kind, chain, dev)
    with torch.no_grad():
        table = model.netG._segment_table(keep, B).contiguous()
    ms_k = timed(lambda: step(model.img, model.gt, table), args.steps, args.warmup) / args.steps
    # ---- forward-only (fused inference, the other half of BASELINE.json's metric): read raw, write BGR = 16 B/px ----
    with torch.no_grad():
        ms_f = timed(lambda: ops.pipeline_fwd(model.img, dm_kind, chain, table), max(10, args.steps // 4), args.warmup) / max(10, args.steps // 4)
    fwd_rate = world * px_per_step_rank / 1e6 / (ms_f / 1e3)
    clocks = sampler.summary()          # sampled across all timed legs (value, e2e, e2e_codec, roofline, fwd)
    peak, peak_src = measured_peak_gbs()
    achieved = ALGO_BYTES_PER_PX * px_per_step_rank / (ms_k / 1e3) / 1e9

    line = None
    if rank == 0:
        line = {'metric': METRIC, 'value': round(value, 1), 'unit': 'MP/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': round(ms_step, 4), 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': 'configs[1]: fixed pipeline %s fwd+bwd (proxy tuning), %d x 12MP RGGB raws per GPU' % (ARCH, B),
                           'frame': [H, W], 'frames_per_gpu': B, 'parallelism': 'dp%d (frames sharded, one flattened grad all-reduce)' % world,
                           'l2': 'inputs %.0f MB per step > 126 MB L2 (no flush needed)' % ((4 + 12) * px_per_step_rank / 1e6),
                           'peak_source': peak_src},
                'clocks': clocks,
                'e2e': {'value': round(e2e, 1), 'unit': 'MP/s', 'h2d_bytes_per_step': 16 * px_per_step_rank,
                        'd2h_bytes_per_step': 4, 'ms_per_step': round(ms_e2e, 3)},
                'e2e_codec': {'value': round(e2e_codec, 1), 'unit': 'MP/s', 'h2d_bytes_per_step': 5 * px_per_step_rank,
                              'd2h_bytes_per_step': 4, 'ms_per_step': round(ms_codec, 3),
                              'note': 'same step; raw as 10-bit codes (int16) and GT as uint8 cross PCIe, normalised on the device'},
                'fwd': {'value': round(fwd_rate, 1), 'unit': 'MP/s', 'ms_per_step': round(ms_f, 4),
                        'roofline_frac': round(ALGO_BYTES_PER_PX * px_per_step_rank / (ms_f / 1e3) / 1e9 / measured_peak_gbs()[0], 4),
                        'note': 'fused inference of the same pipeline (demosaic + 4 stages in one pass), data resident'},
                'gpu_launches': 2 * args.steps,          # per step: risp::pipeline_kernel + risp::finalize_rows_kernel
                'roofline': {'bound': 'hbm', 'achieved': round(achieved, 1), 'peak': peak, 'unit': 'GB/s',
                             'frac': round(achieved / peak, 4), 'traffic': None, 'kernel': 'risp::pipeline_kernel<BILINEAR, STEP, sigA>',
                             'ms_per_launch': round(ms_k, 4), 'algorithmic_bytes_per_px': ALGO_BYTES_PER_PX}}
        traffic_file = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(traffic_file):
            try:
                t = json.load(open(traffic_file))
                line['roofline']['traffic'] = t.get('dram_bytes_per_px', None) and t['dram_bytes_per_px'] * px_per_step_rank
            except Exception:
                pass
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, cores, dt = cpu_reference_rate((H, W), 8)
        line['cpu_baseline'] = {'value': round(rate, 3), 'unit': 'MP/s', 'cores': cores, 'kind': 'port',
                                'sample': '1 frame %dx%d of the workload, 1 warm-up + 8 reps of fwd+bwd = %.1f s '
                                          '(oracle/pipeline_oracle.py FixedPipeline, torch CPU)' % (H, W, 9 * dt)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # teardown must never hold the job: release the captured graphs first, and leave hard if NCCL teardown stalls
        model._graph = None
        torch.cuda.synchronize()
        watchdog = threading.Timer(60.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.barrier()
        dist.destroy_process_group()
        watchdog.cancel()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--frames', type=int, default=FRAMES_PER_GPU, help='12MP frames per GPU per step')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        return run_reference(args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_b200(args)


if __name__ == '__main__':
    sys.exit(main())
