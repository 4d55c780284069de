"""Data-parallel plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink / NVSwitch).

The only exchange step on this path is the all-reduce of the module-parameter and architecture-weight
gradients: 216 floats = 864 B for n_step = 3.  It is latency-bound, so all gradients
of one backward pass travel as ONE contiguous fp32 buffer in ONE collective.  (The reference reduces only
the DDP-wrapped pass and leaves the alpha gradients rank-local; north_star asks for both.)"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_mean_flat(tensors):
    """Average a list of (small) gradient tensors across ranks with one collective.  None entries are
    treated as zeros of unknown shape and returned as None (they are None on every rank alike)."""
    if not is_dist():
        return list(tensors)
    live = [t for t in tensors if t is not None]
    if not live:
        return list(tensors)
    flat = torch.cat([t.reshape(-1).float() for t in live])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    out, o = [], 0
    for t in tensors:
        if t is None:
            out.append(None)
        else:
            out.append(flat[o:o + t.numel()].view_as(t))
            o += t.numel()
    return out


def allreduce_mean_(flat):
    """In-place average of one contiguous gradient buffer across ranks: a single collective, nothing else."""
    if is_dist():
        if dist.get_backend() == 'nccl':
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:                                   # gloo (CPU tests) has no AVG
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat /= dist.get_world_size()
    return flat


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off BEFORE it allocates pinned host buffers, so that
    the H2D copies of 8 ranks do not all stream out of one socket's DRAM (e2e scaling).  Best effort: returns the node or None."""
    try:
        import torch.cuda as tc
        prop = tc.get_device_properties(local_rank)
        bus = '%04x:%02x:%02x.0' % (getattr(prop, 'pci_domain_id', 0), prop.pci_bus_id, prop.pci_device_id)
        path = '/sys/bus/pci/devices/%s/numa_node' % bus
        if not os.path.exists(path):
            import subprocess
            q = subprocess.run(['nvidia-smi', '-i', str(local_rank), '--query-gpu=pci.bus_id', '--format=csv,noheader'],
                               capture_output=True, text=True, timeout=10).stdout.strip().lower()
            path = '/sys/bus/pci/devices/%s/numa_node' % (q[4:] if len(q) > 12 else q)     # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            a, _, b = part.partition('-')
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def shard_batch(n_items, rank, world):
    """Per-rank slice of a batch: batch_size // world_size items each (data/__init__.py:15-16)."""
    per = n_items // world
    return slice(rank * per, (rank + 1) * per)
