"""H2D bandwidth of one GPU: torch pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined)."""
import ctypes, sys, os, time
import torch
rt = ctypes.CDLL('libcudart.so.12') if os.path.exists('/usr/local/cuda/lib64/libcudart.so.12') else ctypes.CDLL('libcudart.so')
n = 240_000_000
dev = torch.empty(n, dtype=torch.uint8, device='cuda')
pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
p = ctypes.c_void_p()
rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(4))          # cudaHostAllocWriteCombined = 4
print('cudaHostAlloc WC rc', rc)
buf = (ctypes.c_ubyte * n).from_address(p.value)
wc = torch.frombuffer(buf, dtype=torch.uint8)
print('is_pinned', pinned.is_pinned(), wc.is_pinned())


def bw(src):
    for _ in range(3):
        dev.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dev.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return n * 20 / (e0.elapsed_time(e1) / 1e3) / 1e9


print('torch pinned  %.1f GB/s' % bw(pinned))
print('write-combined %.1f GB/s' % bw(wc))
