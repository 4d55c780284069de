"""Tiled inference of large frames (codes/test_split.py:57-108 + utils/util_path_restore.py:47-134).

The reference tiles the frame on the host, pushes the tiles through the model ONE BY ONE (H2D + D2H per tile,
`test_split.py:88-100`) and blends on the host with a linear-ramp mask.  Here the split, the model pass over
ALL tiles as one batch (optionally in chunks, optionally sharded over ranks) and the blend are device-side:
two kernel launches around the model, no host round trips."""
import torch

from . import dist as D
from . import ops


def split_inference(model_fn, frame, patch_size, patch_stride, chunk=None, clip01=True, out_channels=3, uint8=False):
    """frame: (1,C,H,W) or (C,H,W) CUDA tensor in [0,1]; model_fn: (T,C,h,w) -> (T,C',h,w).
    Returns the blended (C',H,W) frame (clipped to [0,1] like test_split.py:107 when clip01), or with `uint8` the
    (H,W,C') 8-bit image of test_split.py:107 packed by the blend kernel itself (4x less to copy back)."""
    f = frame[0] if frame.dim() == 4 else frame
    C, H, W = f.shape
    size, stride = (patch_size, patch_size), (patch_stride, patch_stride)
    tiles, _ = ops.whole2patch(f, size, stride)
    T = tiles.shape[0]
    chunk = chunk or T
    outs = []
    with torch.no_grad():
        for s in range(0, T, chunk):
            outs.append(model_fn(tiles[s:s + chunk]))
    out_tiles = torch.cat(outs) if len(outs) > 1 else outs[0]
    assert out_tiles.shape[1] == out_channels
    if uint8:
        return ops.patch2whole_u8(out_tiles.contiguous(), (H, W), stride)
    return ops.patch2whole(out_tiles.contiguous(), (H, W), stride, clip01=clip01)


def split_inference_sharded(model_fn, frame, patch_size, patch_stride, clip01=True):
    """Tiles round-robin over the ranks of the process group; every rank ends with the full blended frame
    (one all-gather of the processed tiles: 63 x 512^2 x 12 B = 198 MB per 12 MP frame)."""
    import torch.distributed as dist
    f = frame[0] if frame.dim() == 4 else frame
    C, H, W = f.shape
    size, stride = (patch_size, patch_size), (patch_stride, patch_stride)
    tiles, _ = ops.whole2patch(f, size, stride)
    T = tiles.shape[0]
    if not D.is_dist():
        return split_inference(model_fn, frame, patch_size, patch_stride, clip01=clip01)
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (T + world - 1) // world
    idx = list(range(rank * per, min(T, (rank + 1) * per)))
    with torch.no_grad():
        mine = model_fn(tiles[idx[0]:idx[-1] + 1]) if idx else tiles.new_zeros((0, 3) + tuple(tiles.shape[2:]))
    pad = tiles.new_zeros((per,) + tuple(mine.shape[1:]))
    pad[:mine.shape[0]] = mine
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    out_tiles = torch.cat(gathered)[:T].contiguous()
    return ops.patch2whole(out_tiles, (H, W), stride, clip01=clip01)
