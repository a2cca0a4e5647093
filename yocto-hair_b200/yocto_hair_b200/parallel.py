"""Multi-GPU plumbing: one process per GPU, image tiles round-robin over ranks,
ONE reduce of the fp32 accumulation buffers at the end.

The reference parallelises over image rows with threads
(parallel_for, libs/yocto_pathtrace/yocto_pathtrace.cpp:1954-1989); pixels are
independent (own PCG32 stream each), so ranks need no exchange while rendering.
`torch.distributed` is used for the rendezvous and the single collective only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi as capi


def env_rank():
    """(rank, local_rank, world) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_pixels(width: int, height: int, rank: int, world: int) -> np.ndarray:
    """Pixel indices (j * width + i) owned by `rank`: the map yh_state_set_shard uses."""
    lib = capi.load()
    n = C.c_int64()
    capi.check(lib.yh_shard_pixels(width, height, rank, world, None, C.byref(n)))
    px = np.zeros(n.value, np.int32)
    capi.check(lib.yh_shard_pixels(width, height, rank, world, px.ctypes.data_as(capi.c_int_p), C.byref(n)))
    return px


def shard_mask(width: int, height: int, rank: int, world: int) -> np.ndarray:
    m = np.zeros(width * height, bool)
    m[shard_pixels(width, height, rank, world)] = True
    return m.reshape(height, width)


class _DevArray:
    """__cuda_array_interface__ view of a library-owned device buffer."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def state_tensors(state, device):
    """torch views (no copy) of a DeviceState's accumulation (H, W, 4) f32 and sample-count (H, W) i32 buffers."""
    import torch
    a_ptr, s_ptr = state.buffers_dev()
    accum = torch.as_tensor(_DevArray(a_ptr, (state.height, state.width, 4), "<f4"), device=device)
    samples = torch.as_tensor(_DevArray(s_ptr, (state.height, state.width), "<i4"), device=device)
    return accum, samples


def reduce_image(accum, samples, dst: int = 0, out=None):
    """The one collective of a render: the per-rank accumulation and sample-count buffers summed on rank `dst`.
    Tile shards: every pixel is non-zero on exactly one rank, so the sum is exact; sample shards: the sum of the
    ranks' partial images.  The inputs (views of library-owned buffers) are left untouched - the sum goes into
    scratch tensors (`out`, or fresh ones), so tracing on and reducing again does not count anything twice.
    Works on CUDA tensors (nccl) and CPU tensors (gloo).  Returns (accum_sum, samples_sum), valid on `dst`."""
    import torch.distributed as dist
    a, s = out if out is not None else (accum.clone(), samples.clone())
    if out is not None:
        a.copy_(accum), s.copy_(samples)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(a, dst, op=dist.ReduceOp.SUM)
        dist.reduce(s, dst, op=dist.ReduceOp.SUM)
    return a, s


def resolve_dev(accum, samples, render):
    """state->render = accumulated / samples with the library's kernel (yh_resolve_dev) on CUDA tensors."""
    import torch
    torch.cuda.current_stream().synchronize()
    capi.check(capi.load().yh_resolve_dev(accum.data_ptr(), samples.data_ptr(), samples.numel(), render.data_ptr()))
    return render


def resolve(accum, samples):
    """state->render = accumulated / samples (yocto_pathtrace.cpp:1688); 0 where no sample landed."""
    import torch
    s = samples.to(accum.dtype).unsqueeze(-1)
    return torch.where(s > 0, accum / s.clamp(min=1), torch.zeros_like(accum))
