"""Multi-GPU plumbing: one process per GPU, samples-per-pixel split across ranks, one NCCL sum-reduce of the exact
(int64 fixed-point) accumulators over NVLink. Mirrors what the reference does across threads: every thread renders whole
frames and Image::merge adds them (src/main.rs:235-294, 629-638).

Because sample s of pixel p always draws from the Philox stream keyed (seed, p, s, bounce) and the accumulators are integers,
the reduced image is bit-identical for any world size."""
import ctypes as C


def sample_range(rank, world_size, spp):
    """Rank r renders samples [begin, begin+count) of every pixel: contiguous, near-equal, covering [0, spp) exactly once."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    begin = (spp * rank) // world_size
    end = (spp * (rank + 1)) // world_size
    return begin, end - begin


class _CudaArray:
    """Minimal __cuda_array_interface__ holder so torch can wrap the renderer's device accumulators without a copy."""

    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


def accumulators_as_tensor(renderer, device):
    """torch.int64 view (w*h*4) of the renderer's device accumulators {r, g, b in 2^-32 units, bounce sum}."""
    import torch

    ptr, n = renderer.accum_device_ptr()
    return torch.as_tensor(_CudaArray(ptr, n), device=device)


def reduce_accumulators(tensor, dst=0, group=None):
    """The one collective of the path: integer SUM reduce to `dst` (exact, order-independent)."""
    import torch.distributed as dist

    dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return tensor


def fixed_to_float(acc_i64, w, h):
    """Host-side resolve of reduced accumulators: (sum_rgb f32 (h,w,3), sum_bounces u32 (h,w)) like mrt_accum_download."""
    import numpy as np

    a = np.asarray(acc_i64, dtype=np.int64).reshape(h, w, 4)
    rgb = (a[..., :3].astype(np.float64) * 2.0 ** -32).astype(np.float32)
    return rgb, a[..., 3].astype(np.uint32)
