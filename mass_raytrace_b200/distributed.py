"""Multi-GPU host glue for one-process-per-GPU launches (torchrun). The collective itself lives behind the C ABI
(mrt_comm_* in include/mrt.h, NCCL over NVLink inside libmrt_cuda.so); what is left for the host is to carry the 128-byte NCCL
id from rank 0 to the other ranks -- here through torch.distributed -- and to know the split rule.

Mirrors what the reference does across threads: every thread renders whole frames and Image::merge adds them
(src/main.rs:235-294, 629-638). Sample s of pixel p always draws from the Philox stream keyed (seed, p, s, bounce) and the
accumulators are integers, so the merged image is bit-identical for any world size."""
import ctypes as C


def sample_range(rank, world_size, spp, spp_begin=0):
    """Rank r renders samples [begin, begin+count) of every pixel: contiguous, near-equal, covering the range exactly once.
    The rule of mrt_sample_range (the library applies it itself when mrt_render_accumulate is called on a communicator)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    begin = (spp * rank) // world_size
    end = (spp * (rank + 1)) // world_size
    return spp_begin + begin, end - begin


def library_sample_range(rank, world_size, spp, spp_begin=0):
    """The same through the C ABI (no GPU needed)."""
    from . import _ffi

    b, c = C.c_uint32(0), C.c_uint32(0)
    rc = _ffi.cuda_lib().mrt_sample_range(rank, world_size, spp_begin, spp, C.byref(b), C.byref(c))
    if rc != 0:
        raise ValueError("bad rank/world_size")
    return b.value, c.value


def join_communicator(renderer, rank, world_size, group=None):
    """Every rank of an initialised torch.distributed job calls this once: rank 0 creates the NCCL id (mrt_comm_unique_id), the
    job's own backend broadcasts it, every rank joins with mrt_comm_init_rank. After that renderer.accumulate() is collective."""
    import torch.distributed as dist

    if world_size == 1:
        return
    box = [renderer.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    renderer.comm_init_rank(box[0], rank, world_size)


def fixed_to_float(acc_i64, w, h):
    """Host-side resolve of reduced accumulators: (sum_rgb f32 (h,w,3), sum_bounces u32 (h,w)) like mrt_accum_download."""
    import numpy as np

    a = np.asarray(acc_i64, dtype=np.int64).reshape(h, w, 4)
    rgb = (a[..., :3].astype(np.float64) * 2.0 ** -32).astype(np.float32)
    return rgb, a[..., 3].astype(np.uint32)
