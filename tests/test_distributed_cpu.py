"""The N>1 path on CPU: world_size-2 gloo processes. Each rank owns a sample range (the library's mrt_sample_range), fills exact
int64 accumulators for its samples -- with the per-pixel non-finite flags folded into the bounce lane the way the library does
before its NCCL reduce (csrc/mrt_comm.inc) -- and one SUM reduce must reproduce the single-process image bit for bit: the
property the GPU path relies on (integer accumulators + a sample-keyed RNG make the result independent of the split). The
128-byte communicator id travels from rank 0 to the others through the job's own backend (distributed.join_communicator)."""
import os
import socket
import sys

import numpy as np
import pytest

from mass_raytrace_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_range_partitions_exactly():
    for spp in (0, 1, 7, 10, 256, 1000, 4096):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                b, c = D.sample_range(r, world, spp)
                covered += list(range(b, b + c))
            assert covered == list(range(spp))
            sizes = [D.sample_range(r, world, spp)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.sample_range(2, 2, 10)


def test_library_split_rule_is_the_host_rule():
    """mrt_sample_range (what mrt_render_accumulate applies on a communicator) == distributed.sample_range, offsets included."""
    for spp, begin in ((0, 0), (1, 5), (10, 0), (255, 1000), (4096, 1 << 20), (1000, 0xFFFFF000)):
        for world in (1, 2, 3, 8, 255):
            for r in range(world):
                assert D.library_sample_range(r, world, spp, begin) == D.sample_range(r, world, spp, begin)
    with pytest.raises(ValueError):
        D.library_sample_range(3, 3, 10)


FLAG_SHIFT = 40  # csrc/mrt_comm.inc kFlagShift


def fold_flags(acc, flags):
    """numpy restatement of k_fold_flags: three 8-bit member counters above the bounce sum."""
    a = acc.reshape(-1, 4).copy()
    for k in range(3):
        a[:, 3] += ((flags >> k) & 1).astype(np.int64) << (FLAG_SHIFT + 8 * k)
    return a.reshape(-1)


def unfold_flags(acc):
    a = acc.reshape(-1, 4).copy()
    v = a[:, 3].astype(np.uint64)
    flags = np.zeros(len(a), np.uint32)
    for k in range(3):
        flags |= (((v >> np.uint64(FLAG_SHIFT + 8 * k)) & np.uint64(0xFF)) != 0).astype(np.uint32) << k
    a[:, 3] = (v & np.uint64((1 << FLAG_SHIFT) - 1)).astype(np.int64)
    return a.reshape(-1), flags


def _sample_flags(w, h, s):
    """a few pixels get a non-finite channel in a few samples"""
    p = np.arange(w * h, dtype=np.uint32)
    return (((p * 7 + s * 13) % 29 == 0).astype(np.uint32) << ((p + s) % 3)).astype(np.uint32)


class _FakeRenderer:
    def __init__(self):
        self.joined = None

    @staticmethod
    def comm_unique_id():
        return os.urandom(128)

    def comm_init_rank(self, uid, rank, n):
        self.joined = (bytes(uid), rank, n)


def _sample_contribution(w, h, s):
    """Stand-in for one sample of every pixel: a deterministic function of (pixel, sample) only, like the Philox-keyed render."""
    p = np.arange(w * h, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (p * np.uint64(0x9E3779B97F4A7C15) + np.uint64(s) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFF)
    vals = np.stack([x % np.uint64(1 << 34), (x >> np.uint64(3)) % np.uint64(1 << 33), (x >> np.uint64(7)) % np.uint64(1 << 35), x % np.uint64(51)], axis=1)
    return vals.astype(np.int64).reshape(-1)


def _worker(rank, world, port, w, h, spp, out_path):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fake = _FakeRenderer()
    D.join_communicator(fake, rank, world)
    ids = [None] * world
    dist.all_gather_object(ids, fake.joined)
    assert all(i[0] == ids[0][0] and len(i[0]) == 128 for i in ids) and [i[1] for i in ids] == list(range(world))
    begin, count = D.library_sample_range(rank, world, spp)
    acc = np.zeros(w * h * 4, np.int64)
    flags = np.zeros(w * h, np.uint32)
    for s in range(begin, begin + count):
        acc += _sample_contribution(w, h, s)
        flags |= _sample_flags(w, h, s)
    t = torch.from_numpy(fold_flags(acc, flags))
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)  # the library's ncclReduce(int64, sum, root 0)
    if rank == 0:
        merged, mflags = unfold_flags(t.numpy())
        np.save(out_path, merged)
        np.save(out_path + ".flags.npy", mflags)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_two_rank_gloo_reduce_equals_single_process(tmp_path):
    import torch.multiprocessing as mp

    w, h, spp = 16, 8, 11
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, _free_port(), w, h, spp, out), nprocs=2, join=True)
    single = np.zeros(w * h * 4, np.int64)
    for s in range(spp):
        single += _sample_contribution(w, h, s)
    flags = np.zeros(w * h, np.uint32)
    for s in range(spp):
        flags |= _sample_flags(w, h, s)
    assert flags.any()
    reduced = np.load(out)
    assert np.array_equal(reduced, single)
    assert np.array_equal(np.load(out + ".flags.npy"), flags)
    rgb, bounces = D.fixed_to_float(reduced, w, h)
    assert rgb.shape == (h, w, 3) and bounces.shape == (h, w) and bounces.dtype == np.uint32
    assert np.array_equal(bounces.reshape(-1), single.reshape(-1, 4)[:, 3].astype(np.uint32))
    np.testing.assert_allclose(rgb.reshape(-1, 3), single.reshape(-1, 4)[:, :3] / 2.0 ** 32, rtol=1e-6)
