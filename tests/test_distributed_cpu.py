"""The N>1 path on CPU: world_size-2 gloo processes. Each rank owns a sample range (distributed.sample_range), fills exact
int64 accumulators for its samples, and one SUM reduce must reproduce the single-process image bit for bit -- the property the
GPU path relies on (integer accumulators + a sample-keyed RNG make the result independent of the split)."""
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
    begin, count = D.sample_range(rank, world, spp)
    acc = np.zeros(w * h * 4, np.int64)
    for s in range(begin, begin + count):
        acc += _sample_contribution(w, h, s)
    t = torch.from_numpy(acc)
    D.reduce_accumulators(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy())
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
    reduced = np.load(out)
    assert np.array_equal(reduced, single)
    rgb, bounces = D.fixed_to_float(reduced, w, h)
    assert rgb.shape == (h, w, 3) and bounces.shape == (h, w) and bounces.dtype == np.uint32
    assert np.array_equal(bounces.reshape(-1), single.reshape(-1, 4)[:, 3].astype(np.uint32))
    np.testing.assert_allclose(rgb.reshape(-1, 3), single.reshape(-1, 4)[:, :3] / 2.0 ** 32, rtol=1e-6)
