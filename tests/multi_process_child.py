"""Child of tests/test_gpu_multi.py::test_one_process_per_gpu_*: one rank of a torch.distributed.run launch. Joins the library's
communicator, uploads the scene collectively (only rank 0 passes one), renders collectively and writes rank 0's image hash."""
import ctypes as C
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from mass_raytrace_b200 import NativeScene, Renderer, scenes
from mass_raytrace_b200 import distributed as D

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
out_path, mesh_path = sys.argv[1], sys.argv[2]
torch.cuda.set_device(rank)
dist.init_process_group("gloo")  # the host's own transport only carries the 128-byte NCCL id
r = Renderer(rank)
D.join_communicator(r, rank, world)
assert r.comm_rank() == (rank, world)
lines = []
n, md = scenes.write_synthetic_ply(mesh_path + f".{rank}", 256, 128, seed=4)
for name, (w, c), W, H, spp in (("cornell", scenes.cornell_box(1.0), 96, 96, 13), ("mesh", scenes.lucy_layout(mesh_path + f".{rank}", md, grid=0), 160, 90, 5)):
    host = NativeScene(w, c)
    if rank == 0:
        r.set_scene(host)
    else:  # the other ranks' scene is not read: NULL is allowed; the camera is set per rank
        r._check(r.lib.mrt_scene_upload(r._h, None), "mrt_scene_upload(NULL) on a non-root rank")
        r._check(r.lib.mrt_camera_set(r._h, host.camera_struct()), "mrt_camera_set")
    rgb, b, cnt = r.render(W, H, spp, 50, seed=11)
    if rank == 0:
        assert cnt == spp
        lines.append(f"{name} {hashlib.sha256(rgb.tobytes() + b.tobytes()).hexdigest()} {r.stats()['scene_bytes']}")
    else:
        assert cnt == 0
# a scene the root rejects fails on every rank instead of hanging the others
host = NativeScene(*scenes.cornell_box(1.0))
desc = host.desc()
desc.contents.abi_version = 12345
rc = r.lib.mrt_scene_upload(r._h, desc if rank == 0 else None)
assert rc != 0, rc
lines.append(f"rejected {rc} {r.lib.mrt_last_error(r._h).decode()}")
with open(f"{out_path}.{rank}", "w") as f:
    f.write("\n".join(lines) + "\n")
r.close()
dist.barrier()
dist.destroy_process_group()
