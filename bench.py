#!/usr/bin/env python
"""bench.py -- Mpaths/s and Mrays/s of the path-tracing hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cornell|book1|mesh1m|book2|mesh10m] [--spp S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm (CPU oracle) on the host cores, same workload

A step = one pass of the hot path over one batch: every rank renders `spp` samples of every pixel of the workload image
(rank r takes samples [r*spp, (r+1)*spp), i.e. weak scaling: N GPUs deliver N*spp samples per pixel), the exact int64
accumulators are sum-reduced to rank 0 with NCCL, and the step ends there. `value` = paths of all ranks / max-over-ranks device
time, inputs resident in HBM. `e2e` = the same through the C ABI with host buffers: scene + camera upload, render, reduce,
device->host copy of the image, all inside the timed region. Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY.md §8d algorithmic bytes: 32 B per AABB test (one 64-B node fetch = two child boxes), 36 B per triangle test,
# 16 B per sphere test, 64 B per instance entry; extend's own queue traffic per ray: the 48-byte ray record read, the 64-byte
# shade-queue entry (ray record + hit) written.
B_NODE_VISIT, B_TRI, B_SPHERE, B_INSTANCE, B_QUEUE_EXTEND = 64, 36, 16, 64, 112

WORKLOADS = {
    # name: (BASELINE.json config index, width, height, spp, description)
    "book1": (0, 1200, 800, 10, "RTIOW book-1 final random-spheres scene, 1200x800, 10 spp, max depth 50"),
    "cornell": (1, 1024, 1024, 1000, "Cornell box (reference scenes/cornell.rs: cube.ply instances, light, rotated block, glass sphere), 1024x1024, 1000 spp, max depth 50"),
    "mesh1m": (2, 1920, 1080, 256, "cube.ply + synthetic tessellated PLY mesh (1,048,576 triangles) under the triangle BVH, 1920x1080, 256 spp"),
    "book2": (3, 1920, 1080, 1000, "RTIOW book-2 final scene restated with reference parts (1024 box instances, volumes, image texture), 1920x1080, 1000 spp"),
    "menger": (-1, 1920, 1080, 64, "extra (not a BASELINE config): Menger sponge of 160,000 cube instances (reference scenes/menger.rs at 4 levels), 1920x1080, 64 spp"),
    "mesh10m": (4, 3840, 2160, 4096, "10 synthetic meshes x 1,048,576 triangles, 3840x2160, 4096 spp split by spp across GPUs"),
}


def build_workload(name, tmpdir):
    from mass_raytrace_b200 import scenes

    if name == "book1":
        return scenes.book1_spheres(1.5, aperture=0.1)
    if name == "cornell":
        return scenes.cornell_box(1.0)
    if name == "book2":
        return scenes.book2_final()
    if name == "menger":
        return scenes.menger(levels=4)
    if name == "mesh1m":
        path = os.path.join(tmpdir, "mesh1m.ply")
        n, md = scenes.write_synthetic_ply(path, 1024, 512, seed=1)
        return scenes.lucy_layout(path, md, grid=0)
    if name == "mesh10m":
        paths, mds = [], []
        for i in range(10):
            p = os.path.join(tmpdir, f"mesh10m_{i}.ply")
            n, md = scenes.write_synthetic_ply(p, 1024, 512, seed=100 + i)
            paths.append(p)
            mds.append(md)
        return scenes.multi_mesh(paths, mds)
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for k, nme in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_sample(scene, w, h, threads, target_s, seed):
    """Times the oracle on a bounded sample of the workload: same scene and camera at 1/4 resolution per axis, spp a multiple of
    the thread count (the reference's unit of parallel work is a whole frame per thread, main.rs:251-273)."""
    sw, sh = max(w // 4, 64), max(h // 4, 64)
    t0 = time.perf_counter()
    scene.render(sw, sh, threads, 50, seed=seed, threads=threads)
    t1 = max(time.perf_counter() - t0, 1e-3)
    k = int(max(1, min(32, target_s / t1)))
    t0 = time.perf_counter()
    _, _, cnt = scene.render(sw, sh, threads * k, 50, seed=seed + 1, threads=threads)
    dt = time.perf_counter() - t0
    return cnt, dt, f"{sw}x{sh} at {threads * k} spp ({cnt['paths']} paths in {dt:.1f} s)"


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; /root/reference is Rust and cannot be built here) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_backend import OracleScene

    cfg, w, h, spp, desc = WORKLOADS[args.workload]
    with tempfile.TemporaryDirectory() as tmp:
        world, camera = build_workload(args.workload, tmp)
        scene = OracleScene(world, camera)
    cores = os.cpu_count() or 1
    budget = 150.0 / max(args.steps + args.warmup, 1)
    times, rays, paths, sample = [], 0, 0, ""
    for i in range(args.warmup + args.steps):
        cnt, dt, sample = cpu_sample(scene, w, h, cores, min(budget * 0.6, 20.0), seed=100 + 2 * i)
        if i >= args.warmup:
            times.append(dt)
            rays += cnt["rays"]
            paths += cnt["paths"]
    total = sum(times)
    v = paths / total / 1e6
    sample = f"per step {sample} of the {w}x{h}x{spp}spp workload; {cores} threads each rendering whole frames and merging (main.rs:235-294)"
    print(json.dumps({
        "impl": "reference", "metric": "Mpaths/s", "value": v, "unit": "Mpaths/s", "mrays_per_s": rays / total / 1e6, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[{cfg}]: {desc}", "name": args.workload, "width": w, "height": h, "max_depth": 50},
        "cpu_baseline": {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cornell", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel per rank per step (default: the BASELINE config's spp)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bvh", default="default", choices=["default", "sah"],
                    help="default: meshes of >= 16384 triangles get their BLAS built on the GPU (LBVH, milliseconds); sah: the host's SAH builder for every mesh")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print(f"note: --warmup {args.warmup} < 3 breaks the timing rules; using 3", file=sys.stderr)
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from mass_raytrace_b200 import NativeScene, Renderer
    from mass_raytrace_b200 import distributed as D

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size != args.gpus:
        if world_size == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
        args.gpus = world_size
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line even under NCCL_DEBUG=VERSION/INFO
        dist.init_process_group("nccl", device_id=dev)

    cfg, w, h, spp_default, desc = WORKLOADS[args.workload]
    spp = args.spp or spp_default
    tmp = tempfile.TemporaryDirectory()
    world, camera = build_workload(args.workload, tmp.name)
    t0 = time.perf_counter()
    # the backend builds its own acceleration structure at upload, so the host skips the reference's median-split build per mesh
    host = NativeScene(world, camera, defer_mesh_bvh=True)
    host.desc()
    build_s = time.perf_counter() - t0

    stream = torch.cuda.Stream(device=dev)
    r = Renderer(local_rank, stream=stream.cuda_stream)
    if args.bvh == "sah":
        r.set_option(Renderer.OPT_DEVICE_BUILD, 0)
    r.set_scene(host)
    npix = w * h
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    acc = None

    def step(i):
        """One pass: this rank's `spp` samples of every pixel, then the integer sum-reduce to rank 0."""
        nonlocal acc
        begin = (i * world_size + rank) * spp  # fresh samples every step; rank r of step i renders [begin, begin + spp)
        with torch.cuda.stream(stream):
            r.reset(w, h)
            r.accumulate(begin % (1 << 31), spp, 50, seed=2024)
            if world_size > 1:
                if acc is None:
                    acc = D.accumulators_as_tensor(r, dev)
                D.reduce_accumulators(acc, dst=0)
        return r.stats()

    # ---- device-resident timing --------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    step_ms, stats = [], []
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            sampler.start()
        with torch.cuda.stream(stream):
            flush.zero_()  # L2 flush between steps (not timed)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = step(i)
        e1.record(stream)
        barrier()
        if i >= args.warmup:
            step_ms.append(e0.elapsed_time(e1))
            stats.append(st)
    clocks = sampler.stop()
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    counts = torch.tensor([sum(s["paths"] for s in stats), sum(s["rays"] for s in stats), sum(s["kernel_launches"] for s in stats)], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_s = float(total_ms.item()) / 1e3
    paths, rays, launches = (float(x) for x in counts.tolist())
    value = paths / total_s / 1e6

    # ---- roofline of the dominant kernel (extend): algorithmic bytes per launch / mean launch duration ----------
    # Two more steps of the same workload with a CUDA-event pair around every generate / extend / shade launch on the render
    # stream. They are kept out of the steps that produce `value` because ~6 event records per wavefront iteration cost ~2-3 %.
    r.set_option(Renderer.OPT_TIME_KERNELS, 1)
    tstats, tstep_ms = [], []
    for i in range(2):
        with torch.cuda.stream(stream):
            flush.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        tstats.append(step(args.warmup + args.steps + i))
        e1.record(stream)
        barrier()
        tstep_ms.append(e0.elapsed_time(e1))
    ext_ms = sum(s["extend_ms"] for s in tstats)
    ext_launches = sum(s["iterations"] for s in tstats)  # launches that had rays (the speculative tail launches exit at once)
    rays_local = sum(s["rays"] for s in tstats)
    r.set_option(Renderer.OPT_TIME_KERNELS, 0)
    r.set_option(Renderer.OPT_COUNT_VISITS, 1)
    with torch.cuda.stream(stream):
        r.reset(w, h)
        r.accumulate(0, max(1, min(spp, 4)), 50, seed=2024)  # instrumented pass on a sample of the same workload
    cs = r.stats()
    r.set_option(Renderer.OPT_COUNT_VISITS, 0)
    per_ray = {k: cs[k] / cs["rays"] for k in ("node_visits", "tri_tests", "sphere_tests", "instance_tests")}
    b_ray = (B_NODE_VISIT * per_ray["node_visits"] + B_TRI * per_ray["tri_tests"] + B_SPHERE * per_ray["sphere_tests"] +
             B_INSTANCE * per_ray["instance_tests"] + B_QUEUE_EXTEND)
    peak, peak_src = measured_peak()
    achieved = (b_ray * rays_local / max(ext_launches, 1)) / (ext_ms / max(ext_launches, 1) * 1e-3) / 1e9 if ext_ms > 0 else 0.0
    traffic = None  # DRAM bytes per launch: the per-ray figure of an ncu capture (profiles/extend_dram_bytes.json) x this run's rays per launch
    tpath = os.path.join(ROOT, "profiles", "extend_dram_bytes.json")
    if os.path.exists(tpath):
        try:
            per_ray_dram = json.load(open(tpath)).get(args.workload)
            traffic = per_ray_dram * rays_local / max(ext_launches, 1) if per_ray_dram else None
        except Exception:
            traffic = None
    limiter = None  # the ncu view of the same kernel (issue slots, lanes per instruction): static figures of the committed captures
    try:
        limiter = json.load(open(os.path.join(ROOT, "profiles", "extend_issue.json"))).get(args.workload)
    except Exception:
        limiter = None
    roofline = {"bound": "hbm", "kernel": "k_extend", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "alg_bytes_per_ray": b_ray, "per_ray": per_ray, "rays_per_launch": rays_local / max(ext_launches, 1),
                "mean_launch_ms": ext_ms / max(ext_launches, 1), "extend_share_of_step": ext_ms / sum(tstep_ms),
                "shade_share_of_step": sum(s["shade_ms"] for s in tstats) / sum(tstep_ms),
                "generate_share_of_step": sum(s["generate_ms"] for s in tstats) / sum(tstep_ms),
                "ncu_limiter": limiter,
                "note": "algorithmic bytes are served mostly by L1/L2 when the acceleration structure is cache-resident, so achieved can exceed the HBM peak; "
                        "the kernel is issue-bound on divergent code (ncu_limiter)"}

    # ---- end to end through the C ABI with host buffers ----------------------------------------------------------
    out_rgb = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
    out_b = torch.empty((h, w), dtype=torch.int32).pin_memory()
    rgb_np, b_np = out_rgb.numpy(), out_b.numpy().view(np.uint32)
    e2e_t = []
    for i in range(1 + args.e2e_steps):
        barrier()
        t0 = time.perf_counter()
        r.set_scene(host)  # H2D: flattened scene + camera, every step
        begin = ((1000 + i) * world_size + rank) * spp
        if world_size == 1:
            r.render(w, h, spp, 50, seed=2024, spp_begin=begin % (1 << 31), out=(rgb_np, b_np))  # render + D2H of sum_rgb / sum_bounces
        else:
            with torch.cuda.stream(stream):
                r.reset(w, h)
                r.accumulate(begin % (1 << 31), spp, 50, seed=2024)
                acc = D.accumulators_as_tensor(r, dev)
                D.reduce_accumulators(acc, dst=0)
            if rank == 0:
                r.size = (w, h)
                rgb_np[...], b_np[...], _ = r.download()
        barrier()
        if i >= 1:
            e2e_t.append(time.perf_counter() - t0)
    e2e_s = torch.tensor([sum(e2e_t)], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    scene_bytes = r.stats()["scene_bytes"] + 76
    e2e = None
    if e2e_t:
        e2e_value = (npix * spp * world_size * len(e2e_t)) / float(e2e_s.item()) / 1e6
        e2e = {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(npix * 16),
               "ms_per_step": 1e3 * float(e2e_s.item()) / len(e2e_t)}

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) -----------------------------------------------------
    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_backend import OracleScene

        orc = OracleScene(world, camera)
        cores = os.cpu_count() or 1
        threads = max(cores - 2, 1)  # the reference's policy, main.rs:159-160
        cnt, dt, sample = cpu_sample(orc, w, h, threads, 12.0, seed=1)
        cpu = {"value": cnt["paths"] / dt / 1e6, "unit": "Mpaths/s", "mrays_per_s": cnt["rays"] / dt / 1e6, "cores": threads, "host_cores": cores, "kind": "port",
               "sample": f"{sample}: C++ restatement of the reference algorithm, {threads} threads = max(cores-2,1) each rendering whole frames "
                         f"and merging (main.rs:159-160, 235-294)"}

    if rank == 0:
        line = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "mrays_per_s": rays / total_s / 1e6, "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[{cfg}]: {desc}", "name": args.workload, "width": w, "height": h, "spp_per_gpu_per_step": spp, "max_depth": 50,
                       "partition": f"samples per pixel split over {world_size} rank(s), one NCCL int64 sum-reduce per step" if world_size > 1 else "single GPU",
                       "l2": "flushed between steps (256 MiB write); the ray / shade queues of one iteration (1.8 GB at 16M paths in flight) also exceed the 126 MB L2",
                       "scene_build_s": build_s, "host_mesh_bvh": "deferred (mrth_defer_mesh_bvh)", "rays_per_path": rays / paths,
                       "bvh": "host SAH for every mesh" if args.bvh == "sah" else "GPU LBVH for meshes of >= 16384 triangles, host SAH otherwise"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    r.close()
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
