#!/usr/bin/env python
"""bench.py -- Mpaths/s and Mrays/s of the path-tracing hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mesh1m|book1|cornell|book2|mesh10m|menger] [--spp S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm (CPU oracle) on the host cores, same workload

A step = one pass of the hot path over one batch: the workload's image at the workload's samples per pixel. With N ranks the
samples of every pixel are SPLIT over the ranks (strong scaling: the job is the same at every N, `distributed.sample_range`),
each rank accumulates its share into its exact int64 image and ONE NCCL sum-reduce inside the library (mrt_comm_*, include/mrt.h)
merges them onto rank 0 before the step ends. `value` = paths of the whole job / max-over-ranks device time, inputs resident in
HBM. `e2e` = the same through the C ABI with host buffers: scene + camera upload (with N ranks the root uploads and builds once and
the library broadcasts the device arrays over NVLink), render, reduce, device->host copy of the image, all inside the timed region. The headline workload is configs[2] (the 1 M-triangle mesh north_star names); the other
BASELINE configs are measured in the same run with fewer steps and reported under `per_config`. `parity` = a fixed small job of
the headline scene rendered by all ranks together: its SHA-256 must be the same at every N, and rank 0 checks it against the
CPU oracle. Prints ONE JSON line (rank 0).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic bytes of k_extend per unit of work: one node visit fetches the 64-byte node record, a
# triangle test its 3 vertices (36 B of payload), a sphere test 16 B, an instance entry the inverse matrix + meta (64 B); per
# ray the 48-byte ray record is read and the 64-byte shade-queue entry (ray record + hit) written.
B_TRI, B_SPHERE, B_INSTANCE, B_QUEUE_EXTEND = 36, 16, 64, 112

WORKLOADS = {
    # name: (BASELINE.json config index, width, height, spp, description)
    "book1": (0, 1200, 800, 10, "RTIOW book-1 final random-spheres scene, 1200x800, 10 spp, max depth 50"),
    "cornell": (1, 1024, 1024, 1000, "Cornell box (reference scenes/cornell.rs: cube.ply instances, light, rotated block, glass sphere), 1024x1024, 1000 spp, max depth 50"),
    "mesh1m": (2, 1920, 1080, 256, "cube.ply + synthetic tessellated PLY mesh (1,048,576 triangles) under the triangle BVH, 1920x1080, 256 spp"),
    "book2": (3, 1920, 1080, 1000, "RTIOW book-2 final scene restated with reference parts (1024 box instances, volumes, image texture), 1920x1080, 1000 spp"),
    "menger": (-1, 1920, 1080, 64, "extra (not a BASELINE config): Menger sponge of 160,000 cube instances (reference scenes/menger.rs at 4 levels), 1920x1080, 64 spp"),
    "mesh10m": (4, 3840, 2160, 4096, "10 synthetic meshes x 1,048,576 triangles, 3840x2160, 4096 spp split by spp across GPUs"),
}
PER_CONFIG = (("book1", 0), ("cornell", 0), ("book2", 0), ("mesh10m", 128))  # (workload, spp per step; 0 = the config's own)


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


def cpu_threads():
    """ONE thread policy for every CPU number this file prints (cpu_baseline and --impl reference): the reference's own,
    max(num_cpus - 2, 1) render threads (main.rs:159-160)."""
    cores = os.cpu_count() or 1
    return max(cores - 2, 1), cores


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
    """--impl reference: the reference's CPU algorithm (oracle port; /root/reference is Rust and cannot be built here) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_backend import OracleScene

    cfg, w, h, spp, desc = WORKLOADS[args.workload]
    with tempfile.TemporaryDirectory() as tmp:
        world, camera = build_workload(args.workload, tmp)
        scene = OracleScene(world, camera)
    threads, cores = cpu_threads()
    budget = 150.0 / max(args.steps + args.warmup, 1)
    times, rays, paths, sample = [], 0, 0, ""
    for i in range(args.warmup + args.steps):
        cnt, dt, sample = cpu_sample(scene, w, h, threads, min(budget * 0.6, 20.0), seed=100 + 2 * i)
        if i >= args.warmup:
            times.append(dt)
            rays += cnt["rays"]
            paths += cnt["paths"]
    total = sum(times)
    v = paths / total / 1e6
    sample = (f"per step {sample} of the {w}x{h}x{spp}spp workload; {threads} threads = max(cores-2,1) of {cores} host cores, each rendering whole "
              f"frames and merging (main.rs:159-160, 235-294)")
    print(json.dumps({
        "impl": "reference", "metric": "Mpaths/s", "value": v, "unit": "Mpaths/s", "mrays_per_s": rays / total / 1e6, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[{cfg}]: {desc}", "name": args.workload, "width": w, "height": h, "spp_per_step": args.spp or spp, "max_depth": 50,
                   "cpu_thread_policy": "max(host cores - 2, 1) (main.rs:159-160) for cpu_baseline and --impl reference alike",
                   "sample": "each step renders a bounded sample of this workload (cpu_baseline.sample); the value is the normalised rate"},
        "cpu_baseline": {"value": v, "unit": "Mpaths/s", "cores": threads, "host_cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="mesh1m", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel per step for the whole job (default: the BASELINE config's spp)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true", help="skip the sub-records of the other BASELINE configs")
    ap.add_argument("--no-parity", action="store_true")
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

    stream = torch.cuda.Stream(device=dev)
    r = Renderer(local_rank, stream=stream.cuda_stream)
    D.join_communicator(r, rank, world_size)  # the library's own NCCL communicator: rank 0's id travels through torch.distributed
    if args.bvh == "sah":
        r.set_option(Renderer.OPT_DEVICE_BUILD, 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    tmp = tempfile.TemporaryDirectory()
    peak, peak_src = measured_peak()

    def barrier():
        torch.cuda.synchronize(dev)
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allreduce(vals, op):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(t, op=op)
        return [float(x) for x in t.tolist()]

    ncu_side = {}
    try:
        ncu_side = json.load(open(os.path.join(ROOT, "profiles", "extend_ncu.json")))
    except Exception:
        ncu_side = {}

    def measure(name, spp_override, steps, warmup, detail):
        """All measurements of one workload. Returns (record, world, camera, host scene)."""
        cfg, w, h, spp_default, desc = WORKLOADS[name]
        spp = spp_override or spp_default
        npix = w * h
        world, camera = build_workload(name, tmp.name)
        t0 = time.perf_counter()
        host = NativeScene(world, camera, defer_mesh_bvh=True)  # the backend builds its own acceleration structure at upload
        host.desc()
        build_s = time.perf_counter() - t0
        r.set_scene(host)

        def step(i, split=True):
            """One pass. split: the job's spp divided over the ranks + merge (strong); else every rank renders spp of its own (weak)."""
            with torch.cuda.stream(stream):
                r.reset(w, h)
                if split or world_size == 1:
                    r.accumulate((i * spp) % (1 << 30), spp, 50, seed=2024)  # collective on a communicator: split + one reduce inside
                else:
                    r.accumulate(((i * world_size + rank) * spp) % (1 << 30), spp, 50, seed=2024)
                    r.comm_reduce()
            return r.stats()

        def timed(n_warm, n_steps, split=True, first=0):
            ms, stats = [], []
            for i in range(n_warm + n_steps):
                with torch.cuda.stream(stream):
                    flush.zero_()  # L2 flush between steps (not timed)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                st = step(first + i, split)
                e1.record(stream)
                barrier()
                if i >= n_warm:
                    ms.append(e0.elapsed_time(e1))
                    stats.append(st)
            total_ms = allreduce([sum(ms)], dist.ReduceOp.MAX)[0]
            sums = allreduce([sum(s["paths"] for s in stats), sum(s["rays"] for s in stats), sum(s["kernel_launches"] for s in stats)], dist.ReduceOp.SUM)
            timed.last_steps_ms = [round(x, 3) for x in ms]  # this rank's steps one by one: a step far off the others is a path that met a pathological ray
            return total_ms / 1e3, sums

        # ---- device-resident timing (strong scaling) ------------------------------------------------------------------------
        total_s, (paths, rays, launches) = timed(warmup, steps)
        rec = {"value": paths / total_s / 1e6, "unit": "Mpaths/s", "mrays_per_s": rays / total_s / 1e6, "ms_per_step": 1e3 * total_s / steps, "steps": steps,
               "warmup": warmup, "workload": f"configs[{cfg}]: {desc}", "width": w, "height": h, "spp_per_step": spp, "rays_per_path": rays / max(paths, 1),
               "scene_build_s": build_s, "steps_ms": timed.last_steps_ms}
        rec["_launches"] = launches + (4 * world_size * steps if world_size > 1 else 0)  # + fold / cell / 2 NCCL kernels / unfold per rank and step

        # ---- roofline of the dominant kernel (extend) -------------------------------------------------------------------------
        # Two more steps with a CUDA-event pair around every generate / extend / shade launch on the render stream (kept out of
        # the steps that produce `value`: ~6 event records per wavefront iteration cost 2-3 %), then one instrumented pass that
        # counts node visits and primitive tests per ray.
        r.set_option(Renderer.OPT_TIME_KERNELS, 1)
        tstats, tstep_ms = [], []
        for i in range(2):
            with torch.cuda.stream(stream):
                flush.zero_()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            tstats.append(step(warmup + steps + i))
            e1.record(stream)
            barrier()
            tstep_ms.append(e0.elapsed_time(e1))
        r.set_option(Renderer.OPT_TIME_KERNELS, 0)
        ext_ms = sum(s["extend_ms"] for s in tstats)
        ext_launches = max(sum(s["iterations"] for s in tstats), 1)  # launches that had rays (the speculative tail launches exit at once)
        rays_local = sum(s["rays"] for s in tstats)
        r.set_option(Renderer.OPT_COUNT_VISITS, 1)
        r.set_option(Renderer.OPT_COMM_SPLIT, 0)
        with torch.cuda.stream(stream):
            r.reset(w, h)
            r.accumulate(0, max(1, min(spp, 4)), 50, seed=2024)
        cs = r.stats()
        r.set_option(Renderer.OPT_COMM_SPLIT, 1)
        r.set_option(Renderer.OPT_COUNT_VISITS, 0)
        per_ray = {k: cs[k] / max(cs["rays"], 1) for k in ("node_visits", "tri_tests", "sphere_tests", "instance_tests")}
        b_node = cs.get("node_bytes", 64)
        b_ray = b_node * per_ray["node_visits"] + B_TRI * per_ray["tri_tests"] + B_SPHERE * per_ray["sphere_tests"] + B_INSTANCE * per_ray["instance_tests"] + B_QUEUE_EXTEND
        achieved = (b_ray * rays_local) / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        side = ncu_side.get(name) or {}
        rays_per_launch = rays_local / ext_launches
        roofline = {
            # What binds the kernel is instruction issue on divergent code, not bandwidth (ncu: profiles/). `achieved` / `frac` stay the
            # algorithmic-bytes figure north_star asks for; issue_frac and dram_frac, from the ncu capture of this build named in
            # `ncu`, say how close the kernel is to the limits that actually exist.
            "bound": "issue", "kernel": "k_extend", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
            "traffic": side.get("dram_bytes_per_ray") * rays_per_launch if side.get("dram_bytes_per_ray") else None,
            "issue_frac": side.get("issue_frac"), "dram_frac": side.get("dram_frac"), "ncu": side or None,
            "alg_bytes_per_ray": b_ray, "node_bytes": b_node, "per_ray": per_ray, "rays_per_launch": rays_per_launch, "mean_launch_ms": ext_ms / ext_launches,
            "extend_share_of_step": ext_ms / sum(tstep_ms), "shade_share_of_step": sum(s["shade_ms"] for s in tstats) / sum(tstep_ms),
            "generate_share_of_step": sum(s["generate_ms"] for s in tstats) / sum(tstep_ms),
            "note": "achieved = algorithmic bytes x rays / k_extend time (CUDA events, this run). The acceleration structure is served mostly by L1/L2, so "
                    "the figure can exceed the HBM peak; issue_frac = issue slots busy x active lanes / 32 and dram_frac = DRAM throughput / peak come from "
                    "the ncu --set full capture of this build (static, regenerated by tools/ncu_side_files.py)"}
        rec["roofline"] = roofline

        # ---- end to end through the C ABI with host buffers: same step count, same flush ----------------------------------------
        out_rgb = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
        out_b = torch.empty((h, w), dtype=torch.int32).pin_memory()
        rgb_np, b_np = out_rgb.numpy(), out_b.numpy().view(np.uint32)
        e2e_t = []
        for i in range(1 + steps):
            with torch.cuda.stream(stream):
                flush.zero_()
            barrier()
            t0 = time.perf_counter()
            r.set_scene(host)  # H2D: flattened scene + camera, every step (N > 1: rank 0 uploads and builds, the finished arrays are broadcast over NVLink)
            with torch.cuda.stream(stream):
                r.render(w, h, spp, 50, seed=2024, spp_begin=((2000 + i) * spp) % (1 << 30), out=(rgb_np, b_np))  # render (+ reduce) + D2H on the root
            barrier()
            if i >= 1:
                e2e_t.append(time.perf_counter() - t0)
        e2e_s = allreduce([sum(e2e_t)], dist.ReduceOp.MAX)[0]
        scene_bytes = r.stats()["scene_bytes"] + 76
        rec["e2e"] = {"value": (npix * spp * len(e2e_t)) / e2e_s / 1e6, "unit": "Mpaths/s", "h2d_bytes_per_step": int(scene_bytes) + 76 * (world_size - 1),
                      "d2h_bytes_per_step": int(npix * 16), "ms_per_step": 1e3 * e2e_s / len(e2e_t), "steps": len(e2e_t)}

        if detail and world_size > 1:  # weak scaling beside it: every rank renders the config's full spp, N x the samples per step
            r.set_option(Renderer.OPT_COMM_SPLIT, 0)  # ranges chosen here, merge called explicitly (mrt_comm_reduce)
            weak_s, (wpaths, wrays, _) = timed(3, 3, split=False, first=5000)
            r.set_option(Renderer.OPT_COMM_SPLIT, 1)
            rec["weak"] = {"value": wpaths / weak_s / 1e6, "unit": "Mpaths/s", "mrays_per_s": wrays / weak_s / 1e6, "ms_per_step": 1e3 * weak_s / 3,
                           "spp_per_gpu_per_step": spp, "steps": 3}
        return rec, world, camera, host

    # ---- the headline workload -------------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    main_rec, world, camera, host = measure(args.workload, args.spp, args.steps, args.warmup, detail=True)
    clocks = sampler.stop()
    cfg, w, h, spp_default, desc = WORKLOADS[args.workload]

    # ---- parity: one fixed small job of the same scene, rendered by all ranks together -----------------------------------------
    parity = None
    if not args.no_parity:
        pw = 256
        ph = max(2, int(round(pw * h / w)))
        pspp = 64
        r.set_scene(host)
        with torch.cuda.stream(stream):
            r.reset(pw, ph)
            r.accumulate(0, pspp, 50, seed=4242)
        if rank == 0:
            rgb, bnc, cnt = r.download()
            digest = hashlib.sha256(rgb.tobytes() + bnc.tobytes()).hexdigest()
            parity = {"job": f"{args.workload} scene, {pw}x{ph}, {pspp} spp, seed 4242, samples split over {world_size} rank(s), one NCCL reduce",
                      "image_sha256": digest, "count": int(cnt)}
            # the checker: the CPU oracle renders the same job twice (two seeds); the GPU image must be as close to an oracle image
            # as two oracle images are to each other (criterion of tests/test_gpu_parity.py::stat_compare)
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from oracle_backend import OracleScene

            orc = OracleScene(world, camera)
            threads, _ = cpu_threads()
            a_rgb, a_b, _ = orc.render(pw, ph, pspp, 50, seed=4242, threads=threads)
            b_rgb, b_b, _ = orc.render(pw, ph, pspp, 50, seed=4243, threads=threads)
            Y = np.array([0.2126, 0.7152, 0.0722])
            rmse = lambda x, y: float(np.sqrt(np.mean((x.astype(np.float64) / pspp - y.astype(np.float64) / pspp) ** 2)))
            lum = lambda x: float((x.astype(np.float64) * Y).sum(-1).mean() / pspp)
            rmse_oo = rmse(a_rgb, b_rgb)
            rmse_go = 0.5 * (rmse(rgb, a_rgb) + rmse(rgb, b_rgb))
            lum_o = 0.5 * (lum(a_rgb) + lum(b_rgb))
            mb_o = 0.5 * (a_b.mean() + b_b.mean()) / pspp
            vs = {"rmse_gpu_oracle": rmse_go, "rmse_oracle_oracle": rmse_oo, "rmse_ratio": rmse_go / max(rmse_oo, 1e-12), "lum_gpu": lum(rgb), "lum_oracle": lum_o,
                  "dlum_rel": (lum(rgb) - lum_o) / max(lum_o, 1e-12), "bounces_gpu": float(bnc.mean() / pspp), "bounces_oracle": float(mb_o)}
            vs["ok"] = bool(vs["rmse_ratio"] <= 1.15 and abs(vs["dlum_rel"]) <= 0.02 and abs(vs["bounces_gpu"] - mb_o) <= 0.02 * mb_o and cnt == pspp)
            parity["vs_oracle"] = vs
        barrier()

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) -----------------------------------------------------
    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_backend import OracleScene

        orc = OracleScene(world, camera)
        threads, cores = cpu_threads()
        cnt, dt, sample = cpu_sample(orc, w, h, threads, 12.0, seed=1)
        cpu = {"value": cnt["paths"] / dt / 1e6, "unit": "Mpaths/s", "mrays_per_s": cnt["rays"] / dt / 1e6, "cores": threads, "host_cores": cores, "kind": "port",
               "sample": f"{sample}: C++ restatement of the reference algorithm, {threads} threads = max(cores-2,1) each rendering whole frames "
                         f"and merging (main.rs:159-160, 235-294); the same policy as --impl reference"}

    # ---- the other BASELINE configs, same run, fewer steps -------------------------------------------------------------
    per_config = {}
    launches_total = main_rec.pop("_launches")
    if not args.no_per_config:
        del world, camera, host
        for name, spp_step in PER_CONFIG:
            if name == args.workload:
                continue
            rec, _, _, hst = measure(name, spp_step, 3, 3, detail=False)
            rec.pop("_launches")
            hst.close()
            key = name if not spp_step else f"{name}@{spp_step}spp"
            per_config[key] = rec

    if rank == 0:
        spp = args.spp or spp_default
        line = {
            "metric": "Mpaths/s", "value": main_rec["value"], "unit": "Mpaths/s", "mrays_per_s": main_rec["mrays_per_s"], "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_rec["ms_per_step"], "steps_ms": main_rec["steps_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[{cfg}]: {desc}", "name": args.workload, "width": w, "height": h, "spp_per_step": spp, "max_depth": 50,
                       "partition": (f"the {spp} samples of every pixel split over {world_size} ranks (mrt_sample_range), one NCCL int64 sum-reduce per step inside "
                                     f"the library" if world_size > 1 else "single GPU"),
                       "l2": "flushed between steps (256 MiB write); the ray / shade queues of one iteration (3.6 GB at 32M paths in flight) also exceed the 126 MB L2",
                       "scene_build_s": main_rec["scene_build_s"], "host_mesh_bvh": "deferred (mrth_defer_mesh_bvh)", "rays_per_path": main_rec["rays_per_path"],
                       "bvh": "host SAH for every mesh" if args.bvh == "sah" else "GPU-built for meshes of >= 16384 triangles, host SAH otherwise",
                       "cpu_thread_policy": "max(host cores - 2, 1) (main.rs:159-160) for cpu_baseline and --impl reference alike"},
            "clocks": clocks, "e2e": main_rec["e2e"], "gpu_launches": int(launches_total), "roofline": main_rec["roofline"], "cpu_baseline": cpu,
            "parity": parity, "weak": main_rec.get("weak"), "per_config": per_config or None,
        }
        print(json.dumps(line), flush=True)
    r.close()
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
