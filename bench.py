#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 ray-tracing backend (BASELINE.json metric).

A "step" is one C3 pass (BASELINE config 2): 1024x1024 primary rays generated on the device ->
closest hit against the 10,008,338-triangle displaced grid -> shadow rays to the point light (any-hit) and
cosine-bounce rays (incoherent closest hit) spawned from the hits -> both traced.  3 x 1,048,576 rays per step.

  value     = rays traced / device time, inputs resident in HBM (CUDA events on the launching stream); the shadow and the
              bounce batch of a step are independent and go out on two streams (joined before the step's end event)
  e2e       = the same three batches through the host-buffer C ABI from pinned host memory, H2D + D2H inside the timed
              region of every step: enqueued back to back (pb2_intersect_async / pb2_intersect_p_async), one step kept in
              flight while the previous step's results are retired (pb2_scene_wait_until), the last step waited for in
              full; the synchronous calls (pb2_intersect / pb2_intersect_p, one drain per batch) are timed beside it
  roofline  = algorithmic bytes of the incoherent closest-hit launch (oracle-counted nodes/triangles in reference order,
              SURVEY §8d) / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference = the CPU restatement of the reference (oracle/, all host threads) on a bounded sample

Launch: `python bench.py --gpus 1` or `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N`.
With N > 1 every rank traces its own full pass from a camera rotated about the y axis (weak scaling, BVH replicated,
no data-path collective for ray casting).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid-n", type=int, default=2237, help="quads per side of the C3 grid (2237 -> 10,008,338 triangles)")
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--write-c5-anchor", action="store_true", help="N = 1 only: write profiles/c5_n1_anchor.json from this run's C5 frame")
    ap.add_argument("--no-path", action="store_true", help="tuning / profiling runs: C3 traversal only (the default line carries C2, C4 and C5 too)")
    ap.add_argument("--cpu-stride", type=int, default=1, help="cpu sample = every k-th ray of each ray set")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline: repeat the sample until this much traversal time")
    return ap.parse_args()


def rank_info():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def camera_for_rank(scenes, rank, res):
    cam = dict(scenes.C3_CAMERA, res=(res, res))
    if rank:
        a = 2.0 * np.pi * rank / 8.0
        x, y, z = cam["pos"]
        cam["pos"] = (float(x * np.cos(a) - z * np.sin(a)), y, float(x * np.sin(a) + z * np.cos(a)))
    return cam


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def algorithmic_bytes(n_rays, nodes, tris, hit_bytes):
    # B(r) = 32 (ray in) + 16 or 4 (hit / any-hit out) + 32 N_nodes(r) + 36 N_tris(r)
    return n_rays * (32 + hit_bytes) + 32 * int(nodes) + 36 * int(tris)


PATH_STATE_BYTES_8D = 160     # SoA path state read + write per path segment (ray 32 + beta 12 + L 12 + pixel 4 + rng 16 + flags 4, each way)


def path_roofline(rays_gpu, n_samples, ms, cnt, peak, peak_src, traffic_key):
    """SURVEY §8d, per path sample: B = sum over its rays of B(r) + 160 x (path segments) + 16 (film).  rays_gpu = the device's
    own ray counts for the timed frame; nodes / triangles per ray of each kind from the oracle's instrumented render of a
    bounded sample range of the same frame (reference traversal order).  The achieved figure is whole-frame: every kernel
    of the wavefront (extend, shade, shadow, compaction, film) runs inside `ms`."""
    kinds = (("extend", "extend_rays", 16), ("shadow", "shadow_rays", 4), ("mis", "mis_rays", 16))
    ray_bytes, per_kind = 0.0, {}
    for k, gk, hb in kinds:
        r = max(1, cnt["rays"][k])
        npr, tpr = cnt["nodes"][k] / r, cnt["tris"][k] / r
        b = rays_gpu[gk] * (32 + hb + 32 * npr + 36 * tpr)
        ray_bytes += b
        per_kind[k] = {"rays": rays_gpu[gk], "nodes_per_ray": npr, "tris_per_ray": tpr, "bytes": b}
    state_bytes = PATH_STATE_BYTES_8D * rays_gpu["extend_rays"]
    total = ray_bytes + state_bytes + 16.0 * n_samples
    gbs = total / (ms * 1e-3) / 1e9
    out = {"bound": "hbm", "kernel": "whole wavefront frame (k_extend + k_shade + k_shadow + queue compaction + film)", "achieved": gbs, "peak": peak,
           "unit": "GB/s", "frac": gbs / peak, "peak_source": peak_src, "algorithmic_bytes": total,
           "bytes_per_sample": total / n_samples, "ray_bytes": ray_bytes, "state_bytes_8d": state_bytes, "film_bytes": 16.0 * n_samples,
           "per_ray_kind": per_kind, "segments_per_sample": rays_gpu["extend_rays"] / n_samples,
           "state_bytes_actual_per_segment": 250,
           "note": "the §8d state term counts 160 B per segment; the SoA arrays this build reads and writes per segment (ray, hit, beta, L, rng, "
                   "NEE record, state byte, queue entries) come to ~250 B, so real HBM traffic for state is ~1.6x the term above",
           "traffic": None}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        t = tj.get("path", {}).get(traffic_key)
        if t:
            # dram bytes of one ncu-captured batch, scaled to the timed frame by camera samples
            out["traffic"] = t["dram_bytes"] * (n_samples / t["camera_samples"])
            out["traffic_source"] = t.get("source")
            out["dram_frac"] = out["traffic"] / (ms * 1e-3) / 1e9 / peak
            out["kernel_share"] = t.get("kernel_share")
    except (OSError, KeyError, ValueError):
        pass
    return out


def load_peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    return peak, ("measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)")


def bench_config(n_tris, res):
    """The `config` object of both arms (the repo's and --impl reference): same keys, same values."""
    return {"workload": f"C3: {n_tris}-triangle displaced grid, {res}x{res} primary closest-hit + shadow any-hit + "
                        "incoherent bounce closest-hit, SAH BVH max_prims_in_node=4", "rays_per_step_per_gpu": 3 * res * res}


def host_info():
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return os.cpu_count() or 1, model


def path_steps(args):
    return max(1, min(args.steps, 5)), max(1, min(args.warmup, 3))


def bench_path_c2(pb2, scenes, torch, args, dist, world):
    """Path-traced Msamples/s on BASELINE config 1 (Cornell box, maxdepth 5, 512x512 @ 64 spp, box filter): one step = one
    full frame (16.8 M camera samples) rendered into a device-resident Film.  e2e adds the film read-back to the host."""
    sc = scenes.scene_c2()
    cam = scenes.C2_CAMERA
    pk = scenes.C2_PATH
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4)
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    integ = pb2.PathIntegrator(accel, camera, **pk)
    film = pb2.Film(cam["res"])
    stream = torch.cuda.current_stream().cuda_stream
    n_samples = cam["res"][0] * cam["res"][1] * pk["spp"]
    steps, warmup = path_steps(args)
    for _ in range(warmup):
        integ.render(film, stream=stream)
    torch.cuda.synchronize()
    c0 = integ.counters()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        film.clear()
        a.record()
        integ.render(film, stream=stream)
        b.record()
    torch.cuda.synchronize()
    c1 = integ.counters()
    ms = [a.elapsed_time(b) for a, b in ev]
    t0 = time.perf_counter()
    for _ in range(steps):
        film.clear()
        integ.render(film, stream=stream)
        xyzw = film.read_xyzw()
    e2e_s = time.perf_counter() - t0
    tot = float(sum(ms))
    if world > 1:
        t = torch.tensor([tot, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot, e2e_s = float(t[0]), float(t[1])
    rays = {k: (c1[k] - c0[k]) / steps for k in ("extend_rays", "shadow_rays", "mis_rays")}
    return {"workload": "C2: Cornell box (32 triangles, matte, quad area light), PathIntegrator maxdepth=5, 512x512 @ 64 spp, RandomSampler "
                        "streams per (pixel, sample), box filter", "unit": "Msamples/s", "value": world * n_samples * steps / (tot * 1e-3) / 1e6,
            "ms_per_frame": tot / steps, "samples_per_frame": n_samples,
            "e2e": {"value": world * n_samples * steps / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 128,
                    "d2h_bytes_per_step": int(xyzw.nbytes), "api": "pb2_render_path + pb2_film_read_xyzw"},
            "rays_per_frame": rays, "mrays_per_s": sum(rays.values()) / (tot / steps * 1e-3) / 1e6,
            "kernel_launches_per_frame": (c1["kernel_launches"] - c0["kernel_launches"]) / steps,
            "mean_rgb": [float(v) for v in pb2_mean_rgb(film)]}, (accel, camera, integ, film, sc)


def bench_path_c4(pb2, scenes, torch, args, dist, world, spp_timed=256):
    """BASELINE config 3 (mixed matte / plastic / glass, point + area light, maxdepth 8, 1920x1080 @ 256 spp, power light
    distribution, material-sorted shading): a step renders the whole frame, all 256 samples of every pixel (530.8 M camera
    samples, ~1.4 s on one B200)."""
    sc = scenes.scene_c4()
    cam = scenes.C4_CAMERA
    pk = dict(scenes.C4_PATH)
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4)
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    integ = pb2.PathIntegrator(accel, camera, **pk)
    film = pb2.Film(cam["res"])
    stream = torch.cuda.current_stream().cuda_stream
    integ.render(film, 0, 4, stream=stream)
    torch.cuda.synchronize()
    c0 = integ.counters()
    steps = 2
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        film.clear()
        a.record()
        integ.render(film, 0, spp_timed, stream=stream)
        b.record()
    torch.cuda.synchronize()
    c1 = integ.counters()
    tot = float(sum(a.elapsed_time(b) for a, b in ev))
    # e2e: the call a user makes — render the frame, then read the film back to the host (wall clock, copy included)
    film.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    integ.render(film, 0, spp_timed, stream=stream)
    xyzw = film.read_xyzw()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([tot, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot, e2e_s = float(t[0]), float(t[1])
    n_samples = cam["res"][0] * cam["res"][1] * spp_timed
    rays = {k: (c1[k] - c0[k]) / steps for k in ("extend_rays", "shadow_rays", "mis_rays")}
    return {"workload": f"C4: {len(sc['idx'])}-triangle room, matte / plastic / glass spheres, area + point light, PathIntegrator maxdepth=8, "
                        f"1920x1080, sample indices [0,{spp_timed}) of 256 spp per step, power light distribution",
            "unit": "Msamples/s", "value": world * n_samples * steps / (tot * 1e-3) / 1e6, "ms_per_step": tot / steps,
            "samples_per_step": n_samples, "rays_per_step": rays, "mrays_per_s": sum(rays.values()) / (tot / steps * 1e-3) / 1e6,
            "e2e": {"value": world * n_samples / e2e_s / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 192,
                    "d2h_bytes_per_step": int(xyzw.nbytes), "api": "pb2_render_path + pb2_film_read_xyzw (wall clock, one frame)"},
            "kernel_launches_per_step": (c1["kernel_launches"] - c0["kernel_launches"]) / steps,
            "mean_rgb": [float(v) for v in pb2_mean_rgb(film)]}, (accel, camera, integ, film, sc)


def pb2_mean_rgb(film):
    return film.resolve_rgb().mean(axis=(0, 1))


PATH_CPU_SPP = 16


def bench_path_cpu(orc_mod, scenes, sc, gpu_film_xyzw):
    """CPU baseline for path tracing: the oracle's SamplerIntegrator::render on a bounded sample (16 of the 64 spp of C2)
    with all host threads, and a parity check of that sample range against the GPU."""
    from oracle import oracle_path as OP
    cam = scenes.C2_CAMERA
    pk = dict(scenes.C2_PATH)
    ref = OP.Scene(sc, 4)
    fd = OP.film_desc(cam["res"])
    pd = OP.path_desc(sample_begin=0, sample_end=PATH_CPU_SPP, **pk)
    xyzw, dt = ref.render(cam, fd, pd, mode=1)
    n = cam["res"][0] * cam["res"][1] * PATH_CPU_SPP
    cores, model = host_info()
    return {"value": n / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port", "cpu_model": model,
            "sample": f"sample indices [0,{PATH_CPU_SPP}) of the 64 spp of every pixel ({n:,} camera samples), render time only"}, xyzw


def bench_path_c4_cpu(scenes, keep, spp_cpu=2):
    """C4 on the CPU: the oracle renders sample indices [0, spp_cpu) of every pixel (all host threads) = cpu_baseline; the GPU
    film of the same range must equal it bit for bit = parity; a counted render of sample index 0 feeds the roofline."""
    from oracle import oracle_path as OP
    accel, camera, integ, film, sc = keep
    cam = scenes.C4_CAMERA
    pk = dict(scenes.C4_PATH)
    ref = OP.Scene(sc, 4)
    fd = OP.film_desc(cam["res"])
    ref_xyzw, dt = ref.render(cam, fd, OP.path_desc(sample_begin=0, sample_end=spp_cpu, **pk), mode=1)
    film.clear()
    integ.render(film, 0, spp_cpu)
    g_xyzw = film.read_xyzw()
    n = cam["res"][0] * cam["res"][1] * spp_cpu
    cores, model = host_info()
    cpu = {"value": n / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port", "cpu_model": model,
           "sample": f"sample indices [0,{spp_cpu}) of the 256 spp of every pixel ({n:,} camera samples), render time only"}
    parity = {"pixels_checked": int(g_xyzw.shape[0] * g_xyzw.shape[1]),
              "pixels_differing": int((g_xyzw.view(np.uint32) != ref_xyzw.view(np.uint32)).any(axis=2).sum()),
              "checked_against": f"oracle SamplerIntegrator::render, per-(pixel,sample) sampler streams, samples [0,{spp_cpu}), bitwise"}
    _, _, cnt = ref.render_counted(cam, fd, OP.path_desc(sample_begin=0, sample_end=1, **pk), mode=1)
    return cpu, parity, cnt


def bench_path_c5(pb2, scenes, torch, args, dist, rank, world):
    """BASELINE config 4 (C5): the C4 scene (matte / plastic / glass spheres, area + point light, maxdepth 8, power light
    distribution) at 3840x2160 @ 1024 spp, the sample indices of every pixel split across the GPUs (partition_samples), the
    per-GPU films summed with one ncclReduce to rank 0.  A step is the whole frame (8.49 G camera samples)."""
    sc = scenes.scene_c4()
    cam = scenes.C5_CAMERA
    pk = dict(scenes.C5_PATH)
    spp = pk["spp"]
    accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4)
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    integ = pb2.PathIntegrator(accel, camera, **pk)
    film = pb2.Film(cam["res"])
    stream = torch.cuda.current_stream().cuda_stream
    if world > 1:
        uid = [pb2.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        pb2.nccl_init(uid[0], rank, world)
    steps = 2 if world > 1 else 1               # N=1: one 14 s frame is the anchor of the strong-scaling curve
    s0, s1 = pb2.partition_samples(spp, rank, world)

    def frame(ev=None, end=None):
        film.clear()
        if ev:
            ev[0].record()
        integ.render(film, s0, s1 if end is None else end, stream=stream)
        if ev:
            ev[1].record()
        if world > 1:
            film.reduce(0, stream=stream)
        if ev:
            ev[2].record()

    for _ in range(3 if world > 1 else 1):
        frame(end=min(s1, s0 + 8))              # warm-up: 8 sample indices per GPU (same kernels, same buffers)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for ev in evs:
        frame(ev)
    torch.cuda.synchronize()
    render_ms = float(sum(e[0].elapsed_time(e[1]) for e in evs))
    reduce_ms = float(sum(e[1].elapsed_time(e[2]) for e in evs))
    total_ms = float(sum(e[0].elapsed_time(e[2]) for e in evs))
    if world > 1:
        t = torch.tensor([render_ms, reduce_ms, total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        render_ms, reduce_ms, total_ms = (float(v) for v in t)
    # the reduce alone: ranks aligned first, so the wait for the slowest renderer (which the per-frame figure includes) is excluded
    reduce_alone_ms = None
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        film.reduce(0, stream=stream)
        a.record()
        for _ in range(4):
            film.reduce(0, stream=stream)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 4.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        reduce_alone_ms = float(t[0])
    n_samples = cam["res"][0] * cam["res"][1] * spp
    out = {"workload": f"C5: {len(sc['idx'])}-triangle matte/plastic/glass room, maxdepth 8, 3840x2160 @ {spp} spp, sample indices split over "
                       f"{world} GPUs ({s1 - s0} per GPU), one ncclReduce of the float4 film (132.7 MB) to rank 0",
           "unit": "Msamples/s", "scaling": "strong", "value": n_samples * steps / (total_ms * 1e-3) / 1e6, "ms_per_frame": total_ms / steps,
           "samples_per_frame": n_samples, "frames_timed": steps,
           "render_ms": render_ms / steps, "film_reduce_ms": reduce_ms / steps, "film_reduce_frac_of_frame": reduce_ms / total_ms,
           "film_reduce_note": "film_reduce_ms runs from this rank's last render kernel to the end of ncclReduce, max over ranks: it includes the wait "
                               "for the slowest renderer; film_reduce_alone_ms is the collective with the ranks aligned",
           "film_reduce_alone_ms": reduce_alone_ms,
           "film_reduce_alone_frac_of_frame": None if reduce_alone_ms is None else reduce_alone_ms / (total_ms / steps)}
    # parity of the split + ncclReduce: sample indices [0, 2N) of the 1024 split over the N ranks and reduced to
    # rank 0, against the same N ranges rendered one after the other into one film by rank 0 alone.  Weights are sums of
    # ones (exact in any order); XYZ sums differ only by the association of N float adds.
    if world > 1:
        p0, p1 = pb2.partition_samples(2 * world, rank, world)
        film.clear()
        integ.render(film, p0, p1, stream=stream)
        film.reduce(0, stream=stream)
        torch.cuda.synchronize()
        if rank == 0:
            got = film.read_xyzw()
            film.clear()
            for r in range(world):
                a, b = pb2.partition_samples(2 * world, r, world)
                integ.render(film, a, b, stream=stream)
            torch.cuda.synchronize()
            want = film.read_xyzw()
            den = np.maximum(np.abs(want[..., :3]), 1e-6)
            rel = float((np.abs(got[..., :3] - want[..., :3]) / den).max())
            out["parity"] = {"pixels_checked": int(want.shape[0] * want.shape[1]), "max_rel_diff": rel,
                             "weights_equal": bool(np.array_equal(got[..., 3], want[..., 3])),
                             "bitwise_equal_pixels": int((got.view(np.uint32) == want.view(np.uint32)).all(axis=2).sum()),
                             "ok": bool(rel <= 1e-6 and np.array_equal(got[..., 3], want[..., 3])),
                             "checked_against": f"rank 0 rendering the same {world} sample ranges of [0,{2 * world}) alone into one film; tolerance 1e-6 "
                                                "relative on X, Y, Z (float association of the reduce), weights exact"}
        dist.barrier()
        pb2.nccl_shutdown()
    # N=1 anchor of the strong-scaling curve (measured by this bench at --gpus 1 and committed; the driver computes its own ratios)
    try:
        n1 = json.load(open(os.path.join(ROOT, "profiles", "c5_n1_anchor.json")))
        out["n1_anchor"] = n1
        if world > 1:
            out["efficiency_vs_n1"] = out["value"] / (world * n1["msamples_s"])
    except (OSError, KeyError, ValueError):
        pass
    return out


def bench_path_extras(pb2, scenes, torch, rank, no_cpu):
    """The two shapes / integrators added after the BASELINE configs, so that they have measured numbers
    too: the Cornell room with ANALYTIC spheres (shapes/sphere.rs: EFloat quadratic in k_extend_spheres / k_shadow_spheres, a
    spherical area light) under the wavefront PathIntegrator, and the fog + smoke scene under the VolPathIntegrator
    (integrators/volpath.rs, media/homogeneous.rs: the wavefront stages of wavefront_volpath.cu).  Each with a bit-for-bit parity
    check of sample index 0 of every pixel against the oracle."""
    out = {}
    stream = torch.cuda.current_stream().cuda_stream
    for key, sc, kw, res, spp in (("spheres", scenes.scene_spheres(), dict(max_depth=5, rr_threshold=1.0, light_strategy="power"), (1024, 1024), 16),
                                  ("volpath", scenes.scene_media(), dict(max_depth=8, rr_threshold=1.0, light_strategy="power", integrator="volpath"), (512, 512), 16)):
        cam = dict(scenes.C2_CAMERA, res=res)
        accel = pb2.BVHAccel(pb2.scene_from_dict(sc), max_prims_in_node=4)
        camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
        integ = pb2.PathIntegrator(accel, camera, spp=spp, **kw)
        film = pb2.Film(cam["res"])
        integ.render(film, 0, 2, stream=stream)
        torch.cuda.synchronize()
        c0 = integ.counters()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        film.clear()
        a.record()
        integ.render(film, stream=stream)
        b.record()
        torch.cuda.synchronize()
        c1 = integ.counters()
        ms = a.elapsed_time(b)
        n = res[0] * res[1] * spp
        rays = {k: int(c1[k] - c0[k]) for k in ("extend_rays", "shadow_rays", "mis_rays")}
        e = {"workload": {"spheres": f"Cornell room, {len(sc['spheres'])} analytic spheres (matte / plastic / glass, ellipsoid, partial sphere, spherical area light), "
                                     f"PathIntegrator maxdepth 5, {res[0]}x{res[1]} @ {spp} spp",
                          "volpath": f"fog-filled Cornell room + smoke box behind a material-less interface + glass sphere, VolPathIntegrator maxdepth 8, "
                                     f"{res[0]}x{res[1]} @ {spp} spp"}[key],
             "unit": "Msamples/s", "value": n / (ms * 1e-3) / 1e6, "ms_per_frame": ms, "rays_per_frame": rays,
             "mrays_per_s": sum(rays.values()) / (ms * 1e-3) / 1e6}
        if rank == 0 and not no_cpu:
            from oracle import oracle_path as OP
            okw = dict(kw)
            ref = OP.Scene(sc, 4)
            want, dt = ref.render(cam, OP.film_desc(cam["res"]), OP.path_desc(spp=spp, sample_begin=0, sample_end=1, **okw), mode=1)
            film.clear()
            integ.render(film, 0, 1)
            got = film.read_xyzw()
            cores, model = host_info()
            e["cpu_baseline"] = {"value": res[0] * res[1] / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port", "cpu_model": model,
                                 "sample": f"sample index 0 of every pixel ({res[0] * res[1]:,} camera samples), render time only"}
            e["parity"] = {"pixels_checked": int(got.shape[0] * got.shape[1]),
                           "pixels_differing": int((got.view(np.uint32) != want.view(np.uint32)).any(axis=2).sum()),
                           "checked_against": "oracle SamplerIntegrator::render, sample index 0, bitwise"}
        out[key] = e
    return out


def run_reference(args):
    """--impl reference: the CPU restatement of the reference hot path (oracle/; the Rust crate cannot be built —
    no rustc/cargo in the image), all host threads, on a bounded sample of the C3 pass."""
    rank, _, world = rank_info()
    if rank != 0:
        return
    scenes, orc = ge.load_scenes(), ge.load_oracle()
    verts, idx = scenes.scene_c3(args.grid_n)
    cam = camera_for_rank(scenes, 0, args.res)
    ref = orc.BVHAccel(verts, idx, 4)
    rays = orc.camera_primary_rays(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    hits, b0, _ = ref.intersect(rays, want_b0=True)
    srays = orc.spawn_shadow_rays(ref, hits, b0, scenes.C3_POINT_LIGHT)
    brays = orc.spawn_bounce_rays(ref, rays, hits, b0)
    k = max(1, args.cpu_stride)
    sample = [np.ascontiguousarray(rays[::k]), np.ascontiguousarray(srays[::k]), np.ascontiguousarray(brays[::k])]
    n_sample = sum(len(s) for s in sample)
    cores, model = host_info()

    def step():
        return ref.intersect(sample[0])[-1] + ref.intersect_p(sample[1])[-1] + ref.intersect(sample[2])[-1]

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    total = float(sum(times))
    value = n_sample * args.steps / total / 1e6
    path_cpu, _ = bench_path_cpu(orc, scenes, scenes.scene_c2(), None)
    line = {
        "impl": "reference", "metric": "closest-hit + any-hit traversal throughput (C3 pass)", "value": value, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(len(idx), args.res),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "cpu_model": model,
                         "sample": f"every {k}-th ray of each of the 3 ray sets ({n_sample} rays per step), traversal time only"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "path": {"workload": "C2: Cornell box, PathIntegrator maxdepth=5, 512x512, RandomSampler, box filter", "unit": "Msamples/s",
                 "value": path_cpu["value"], "cpu_baseline": path_cpu},
        "note": "CPU restatement of pbrt-rs BVHAccel/Triangle/PathIntegrator (oracle/); the Rust reference itself is not compilable here",
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    rank, local_rank, world = rank_info()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pb2, scenes = ge.load_package(), ge.load_scenes()
    pb2.init(local_rank)                    # raises if the extension or the GPU is missing: no fallback
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream

    t0 = time.time()
    verts, idx = scenes.scene_c3(args.grid_n)
    t_gen = time.time() - t0
    t0 = time.time()
    accel = pb2.BVHAccel(verts, idx, max_prims_in_node=4)
    t_build = time.time() - t0
    n_nodes, n_prims, depth = accel.info()
    cam = camera_for_rank(scenes, rank, args.res)
    camera = pb2.PerspectiveCamera(cam["pos"], cam["look"], cam["up"], cam["fov"], cam["res"])
    n = args.res * args.res

    def dbuf(nbytes):
        return torch.empty(nbytes, dtype=torch.uint8, device=dev)

    d_rays, d_hits, d_b0 = dbuf(n * 32), dbuf(n * 16), dbuf(n * 4)
    d_srays, d_brays, d_occ, d_bhits = dbuf(n * 32), dbuf(n * 32), dbuf(n), dbuf(n * 16)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2

    names = ["raygen", "closest_primary", "spawn", "any_shadow", "closest_bounce"]

    side = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()

    def device_step(events=None, overlap=True):
        # overlap: the shadow batch and the bounce batch do not depend on each other, so they go out on two streams (the
        # `stream` argument of the C ABI) and the tail of each persistent launch is filled by the other; overlap=False runs
        # the five launches one after the other, which is what the per-kernel times (kernel_ms, roofline) are taken from.
        def mark(i):
            if events is not None:
                events[i].record()
        mark(0)
        camera.primary_rays_device(d_rays.data_ptr(), stream)
        mark(1)
        accel.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), d_b0.data_ptr(), stream)
        mark(2)
        accel.spawn_shadow_bounce_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n, scenes.C3_POINT_LIGHT, d_srays.data_ptr(), d_brays.data_ptr(), stream)
        mark(3)
        if overlap:
            fork, join = torch.cuda.Event(), torch.cuda.Event()
            fork.record(main_stream)
            side.wait_event(fork)
            accel.intersect_p_device(d_srays.data_ptr(), n, d_occ.data_ptr(), side.cuda_stream)
            accel.intersect_device(d_brays.data_ptr(), n, d_bhits.data_ptr(), None, stream)
            join.record(side)
            main_stream.wait_event(join)
        else:
            accel.intersect_p_device(d_srays.data_ptr(), n, d_occ.data_ptr(), stream)
            mark(4)
            accel.intersect_device(d_brays.data_ptr(), n, d_bhits.data_ptr(), None, stream)
        mark(5)
    launches_per_step = 5

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step()
        flush.zero_()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    all_events = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        device_step(all_events[s])
        flush.zero_()                       # L2 flush between timed iterations (outside the event pairs)
    barrier()
    step_ms = [ev[0].elapsed_time(ev[5]) for ev in all_events]
    total_ms = float(sum(step_ms))
    # per-kernel times: the same steps once more with the five launches strictly one after the other (not part of `value`)
    serial_events = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(args.steps)]
    for s in range(args.steps):
        device_step(serial_events[s], overlap=False)
        flush.zero_()
    barrier()
    kernel_ms = {names[i]: float(np.mean([ev[i].elapsed_time(ev[i + 1]) for ev in serial_events])) for i in range(5)}
    serial_step_ms = float(np.mean([ev[0].elapsed_time(ev[5]) for ev in serial_events]))

    # ---- e2e: host buffers through the C ABI (pinned), copies inside the timed region ----
    h_rays = torch.empty(n * 8, dtype=torch.float32).pin_memory()
    h_srays, h_brays = torch.empty_like(h_rays).pin_memory(), torch.empty_like(h_rays).pin_memory()
    h_hits, h_bhits = torch.empty(n * 4, dtype=torch.int32).pin_memory(), torch.empty(n * 4, dtype=torch.int32).pin_memory()
    h_occ = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_rays.copy_(d_rays.view(torch.float32))
    h_srays.copy_(d_srays.view(torch.float32))
    h_brays.copy_(d_brays.view(torch.float32))
    torch.cuda.synchronize()
    L = pb2.lib()

    def e2e_step_sync():
        # the synchronous entry points: every call returns with its results in the host buffer (one ring drain per batch)
        pb2.check(L.pb2_intersect(accel.h, h_rays.data_ptr(), n, h_hits.data_ptr(), None))
        pb2.check(L.pb2_intersect_p(accel.h, h_srays.data_ptr(), n, h_occ.data_ptr()))
        pb2.check(L.pb2_intersect(accel.h, h_brays.data_ptr(), n, h_bhits.data_ptr(), None))

    # a second set of result buffers: step k + 1 is enqueued before step k's results are waited for, as a renderer that
    # produces ray batches continuously would (every step's H2D and D2H copies are still inside the timed region)
    h_out2 = (torch.empty(n * 4, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory(),
              torch.empty(n * 4, dtype=torch.int32).pin_memory())
    e2e_count = [0]

    def e2e_step(last=False):
        # the three host batches enqueued back to back on the scene's ring; the wait leaves this step's three batches in flight
        # and returns when the previous step's results are in the host buffers, so the ring never drains between steps
        hh, ho, hb = (h_hits, h_occ, h_bhits) if e2e_count[0] % 2 == 0 else h_out2
        e2e_count[0] += 1
        pb2.check(L.pb2_intersect_async(accel.h, h_rays.data_ptr(), n, hh.data_ptr(), None))
        pb2.check(L.pb2_intersect_p_async(accel.h, h_srays.data_ptr(), n, ho.data_ptr()))
        pb2.check(L.pb2_intersect_async(accel.h, h_brays.data_ptr(), n, hb.data_ptr(), None))
        pb2.check(L.pb2_scene_wait_until(accel.h, 0 if last else 3))

    for _ in range(args.warmup):
        e2e_step_sync()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step_sync()
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    for i in range(args.warmup):
        e2e_step(last=i == args.warmup - 1)
    # what the link itself delivers: one plain pinned H2D / D2H copy of a ray-set-sized buffer (explains e2e vs value)
    cp = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    cp[0].record()
    d_rays.view(torch.float32).copy_(h_rays, non_blocking=True)
    cp[1].record()
    h_rays.copy_(d_rays.view(torch.float32), non_blocking=True)
    cp[2].record()
    torch.cuda.synchronize()
    pcie = {"h2d_gbs": n * 32 / (cp[0].elapsed_time(cp[1]) * 1e-3) / 1e9, "d2h_gbs": n * 32 / (cp[1].elapsed_time(cp[2]) * 1e-3) / 1e9}
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(last=i == args.steps - 1)                # (the last step waits for everything)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if e2e_count[0] % 2 == 0:                              # the last step wrote the second set: compare that one
        h_hits, h_occ, h_bhits = h_out2
    clock_rec = clocks.stop()
    import zlib
    hits_crc = "%08x" % zlib.crc32(d_occ.cpu().numpy().tobytes(), zlib.crc32(d_bhits.cpu().numpy().tobytes(), zlib.crc32(d_hits.cpu().numpy().tobytes())))
    e2e_parity = bool(np.array_equal(h_hits.numpy(), d_hits.view(torch.int32).cpu().numpy())
                      and np.array_equal(h_occ.numpy(), d_occ.cpu().numpy())
                      and np.array_equal(h_bhits.numpy(), d_bhits.view(torch.int32).cpu().numpy()))

    # ---- path tracing (second half of the BASELINE metric) ----
    dist_mod = dist if world > 1 else None
    path_c2 = path_c4 = path_c5 = path_extras = None
    path_launches = 0
    if not args.no_path:
        path_c2, path_keep = bench_path_c2(pb2, scenes, torch, args, dist_mod, world)
        path_c4, path_c4_keep = bench_path_c4(pb2, scenes, torch, args, dist_mod, world)
        path_c5 = bench_path_c5(pb2, scenes, torch, args, dist_mod, rank, world)
        path_launches = int(path_c2["kernel_launches_per_frame"] * path_steps(args)[0] + path_c4["kernel_launches_per_step"] * 2)
        path_extras = bench_path_extras(pb2, scenes, torch, rank, args.no_cpu_baseline) if rank == 0 else None

    # ---- BVH build: SplitMethod::HLBVH built on the GPU next to the host SAH build, and what the C3
    # ray sets cost on that tree (same rays, same hits: tests/ compare ids and t bits)
    bvh_build = None
    if rank == 0:
        t0 = time.time()
        accel_h = pb2.BVHAccel(verts, idx, max_prims_in_node=4, split_method=1)
        t_hl = time.time() - t0
        stages = accel_h.build_stats()

        def hl_step():
            accel_h.intersect_device(d_rays.data_ptr(), n, d_hits.data_ptr(), d_b0.data_ptr(), stream)
            accel_h.intersect_p_device(d_srays.data_ptr(), n, d_occ.data_ptr(), stream)
            accel_h.intersect_device(d_brays.data_ptr(), n, d_bhits.data_ptr(), None, stream)
        prim_before = d_hits.view(torch.int32)[::4].clone()
        hl_step()
        flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        hl_step()
        ev[1].record()
        torch.cuda.synchronize()
        same = bool(torch.equal(prim_before, d_hits.view(torch.int32)[::4]))
        hn, _, hdepth = accel_h.info()
        bvh_build = {"sah_host_build_upload_s": t_build, "hlbvh_gpu_build_upload_s": t_hl,
                     "hlbvh_stage_ms": dict(zip(["upload_bounds_morton", "sort", "treelets", "upper_sah_host", "flatten", "device_repack"],
                                                [round(x, 3) for x in stages])),
                     "hlbvh_nodes": hn, "hlbvh_depth": hdepth, "hlbvh_c3_pass_mrays_s": 3 * n / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e6,
                     "hlbvh_primary_hit_ids_equal_sah": same}
        del accel_h

    # ---- max over ranks ----
    if world > 1:
        t = torch.tensor([total_ms, e2e_s, e2e_sync_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, e2e_sync_s = float(t[0]), float(t[1]), float(t[2])
    rays_per_step = 3 * n
    value = world * rays_per_step * args.steps / (total_ms * 1e-3) / 1e6
    e2e_value = world * rays_per_step * args.steps / e2e_s / 1e6

    # ---- roofline + cpu baseline (rank 0, after the GPU timing so host threads do not perturb it) ----
    peak, peak_src = load_peak()
    roofline, cpu_baseline, parity = None, None, None
    if rank == 0 and not args.no_cpu_baseline:
        orc = ge.load_oracle()
        ref = orc.BVHAccel(verts, idx, 4)
        g_rays = d_rays.view(torch.float32).cpu().numpy().reshape(-1, 8)
        g_srays = d_srays.view(torch.float32).cpu().numpy().reshape(-1, 8)
        g_brays = d_brays.view(torch.float32).cpu().numpy().reshape(-1, 8)
        rh, rb0, c_prim, _ = ref.intersect(g_rays, counters=True, want_b0=True)
        ro, c_shadow, _ = ref.intersect_p(g_srays, counters=True)
        rbh, c_bounce, _ = ref.intersect(g_brays, counters=True)
        g_hits = d_hits.cpu().numpy().view(pb2.HIT_DTYPE)
        g_bhits = d_bhits.cpu().numpy().view(pb2.HIT_DTYPE)
        mism = int((g_hits["prim_id"] != rh["prim_id"]).sum() + (g_hits["t"].view(np.uint32) != rh["t"].view(np.uint32)).sum()
                   + (d_occ.cpu().numpy() != ro).sum() + (g_bhits["prim_id"] != rbh["prim_id"]).sum()
                   + (g_bhits["t"].view(np.uint32) != rbh["t"].view(np.uint32)).sum())
        parity = {"rays_checked": 3 * n, "mismatches": mism, "checked_against": "oracle (CPU restatement), same ray batch",
                  "e2e_equals_device": e2e_parity}
        per_kernel = {}
        for name, cnt, hb in (("closest_primary", c_prim, 16), ("any_shadow", c_shadow, 4), ("closest_bounce", c_bounce, 16)):
            bts = algorithmic_bytes(n, cnt[0], cnt[1], hb)
            per_kernel[name] = {"ms": kernel_ms[name], "algorithmic_bytes": bts, "gbs": bts / (kernel_ms[name] * 1e-3) / 1e9,
                                "nodes_per_ray": float(cnt[0]) / n, "tris_per_ray": float(cnt[1]) / n,
                                "mrays_s": n / (kernel_ms[name] * 1e-3) / 1e6}
        dom = max(("closest_primary", "any_shadow", "closest_bounce"), key=lambda k: kernel_ms[k])
        # DRAM bytes per launch of the same kernels from the committed `ncu --set full` capture (profiles/traffic.json,
        # written by tools/ncu_traffic.py from the .ncu-rep; dram__bytes_read.sum + dram__bytes_write.sum)
        traffic, traffic_src, lanes = None, None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic, traffic_src = tj["launches"].get(dom), tj.get("source")
            lanes = tj.get("lanes_per_inst", {}).get(dom)
            for name in per_kernel:
                if tj["launches"].get(name):
                    per_kernel[name]["dram_bytes"] = tj["launches"][name]
                    per_kernel[name]["dram_frac"] = tj["launches"][name] / (kernel_ms[name] * 1e-3) / 1e9 / peak
                    per_kernel[name]["lanes_per_inst"] = tj.get("lanes_per_inst", {}).get(name)
        except (OSError, KeyError, ValueError):
            pass
        roofline = {"bound": "hbm", "kernel": f"k_closest_hit/k_any_hit [{dom}]", "achieved": per_kernel[dom]["gbs"], "peak": peak,
                    "unit": "GB/s", "frac": per_kernel[dom]["gbs"] / peak, "traffic": traffic, "traffic_source": traffic_src,
                    # what the DRAM counters saw (ncu capture of the same launch) against the same peak, and the mean number of
                    # active lanes per executed warp instruction: `frac` above follows SURVEY §8d's byte definition, which counts
                    # every node / triangle visit as HBM traffic although the upper tree is served by L1 / L2
                    "dram_frac": None if traffic is None else traffic / (kernel_ms[dom] * 1e-3) / 1e9 / peak,
                    "lanes_per_inst": lanes,
                    "peak_source": peak_src, "per_kernel": per_kernel}
        k = max(1, args.cpu_stride)
        sample = [np.ascontiguousarray(g_rays[::k]), np.ascontiguousarray(g_srays[::k]), np.ascontiguousarray(g_brays[::k])]
        n_sample = sum(len(s) for s in sample)
        ref.intersect(sample[0])
        cpu_s, cpu_reps = 0.0, 0
        while cpu_s < args.cpu_seconds and cpu_reps < 200:          # ~10 s of CPU work on the same rays
            cpu_s += ref.intersect(sample[0])[-1] + ref.intersect_p(sample[1])[-1] + ref.intersect(sample[2])[-1]
            cpu_reps += 1
        cores, model = host_info()
        cpu_baseline = {"value": cpu_reps * n_sample / cpu_s / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port", "cpu_model": model,
                        "sample": f"every {k}-th ray of each of the 3 ray sets ({n_sample} rays) x {cpu_reps} passes = {cpu_s:.1f} s, "
                                  "traversal time only"}
    if rank == 0 and not args.no_cpu_baseline and not args.no_path:
        orc = ge.load_oracle()
        # path tracing: oracle on sample indices [0,2) of C2, and the GPU film of the same range must equal it bit for bit
        accel2, camera2, integ2, film2, sc2 = path_keep
        path_cpu, ref_xyzw = bench_path_cpu(orc, scenes, sc2, None)
        film2.clear()
        integ2.render(film2, 0, PATH_CPU_SPP)
        g_xyzw = film2.read_xyzw()
        path_c2["cpu_baseline"] = path_cpu
        path_c2["parity"] = {"pixels_checked": int(g_xyzw.shape[0] * g_xyzw.shape[1]),
                             "pixels_differing": int((g_xyzw.view(np.uint32) != ref_xyzw.view(np.uint32)).any(axis=2).sum()),
                             "checked_against": f"oracle SamplerIntegrator::render, per-(pixel,sample) sampler streams, samples [0,{PATH_CPU_SPP})"}
        from oracle import oracle_path as OP
        _, _, cnt2 = OP.Scene(sc2, 4).render_counted(scenes.C2_CAMERA, OP.film_desc(scenes.C2_CAMERA["res"]),
                                                    OP.path_desc(sample_begin=0, sample_end=4, **scenes.C2_PATH), mode=1)
        path_c2["roofline"] = path_roofline(path_c2["rays_per_frame"], path_c2["samples_per_frame"], path_c2["ms_per_frame"], cnt2, peak, peak_src, "c2")
        c4_cpu, c4_parity, cnt4 = bench_path_c4_cpu(scenes, path_c4_keep)
        path_c4["cpu_baseline"], path_c4["parity"] = c4_cpu, c4_parity
        path_c4["roofline"] = path_roofline(path_c4["rays_per_step"], path_c4["samples_per_step"], path_c4["ms_per_step"], cnt4, peak, peak_src, "c4")
        if path_c5 is not None:
            # same scene, same integrator settings, 4x the pixels: per-ray node / triangle counts of the C4 sample apply
            path_c5["roofline"] = path_roofline({k: path_c4["rays_per_step"][k] * (path_c5["samples_per_frame"] / path_c4["samples_per_step"])
                                                 for k in path_c4["rays_per_step"]}, path_c5["samples_per_frame"], path_c5["ms_per_frame"], cnt4,
                                                peak * world, peak_src + f" x {world} GPUs", "c4")
            path_c5["roofline"]["note_c5"] = "ray counts scaled from the C4 frame of the same run (same scene and integrator, 4x the pixels, 4x the spp)"
            path_c5["cpu_baseline"] = dict(c4_cpu, note="the C4 sample: same scene and settings; the oracle's rate per camera sample does not depend on resolution")
    if rank == 0:
        line = {
            "metric": "closest-hit + any-hit traversal throughput (C3 pass)", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(n_prims, args.res),
            "scene": {"bvh_nodes": n_nodes, "bvh_depth": depth, "scene_gen_s": t_gen, "bvh_build_upload_s": t_build,
                      "l2": "BVH+triangles (~1 GB) exceed the 126 MB L2 and a 512 MB buffer is rewritten between timed steps"},
            "kernel_ms": kernel_ms, "kernel_ms_note": "each launch alone on the device (the same steps repeated with the five launches in series, "
                                                      "outside the timed region); in the timed steps the shadow and bounce batches share the device on two streams",
            "serial_ms_per_step": serial_step_ms, "hits_crc32": hits_crc,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 3 * n * 32, "d2h_bytes_per_step": n * 16 + n + n * 16,
                    "api": "pb2_intersect_async x2 + pb2_intersect_p_async per step, pb2_scene_wait_until(3): one step stays in flight while the previous one is retired; pinned host buffers",
                    "synchronous_calls_mrays_s": world * rays_per_step * args.steps / e2e_sync_s / 1e6,
                    "synchronous_api": "pb2_intersect / pb2_intersect_p, each returning with its results on the host",
                    "pcie_measured": pcie,
                    "link_bound_mrays_s": world * rays_per_step / (3 * n * 32 / (pcie["h2d_gbs"] * 1e9)) / 1e6},
            "gpu_launches": launches_per_step * args.steps + path_launches,
            "clocks": clock_rec, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
            "path": path_c2, "path_c4": path_c4, "path_multi_gpu": path_c5, "path_extras": path_extras, "bvh_build": bvh_build,
        }
        print(json.dumps(line), flush=True)
        if args.write_c5_anchor and world == 1 and path_c5:
            json.dump({"msamples_s": path_c5["value"], "ms_per_frame": path_c5["ms_per_frame"], "samples_per_frame": path_c5["samples_per_frame"],
                       "measured_by": "python bench.py --gpus 1 --write-c5-anchor (one B200, whole 3840x2160 @ 1024 spp frame)"},
                      open(os.path.join(ROOT, "profiles", "c5_n1_anchor.json"), "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
