#!/usr/bin/env python
"""bench.py -- throughput of the render() hot path on N B200s (one process per GPU).

    python bench.py --gpus N --steps K --warmup W [--workload c1|c2|c3|c4|c5] [--scale F]
    python bench.py --impl reference ...        # the reference's CPU algorithm on host cores

A "step" is one pass of the hot path over one frame of synthetic input: every rank renders
`spp` samples per pixel of the workload's frame (its own disjoint range of global sample
indices -- weak scaling: per-GPU work is fixed), the per-GPU float sums are reduced to rank 0
by ONE NCCL reduce, and rank 0 runs the gamma/quantise kernel.  `value` is whole-job
Mrays/s with the scene and BVH already resident in HBM (ray = one trace_path invocation, the
reference's ray_count, raytracer.c:484).  `e2e` is the same metric through the
reference-facing call with HOST buffers: scene upload + BVH build + render + framebuffer
read-back inside the timed region, every step.

Workloads (BASELINE.json configs / SURVEY.md 8d):
  c1  reference main.c default scene, 320x180, 50 spp, depth 5
  c2  10k random spheres + walls, 1920x1080, 256 spp, depth 8
  c3  ~1M-triangle height-field mesh room, 1920x1080, 128 spp, depth 5   (default: the
      configuration the target ">= 1 Grays/s on 1 B200 for the 1080p mesh scene" is quoted on)
  c4  dielectric/metal-heavy 10k spheres, 3840x2160, 512 spp per GPU, depth 8
  c5  all-dielectric deep-bounce stress, 512x512, 64 spp, depth 64
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

WORKLOADS = {
    #      W     H    spp depth  description
    "c1": (320, 180, 50, 5, "reference main.c default scene (38 spheres), 320x180, 50 spp, depth 5"),
    "c2": (1920, 1080, 256, 8, "10k random spheres + 6 walls + 2 lights, 1920x1080, 256 spp, depth 8"),
    "c3": (1920, 1080, 128, 5, "1,002,528-triangle height-field mesh + 12 spheres, 1920x1080, 128 spp, depth 5"),
    "c4": (3840, 2160, 512, 8, "10k spheres 40% dielectric / 40% mirror, 3840x2160, 512 spp per GPU, depth 8"),
    "c5": (512, 512, 64, 64, "2k all-dielectric spheres deep-bounce stress, 512x512, 64 spp, depth 64"),
}
SEED = 1666943821  # main.c:182


def build_host_scene(api, name, W, H, pinned=False):
    """-> (source for api.Scene, n_bvh_levels L, bytes per primitive S, flops per test P, h2d bytes)"""
    if name == "c1":
        objs = api.scene_default(W, H)
        return objs, None, 16, 20, objs.nbytes
    if name in ("c2", "c4", "c5"):
        mix = {"c2": (0.5, 0.2, 0.2), "c4": (0.2, 0.4, 0.4), "c5": (0.1, 0.9, 0.0)}[name]
        count = 2000 if name == "c5" else 10000
        objs = api.scene_sphere_field(count, W, H, mix=mix, seed=SEED)
        return objs, len(objs), 16, 20, objs.nbytes
    if name == "c3":
        verts = api.heightfield_mesh(708, 20 * W / H * 0.98)
        if pinned:
            import torch
            buf = torch.empty(verts.nbytes, dtype=torch.uint8).pin_memory()
            pinned_np = buf.numpy().view(verts.dtype)
            pinned_np[:] = verts
            holder = api.mesh_room(pinned_np, W, H)
            holder._keep.append(buf)
        else:
            holder = api.mesh_room(verts, W, H)
        return holder, len(verts) // 3 + 12, 48, 51, verts.nbytes + 12 * 96
    raise SystemExit(f"unknown workload {name}")


def workload_config(name, spp, world):
    W, H, _, depth, text = WORKLOADS[name]
    return {"workload": f"{name}: {text}", "width": W, "height": H, "spp_per_gpu": spp, "max_depth": depth, "seed": SEED,
            "l2": "512 MB memset between timed steps (L2 flushed)",
            "parallelism": f"spp-sharded x{world}, one NCCL reduce",
            "kernel": "wavefront (k_wf_generate / k_wf_trace / k_wf_shade), compressed BVH4"}


def algorithmic_cost(name, n_prims, S, P):
    """per INTERSECTED ray.  flops = 2*L*22 + 4*P + 60, bytes = L*64 + 4*S,
    L = ceil(log2(N/4)); C1 is the brute-force case 38*20+60 flop, 38*16 B."""
    if name == "c1":
        return 38 * 20 + 60, 38 * 16
    L = max(1, math.ceil(math.log2(max(n_prims, 8) / 4.0)))
    return 2 * L * 22 + 4 * P + 60, L * 64 + 4 * S


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nme in enumerate(names):
                if len(r) > 4 + k and r[4 + k].lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes (read + write) summed over the k_wf_trace launches of ONE step, from the
    committed ncu capture of the same command (profiles/ncu_traffic.json), if any"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if d.get("kernel") == "k_wf_trace":
            return d.get(workload)
    return None


# ---- CPU arm -------------------------------------------------------------------------------

def cpu_sample(name, W, H, depth):
    """a bounded sample of the same workload for the host cores: same scene, reduced frame"""
    # sized for roughly 10-20 s of host work per pass (the brute-force loop visits every primitive)
    return {"c1": (320, 180, 2000), "c2": (960, 540, 1), "c3": (64, 36, 1), "c4": (960, 540, 1), "c5": (512, 512, 8)}[name]


def run_cpu_step(api, ol, name, source, depth, step, threads):
    """one pass of the reference algorithm over the bounded sample -> (rays, seconds, kind).
    The oracle port (oracle/oracle.c: the reference's brute-force intersect() and trace_path()
    restated, bit-identical to the reference for sphere scenes) with OpenMP over rows and the
    keyed RNG.  The unmodified reference cannot run these workloads: it has no live mesh path
    (c3), its dielectric split is 2^depth (c5) and its OpenMP build serialises on rand()
    (SURVEY.md section 6)."""
    w, h, spp = cpu_sample(name, 0, 0, depth)
    cam = api.init_camera(w, h)
    t0 = time.perf_counter()
    _, (rays, _) = ol.render_sum(source, cam, w, h, spp, rng="philox", dielectric="stochastic", max_depth=depth,
                                 seed=SEED, sample_offset=step * spp, threads=threads)
    return rays, time.perf_counter() - t0, "port", f"{w}x{h}x{spp}spp of the same scene, oracle port, {threads} OpenMP threads"


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pkg = entry.load_package()
    api = pkg.api
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    W, H, spp, depth, text = WORKLOADS[args.workload]
    w, h, s = cpu_sample(args.workload, W, H, depth)
    # the bounded sample keeps the FULL scene (brute force over all primitives, like the reference)
    source, _, _, _, _ = build_host_scene(api, args.workload, W, H)
    threads = os.cpu_count() or 1
    for i in range(args.warmup):
        run_cpu_step(api, ol, args.workload, source, depth, i, threads)
    rays_total, t_total, kind, sample = 0, 0.0, "port", ""
    for i in range(args.steps):
        rays, dt, kind, sample = run_cpu_step(api, ol, args.workload, source, depth, args.warmup + i, threads)
        rays_total += rays
        t_total += dt
    value = rays_total / t_total / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args.workload, spp, args.gpus), sample=sample,
                       note="the reference arm runs the bounded sample named in `sample` of this workload, on the host cores of rank 0"),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---- GPU arm -------------------------------------------------------------------------------

def main_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    pkg = entry.load_package()
    api, abi = pkg.api, pkg.abi
    W, H, spp, depth, text = WORKLOADS[args.workload]
    spp = max(1, int(round(spp * args.scale)))
    source, n_prims, S, P, h2d_bytes = build_host_scene(api, args.workload, W, H, pinned=True)
    cam = api.init_camera(W, H)
    scene = api.Scene(source, device=local)
    info = scene.info

    accum = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    fb = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
    fb_host = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    s_begin, s_end, total_spp = pkg.sharding.shard_weak(rank, world, spp)
    desc = api.make_desc(W, H, s_begin, s_end, max_depth=depth, seed=SEED)

    def step():
        scene.render_accum(cam, desc, accum.data_ptr(), stream=stream.cuda_stream)
        pkg.sharding.reduce_to_root(accum, world)
        if rank == 0:
            api.tonemap(accum.data_ptr(), W, H, total_spp, fb.data_ptr(), device=local, stream=stream.cuda_stream)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # warm-up; one extra pass with counters gives the (deterministic) ray count of a step
    ctr = scene.render_accum(cam, desc, accum.data_ptr(), stream=stream.cuda_stream, want_counters=True)
    rays_rank, hit_rank, kernel_ms_probe = ctr.rays, ctr.rays_intersected, ctr.gpu_ms
    launches_per_step = int(ctr.launches) + (1 if rank == 0 else 0)
    for _ in range(max(3, args.warmup)):
        step()
    sync_all()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    sync_all()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # evict L2 between timed steps (not timed)
        if world > 1:
            dist.barrier()
        ev[i][0].record(stream)
        kev[i][0].record(stream)
        scene.render_accum(cam, desc, accum.data_ptr(), stream=stream.cuda_stream)
        kev[i][1].record(stream)
        if world > 1:
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            api.tonemap(accum.data_ptr(), W, H, world * spp, fb.data_ptr(), device=local, stream=stream.cuda_stream)
        ev[i][1].record(stream)
    sync_all()
    clocks = sampler.stop() if sampler else None
    ms_local = sum(a.elapsed_time(b) for a, b in ev)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps

    t = torch.tensor([ms_local, float(rays_rank), float(hit_rank)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, rays_step, hits_step = tmax[0].item(), tsum[1].item(), tsum[2].item()
    else:
        ms_total, rays_step, hits_step = t[0].item(), t[1].item(), t[2].item()
    ms_per_step = ms_total / args.steps
    value = rays_step / ms_per_step / 1e3  # Mrays/s, whole job
    paths_per_s = world * W * H * spp / ms_per_step * 1e3

    # ---- e2e: the reference-facing call with HOST buffers, every step ------------------------
    e2e_steps = max(1, min(args.steps, 3))
    opt = abi.Options()
    opt.width, opt.height, opt.samples = W, H, spp
    rp = abi.RenderParams()
    _, host = pkg.load()
    host.render_params_default(rp)
    rp.max_depth, rp.seed, rp.device = depth, SEED, local
    rp.sample_offset = rank * spp
    import ctypes as C

    def e2e_step():
        if world == 1:
            # the drop-in: render()/render_scene() of the C99 host library (raytracer.h:156)
            if isinstance(source, abi.SceneHolder):
                host.render_scene(fb_host.data_ptr(), C.addressof(source.objects), source.n, C.byref(cam), C.byref(opt), C.byref(rp))
            else:
                host.render_ex(fb_host.data_ptr(), source.ctypes.data, len(source), C.byref(cam), C.byref(opt), C.byref(rp))
        else:
            sc = api.Scene(source, device=local)  # upload + marshal + BVH build
            sc.render_accum(cam, desc, accum.data_ptr(), stream=stream.cuda_stream)
            dist.reduce(accum, dst=0, op=dist.ReduceOp.SUM)
            if rank == 0:
                api.tonemap(accum.data_ptr(), W, H, world * spp, fb.data_ptr(), device=local, stream=stream.cuda_stream)
                fb_host.copy_(fb, non_blocking=True)
            torch.cuda.synchronize(dev)
            sc.close()

    e2e_step()  # warm
    sync_all()
    t0 = time.perf_counter()
    e2e_calls = []
    for _ in range(e2e_steps):
        tc = time.perf_counter()
        e2e_step()
        e2e_calls.append((time.perf_counter() - tc) * 1e3)
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = rays_step / te[0].item() / 1e3

    # ---- roofline of the dominant kernel (k_wf_trace) -----------------------------------------
    # its launches of one step (one per bounce and wave) are timed live with CUDA events on the
    # launching stream (rtb_counters.trace_ms); L2 flushed before each pass like the timed steps
    trace_ms_list, step_ms_list, trace_launches = [], [], 0
    for i in range(min(args.steps, 3)):
        flush.fill_(i & 0xFF)
        c = scene.render_accum(cam, desc, accum.data_ptr(), stream=stream.cuda_stream, want_counters=True)
        trace_ms_list.append(c.trace_ms)
        step_ms_list.append(c.gpu_ms)
        trace_launches = int(c.trace_launches)
    trace_ms = sum(trace_ms_list) / len(trace_ms_list)
    trace_share = trace_ms / (sum(step_ms_list) / len(step_ms_list))
    flops_ray, bytes_ray = algorithmic_cost(args.workload, n_prims or 38, S, P)
    peak, peak_src = measured_peaks()
    achieved_gbs = hit_rank * bytes_ray / (trace_ms * 1e-3) / 1e9
    achieved_tflops = hit_rank * flops_ray / (trace_ms * 1e-3) / 1e12
    l2_peak = api.probe_l2_bandwidth(32 << 20, 50, device=local) if rank == 0 else 0.0  # measured now, GB/s
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12

    line = None
    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as ol
            threads = os.cpu_count() or 1
            # the bounded sample keeps the full scene (brute force over every primitive)
            rays, dt, kind, sample = run_cpu_step(api, ol, args.workload, source, depth, 0, threads)
            cpu_baseline = {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind,
                            "sample": sample, "seconds": dt}
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 geometry / f32 colour", "data": "synthetic",
            "config": workload_config(args.workload, spp, world),
            "paths_per_s": paths_per_s, "rays_per_step": rays_step, "rays_per_path": rays_step / (world * W * H * spp),
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_step": te[0].item(), "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(W * H * 3), "ms_per_call": [round(x, 2) for x in e2e_calls], "api": "render_scene()/render_ex() of libraytracer_b200.so" if world == 1
                    else "rtb_scene_create + rtb_render_accum + NCCL reduce + rtb_tonemap + D2H"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peak, "unit": "GB/s", "frac": achieved_gbs / peak,
                         "traffic": ncu_traffic(args.workload), "peak_source": peak_src, "kernel": "k_wf_trace",
                         "kernel_ms_per_step": trace_ms, "launches_per_step": trace_launches,
                         "avg_launch_ms": trace_ms / max(1, trace_launches), "share_of_step": trace_share,
                         "step_kernels_ms": kernel_ms, "bytes_per_intersected_ray": bytes_ray,
                         "intersected_rays_per_step": hit_rank,
                         "achieved_def": "intersected rays of one step x SURVEY 8(d) bytes per ray / summed k_wf_trace time of the step",
                         "note": "algorithmic bytes are L2/L1-resident BVH+primitive fetches (SURVEY 8d), so frac against HBM can exceed 1; see roofline_l2 for the L2 denominator and traffic for the DRAM bytes ncu measured"},
            "roofline_l2": {"bound": "l2", "achieved": achieved_gbs, "peak": l2_peak, "unit": "GB/s",
                            "frac": achieved_gbs / l2_peak if l2_peak else None,
                            "peak_source": "measured in this run: rtb_probe_l2_bandwidth, 32 MB L2-resident buffer, ld.global.cg.v4 from all SMs",
                            "note": "the denominator SURVEY 8(d) names for the walk's bytes (node and primitive fetches are served by L1/L2)"},
            "roofline_fp32": {"bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                              "frac": achieved_tflops / fp32_peak, "flops_per_intersected_ray": flops_ray,
                              "peak_source": f"148 SM x 128 lanes x 2 x {sm_mhz:.0f} MHz (median clock under load)"},
            "scene": {"objects": int(info.n_objects), "spheres": int(info.n_spheres), "triangles": int(info.n_triangles),
                      "bvh_nodes": int(info.n_bvh_nodes), "bvh_depth": int(info.bvh_depth), "big_prims": int(info.n_big_prims),
                      "device_bytes": int(info.device_bytes), "build_ms": float(info.build_ms)},
            "clocks": clocks, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    scene.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="scale spp (debug only; 1.0 = the named config)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
