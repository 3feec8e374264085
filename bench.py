#!/usr/bin/env python
"""bench.py -- Mpaths/s of the path-tracing hot path on B200, next to the reference's CPU path.

    python bench.py --gpus N --steps K --warmup W [--workload c1|c2|c3|c4|c5]   (N>1: under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W [--workload ...]

A "step" is one full render of the workload; a path = one camera sample (one iteration of render.hpp:95-101).
Workloads are BASELINE.json's configs; the default, c1, is the one the metric is quoted on:

  c1  the reference's default src/main.cpp scene, 800x480, 100 spp (scene + camera captured from the unmodified
      main.cpp: tests/golden/c1_scene.ptsc.gz)
  c2  RTIOW random spheres, 1920x1080, 64 spp            c3  Cornell box with smoke boxes, 1024x1024, 1024 spp
  c4  10 002-triangle pyramid mesh, 1920x1080, 256 spp   c5  4K motion blur + depth of field, 3840x2160, 4096 spp
  (c2..c5 are built by tests/scenes.py with numpy's RandomState, not LocalPseudoRNG: "like", not identical to, the
  scenes SURVEY.md sketches -- the oracle and the GPU get the same vectors either way.)

  value   whole-job Mpaths/s with scene and framebuffer resident in HBM (device-resident C-ABI), timed with CUDA
          events on the launching stream, max over ranks.
  N > 1   the path shards by pixels with no data-path collective (seeds are global linear ids).  `--scaling weak`
          (default): every GPU renders the workload's row count -- N GPUs render the same scene with the same
          camera at height x N, rows interleaved (row r -> rank r mod N), landing in rank 0's framebuffer
          through peer stores over NVLink (or an NCCL gather).  `--scaling strong`: the fixed image over N GPUs.
          Every N > 1 line also carries a `strong` sub-record (the fixed image, same steps).
  e2e     the same metric through the blocking host-buffer entry point (pt_render / the per-rank launcher):
          scene upload from pinned host memory + render + framebuffer download every step.
  roofline  FP32: achieved = value x W, W = algorithmic flop per path from the oracle's work counters of this
          exact workload and the per-test constants of SURVEY.md appendix D; peak = N x the FFMA rate measured in
          this run by a register-resident micro-kernel (MEASURED_PEAKS.json has no fp32 entry).  W is the work of
          the reference's brute-force scan; culling skips most of it, so the line also carries the flops really
          executed (ncu counters, profiles/traffic.json) -- `achieved` is an algorithmic rate, not an issue rate.
  cpu_baseline  the reference's CPU path (oracle/_ref/libptref.so = unmodified reference headers where it has the
          workload's template instantiation, else the C port) on ALL of this box's host cores, on a bounded
          sample of the same workload at full spp.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def host_threads():
    """The host cores this process may use (torchrun exports OMP_NUM_THREADS=1: never ask OpenMP)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# name -> description, builder, default --steps, CPU sample (a function of width, height -> pt_region + words)
def _rows(stride):
    def f(w, h):
        from oracle.pyoracle import rows_region
        return rows_region(w, h, 0, stride), "rows 0::%d of %d" % (stride, h)
    return f


def _tile(tw, th):
    def f(w, h):
        from path_tracer_b200 import abi
        x0, y0 = (w - tw) // 2, (h - th) // 2
        return abi.pt_region(x0, y0, tw, th, 1), "the %dx%d tile at (%d, %d)" % (tw, th, x0, y0)
    return f


WORKLOADS = {
    "c1": ("default src/main.cpp scene, 800x480, 100 spp, depth 50 (BASELINE config 1)", 10, _rows(4)),
    "c2": ("RTIOW-like random spheres (488 spheres, numpy-RNG layout), 1920x1080, 64 spp, depth 50 (BASELINE config 2)", 10, _rows(8)),
    "c3": ("Cornell box of thin boxes + xy_rect with two constant_medium smoke boxes and a light, 1024x1024, 1024 spp, "
           "depth 50 (BASELINE config 3)", 3, _tile(128, 96)),
    "c4": ("10 002-triangle pyramid mesh, checker + image textures (numpy-RNG layout), 1920x1080, 256 spp, depth 50 "
           "(BASELINE config 4)", 3, _tile(96, 48)),
    "c5": ("4K motion blur + depth of field, 388 moving of 485 spheres (numpy-RNG layout), 3840x2160, 4096 spp, depth 50 "
           "(BASELINE config 5)", 1, _tile(96, 64)),
}


def load_workload(name):
    import scenes
    if name == "c1":
        sc, cam, (w, h, spp, d) = scenes.load_c1()
        return sc, cam, w, h, spp, d
    w, h, spp = {"c2": (1920, 1080, 64), "c3": (1024, 1024, 1024), "c4": (1920, 1080, 256), "c5": (3840, 2160, 4096)}[name]
    build = {"c2": scenes.rtiow, "c3": scenes.cornell, "c4": scenes.c4_mesh, "c5": scenes.motion_blur}[name]
    sc, cam = build(w / h)
    return sc, cam, w, h, spp, 50


def static_config(args, w, h, spp, d, world, h_total):
    """The workload's description: identical in both arms (measured values live in `measured`)."""
    return {"workload": WORKLOADS[args.workload][0] + ("" if world == 1 else "; %s scaling: image %dx%d over %d GPUs" % (args.scaling, w, h_total, world)),
            "width": w, "height": h_total, "spp": spp, "depth": d, "paths_per_step": w * h_total * spp,
            "partition": "rows interleaved over %d rank(s)" % world, "gather": args.gather if world > 1 else "none",
            "l2": "256 MiB memset between timed iterations"}


# ---- algorithmic work per path ("hoisted minimum" column) ------------------
F_SPHERE, F_MOVING_SPHERE, F_ROOT, F_SPHERE_ACCEPT = 17.0, 25.0, 3.0, 31.0
F_RECT, F_TRIANGLE, F_BOX, F_MEDIUM = 6.0, 22.0, 36.0, 60.0
F_CAMERA, F_SKY = 49.0, 21.0
F_SCATTER = [22.0 + 8.0, 50.0, 63.0, 0.0, 19.0]  # lambertian (+ texture), metal, dielectric, light, isotropic


def flops_per_path(c):
    """c: oracle counters (dict).  Least arithmetic any bit-identical brute-force scan must do."""
    t = c["tests"]
    static = t[0] - c["moving_sphere_tests"]
    w = static * F_SPHERE + c["moving_sphere_tests"] * F_MOVING_SPHERE
    w += t[1] * F_RECT + t[2] * F_TRIANGLE + t[3] * F_BOX + t[4] * F_MEDIUM
    w += c["accepts"][0] * (F_ROOT + F_SPHERE_ACCEPT) + sum(c["accepts"][1:]) * 23.0
    w += sum(n * f for n, f in zip(c["scatters"], F_SCATTER)) + c["sky"] * F_SKY + c["paths"] * F_CAMERA
    return w / max(c["paths"], 1)


# ---- clocks during the timed region ---------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([f.strip() for f in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        mhz, mx, reasons, watts = [], None, set(), []
        for s in self.samples:
            try:
                mhz.append(float(s[0]))
                mx = float(s[1])
                watts.append(float(s[6]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(mhz)[len(mhz) // 2:] if mhz else []  # the upper half = samples under load
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w_max": max(watts) if watts else None, "samples": len(mhz)}


# ---- the reference arm / cpu baseline ---------------------------------------------------------------
def pick_cpu_oracle(w, h, spp, d):
    from oracle.pyoracle import CPort, Ref
    if Ref.available():
        r = Ref()
        if r.supported(w, h, spp, d):
            return r
    return CPort()


def cpu_sample(oracle, sc, cam, w, h, spp, d, region, dynamic):
    """Render `region` at full spp on all host cores.  -> (Mpaths/s, seconds, paths, counters or None)"""
    t0 = time.perf_counter()
    res = oracle.render_region(sc, cam, w, h, spp, d, region, dynamic=dynamic, nthreads=host_threads())
    dt = time.perf_counter() - t0
    paths = region.w * region.h * spp
    counters = res[1].as_dict() if isinstance(res, tuple) else None
    return paths / dt / 1e6, dt, paths, counters


def cpu_baseline_record(args, sc, cam, w, h, spp, d, with_dynamic=True):
    """-> (record, oracle counters of the sample or None)"""
    oracle = pick_cpu_oracle(w, h, spp, d)
    region, words = WORKLOADS[args.workload][2](w, h)
    if args.workload == "c1" and args.cpu_stride:
        region, words = _rows(args.cpu_stride)(w, h)
    v_static, dt, paths, counters = cpu_sample(oracle, sc, cam, w, h, spp, d, region, False)
    rec = {"value": v_static, "unit": "Mpaths/s", "cores": host_threads(), "kind": oracle.kind,
           "sample": "%s at full spp (%d paths), OpenMP schedule(static) over rows like triSYCL's host parallel_for; %.1f s"
                     % (words, paths, dt)}
    if with_dynamic:
        v_dyn, dt2, _, _ = cpu_sample(oracle, sc, cam, w, h, spp, d, region, True)
        rec["value_dynamic_schedule"] = v_dyn
        rec["sample"] += " + %.1f s (dynamic)" % dt2
    return rec, counters


def run_reference_arm(args):
    """The reference's own CPU implementation of the path on this box's host cores, same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = args.gpus
    sc, cam, w, h, spp, d = load_workload(args.workload)
    h_total = h * world if args.scaling == "weak" else h
    oracle = pick_cpu_oracle(w, h, spp, d)
    region, words = WORKLOADS[args.workload][2](w, h)
    if args.workload == "c1" and args.cpu_stride:
        region, words = _rows(args.cpu_stride)(w, h)
    from path_tracer_b200 import abi
    small = abi.pt_region(region.x0, region.y0, region.w, max(region.h // 4, 1), region.y_stride)
    for _ in range(args.warmup):
        cpu_sample(oracle, sc, cam, w, h, spp, d, small, False)
    t_total, paths_total, full_step = 0.0, 0, None
    for it in range(args.steps):
        if it == 0 and args.workload == "c1" and oracle.kind == "reference" and hasattr(oracle, "render_full") and d == 50:
            # (the default scene only: 21 s on 16 cores.  The whole frame of config 4 would take the CPU three hours)
            # the first timed step goes through the reference's OWN entry point, render<W,H,S>() (render.hpp:141-160),
            # on the whole frame; the others through render_pixel<> on the bounded sample
            t0 = time.perf_counter()
            oracle.render_full(sc, cam, w, h, spp, dynamic=False, nthreads=host_threads())
            dt, paths = time.perf_counter() - t0, w * h * spp
            full_step = {"entry": "render<%d,%d,%d>()" % (w, h, spp), "seconds": dt, "Mpaths/s": paths / dt / 1e6}
        else:
            _, dt, paths, _ = cpu_sample(oracle, sc, cam, w, h, spp, d, region, False)
        t_total += dt
        paths_total += paths
    value = paths_total / t_total / 1e6
    line = {
        "impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (reference default scene captured from the unmodified main.cpp / procedural scenes); no external data",
        "config": static_config(args, w, h, spp, d, world, h_total),
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": host_threads(), "kind": oracle.kind,
                         "sample": "%s at full spp per step%s, OpenMP schedule(static) over rows like triSYCL's host "
                                   "parallel_for; the CPU path has one configuration (one box, its host cores): the "
                                   "workload's base image" % (words, ", the first step the whole frame" if full_step else ""),
                         "full_frame_step": full_step},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---- our arm -----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--cpu-stride", type=int, default=0, help="c1: the cpu baseline renders rows 0::stride (default 4)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = the workload's rows per GPU (image height x N), strong = the fixed image")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the fixed-image `strong` sub-record")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = WORKLOADS[args.workload][1]
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    from path_tracer_b200 import dist as ptdist
    from path_tracer_b200 import render as R

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs one process per GPU: launch with python -m torch.distributed.run "
                             "--nproc-per-node %d ..." % (args.gpus, args.gpus))
        raise SystemExit("WORLD_SIZE (%d) != --gpus (%d)" % (world, args.gpus))
    if R.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    sc, cam, w, h_base, spp, d = load_workload(args.workload)
    h = h_base * world if args.scaling == "weak" else h_base  # weak: same scene and camera, N times the rows

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def timed_run(height, steps, sample_clocks):
        """Device-resident renders of the image w x height over the ranks -> dict (timing max over ranks)."""
        rend = ptdist.DistRenderer(sc, cam, w, height, spp, d, rank, world, local_rank, mode=args.gather)
        for _ in range(args.warmup):
            rend.launch()
            rend.gather()
        rend.scene.counters(reset=True)
        launches0 = rend.scene.launch_count()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0 and sample_clocks:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t_wall0 = time.perf_counter()
        for a, b in ev:
            flush.zero_()                       # L2 flush between timed iterations (not inside the event pair)
            a.record(stream)
            rend.launch()
            if args.gather == "nccl" and world > 1:
                rend.gather()
            b.record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop() if rank == 0 and sample_clocks else None
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        _, scans_done = rend.scene.counters()
        launches = rend.scene.launch_count() - launches0
        t = torch.tensor([dev_ms, float(scans_done), float(launches)], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            dev_ms, scans_total, launches = float(tmax[0]), float(tsum[1]), int(tsum[2])
        else:
            scans_total = float(scans_done)
        ms_per_step = dev_ms / steps
        paths = w * height * spp
        final = rend.gather()
        return {"rend": rend, "ms_per_step": ms_per_step, "value": paths / (ms_per_step * 1e-3) / 1e6, "paths": paths,
                "scans_per_path": scans_total / (paths * steps), "launches": launches // max(steps, 1), "clocks": clocks,
                "fb_mean": float(final.mean()) if rank == 0 else None, "wall_ms": 1e3 * t_wall / steps}

    # ---------------- device-resident: `value`
    run = timed_run(h, args.steps, True)
    rend = run["rend"]
    paths_per_step = run["paths"]
    value, ms_per_step = run["value"], run["ms_per_step"]

    # ---------------- end to end through the host-buffer API: `e2e`
    fb_host = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
    pinned_tex = torch.from_numpy(np.ascontiguousarray(sc.texture_bytes)).pin_memory()
    sc.texture_bytes = pinned_tex.numpy()
    h2d = d2h = 0
    e2e_times = []
    e2e_steps = args.steps if ms_per_step < 2000 else 1  # (a 20 s frame is not repeated ten times over)
    e2e_warm = args.warmup if ms_per_step < 2000 else 0
    for it in range(e2e_warm + e2e_steps):
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            R.render(sc, cam, w, h, spp, d, out=fb_host.numpy())
            st = R.stats()
            h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]
        else:
            # per step: upload the scene to this rank's GPU, render its rows into the (persistent) shared
            # framebuffer mapping / local rows, gather, and read the image back on rank 0
            scene2 = R.DeviceScene(sc, local_rank)
            ptr, pitch = rend.target()
            scene2.render_region(cam, w, h, spp, d, rend.region, ptr, pitch, torch.cuda.current_stream().cuda_stream)
            full = rend.gather()
            if rank == 0:
                fb_host.copy_(full, non_blocking=False)
            h2d = int(sc.texture_bytes.size) + sum(int(a.nbytes) for a in sc.arrays().values())
            d2h = h * w * 12
            scene2.close()
        barrier()
        if it >= e2e_warm:
            e2e_times.append(time.perf_counter() - t0)
    e2e_t = torch.tensor([sum(e2e_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = paths_per_step * e2e_steps / float(e2e_t[0]) / 1e6
    e2e_mean = float(fb_host.mean())
    rend.close()

    # ---------------- N > 1: the fixed image over N GPUs (strong scaling), and the single-process multi-GPU entry
    strong = single = None
    if world > 1 and not args.no_strong:
        if args.scaling == "weak":
            s_run = timed_run(h_base, args.steps, False)
            s_run["rend"].close()
            strong = {"value": s_run["value"], "unit": "Mpaths/s", "ms_per_step": s_run["ms_per_step"],
                      "image": "%dx%d" % (w, h_base), "paths_per_step": s_run["paths"],
                      "note": "the workload's fixed image, rows interleaved over %d GPUs; compare with the N = 1 line" % world}
        barrier()
        if rank == 0 and R.device_count() >= 2:
            # pt_set_num_gpus(2) -> pt_render: one process driving two GPUs, peer stores into GPU 0's framebuffer
            sspp = max(spp // 25, 1)
            one = R.render(sc, cam, w, h_base, sspp, d)
            R.set_num_gpus(2)
            try:
                t0 = time.perf_counter()
                two = R.render(sc, cam, w, h_base, sspp, d)
                dt = time.perf_counter() - t0
            finally:
                R.set_num_gpus(1)
            single = {"n_gpus": 2, "spp": sspp, "bit_identical_to_one_gpu": bool(np.array_equal(one.view(np.uint32), two.view(np.uint32))),
                      "seconds_incl_upload": dt}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- rank 0: roofline + cpu baseline + the JSON line
    peak_tflops, peak_mhz = R.measure_fp32_peak(local_rank)
    cpu = counters = None
    if not args.no_cpu_baseline:
        cpu, counters = cpu_baseline_record(args, sc, cam, w, h_base, spp, d)
    if counters is None:
        # work counters for W: the C port counts them; a small sample at full spp is plenty
        try:
            from oracle.pyoracle import CPort, rows_region
            from path_tracer_b200 import abi
            if args.workload in ("c1", "c2"):
                reg = rows_region(w, h_base, 0, 16 if args.workload == "c1" else 60)
            else:
                big, _ = WORKLOADS[args.workload][2](w, h_base)
                reg = abi.pt_region(big.x0, big.y0, max(big.w // 4, 1), max(big.h // 4, 1), 1)
            _, cnt = CPort().render_region(sc, cam, w, h_base, spp, d, reg, nthreads=host_threads())
            counters = cnt.as_dict()
        except OSError:
            counters = None
    flop_per_path = flops_per_path(counters) if counters else float("nan")
    achieved = value * 1e6 * flop_per_path / 1e12
    traffic = executed = issued = per_launch = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        prof = json.load(open(tpath))
        traffic = prof.get(args.workload)
        per_launch = prof.get(args.workload + "_paths_per_launch", w * h_base * spp)
        if prof.get(args.workload + "_executed_flop_per_launch"):
            executed = prof[args.workload + "_executed_flop_per_launch"] / per_launch
        if prof.get(args.workload + "_thread_instructions_per_launch"):
            issued = prof[args.workload + "_thread_instructions_per_launch"] / per_launch
    peak_total = peak_tflops * world
    line = {
        "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (reference default scene captured from the unmodified main.cpp / procedural scenes); no external data",
        "config": static_config(args, w, h_base, spp, d, world, h),
        "measured": {"scans_per_path": run["scans_per_path"], "fb_mean": run["fb_mean"], "e2e_fb_mean": e2e_mean,
                     "wall_ms_per_step_incl_flush": run["wall_ms"], "e2e_steps": e2e_steps},
        "clocks": run["clocks"],
        "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * float(e2e_t[0]) / e2e_steps},
        "gpu_launches": int(run["launches"]),  # per step, all ranks: cost probe + tile sort + render kernel on each
        "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_total, "unit": "TFLOP/s",
                     "frac": achieved / peak_total, "traffic": traffic,
                     # the ncu capture's launch (profiles/traffic.json): c3-c5 were captured at reduced spp
                     "traffic_paths_per_launch": per_launch if traffic is not None else None,
                     "flop_per_path": flop_per_path,
                     "peak_per_gpu": peak_tflops, "n_gpus": world,
                     # what the kernel really executes (ncu, profiles/traffic.json): culling skips most of the
                     # reference's brute-force tests, so `achieved` (ALGORITHMIC flops / time) is not an issue rate
                     "executed_flop_per_path": executed,
                     "executed_tflops": None if executed is None else value * 1e6 * executed / 1e12,
                     "executed_frac": None if executed is None else value * 1e6 * executed / 1e12 / peak_total,
                     # every thread instruction (ncu), against the lane-issue peak = SMs x 128 lanes x clock = half the
                     # FFMA flop peak: what actually bounds this kernel
                     "thread_inst_per_path": issued,
                     "issue_frac": None if issued is None else value * 1e6 * issued / (0.5e12 * peak_total),
                     "peak_source": "%d x the FFMA micro-kernel rate measured in this run (%.0f MHz implied); "
                                    "MEASURED_PEAKS.json has no fp32 entry; `achieved` counts the reference's brute-force "
                                    "scan (every object for every ray), most of which culling skips" % (world, peak_mhz)},
        "cpu_baseline": cpu,
    }
    if line["roofline"]["frac"] > 1.0:
        line["roofline"]["note"] = ("frac > 1 is the order-preserving culling at work, not skipped work: W is the reference's "
                                    "brute-force scan (SURVEY.md 8(d): reported unchanged when a culling structure is used) and "
                                    "the kernel reaches the same pixels, bit for bit and with equal scan counts (tests), while "
                                    "testing a fraction of the objects; executed_frac / issue_frac are the hardware's view")
    if strong is not None:
        line["strong"] = strong
    if single is not None:
        line["single_process_multi_gpu"] = single
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
