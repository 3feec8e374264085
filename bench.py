#!/usr/bin/env python
"""bench.py -- Mpaths/s of the path-tracing hot path on B200, next to the reference's CPU path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one full render of the workload: by default BASELINE config 1, the reference's default
src/main.cpp scene at 800x480, 100 spp, depth 50 (38.4 M paths; the scene and camera come from
tests/golden/c1_scene.ptsc.gz, captured from the unmodified main.cpp).  A path = one camera sample.

  value   whole-job Mpaths/s with scene and framebuffer resident in HBM (device-resident C-ABI),
          timed with CUDA events on the launching stream, max over ranks.
  N > 1   the path shards by pixels with no data-path collective (seeds are global linear ids), so
          the scaling run is WEAK: every GPU renders 480 rows.  N GPUs render the default scene with
          the SAME camera at 800 x (480 N) -- N times the rows, interleaved (row r -> rank r mod N) --
          and the rows land in rank 0's framebuffer through peer stores over NVLink (or an NCCL
          gather).  At N = 1 this is exactly BASELINE config 1.  (Strong scaling of the fixed
          384 000-pixel image is capped near 2x by its deepest pixel -- 3 151 serial bounces -- see
          DESIGN.md section 7; `--scaling strong` measures it.)
  e2e     the same metric through the blocking host-buffer entry point (pt_render / the per-rank
          launcher): scene upload from pinned host memory + render + framebuffer download every step.
  roofline  FP32: achieved = value x W, W = algorithmic flop per path from the oracle's work counters
          of this exact workload and the per-test constants of SURVEY.md appendix D;
          peak = FFMA rate measured in this run by a register-resident micro-kernel
          (MEASURED_PEAKS.json has no fp32 entry).  W is the work of the reference's brute-force scan; the
          kernel skips most of it (chunk culling), so the line also carries the flops it really executes
          (ncu counters, profiles/traffic.json) -- `achieved` is an algorithmic rate, not an issue rate.
  cpu_baseline  the reference's CPU path (oracle/_ref/libptref.so = unmodified reference headers, or
          the C port when that library was not built) on this box's host cores, on a bounded sample
          of the same workload (every `stride`-th row at full spp).
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

WORKLOADS = {
    # name: (description, fixture loader)
    "c1": "default src/main.cpp scene, 800x480, 100 spp, depth 50 (BASELINE config 1)",
}


def load_workload(name):
    import scenes
    if name == "c1":
        sc, cam, (w, h, spp, d) = scenes.load_c1()
        return sc, cam, w, h, spp, d
    raise SystemExit("unknown workload %r" % name)


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
def cpu_sample(oracle, sc, cam, w, h, spp, d, stride, dynamic):
    from oracle.pyoracle import rows_region
    region = rows_region(w, h, 0, stride)
    t0 = time.perf_counter()
    res = oracle.render_region(sc, cam, w, h, spp, d, region, dynamic=dynamic, nthreads=0)
    dt = time.perf_counter() - t0
    paths = region.w * region.h * spp
    return paths / dt / 1e6, dt, paths, res


def pick_cpu_oracle(w, h, spp, d):
    from oracle.pyoracle import CPort, Ref
    if Ref.available():
        r = Ref()
        if r.supported(w, h, spp, d):
            return r
    return CPort()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sc, cam, w, h, spp, d = load_workload(args.workload)
    # the reference's CPU path has one configuration (one box, its host cores): BASELINE config 1
    oracle = pick_cpu_oracle(w, h, spp, d)
    stride = args.cpu_stride
    for _ in range(args.warmup):
        cpu_sample(oracle, sc, cam, w, h, spp, d, stride * 4, False)
    t_total, paths_total = 0.0, 0
    for _ in range(args.steps):
        _, dt, paths, _ = cpu_sample(oracle, sc, cam, w, h, spp, d, stride, False)
        t_total += dt
        paths_total += paths
    value = paths_total / t_total / 1e6
    dyn, _, _, _ = cpu_sample(oracle, sc, cam, w, h, spp, d, stride, True)
    cores = oracle.max_threads()
    line = {
        "impl": "reference", "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "reference default scene (fixture captured from the unmodified main.cpp); no external data",
        "config": {"workload": WORKLOADS[args.workload], "width": w, "height": h, "spp": spp, "depth": d},
        "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": oracle.kind,
                         "sample": "rows 0::%d of %d at full spp (%d paths/step), OpenMP schedule(static) over rows "
                                   "like triSYCL's host parallel_for" % (stride, h, paths_total // args.steps),
                         "value_dynamic_schedule": dyn},
        "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---- our arm -----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--cpu-stride", type=int, default=4, help="cpu baseline renders rows 0::stride")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = 480 rows per GPU (image 800 x 480N), strong = the fixed 800x480 image")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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

    sc, cam, w, h, spp, d = load_workload(args.workload)
    if args.scaling == "weak":
        h = h * world  # same scene and camera, N times the rows: every rank renders the base image's row count
    paths_per_step = w * h * spp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident: `value`
    rend = ptdist.DistRenderer(sc, cam, w, h, spp, d, rank, world, local_rank, mode=args.gather)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()
    for _ in range(args.warmup):
        rend.launch()
        rend.gather()
    rend.scene.counters(reset=True)
    launches0 = rend.scene.launch_count()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
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
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    paths_done, scans_done = rend.scene.counters()
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
    ms_per_step = dev_ms / args.steps
    value = paths_per_step / (ms_per_step * 1e-3) / 1e6
    final = rend.gather()
    fb_check = float(final.mean()) if rank == 0 else None

    # ---------------- end to end through the host-buffer API: `e2e`
    fb_host = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
    pinned_tex = torch.from_numpy(np.ascontiguousarray(sc.texture_bytes)).pin_memory()
    sc.texture_bytes = pinned_tex.numpy()
    h2d = d2h = 0
    e2e_times = []
    for it in range(args.warmup + args.steps):
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
        if it >= args.warmup:
            e2e_times.append(time.perf_counter() - t0)
    e2e_t = torch.tensor([sum(e2e_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = paths_per_step * args.steps / float(e2e_t[0]) / 1e6
    e2e_mean = float(fb_host.mean())

    if rank != 0:
        rend.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- rank 0: roofline + cpu baseline + the JSON line
    peak_tflops, peak_mhz = R.measure_fp32_peak(local_rank)
    cpu = None
    counters = None
    if world == 1 and not args.no_cpu_baseline:
        oracle = pick_cpu_oracle(w, h, spp, d)
        v_static, dt, paths, _ = cpu_sample(oracle, sc, cam, w, h, spp, d, args.cpu_stride, False)
        v_dyn, dt2, _, _ = cpu_sample(oracle, sc, cam, w, h, spp, d, args.cpu_stride, True)
        cpu = {"value": v_static, "unit": "Mpaths/s", "cores": oracle.max_threads(), "kind": oracle.kind,
               "sample": "rows 0::%d of %d at full spp (%d paths), OpenMP schedule(static) over rows like triSYCL's "
                         "host parallel_for; %.1f s + %.1f s (dynamic)" % (args.cpu_stride, h, paths, dt, dt2),
               "value_dynamic_schedule": v_dyn}
    # work counters for W: the C port counts them; a 1-in-16 row sample at full spp is plenty
    try:
        from oracle.pyoracle import CPort, rows_region
        _, cnt = CPort().render_region(sc, cam, w, h, spp, d, rows_region(w, h, 0, 16 * world))
        counters = cnt.as_dict()
        flop_per_path = flops_per_path(counters)
    except OSError:
        flop_per_path = 26.4e3  # SURVEY.md section 8(d), default scene
    achieved = value * 1e6 * flop_per_path / 1e12
    traffic = executed = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        prof = json.load(open(tpath))
        traffic = prof.get(args.workload)
        if prof.get(args.workload + "_executed_flop_per_launch") and world == 1:
            executed = prof[args.workload + "_executed_flop_per_launch"] / paths_per_step
    line = {
        "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32",
        "data": "reference default scene (fixture captured from the unmodified main.cpp); no external data",
        "config": {"workload": WORKLOADS[args.workload] + ("" if world == 1 else
                                                             "; %s scaling: image 800x%d over %d GPUs" % (args.scaling, h, world)),
                   "width": w, "height": h, "spp": spp, "depth": d,
                   "paths_per_step": paths_per_step, "partition": "rows interleaved over %d rank(s)" % world,
                   "gather": args.gather if world > 1 else "none", "l2": "256 MiB memset between timed iterations",
                   "scans_per_path": scans_total / (paths_per_step * args.steps), "fb_mean": fb_check,
                   "e2e_fb_mean": e2e_mean, "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * float(e2e_t[0]) / args.steps},
        "gpu_launches": int(launches),  # per step: cost probe + tile sort + render kernel, on every rank
        "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                     "frac": achieved / peak_tflops, "traffic": traffic, "flop_per_path": flop_per_path,
                     # what the kernel really executes (ncu, profiles/traffic.json): chunk culling skips most of the
                     # reference's brute-force sphere tests, so `achieved` (ALGORITHMIC flops / time) is not an issue rate
                     "executed_flop_per_path": executed,
                     "executed_tflops": None if executed is None else value * 1e6 * executed / 1e12,
                     "peak_source": "FFMA micro-kernel measured in this run (%.0f MHz implied); MEASURED_PEAKS.json "
                                    "has no fp32 entry; `achieved` counts the reference's brute-force scan (every object "
                                    "for every ray), of which the kernel executes about one sixth" % peak_mhz},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    rend.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
