"""bench.py's host-side contract (no GPU): the workloads of BASELINE.json's `configs`, one `config` dict for both arms,
CPU samples that lie inside the image, the work model W, and the no-GPU failure mode."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

SIZES = {"c1": (800, 480, 100), "c2": (1920, 1080, 64), "c3": (1024, 1024, 1024), "c4": (1920, 1080, 256), "c5": (3840, 2160, 4096)}


def test_workloads_are_baseline_configs():
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert len(cfgs) == len(bench.WORKLOADS) == 5
    for name, (w, h, spp) in SIZES.items():
        words = bench.WORKLOADS[name][0]
        assert "%dx%d" % (w, h) in words and "%d spp" % spp in words and "depth 50" in words


@pytest.mark.parametrize("name", sorted(SIZES))
def test_workload_scene_and_cpu_sample(name):
    sc, cam, w, h, spp, d = bench.load_workload(name)
    assert (w, h, spp, d) == SIZES[name] + (50,)
    assert sc.n_hittables > 0
    region, words = bench.WORKLOADS[name][2](w, h)
    assert 0 <= region.x0 and region.x0 + region.w <= w and 0 <= region.y0 and region.y_stride >= 1
    assert region.y0 + (region.h - 1) * region.y_stride < h
    assert region.w * region.h * spp <= 30_000_000  # a bounded sample (10-30 s of CPU work on a 16-core box)
    if name == "c4":
        assert len(sc.arrays()["triangles"]) >= 10_000  # "~10 k triangles" (BASELINE config 4)


def test_both_arms_print_the_same_config():
    args = argparse.Namespace(workload="c5", scaling="strong", gather="peer")
    a = bench.static_config(args, 3840, 2160, 4096, 50, 8, 2160)
    b = bench.static_config(args, 3840, 2160, 4096, 50, 8, 2160)
    assert a == b and a["paths_per_step"] == 3840 * 2160 * 4096 and "strong scaling" in a["workload"]
    one = bench.static_config(argparse.Namespace(workload="c1", scaling="weak", gather="peer"), 800, 480, 100, 50, 1, 480)
    assert one["gather"] == "none" and "scaling" not in one["workload"]


def test_work_model_counts_the_brute_force_scan():
    """W of SURVEY.md 8(d) on the oracle's counters of a tiny render of the default scene: every scan tests every object."""
    from oracle.pyoracle import CPort, rows_region
    sc, cam, w, h, spp, d = bench.load_workload("c1")
    _, cnt = CPort().render_region(sc, cam, w, h, 2, d, rows_region(w, h, 0, 60), nthreads=bench.host_threads())
    c = cnt.as_dict()
    assert sum(c["tests"]) == c["scans"] * sc.n_hittables
    wpp = bench.flops_per_path(c)
    per_scan = 490 * 17 + 1 * 6 + 4 * 22 + 36 + 60  # 490 spheres, 1 rect, 4 triangles, 1 box, 1 medium: misses only
    assert per_scan * c["scans"] / c["paths"] < wpp < 1.25 * per_scan * c["scans"] / c["paths"] + 200


def test_host_threads_ignores_omp_num_threads(monkeypatch):
    monkeypatch.setenv("OMP_NUM_THREADS", "1")  # torchrun exports this; the CPU arm must still use every core
    assert bench.host_threads() == len(os.sched_getaffinity(0))


def test_no_gpu_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
