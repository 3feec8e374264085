"""CPU-only: the sphere chunk layout and the chunk boxes the host side computes (pt_pack.cpp) -- the invariants
the kernel's chunk culling relies on, checked without a GPU through the pt_debug_chunk_layout test hook."""
import ctypes as C

import numpy as np
import pytest

import scenes
from path_tracer_b200 import abi
from path_tracer_b200 import render as R

CHUNK, SETS = 16, 4
INT_MIN = -2 ** 31


def chunk_layout(sc, t0, t1):
    s, keep = sc.as_c()
    ns, nm = C.c_int(), C.c_int()
    keys = (C.c_int * 65536)()
    boxes = (C.c_float * (1 << 20))()
    bounds = (C.c_float * 3)()
    L = R.lib()
    L.pt_debug_chunk_layout.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_int, C.c_void_p]
    n = L.pt_debug_chunk_layout(C.addressof(s), t0, t1, C.byref(ns), C.byref(nm), keys, 65536, boxes, 1 << 20, bounds)
    assert n >= 0, L.pt_last_error()
    n_el = ns.value + nm.value
    n_chunks = n_el // CHUNK
    assert n == SETS * n_chunks * 6 and ns.value % CHUNK == 0 and nm.value % CHUNK == 0
    return (ns.value, nm.value, np.array(keys[:n_el]), np.array(boxes[:n]).reshape(SETS, n_chunks, 2, 3),
            np.array(bounds[:]))


CASES = [("c1", 0.0, 1.0), ("rtiow", 0.0, 0.0), ("moving", 0.0, 1.0), ("moving", -0.5, 2.5), ("motion_blur", 0.2, 0.4),
         ("spheres_basic", 0.0, 0.0), ("media", 0.0, 1.0), ("random3", 0.0, 1.0), ("empty", 0.0, 0.0),
         ("single_light", 0.0, 1.0), ("cornell", 0.0, 0.0)]


@pytest.mark.parametrize("name,t0,t1", CASES)
def test_every_sphere_is_inside_its_chunk_box_over_the_shutter(name, t0, t1):
    if name == "c1":
        sc = scenes.load_c1()[0]
    elif name == "random3":
        sc = scenes.random_scene(3, n_objects=80)[0]
    else:
        sc = getattr(scenes, name)(4 / 3)[0]
    ns, nm, keys, boxes, bounds = chunk_layout(sc, t0, t1)
    order, spheres = sc.order, sc.spheres
    top_level = [i for i in range(len(order)) if order[i]["kind"] == abi.HIT_SPHERE]
    # every top-level sphere sits in exactly one element; the rest is padding
    real = keys[keys != INT_MIN]
    assert sorted(-1 - real) == top_level
    assert bounds[0] <= bounds[1] <= bounds[2]
    culled = 0
    for el, key in enumerate(keys):
        if key == INT_MIN:
            continue
        sp = spheres[order[-1 - key]["index"]]
        is_moving = el >= ns
        assert is_moving == (sp["time0"] != sp["time1"])
        c0, c1 = sp["center0"].astype(np.float64), sp["center1"].astype(np.float64)
        r = abs(float(sp["radius"]))
        fs = [0.0]
        if is_moving:
            den = float(np.float32(sp["time1"]) - np.float32(sp["time0"]))
            fs = [(t - float(sp["time0"])) / den for t in np.linspace(min(t0, t1), max(t0, t1), 9)]
        chunk = el // CHUNK
        for s in range(SETS):
            lo, hi = boxes[s, chunk, 0], boxes[s, chunk, 1]
            if s == SETS - 1:
                assert np.all(np.isneginf(lo)) and np.all(np.isposinf(hi))  # the last set never culls
            if np.all(np.isinf(lo)):
                continue
            culled += 1
            for f in fs:
                c = c0 + f * (c1 - c0)
                assert np.all(c - r > lo) and np.all(c + r < hi), (name, el, s)
            # a culled box is finite and grows with the set (a farther origin class needs a larger margin)
            assert np.all(np.isfinite(lo)) and np.all(np.isfinite(hi))
            if s > 0 and np.all(np.isfinite(boxes[s - 1, chunk, 0])):
                assert np.all(lo <= boxes[s - 1, chunk, 0]) and np.all(hi >= boxes[s - 1, chunk, 1])
    if name in ("c1", "rtiow", "motion_blur"):
        assert culled > 0.9 * 3 * len(top_level)  # all but the ground sphere are culled in sets 0..2
        # the ground sphere (radius 1000) is outsized: its chunk is never culled
        ground = [el for el, key in enumerate(keys) if key != INT_MIN and abs(spheres[order[-1 - key]["index"]]["radius"]) >= 500]
        assert ground and all(np.all(np.isinf(boxes[0, el // CHUNK, 0])) for el in ground)


def test_boxes_without_a_usable_shutter_do_not_cull():
    sc = scenes.moving(4 / 3)[0]
    _, _, _, boxes, bounds = chunk_layout(sc, float("nan"), 1.0)
    assert np.all(np.isinf(boxes)) and np.all(bounds == 0)


def test_chunks_are_compact():
    """k-d ordering: the boxes of the default scene's chunks are small next to the scene (that is the whole point)."""
    sc = scenes.load_c1()[0]
    _, _, _, boxes, _ = chunk_layout(sc, 0.0, 1.0)
    lo, hi = boxes[0, :, 0], boxes[0, :, 1]
    finite = np.all(np.isfinite(lo), axis=1)
    area = (hi[finite, 0] - lo[finite, 0]) * (hi[finite, 2] - lo[finite, 2])
    assert finite.sum() >= 30 and np.median(area) < 60.0  # the grid of small spheres spans 22 x 22
