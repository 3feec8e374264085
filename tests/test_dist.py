"""CPU-only, world_size 2 over gloo: the row-interleave partition and the gather to rank 0 reproduce the
single-process image exactly.  The renderer is stood in by the C oracle (test code may call it)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenes
from path_tracer_b200 import dist as ptdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, spp, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from oracle.pyoracle import CPort
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc, cam = scenes.spheres_basic(w / h)
    region = ptdist.rank_region(w, h, rank, world)
    rows, _ = CPort().render_region(sc, cam, w, h, spp, 50, region, nthreads=1)
    full = ptdist.gather_rows(torch.from_numpy(rows), w, h, rank, world)
    if rank == 0:
        np.save(out_path, full.numpy())
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,h", [(2, 13), (2, 12), (3, 7)])
def test_row_interleave_gather_gloo(tmp_path, cport, world, h):
    w, spp = 16, 2
    out = str(tmp_path / "fb.npy")
    mp.spawn(_worker, args=(world, _free_port(), w, h, spp, out), nprocs=world, join=True)
    got = np.load(out)
    sc, cam = scenes.spheres_basic(w / h)
    want, _ = cport.render(sc, cam, w, h, spp, 50)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_partition_covers_every_row_once():
    for h in (1, 2, 7, 480, 1080):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(h, int)
            for r in range(world):
                first, n = ptdist.rank_rows(h, r, world)
                reg = ptdist.rank_region(5, h, r, world)
                assert (reg.y0, reg.h, reg.y_stride) == (first, n, world) and n <= ptdist.max_rows(h, world)
                seen[first::world][:n] += 1
            assert (seen == 1).all()
