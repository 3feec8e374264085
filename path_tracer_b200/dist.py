"""One-process-per-GPU launcher for the render hot path (torch.distributed plumbing).

Pixels are independent and seeded by their GLOBAL linear id (reference
include/render.hpp:130-132), so the image shards with no data-path collective:
rank r renders rows r, r+N, r+2N, ... (contiguous bands are badly imbalanced,
SURVEY.md section 6; interleaved rows are within 1.2 % of each other).  The only
exchange is the final gather of the framebuffer to rank 0, done one of two ways:

  "peer"   rank 0 allocates the framebuffer, exports a CUDA IPC handle, every
           other rank maps it and its render kernel stores finished pixels
           STRAIGHT into rank 0's HBM over NVLink (no staging, no extra kernel);
  "nccl"   every rank renders its rows into a local buffer and the rows are
           gathered with torch.distributed (NCCL on GPUs, gloo in CPU tests).

torch is used for process-group plumbing and device buffers only.
"""
import ctypes as C

import numpy as np

from . import abi


def rank_rows(height, rank, world):
    """Rows owned by `rank`: rank, rank+world, ...  -> (first, count)."""
    count = (height - rank + world - 1) // world if rank < height else 0
    return rank, count


def rank_region(width, height, rank, world):
    first, count = rank_rows(height, rank, world)
    return abi.pt_region(0, first, width, count, world)


def max_rows(height, world):
    return (height + world - 1) // world


def gather_rows(local_rows, width, height, rank, world, dst=0, group=None):
    """Gather row-interleaved shards to `dst`.

    local_rows: torch tensor [rows_of(rank), width, 3] (fp32) on this rank's device.
    Returns the full [height, width, 3] tensor on `dst`, None elsewhere.
    """
    import torch
    import torch.distributed as dist

    pad = max_rows(height, world)
    send = torch.zeros((pad, width, 3), dtype=torch.float32, device=local_rows.device)
    send[: local_rows.shape[0]] = local_rows
    if world == 1:
        return local_rows.clone()
    if rank == dst:
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, recv, dst=dst, group=group)
        full = torch.empty((height, width, 3), dtype=torch.float32, device=local_rows.device)
        for r in range(world):
            _, n = rank_rows(height, r, world)
            full[r::world] = recv[r][:n]
        return full
    dist.gather(send, None, dst=dst, group=group)
    return None


class DistRenderer:
    """Per-rank renderer: scene resident on this rank's GPU, rows rank::world per step."""

    def __init__(self, scene, camera, width, height, spp, depth, rank, world, device, mode="peer"):
        import torch
        from . import render as R

        self.R = R
        self.torch = torch
        self.width, self.height, self.spp, self.depth = width, height, spp, depth
        self.rank, self.world, self.device, self.mode = rank, world, device, mode
        self.scene = R.DeviceScene(scene, device)
        self.camera = camera
        self.region = rank_region(width, height, rank, world)
        self.row_floats = width * 3
        self._peer_ptr = None
        self._fb_ptr = None
        self._stream = None  # the stream of the last launch(): gather() waits for THAT one
        torch.cuda.set_device(device)
        if world == 1 or mode == "nccl":
            first, count = rank_rows(height, rank, world)
            self.local = torch.empty((max(count, 1), width, 3), dtype=torch.float32, device="cuda:%d" % device)
            if world == 1:
                self.mode = "local"
        else:
            self._setup_peer()

    # -- peer mode: everyone stores into rank 0's framebuffer over NVLink
    def _setup_peer(self):
        import torch.distributed as dist
        L = self.R.lib()
        handle = [None]
        if self.rank == 0:
            p = C.c_void_p()
            self.R._check(L.pt_fb_alloc(self.device, self.height * self.row_floats * 4, C.byref(p)))
            self._fb_ptr = p.value
            buf = (C.c_ubyte * 64)()
            self.R._check(L.pt_fb_export(C.c_void_p(p.value), buf))
            handle[0] = bytes(buf)
        dist.broadcast_object_list(handle, src=0)
        if self.rank != 0:
            buf = (C.c_ubyte * 64).from_buffer_copy(handle[0])
            p = C.c_void_p()
            self.R._check(L.pt_fb_open(self.device, buf, C.byref(p)))
            self._peer_ptr = p.value

    def target(self):
        """(device pointer, row pitch in floats) this rank's kernel writes through."""
        if self.mode == "peer":
            base = self._fb_ptr if self.rank == 0 else self._peer_ptr
            return base + self.rank * self.row_floats * 4, self.world * self.row_floats
        return self.local.data_ptr(), self.row_floats

    def launch(self, stream=None):
        """Asynchronous: enqueue this rank's rows on `stream` (torch stream or None = current)."""
        self._stream = stream or self.torch.cuda.current_stream()
        st = self._stream.cuda_stream
        ptr, pitch = self.target()
        if self.region.h > 0:
            self.scene.render_region(self.camera, self.width, self.height, self.spp, self.depth, self.region, ptr,
                                     pitch, st)

    def gather(self):
        """Complete the step: full framebuffer on rank 0 (torch tensor or numpy view), None elsewhere."""
        import torch.distributed as dist
        torch = self.torch
        launched_on = self._stream or torch.cuda.current_stream()
        if self.mode == "local":
            launched_on.synchronize()
            return self.local
        if self.mode == "nccl":
            if launched_on != torch.cuda.current_stream():
                torch.cuda.current_stream().wait_stream(launched_on)  # the collective runs on the current stream
            _, n = rank_rows(self.height, self.rank, self.world)
            return gather_rows(self.local[:n], self.width, self.height, self.rank, self.world)
        # peer: the stores already landed in rank 0's HBM once every rank's kernel retired
        launched_on.synchronize()
        dist.barrier()
        return self.fb_view() if self.rank == 0 else None

    def fb_view(self):
        """Rank 0, peer mode: the framebuffer as a torch tensor aliasing the pt_fb_alloc memory."""
        torch = self.torch

        class _Ext:
            pass

        holder = _Ext()
        holder.__cuda_array_interface__ = {
            "shape": (self.height, self.width, 3), "typestr": "<f4", "data": (self._fb_ptr, False), "version": 2}
        return torch.as_tensor(holder, device="cuda:%d" % self.device)

    def close(self):
        L = self.R.lib()
        if self.mode == "peer" and self.world > 1:
            # nobody may still be storing into, or have mapped, rank 0's framebuffer when it is freed
            import torch.distributed as dist
            if self._stream is not None:
                self._stream.synchronize()
            if self._peer_ptr:
                L.pt_fb_close(C.c_void_p(self._peer_ptr))
                self._peer_ptr = None
            dist.barrier()
        if self._peer_ptr:
            L.pt_fb_close(C.c_void_p(self._peer_ptr))
            self._peer_ptr = None
        if self._fb_ptr:
            L.pt_fb_free(self.device, C.c_void_p(self._fb_ptr))
            self._fb_ptr = None
        self.scene.close()
