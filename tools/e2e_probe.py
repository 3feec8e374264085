import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist, scenes
from path_tracer_b200 import render as R, dist as ptdist
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
sc, cam, (w, h, spp, d) = scenes.load_c1(); h *= world
rend = ptdist.DistRenderer(sc, cam, w, h, spp, d, rank, world, lr, mode="peer")
fb_host = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    dist.barrier(); t0 = T()
    scene2 = R.DeviceScene(sc, lr); t1 = T()
    ptr, pitch = rend.target()
    scene2.render_region(cam, w, h, spp, d, rend.region, ptr, pitch, torch.cuda.current_stream().cuda_stream); t2 = T()
    dist.barrier(); t3 = T()
    if rank == 0:
        full = rend.fb_view(); t4 = T()
        fb_host.copy_(full); t5 = T()
    else:
        t4 = t5 = T()
    scene2.close(); t6 = T()
    print("rank %d it %d: upload %.1f render %.1f barrier %.1f view %.1f d2h %.1f free %.1f ms" % (rank, it, *(1e3*(b-a) for a, b in zip((t0,t1,t2,t3,t4,t5),(t1,t2,t3,t4,t5,t6)))), flush=True)
dist.destroy_process_group()
