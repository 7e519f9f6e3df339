"""Timing probe of hc_fb_reduce at 4K (133 MB full-size sum, 16.6 MB per rank for the tile gather at G = 8): python -m torch.distributed.run ... scripts/gpu_reduce_probe.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import hydracore_b200 as hc  # noqa: E402
from hydracore_b200 import multigpu as MG  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {"world": world}
for (w, h) in ((1920, 1080), (3840, 2160)):
    lay = hc.CudaLayer(device=local)
    lay.ResizeScreen(w, h)
    lay.SetTiles(32, rank, world)
    MG.join_communicator(lay, dist, dev)
    for mode in (0, 1):
        ms = []
        for _ in range(8):
            dist.barrier()
            torch.cuda.synchronize()
            ms.append(lay.ReduceFramebuffer(0, mode))
        out["%dx%d_mode%d_ms" % (w, h, mode)] = [round(x, 4) for x in ms]
    # torch's own reduce of the same buffer for comparison
    t = MG.framebuffer_tensor(lay, dev)
    tm = []
    for _ in range(6):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
        e1.record()
        torch.cuda.synchronize()
        tm.append(round(e0.elapsed_time(e1), 4))
    out["%dx%d_torch_reduce_ms" % (w, h)] = tm
    lay.close()
if rank == 0:
    print(json.dumps(out))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "reduce_probe_n%d.json" % world), "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
