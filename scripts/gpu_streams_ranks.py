"""Where does a tile-split frame with sample streams lose time when several ranks run at once?  Under torchrun (N = 2 is enough): every rank
renders its 1/8 of a 1080p C3 frame with 8 streams (one wavefront of 8 passes), first one rank at a time, then all together, and prints host
enqueue / wait / device-span times (HC_PT_LOG=1 lines of the library) and the wall time per call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import hydracore_b200 as hc
from hydracore_b200 import scene as S, multigpu as MG
scn = S.scene_c3(1920, 1080)
lay = hc.CudaLayer(device=local)
lay.SetSampleStreams(8)
lay.LoadScene(scn)
lay.SetTiles(32, rank, 8)
if world > 1 and os.environ.get("WITH_COMM", "0") == "1":
    lay.SetTiles(32, rank, world)
    MG.join_communicator(lay, dist, torch.device("cuda", local))
lay.InitPathTracing(777)
lay.TracingPass(2, 8)
def bar():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
os.environ["HC_PT_LOG"] = "1"
COMM = world > 1 and os.environ.get("WITH_COMM", "0") == "1"
if COMM:
    lay.ReduceFramebuffer(0, 0)
for phase in (("together",) if COMM else ("alone", "together")):          # (a reduce needs every rank: no "alone" turns with the communicator)
    for turn in range(world if phase == "alone" else 1):
        bar()
        if phase == "together" or turn == rank:
            for rep in range(4):
                t0 = time.perf_counter()
                lay.TracingPass(2, 8)
                t1 = time.perf_counter()
                red = lay.ReduceFramebuffer(0, 0) if (world > 1 and os.environ.get("WITH_COMM", "0") == "1") else 0.0
                print("[%s] rank %d rep %d wall %.3f ms, reduce wall %.3f ms (device %.3f)" % (phase, rank, rep, 1e3*(t1 - t0), 1e3*(time.perf_counter() - t1), red), file=sys.stderr, flush=True)
                if phase == "together":
                    bar()
        bar()
lay.close()
if world > 1:
    dist.destroy_process_group()
