"""Multi-GPU exchange inside the library (hc_comm_init / hc_fb_reduce) on N >= 2 GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 scripts/gpu_comm_check.py
Tile partition (MISPT): the image gathered on rank 0 must equal, bit for bit, the frame one GPU renders alone, also when the gather is repeated
after more passes.  Sample partition (MISPT-QMC): the full-size sum on rank 0 must match the single-GPU frame within the accumulation-order
tolerance, and repeating the reduce must not count anything twice."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import hydracore_b200 as hc  # noqa: E402
from hydracore_b200 import multigpu as MG  # noqa: E402
from tests import scenes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {"world": world}
scn = scenes.cornell(200, 152)
lay = hc.CudaLayer(device=local)
lay.LoadScene(scn)
MG.join_communicator(lay, dist, dev)
ver = __import__("ctypes").c_int()
hc.load().hc_comm_version(__import__("ctypes").byref(ver))
out["nccl_version"] = ver.value

# ---- ray casting: one frame split in tiles, records gathered on rank 0 (hc_raycast_pass + hc_comm_gather_raycast) == the frame of one GPU
from hydracore_b200._lib import HC_HOST  # noqa: E402
lay.SetTiles(32, rank, world)
nn = scn.width*scn.height
hits = np.zeros(nn, dtype=[("t", np.float32), ("primId", np.int32), ("instId", np.int32), ("geomId", np.int32)])
vis = np.zeros(nn, np.uint8)
for _ in range(2):
    lay.RaycastPass((0.0, 3.0, 0.0), hits.ctypes.data, vis.ctypes.data, HC_HOST)
if rank == 0:
    solo = hc.CudaLayer(device=local)
    solo.LoadScene(scn)
    h1, v1 = np.zeros_like(hits), np.zeros_like(vis)
    solo.RaycastPass((0.0, 3.0, 0.0), h1.ctypes.data, v1.ctypes.data, HC_HOST)
    out["raycast_split_equals_single_gpu"] = bool(hits.tobytes() == h1.tobytes() and np.array_equal(vis, v1))
    out["raycast_hit_fraction"] = float((h1["primId"] >= 0).mean())
    solo.close()

# ---- tile partition, MISPT
lay.SetTiles(32, rank, world)
lay.InitPathTracing(777)
lay.TracingPass(2, 2)
ms1 = lay.ReduceFramebuffer(0, 0)
img1 = lay.GetSumImage() if rank == 0 else None
lay.TracingPass(2, 1)
ms2 = lay.ReduceFramebuffer(0, 0)
ms3 = lay.ReduceFramebuffer(0, 0)          # repeated: idempotent
img2 = lay.GetSumImage() if rank == 0 else None
hdr2 = lay.GetHDRImage() if rank == 0 else None
if rank == 0:
    solo = hc.CudaLayer(device=local)
    solo.LoadScene(scn)
    solo.SetTiles(32, 0, 1)
    solo.InitPathTracing(777)
    solo.TracingPass(2, 2)
    ref1 = solo.GetSumImage()
    solo.TracingPass(2, 1)
    ref2, refh = solo.GetSumImage(), solo.GetHDRImage()
    out["tiles_equal_after_2_passes"] = bool(np.array_equal(img1, ref1))
    out["tiles_equal_after_3_passes_and_repeated_reduce"] = bool(np.array_equal(img2, ref2))
    out["hdr_equal"] = bool(np.array_equal(hdr2, refh))
    out["tiles_reduce_ms"] = [ms1, ms2, ms3]
    solo.close()

# ---- sample partition, MISPT-QMC (full-size buffers, ncclReduce into a separate buffer)
lay.SetTiles(32, rank, world)
lay.InitPathTracing(555)
lay.TracingPass(3, 2)
msq = lay.ReduceFramebuffer(0, 1)
q1 = lay.GetSumImage() if rank == 0 else None
lay.ReduceFramebuffer(0, 1)
q2 = lay.GetSumImage() if rank == 0 else None
if rank == 0:
    solo = hc.CudaLayer(device=local)
    solo.LoadScene(scn)
    solo.SetTiles(32, 0, 1)
    solo.InitPathTracing(555)
    solo.TracingPass(3, 2)
    rq = solo.GetSumImage()
    d = np.abs(q1 - rq)
    out["qmc_rel_rmse_vs_single_gpu"] = float(np.sqrt((d[..., :3]**2).mean())/max(1e-12, np.sqrt((rq[..., :3]**2).mean())))
    out["qmc_repeat_equal"] = bool(np.array_equal(q1, q2))
    out["qmc_reduce_ms"] = msq
    solo.close()
    ok = out["raycast_split_equals_single_gpu"] and out["tiles_equal_after_2_passes"] and out["tiles_equal_after_3_passes_and_repeated_reduce"] and out["hdr_equal"] and out["qmc_repeat_equal"] and out["qmc_rel_rmse_vs_single_gpu"] < 1e-3
    out["ok"] = bool(ok)
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "comm_check_n%d.json" % world), "w"), indent=1)
lay.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (rank != 0 or out.get("ok")) else 1)
