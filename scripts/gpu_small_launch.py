"""Latency of small traversal launches: primary rays of the C1 scene (test_42, 25,612 triangles) for n = 32k ... 262k rays, and the HC_PT_LOG
per-launch times of one C1 pass.  python scripts/gpu_small_launch.py  (under gpurun)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hydracore_b200 as hc  # noqa: E402
from hydracore_b200 import hydra_scene as HS  # noqa: E402

scn = HS.build_scene(HS.load_fixture(os.path.join(ROOT, "tests", "golden", "hydra_scenes.npz"), "test_42"), 512, 512)
lay = hc.CudaLayer()
lay.LoadScene(scn)
n = 512*512
dev = torch.device("cuda", 0)
rays = torch.empty(n*8, dtype=torch.float32, device=dev)
hits = torch.empty(n*4, dtype=torch.int32, device=dev)
lay.make_eye_rays_device(512, 512, rays.data_ptr())
for m in (n, n//2, n//4, n//8, 4096, 32):
    ms = []
    for _ in range(12):
        lay.trace_closest_device(rays.data_ptr(), m, hits.data_ptr())
        ms.append(lay.last_trace_ms())
    t = float(np.median(ms[2:]))
    print("primary rays %7d: %.1f us  %.2f Grays/s" % (m, 1e3*t, m/t/1e6), flush=True)
os.environ["HC_PT_LOG"] = "1"
lay.InitPathTracing(777)
lay.TracingPass(0, 2)
