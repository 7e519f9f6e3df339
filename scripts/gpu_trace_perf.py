"""Traversal-kernel throughput on the C2 scene (device-resident rays, CUDA events inside the library): primary, shadow and
fully incoherent (cosine-distributed secondary) rays.  Usage under gpurun: python scripts/gpu_trace_perf.py [tag]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hydracore_b200 as hc  # noqa: E402
from hydracore_b200 import scene as S  # noqa: E402


def bench(fn, reps=10):
    ms = []
    for _ in range(reps):
        fn()
        ms.append(lay.last_trace_ms())
    return float(np.median(ms[2:]))


tag = sys.argv[1] if len(sys.argv) > 1 else "perf"
scn = S.scene_c2(1920, 1080)
lay = hc.CudaLayer()
lay.LoadScene(scn)
W, H = 1920, 1080
n = W*H
dev = torch.device("cuda", 0)
rays = torch.empty(n*8, dtype=torch.float32, device=dev)
hits = torch.empty(n*4, dtype=torch.int32, device=dev)
vis = torch.empty(n, dtype=torch.uint8, device=dev)
lay.make_eye_rays_device(W, H, rays.data_ptr())
out = {}
out["primary_ms"] = bench(lambda: lay.trace_closest_device(rays.data_ptr(), n, hits.data_ptr()))
out["primary_mrays"] = n/out["primary_ms"]/1e3
h = hits.view(-1, 4)
hit = h[:, 1] >= 0
t = h[:, 0].view(torch.float32)
r8 = rays.view(-1, 8)
pos = r8[:, 0:3] + r8[:, 4:7]*t[:, None]
nh = int(hit.sum().item())

# shadow rays to the C2 point light
srays = torch.empty(n*8, dtype=torch.float32, device=dev)
lay.make_shadow_rays_device(rays.data_ptr(), hits.data_ptr(), n, S.C2_LIGHT_POS, srays.data_ptr())
out["shadow_ms"] = bench(lambda: lay.trace_shadow_device(srays.data_ptr(), n, vis.data_ptr()))
out["shadow_mrays"] = nh/out["shadow_ms"]/1e3
out["shadow_visible_frac"] = float(vis[hit].float().mean().item())

# incoherent: cosine-distributed directions about +y from the primary hit points (fully incoherent secondary rays)
g = torch.Generator(device=dev)
g.manual_seed(7)
u = torch.rand(nh, 2, device=dev, generator=g)
rr = torch.sqrt(u[:, 0])
phi = 2*np.pi*u[:, 1]
d = torch.stack([rr*torch.cos(phi), torch.sqrt(1 - u[:, 0]).clamp_min(1e-3), rr*torch.sin(phi)], 1)
inc = torch.zeros(nh, 8, device=dev)
inc[:, 0:3] = pos[hit] + torch.tensor([0, 1e-3, 0], device=dev)
inc[:, 4:7] = d/d.norm(dim=1, keepdim=True)
inc[:, 7] = 3.0e38
inc = inc.contiguous()
hits2 = torch.empty(nh*4, dtype=torch.int32, device=dev)
out["incoherent_ms"] = bench(lambda: lay.trace_closest_device(inc.data_ptr(), nh, hits2.data_ptr()))
out["incoherent_mrays"] = nh/out["incoherent_ms"]/1e3
out["incoherent_hit_frac"] = float((hits2.view(-1, 4)[:, 1] >= 0).float().mean().item())
# the same rays in random order (no spatial coherence of origins either)
perm = torch.randperm(nh, device=dev, generator=g)
inc2 = inc[perm].contiguous()
out["incoherent_shuffled_ms"] = bench(lambda: lay.trace_closest_device(inc2.data_ptr(), nh, hits2.data_ptr()))
out["incoherent_shuffled_mrays"] = nh/out["incoherent_shuffled_ms"]/1e3
inc[:, 7] = 5.0
out["incoherent_anyhit_ms"] = bench(lambda: lay.trace_shadow_device(inc.data_ptr(), nh, vis.data_ptr()))
out["incoherent_anyhit_mrays"] = nh/out["incoherent_anyhit_ms"]/1e3
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", tag + "_trace_perf.json"), "w"), indent=1)
