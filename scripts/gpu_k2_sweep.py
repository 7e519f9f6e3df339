"""K2 A/B on the C2 scene (device-resident rays, CUDA events inside the library): first-generation kernel (HC_TRACE_IMPL=1) against
k_trace2 over refill threshold x quad bias, with a hit-for-hit comparison of every ray class.
Usage under gpurun: python scripts/gpu_k2_sweep.py tag [impl:refill:qbias ...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hydracore_b200 as hc  # noqa: E402
from hydracore_b200 import scene as S  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "k2"
configs = sys.argv[2:] or ["1:24:4", "2:8:4", "2:12:4", "2:16:4", "2:24:4", "2:8:2", "2:8:8", "2:4:4"]
scene_name = os.environ.get("HC_SWEEP_SCENE", "c2")
W, H = 1920, 1080
scn = S.scene_c2(W, H) if scene_name == "c2" else S.scene_c4(W, H)
n = W*H
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(7)
state = {}


def med(fn, lay, reps=int(os.environ.get('HC_SWEEP_REPS', '9'))):
    ms = []
    for _ in range(reps):
        fn()
        ms.append(lay.last_trace_ms())
    return float(np.median(ms[2:] if len(ms) > 2 else ms))


def run(cfg):
    impl, refill, qbias = cfg.split(":")
    os.environ["HC_TRACE_IMPL"], os.environ["HC_TRACE_REFILL"], os.environ["HC_TRACE_QBIAS"] = impl, refill, qbias
    lay = hc.CudaLayer()
    lay.LoadScene(scn)
    out = {"cfg": cfg}
    if "rays" not in state:
        rays = torch.empty(n*8, dtype=torch.float32, device=dev)
        lay.make_eye_rays_device(W, H, rays.data_ptr())
        state["rays"] = rays
    rays = state["rays"]
    hits = torch.empty(n*4, dtype=torch.int32, device=dev)
    vis = torch.empty(n, dtype=torch.uint8, device=dev)
    out["primary_ms"] = med(lambda: lay.trace_closest_device(rays.data_ptr(), n, hits.data_ptr()), lay)
    out["primary_mrays"] = n/out["primary_ms"]/1e3
    if "inc" not in state:
        h = hits.view(-1, 4)
        hit = h[:, 1] >= 0
        t = h[:, 0].view(torch.float32)
        r8 = rays.view(-1, 8)
        pos = r8[:, 0:3] + r8[:, 4:7]*t[:, None]
        nh = int(hit.sum().item())
        srays = torch.empty(n*8, dtype=torch.float32, device=dev)
        lay.make_shadow_rays_device(rays.data_ptr(), hits.data_ptr(), n, S.C2_LIGHT_POS, srays.data_ptr())
        u = torch.rand(nh, 2, device=dev, generator=g)
        rr = torch.sqrt(u[:, 0])
        phi = 2*np.pi*u[:, 1]
        # cosine-distributed about the geometric up axis of the terrain (C2) / about +y (C4): fully incoherent secondary rays
        d = torch.stack([rr*torch.cos(phi), torch.sqrt(1 - u[:, 0]).clamp_min(1e-3), rr*torch.sin(phi)], 1)
        inc = torch.zeros(nh, 8, device=dev)
        inc[:, 0:3] = pos[hit] + torch.tensor([0, 1e-3, 0], device=dev)
        inc[:, 4:7] = d/d.norm(dim=1, keepdim=True)
        inc[:, 7] = 3.0e38
        perm = torch.randperm(nh, device=dev, generator=g)
        state.update(inc=inc.contiguous(), inc2=inc[perm].contiguous(), srays=srays, nh=nh, hit=hit)
    inc, inc2, srays, nh, hit = state["inc"], state["inc2"], state["srays"], state["nh"], state["hit"]
    out["shadow_ms"] = med(lambda: lay.trace_shadow_device(srays.data_ptr(), n, vis.data_ptr()), lay)
    out["shadow_mrays"] = nh/out["shadow_ms"]/1e3
    hits2 = torch.empty(nh*4, dtype=torch.int32, device=dev)
    out["incoherent_ms"] = med(lambda: lay.trace_closest_device(inc.data_ptr(), nh, hits2.data_ptr()), lay)
    out["incoherent_mrays"] = nh/out["incoherent_ms"]/1e3
    hits3 = torch.empty(nh*4, dtype=torch.int32, device=dev)
    out["incoherent_shuffled_ms"] = med(lambda: lay.trace_closest_device(inc2.data_ptr(), nh, hits3.data_ptr()), lay)
    out["incoherent_shuffled_mrays"] = nh/out["incoherent_shuffled_ms"]/1e3
    vis2 = torch.empty(nh, dtype=torch.uint8, device=dev)
    inc_s = inc.clone()
    inc_s[:, 7] = 5.0
    out["incoherent_anyhit_ms"] = med(lambda: lay.trace_shadow_device(inc_s.data_ptr(), nh, vis2.data_ptr()), lay)
    out["incoherent_anyhit_mrays"] = nh/out["incoherent_anyhit_ms"]/1e3
    res = dict(primary=hits.clone(), shadow=vis.clone(), inc=hits2.clone(), incs=hits3.clone(), any=vis2.clone())
    if "ref" not in state:
        state["ref"] = res
    else:
        ref = state["ref"]
        for k in res:
            a, b = res[k], ref[k]
            if a.dtype == torch.uint8:
                out["diff_" + k] = int((a != b).sum().item())
            else:
                a4, b4 = a.view(-1, 4), b.view(-1, 4)
                bad = (a4 != b4).any(dim=1)
                ta, tb = a4[:, 0].view(torch.float32), b4[:, 0].view(torch.float32)
                tie = bad & (a4[:, 1] >= 0) & (b4[:, 1] >= 0) & ((ta - tb).abs() <= 1e-5*tb.abs())
                out["diff_" + k] = [int(bad.sum().item()), int((bad & ~tie).sum().item())]
    lay.close()
    print(json.dumps(out), flush=True)
    return out


results = [run(c) for c in configs]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(results, open(os.path.join(ROOT, "gpurun_out", tag + "_k2_sweep.json"), "w"), indent=1)
