"""Experiment: how much does ray ordering buy for incoherent secondary rays?  (device-resident rays, C2 scene)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hydracore_b200 as hc  # noqa: E402
from hydracore_b200 import scene as S  # noqa: E402

scn = S.scene_c2(1920, 1080)
lay = hc.CudaLayer()
lay.LoadScene(scn)
W, H = 1920, 1080
n = W*H
dev = torch.device("cuda", 0)
rays = torch.empty(n*8, dtype=torch.float32, device=dev)
hits = torch.empty(n*4, dtype=torch.int32, device=dev)
lay.make_eye_rays_device(W, H, rays.data_ptr())
lay.trace_closest_device(rays.data_ptr(), n, hits.data_ptr())
h = hits.view(-1, 4)
hit = h[:, 1] >= 0
t = h[:, 0].view(torch.float32)
r8 = rays.view(-1, 8)
pos = r8[:, 0:3] + r8[:, 4:7]*t[:, None]
g = torch.Generator(device=dev)
g.manual_seed(7)
# uniform hemisphere-ish directions about the geometric up vector, from every hit point: diffuse bounce
u = torch.rand(n, 2, device=dev, generator=g)
rr = torch.sqrt(u[:, 0])
phi = 2*np.pi*u[:, 1]
d = torch.stack([rr*torch.cos(phi), torch.sqrt(1 - u[:, 0]).clamp_min(1e-3), rr*torch.sin(phi)], 1)
inc = torch.zeros(n, 8, device=dev)
inc[:, 0:3] = pos + torch.tensor([0, 1e-3, 0], device=dev)
inc[:, 4:7] = d/d.norm(dim=1, keepdim=True)
inc[:, 7] = 3.0e38
pix = torch.arange(n, device=dev)
inc, pix = inc[hit].contiguous(), pix[hit]
m = inc.shape[0]
out = torch.empty(m*4, dtype=torch.int32, device=dev)


def bench(r):
    ms = []
    for _ in range(8):
        lay.trace_closest_device(r.data_ptr(), m, out.data_ptr())
        ms.append(lay.last_trace_ms())
    return float(np.median(ms[2:]))


def report(name, order):
    r = inc[order].contiguous()
    ms = bench(r)
    print("%-40s %.3f ms  %.0f Mrays/s" % (name, ms, m/ms/1e3))


report("pixel order (row-major)", torch.arange(m, device=dev))
report("random order", torch.randperm(m, device=dev, generator=g))
octant = ((inc[:, 4] < 0).long() | ((inc[:, 5] < 0).long() << 1) | ((inc[:, 6] < 0).long() << 2))
px, py = pix % W, pix//W
for ts in (8, 16, 32, 64, 128):
    tile = (py//ts)*((W + ts - 1)//ts) + (px//ts)
    key = tile*8 + octant
    report("tile %dx%d then octant" % (ts, ts), torch.argsort(key, stable=True))
    # 8x4 blocks inside
report("octant only (stable)", torch.argsort(octant, stable=True))
# quantised direction (6 faces x 4x4) within 32x32 tiles
ax = inc[:, 4:7].abs().argmax(1)
sgn = (inc[:, 4:7].gather(1, ax[:, None]).squeeze(1) < 0).long()
face = ax*2 + sgn
tile = (py//32)*((W + 31)//32) + (px//32)
report("tile 32x32 then cube face", torch.argsort(tile*6 + face, stable=True))
