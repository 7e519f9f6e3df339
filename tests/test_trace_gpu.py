"""GPU: K1 (eye rays), K2 (closest hit) and K2s (any hit) through the C ABI against the oracle, the golden vectors of the
reference and — at BASELINE sizes — size-independent properties."""
import os

import numpy as np
import pytest

from tests import refapi, scenes

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same_hits(a, b, rtol=1e-5):
    """bit-exact ids; t bit-exact except documented ties: equal-distance hits (|dt| <= 1e-5 relative) may pick another triangle."""
    exact = (a["primId"] == b["primId"]) & (a["instId"] == b["instId"]) & (a["geomId"] == b["geomId"]) & (a["t"] == b["t"])
    bad = ~exact
    ties = bad & (a["primId"] >= 0) & (b["primId"] >= 0) & (np.abs(a["t"] - b["t"]) <= rtol*np.abs(b["t"]))
    return int(bad.sum()), int((bad & ~ties).sum())


@pytest.fixture(scope="module")
def raycast():
    return np.load(os.path.join(G, "raycast.npz"))


def test_golden_hits(layer, raycast):
    layer.SetAllBVH4(raycast["bvh_nodes"], raycast["bvh_tris"])
    for rays, key in ((raycast["rays_primary"], "hits_primary"), (scenes.incoherent_rays(6000, 11), "hits_incoherent")):
        h = layer.TraceClosest(rays)
        nbad, nhard = _same_hits(h, raycast[key])
        assert nhard == 0 and nbad <= 2, (key, nbad, nhard)
    sh = scenes.incoherent_rays(6000, 11)
    sh[:, 7] = raycast["shadow_tfar"]
    v = layer.TraceShadow(sh)
    assert (v != raycast["vis_incoherent"]).sum() <= 2
    layer.SetAllBVH4(raycast["bvh1_nodes"], raycast["bvh1_tris"])
    h = layer.TraceClosest(raycast["rays_single_leaf"])
    assert _same_hits(h, raycast["hits_single_leaf"]) == (0, 0)


@pytest.mark.parametrize("dof", [False, True])
def test_eye_rays_bit_exact(layer, oracle, dof):
    scn = scenes.instanced_geometry(dof=dof)
    layer.LoadScene(scn)
    W, H = scn.width, scn.height
    offs = (np.random.RandomState(21).rand(W*H, 4)*2 - 1).astype(np.float32)
    got = layer.MakeEyeRays(W, H, offs)
    want = oracle.make_rand_eye_rays(scn.globals_blob, W, H, scenes.pixel_grid(W, H), offs)
    assert np.array_equal(got[:, 0:3], want[:, 0:3]) and np.array_equal(got[:, 4:7], want[:, 3:6])
    got0 = layer.MakeEyeRays(W, H, None)
    want0 = oracle.make_rand_eye_rays(scn.globals_blob, W, H, scenes.pixel_grid(W, H), np.zeros((W*H, 4), np.float32))
    assert np.array_equal(got0[:, 4:7], want0[:, 3:6])


def test_closest_and_shadow_vs_oracle_200k(layer, oracle):
    scn = scenes.instanced_geometry(320, 240)
    layer.LoadScene(scn)
    prim = layer.MakeEyeRays(320, 240, None)
    rays = np.concatenate([prim, scenes.incoherent_rays(120000, 3)])
    h = layer.TraceClosest(rays)
    ho = oracle.trace_closest(scn.bvh["nodes"], scn.bvh["tris"], rays)
    nbad, nhard = _same_hits(h, ho)
    assert nhard == 0 and nbad <= 1e-4*rays.shape[0], (nbad, nhard)
    sh = rays.copy()
    sh[:, 7] = np.random.RandomState(8).uniform(0.1, 20.0, rays.shape[0])
    v = layer.TraceShadow(sh)
    vo = oracle.trace_shadow(scn.bvh["nodes"], scn.bvh["tris"], sh)
    assert (v != vo).sum() <= 1e-4*rays.shape[0]


def test_edge_cases(layer):
    scn = scenes.single_triangle_leaf()
    layer.LoadScene(scn)
    assert layer.TraceClosest(np.zeros((0, 8), np.float32)).shape[0] == 0              # empty
    r = np.zeros((3, 8), np.float32)
    r[:, 0:3] = (0, 5, 0)
    r[0, 4:7] = (0, -1, 0)          # straight down: hit at t = 5
    r[1, 4:7] = (0, 1, 0)           # away: miss
    r[2, 4:7] = (1, 0, 0)           # parallel to the quad: miss (division by zero inside the triangle test must not hit)
    r[:, 7] = 3.0e38
    h = layer.TraceClosest(r)
    assert h["primId"][0] >= 0 and h["t"][0] == 5.0 and h["instId"][0] == 0 and h["geomId"][0] == 0
    assert h["primId"][1] == -1 and h["primId"][2] == -1 and h["geomId"][1] == -1073741824
    s = r.copy()
    s[:, 7] = (4.0, 10.0, 0.0)      # occluder beyond tFar -> visible ; miss -> visible ; tFar == 0 -> visible without tracing
    assert layer.TraceShadow(s).tolist() == [1, 1, 1]
    s[0, 7] = 6.0
    assert layer.TraceShadow(s).tolist() == [0, 1, 1]
    with pytest.raises(Exception):
        layer.SetAllBVH4(np.zeros((4, 8), np.float32), np.zeros((4, 4), np.float32))       # too small / invalid tree


def test_million_triangles_properties(layer, oracle):
    """BASELINE config C2 size: 1M-triangle mesh, 1080p primaries.  Checked through properties + an oracle sample."""
    from hydracore_b200 import scene as S
    scn = S.scene_c2()
    layer.LoadScene(scn)
    rays = layer.MakeEyeRays(1920, 1080, None)
    h = layer.TraceClosest(rays)
    hit = h["primId"] >= 0
    assert 0.8 < hit.mean() <= 1.0
    assert np.all(h["primId"][hit] < 2*708*707) and np.all(h["instId"][hit] == 0) and np.all(h["geomId"][hit] == 0)
    assert np.all(np.isfinite(h["t"][hit])) and np.all(h["t"][hit] > 0)
    # idempotence / determinism
    assert np.array_equal(h.view(np.uint8), layer.TraceClosest(rays).view(np.uint8))
    # a shadow ray stopped just short of the found hit is visible, one going past it is occluded
    s = rays[hit][:200000].copy()
    t = h["t"][hit][:200000]
    s[:, 7] = t*0.999
    assert layer.TraceShadow(s).mean() > 0.999
    s[:, 7] = t*1.01 + 1e-3
    assert layer.TraceShadow(s).mean() < 0.001
    # oracle on a strided sample
    idx = np.arange(0, rays.shape[0], 97)
    ho = oracle.trace_closest(scn.bvh["nodes"], scn.bvh["tris"], rays[idx])
    nbad, nhard = _same_hits(h[idx], ho)
    assert nhard == 0 and nbad <= 3


def test_raycast_pass_vs_oracle(layer, oracle):
    """hc_raycast_pass (K1 -> K2 -> shadow rays -> K2s, rays fetched in 8x4 pixel blocks): same hits as the oracle's BVH4InstTraverse up to
    the documented equal-t ties, same visibility as closest-hit-then-compare."""
    import ctypes as ct
    from hydracore_b200 import scene as S
    from hydracore_b200._lib import HC_HOST
    scn = scenes.instanced_geometry(320, 240)
    layer.LoadScene(scn)
    n = 320*240
    hits = np.empty(n, dtype=layer.TraceClosest(np.zeros((0, 8), np.float32)).dtype)
    vis = np.empty(n, np.uint8)
    light = (3.0, 12.0, 5.0)
    layer.RaycastPass(light, hits.ctypes.data, vis.ctypes.data, HC_HOST)
    rays = layer.MakeEyeRays(320, 240, None)
    want = oracle.trace_closest(scn.bvh["nodes"], scn.bvh["tris"], rays)
    nbad, nhard = _same_hits(hits, want)
    assert nhard == 0 and nbad <= 1e-4*n, (nbad, nhard)
    assert (hits["primId"] >= 0).mean() > 0.3
    srays = layer.MakeShadowRays(rays, want, light)
    vo = oracle.trace_shadow(scn.bvh["nodes"], scn.bvh["tris"], srays)
    same = hits["primId"] == want["primId"]
    assert ((vis != vo) & same).sum() <= 1e-4*n
    assert 0.05 < vis[hits["primId"] >= 0].mean() < 0.999


def test_raycast_pass_fused_equals_unfused_and_two_trees(layer):
    """The eye / shadow rays generated inside the traversal kernels' fetch give exactly the records of the path through ray buffers
    (HC_RAYCAST_UNFUSED=1); with a second (alpha-tested) BVH tree the pass walks both trees and equals hc_trace_closest on the same eye rays."""
    import os
    from hydracore_b200._lib import HC_HOST
    scn = scenes.instanced_geometry(320, 240, dof=False)
    layer.LoadScene(scn)
    n = 320*240
    dt = layer.TraceClosest(np.zeros((0, 8), np.float32)).dtype
    out = {}
    for mode in ("fused", "unfused"):
        if mode == "unfused":
            os.environ["HC_RAYCAST_UNFUSED"] = "1"
        try:
            h, v = np.empty(n, dt), np.empty(n, np.uint8)
            layer.RaycastPass((3.0, 12.0, 5.0), h.ctypes.data, v.ctypes.data, HC_HOST)
            out[mode] = (h, v)
        finally:
            os.environ.pop("HC_RAYCAST_UNFUSED", None)
    assert out["fused"][0].tobytes() == out["unfused"][0].tobytes() and np.array_equal(out["fused"][1], out["unfused"][1])
    cut = scenes.cornell_with_cutout(96, 96)
    layer.LoadScene(cut)
    h, v = np.empty(96*96, dt), np.empty(96*96, np.uint8)
    layer.RaycastPass((0.0, 3.0, 0.0), h.ctypes.data, v.ctypes.data, HC_HOST)
    want = layer.TraceClosest(layer.MakeEyeRays(96, 96, None))
    assert h.tobytes() == want.tobytes()
    assert set(np.unique(h["instId"])) >= {0, 1, 3}                 # hits in both trees (instances 1, 3, 4 live in the alpha-tested tree)
