"""GPU: parity with the reference (oracle/_ref: the reference's own integrators and traversal compiled in place) AT THE BASELINE SIZES
(BASELINE.json configs C1 ... C4).  Per-pixel generators (gen[p] = RandomGenInit(seed + p), carried across passes) make every pixel of a
frame independent of the others, so the reference can render scattered pixel windows of the full-size frame in seconds
(ref_render_pass(x0, y0, x1, y1), IntegratorCommon::DoPass semantics, CPUExp_Integrators_Common.cpp:278-316) and the GPU frame is
compared with them pixel for pixel.  Tolerances as in test_path_gpu.py: MISPT per-pixel relative 1e-5 on >= 99.9 % of the pixels and
relRMSE <= 1e-4 (observed bit-exact), PT relRMSE <= 1e-5; hit records bit-exact up to documented equal-t ties."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MISPT, PT = 2, 0


def _rel_rmse(a, b):
    return float(np.sqrt(((a - b)**2).mean())/max(np.sqrt((b**2).mean()), 1e-12))


def _tiles(width, height, count, seed, size=32):
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(count):
        x0 = int(rs.randint(0, width//size))*size
        y0 = int(rs.randint(0, height//size))*size
        out.append((x0, y0, x0 + size, y0 + size))
    return out


def _compare_tiles(layer, ref, scn, integrator, ref_kind, passes, tiles, seed=777, streams=1):
    layer.SetSampleStreams(streams)
    try:
        layer.LoadScene(scn)
        layer.InitPathTracing(seed)
        layer.TracingPass(integrator, passes)
        got = layer.GetHDRImage()[..., :3]*np.float32(passes)
    finally:
        layer.SetSampleStreams(1)
    rs = ref.scene(scn)
    worst, close_frac, lit = 0.0, 1.0, 0
    try:
        for (x0, y0, x1, y1) in tiles:
            want, npass = rs.render(ref_kind, seed, passes, window=(x0, y0, x1, y1), streams=streams)
            assert npass == passes
            w, g = want[y0:y1, x0:x1, :3], got[y0:y1, x0:x1]
            assert np.isfinite(g).all()
            if float(np.abs(w).max()) == 0.0:
                assert float(np.abs(g).max()) == 0.0
                continue
            lit += 1
            worst = max(worst, _rel_rmse(g, w))
            close_frac = min(close_frac, float((np.abs(g - w) <= 1e-5*np.maximum(np.abs(w), 1e-3)).all(-1).mean()))
    finally:
        rs.close()
    return worst, close_frac, lit


def test_c1_at_its_stated_size_vs_reference(layer, ref):
    """C1 as BASELINE.json states it: hydra_app/tests/test_42, unidirectional PT (IntegratorStupidPT), 512 x 512, 64 spp - the whole frame."""
    from hydracore_b200 import hydra_scene as HS
    scn = HS.build_scene(HS.load_fixture(os.path.join(G, "hydra_scenes.npz"), "test_42"), 512, 512)
    layer.LoadScene(scn)
    layer.InitPathTracing(777)
    layer.TracingPass(PT, 64)
    assert abs(layer.GetSPP() - 64.0) < 1e-4
    got = layer.GetHDRImage()[..., :3]*np.float32(64)
    rs = ref.scene(scn)
    try:
        want, npass = rs.render(0, 777, 64)
    finally:
        rs.close()
    assert npass == 64 and np.isfinite(got).all()
    assert _rel_rmse(got, want[..., :3]) <= 1e-5, _rel_rmse(got, want[..., :3])


def test_c3_full_size_tiles_vs_reference(layer, ref):
    """C3: MISPT, trace_depth 8, mixed materials, 1,001,116 triangles, 1920 x 1080: eight scattered 32 x 32 tiles of the full-size frame, 2 passes."""
    from hydracore_b200 import scene as S
    scn = S.scene_c3(1920, 1080)
    worst, close, lit = _compare_tiles(layer, ref, scn, MISPT, 2, 2, _tiles(1920, 1080, 8, 5))
    assert lit >= 6 and close >= 0.999 and worst <= 1e-4, (worst, close, lit)


def test_c3_full_size_with_sample_streams_vs_reference(layer, ref):
    """C3 at 1080p with 8 sample streams, 8 passes: two wavefronts of four passes each (8.3 M paths in flight, material sort on) against the reference
    integrator driven with the same stream rule, on six scattered tiles - the configuration bench.py times."""
    from hydracore_b200 import scene as S
    scn = S.scene_c3(1920, 1080)
    worst, close, lit = _compare_tiles(layer, ref, scn, MISPT, 2, 8, _tiles(1920, 1080, 6, 21), streams=8)
    assert lit >= 4 and close >= 0.999 and worst <= 1e-4, (worst, close, lit)


def test_c4_full_size_tiles_vs_reference(layer, ref):
    """C4: 200 rigid instances of a 100,352-triangle patch (20 M instanced triangles), five diffuse materials (material sort on), MISPT,
    1920 x 1080: twelve scattered tiles of the full-size frame."""
    from hydracore_b200 import scene as S
    scn = S.scene_c4(1920, 1080)
    worst, close, lit = _compare_tiles(layer, ref, scn, MISPT, 2, 2, _tiles(1920, 1080, 12, 9))       # about half of the frame is sky (black tiles must be black)
    assert lit >= 4 and close >= 0.999 and worst <= 1e-4, (worst, close, lit)


def test_small_c4_whole_frame_vs_reference(layer, ref):
    """The C4 construction with a small patch (200 random rigid instances x 1,152 triangles sharing one sub-tree): the whole 160 x 120 frame,
    MISPT and PT - where instance entry / exit and SafeInverse bugs would show."""
    from hydracore_b200 import scene as S
    scn = S.scene_c4(160, 120, instances=200, grid=(24, 24))
    for integ, kind, tol in ((MISPT, 2, 1e-4), (PT, 0, 1e-5)):
        layer.LoadScene(scn)
        layer.InitPathTracing(31)
        layer.TracingPass(integ, 3)
        got = layer.GetHDRImage()[..., :3]*np.float32(3)
        rs = ref.scene(scn)
        try:
            want, _ = rs.render(kind, 31, 3)
        finally:
            rs.close()
        assert float(np.abs(want).max()) > 0.0 and _rel_rmse(got, want[..., :3]) <= tol, (integ, _rel_rmse(got, want[..., :3]))


def test_million_incoherent_rays_on_c2_vs_oracle(layer, oracle):
    """1 M cosine-distributed secondary rays leaving the primary hit points of the C2 frame (the `mrays_incoherent` workload of bench.py):
    every 50th ray against the oracle's BVH4InstTraverse, ids and t bit-exact up to equal-t ties; any-hit agrees with closest-hit-then-compare."""
    from hydracore_b200 import scene as S
    scn = S.scene_c2()
    layer.LoadScene(scn)
    rays = layer.MakeEyeRays(1920, 1080, None)
    h = layer.TraceClosest(rays)
    hit = np.nonzero(h["primId"] >= 0)[0][:1000000]
    rs = np.random.RandomState(7)
    u = rs.rand(hit.size, 2).astype(np.float32)
    rr, phi = np.sqrt(u[:, 0]), np.float32(2*np.pi)*u[:, 1]
    d = np.stack([rr*np.cos(phi), np.maximum(np.sqrt(1 - u[:, 0]), 1e-3), rr*np.sin(phi)], 1).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    inc = np.zeros((hit.size, 8), np.float32)
    inc[:, 0:3] = rays[hit, 0:3] + rays[hit, 4:7]*h["t"][hit, None] + np.array([0, 1e-3, 0], np.float32)
    inc[:, 4:7] = d
    inc[:, 7] = 3.0e38
    got = layer.TraceClosest(inc)
    assert 0.3 < (got["primId"] >= 0).mean() < 0.8
    idx = np.arange(0, inc.shape[0], 50)
    want = oracle.trace_closest(scn.bvh["nodes"], scn.bvh["tris"], np.ascontiguousarray(inc[idx]))
    g = got[idx]
    exact = (g["primId"] == want["primId"]) & (g["instId"] == want["instId"]) & (g["geomId"] == want["geomId"]) & (g["t"] == want["t"])
    ties = ~exact & (g["primId"] >= 0) & (want["primId"] >= 0) & (np.abs(g["t"] - want["t"]) <= 1e-5*np.abs(want["t"]))
    assert int((~exact & ~ties).sum()) == 0 and int((~exact).sum()) <= 3, (int((~exact).sum()), int((~exact & ~ties).sum()))
    sh = inc[:200000].copy()
    sh[:, 7] = 6.0
    vis = layer.TraceShadow(sh)
    t = got["t"][:200000]
    want_vis = ~((got["primId"][:200000] >= 0) & (t > 0) & (t < 6.0))
    assert (vis.astype(bool) != want_vis).sum() <= 2


def test_shadow_rays_through_the_alpha_tested_tree(layer):
    """hc_pt_set_shadow_trees(1) (the library default, what GPUOCLLayer does, GPUOCLKernels.cpp:959-1000): an any-hit query walks both BVH trees and a
    cut-out occludes exactly where the two-tree closest hit lands on it; mode 0 (the CPU integrators' shadowTrace) sees the first tree only."""
    from tests import scenes
    scn = scenes.cornell_with_cutout(96, 96)
    layer.LoadScene(scn)
    rays = layer.MakeEyeRays(96, 96, None)
    both = layer.TraceClosest(rays)                       # tree 0, then tree 1 with the opacity test
    tree1 = np.isin(both["instId"], (1, 3, 4)) & (both["primId"] >= 0)        # nearest hit on an instance of the alpha-tested tree (cornell_with_cutout)
    assert tree1.sum() > 50
    sh = rays.copy()
    sh[:, 7] = np.where(both["primId"] >= 0, np.minimum(both["t"], np.float32(1.0e30))*np.float32(1.001), np.float32(1.0e30))     # just past the nearest hit of both trees
    try:
        layer.SetShadowTrees(1)
        vis_all = layer.TraceShadow(sh)
        layer.SetShadowTrees(0)
        vis_first = layer.TraceShadow(sh)
    finally:
        layer.SetShadowTrees(0)
    assert np.array_equal(vis_all == 0, both["primId"] >= 0)                  # every ray that hits something in either tree is occluded
    assert (vis_first[tree1] == 1).mean() > 0.9                               # the first tree alone does not see the cut-outs
    assert np.array_equal(vis_first[~tree1], vis_all[~tree1])
    # path tracing: cut-outs that cast shadows darken the frame
    imgs = {}
    for mode in (0, 1):
        layer.SetShadowTrees(mode)
        layer.LoadScene(scn)
        layer.InitPathTracing(3)
        layer.TracingPass(MISPT, 8)
        imgs[mode] = layer.GetHDRImage()[..., :3]
    layer.SetShadowTrees(0)
    assert imgs[1].mean() < imgs[0].mean() and _rel_rmse(imgs[1], imgs[0]) > 1e-3


def test_scene_uploads_after_init_are_validated_again(layer):
    """lights / materials / globals re-uploaded after hc_pt_init must be validated before the next pass, and a material that gains a
    normal map must select the normal-mapping shade kernel - without resetting generators or the framebuffer."""
    import hydracore_b200 as hc
    from tests import scenes
    plain, nm = scenes.cornell(64, 64), scenes.cornell_normal_mapped(64, 64)
    # an unsupported material class slipped in after init is refused at the next pass
    layer.LoadScene(plain)
    layer.InitPathTracing(5)
    layer.TracingPass(MISPT, 1)
    bad = plain.storages["materials"].copy()
    bad.view(np.int32).reshape(-1)[0:192][int(hc.layout.C["PLAIN_MAT_TYPE_OFFSET"])] = 999
    layer.UploadStorage("materials", bad)
    with pytest.raises(hc.HcError, match="material class 999"):
        layer.TracingPass(MISPT, 1)
    # a scene swapped in after init (same screen size) renders like a freshly initialised one, normal maps included
    layer.LoadScene(nm)
    layer.InitPathTracing(5)
    layer.TracingPass(MISPT, 2)
    want = layer.GetHDRImage()
    layer.LoadScene(plain)
    layer.InitPathTracing(5)
    for name in ("textures", "textures_aux", "geom", "materials", "pdfs"):
        layer.UploadStorage(name, nm.storages[name])
    layer.SetAllBVH4(nm.bvh["nodes"], nm.bvh["tris"])
    layer.SetAllInstMatrices(nm.bvh["inv_matrices"])
    layer.SetAllInstLightInstId(nm.inst_light_ids)
    layer.PrepareEngineGlobals(nm.globals_blob)
    layer.TracingPass(MISPT, 2)
    assert np.array_equal(want, layer.GetHDRImage())


def test_hdr_read_back_is_normalised_on_the_device(layer):
    from tests import scenes
    layer.LoadScene(scenes.cornell(64, 64))
    layer.InitPathTracing(11)
    layer.TracingPass(MISPT, 3)
    s, h = layer.GetSumImage(), layer.GetHDRImage()
    assert np.array_equal(h, s*np.float32(1.0/3.0))


def test_single_level_tree_vs_reference(layer, ref):
    """bvhType "triangle4v": no instance level, world-space triangles that carry their own instance id (BVH4Traverse, ctrace.h:669-838, leaf test
    IntersectAllPrimitivesInLeaf1 :63-122).  The tree is made from the mesh sub-tree of a one-instance scene: its root quad copied to quad 1
    (child offsets are absolute quad indices), an instance id written into every triangle's third vertex."""
    from hydracore_b200 import scene as S
    from tests import scenes
    scn = S.scene_c2(320, 180)
    nodes = np.ascontiguousarray(scn.bvh["nodes"], np.float32).copy().reshape(-1, 8)
    tris = np.ascontiguousarray(scn.bvh["tris"], np.float32).copy().reshape(-1, 4)
    un = nodes.view(np.uint32)
    top = un[4:8]                                        # quad 1: the top level over the single instance
    leaf = [c for c in range(4) if not (top[c, 3] == 0xFFFFFFFF and top[c, 7] == 0xFFFFFFFF)]
    assert len(leaf) == 1 and (top[leaf[0], 3] & 0x80000000)
    rec = int(top[leaf[0], 3] & 0x7FFFFFFF)              # instance record quad; its node 0 points at the mesh sub-tree
    sub = int(un[4*rec, 3])
    assert not (sub & 0x80000000)
    nodes[4:8] = nodes[4*sub:4*sub + 4]
    ti = tris.view(np.int32)
    hdr = np.nonzero((ti[:, 2] == -1) & (ti[:, 3] == -1) & (ti[:, 1] > 0) & (ti[:, 1] < 64))[0]          # leaf headers {first, count, -1, -1}
    for h in hdr:
        first, count = int(ti[h, 0]), int(ti[h, 1])
        if first == h + 1:
            ti[first + 2:first + 3*count:3, 3] = 7       # instId in the .w of every triangle's C vertex
    layer.SetAllBVH4(nodes, tris, have_inst=False)
    layer.SetAllInstMatrices(scn.bvh["inv_matrices"])
    layer.ResizeScreen(320, 180)
    layer.PrepareEngineGlobals(scn.globals_blob)
    rays = np.concatenate([layer.MakeEyeRays(320, 180, None), scenes.incoherent_rays(40000, 5, radius=25.0)])
    got = layer.TraceClosest(rays)
    want = np.zeros(rays.shape[0], got.dtype)
    ref.L.ref_trace_closest(nodes.ctypes.data_as(__import__("ctypes").c_void_p), tris.ctypes.data_as(__import__("ctypes").c_void_p), 0,
                            rays.ctypes.data_as(__import__("ctypes").c_void_p), __import__("ctypes").c_longlong(rays.shape[0]), want.ctypes.data_as(__import__("ctypes").c_void_p))
    hit = want["primId"] >= 0
    assert hit.mean() > 0.3 and np.all(want["instId"][hit] == 7)
    exact = (got["primId"] == want["primId"]) & (got["instId"] == want["instId"]) & (got["geomId"] == want["geomId"]) & (got["t"] == want["t"])
    ties = ~exact & (got["primId"] >= 0) & hit & (np.abs(got["t"] - want["t"]) <= 1e-5*np.abs(want["t"]))
    assert int((~exact & ~ties).sum()) == 0 and int((~exact).sum()) <= 3
    sh = rays.copy()
    sh[:, 7] = 12.0
    vis = layer.TraceShadow(sh)
    want_vis = ~(hit & (want["t"] > 0) & (want["t"] < 12.0))
    assert (vis.astype(bool) != want_vis).sum() <= 2
