"""CPU: pins oracle/hydra_oracle.cpp (our restatement) against golden vectors produced by the reference's own code
(tests/golden/make_golden.py over oracle/_ref) and, when oracle/_ref is present, live against it."""
import os

import numpy as np
import pytest

from tests import refapi, scenes

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def samplers():
    return np.load(os.path.join(G, "samplers.npz"))


@pytest.fixture(scope="module")
def raycast():
    return np.load(os.path.join(G, "raycast.npz"))


def test_rng_known_answers(oracle, samplers):
    for i, s in enumerate(samplers["seeds"]):
        st = oracle.rng_init(int(s))
        assert np.array_equal(st, samplers["state0"][i]), f"RandomGenInit({s})"
        assert np.array_equal(oracle.rng_float4(st, 8), samplers["float4"][i])     # bit-exact
        assert np.array_equal(oracle.rng_float1(st, 8), samplers["float1"][i])
        assert np.array_equal(st, samplers["state1"][i])


def test_niederreiter_table_and_samples(oracle, samplers):
    t = oracle.qmc_table()
    assert np.array_equal(t, samplers["qmc_table"])
    for i, p in enumerate(samplers["qmc_pos"]):
        for d in range(11):
            assert oracle.qmc_sobol(p, d, t) == samplers["qmc_val"][i, d]


@pytest.mark.parametrize("name", ["pinhole", "dof"])
def test_eye_rays_match_reference(oracle, raycast, name):
    g = np.zeros(537552//4 + 64, np.int32)
    g[:243] = raycast["globals_" + name]
    W, H = 96, 72
    r = oracle.make_rand_eye_rays(g, W, H, scenes.pixel_grid(W, H), raycast["offs_" + name])
    assert np.array_equal(r, raycast["eye_" + name])                              # bit-exact, DOF included (double sin/cos)
    r2, xy = oracle.make_eye_rays_f4(g, raycast["lens_" + name])
    assert np.array_equal(r2, raycast["eyef4_" + name])
    assert np.array_equal(xy, raycast["eyef4_xy_" + name])


def test_closest_hit_matches_reference(oracle, raycast):
    nodes, tris = raycast["bvh_nodes"], raycast["bvh_tris"]
    for rays, key in ((raycast["rays_primary"], "hits_primary"), (scenes.incoherent_rays(6000, 11), "hits_incoherent")):
        h = oracle.trace_closest(nodes, tris, rays)
        assert np.array_equal(h.view(np.uint8), raycast[key].view(np.uint8)), key   # t, primId, instId, geomId bit-exact
    assert (raycast["hits_primary"]["primId"] >= 0).mean() > 0.3
    h1 = oracle.trace_closest(raycast["bvh1_nodes"], raycast["bvh1_tris"], raycast["rays_single_leaf"])
    assert np.array_equal(h1.view(np.uint8), raycast["hits_single_leaf"].view(np.uint8))


def test_shadow_matches_reference_and_anyhit_kernel(oracle, raycast):
    sh = scenes.incoherent_rays(6000, 11)
    sh[:, 7] = raycast["shadow_tfar"]
    v = oracle.trace_shadow(raycast["bvh_nodes"], raycast["bvh_tris"], sh)
    assert np.array_equal(v, raycast["vis_incoherent"])
    # the OpenCL layer's early-exit kernel (BVH4InstTraverseShadow) answers the same question as the CPU wrapper
    assert np.array_equal(raycast["vis_incoherent"], raycast["vis_incoherent_anyhit"])
    assert 0.05 < v.mean() < 0.95


def test_counters_are_consistent(oracle, raycast):
    rays = raycast["rays_primary"]
    h, cnt = oracle.trace_closest(raycast["bvh_nodes"], raycast["bvh_tris"], rays, count=True)
    q, l, t = (int(c) for c in cnt)
    assert q >= rays.shape[0] and l > 0 and t >= l and t <= 4*l       # every ray fetches quad 1; leaves hold 1..4 triangles


def test_empty_ray_set(oracle, raycast):
    h = oracle.trace_closest(raycast["bvh_nodes"], raycast["bvh_tris"], np.zeros((0, 8), np.float32))
    assert h.shape[0] == 0


def test_live_against_reference_build(oracle, ref):
    """When oracle/_ref is present (it always is in the build container and travels to the GPU box) re-check on fresh inputs."""
    rng = np.random.RandomState(99)
    for s in rng.randint(-2**31, 2**31 - 1, 50):
        a, b = oracle.rng_init(int(s)), ref.rng_init(int(s))
        assert np.array_equal(a, b)
        assert np.array_equal(oracle.rng_float4(a, 4), ref.rng_float4(b, 4))
    scn = scenes.instanced_geometry(64, 48, dof=True)
    xy = scenes.pixel_grid(64, 48)
    offs = (rng.rand(64*48, 4)*2 - 1).astype(np.float32)
    ra, rb = oracle.make_rand_eye_rays(scn.globals_blob, 64, 48, xy, offs), ref.make_rand_eye_rays(scn.globals_blob, 64, 48, xy, offs)
    assert np.array_equal(ra, rb)
    rays = np.concatenate([refapi.rays_from_pos_dir(ra), scenes.incoherent_rays(3000, 5)])
    ha, hb = oracle.trace_closest(scn.bvh["nodes"], scn.bvh["tris"], rays), ref.trace_closest(scn.bvh["nodes"], scn.bvh["tris"], rays)
    assert np.array_equal(ha.view(np.uint8), hb.view(np.uint8))
