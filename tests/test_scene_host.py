"""Host-side scene logic that mirrors RenderDriverRTE (no GPU needed): the alpha-test table, the split into an opaque and an alpha-tested BVH
tree with scene-wide instance ids, light pick probabilities, the aux texture table for normal maps."""
import numpy as np

from tests import scenes


def _instance_records(nodes):
    """{realInstId: meshId} of every instance record reachable from the top level of a flat two-level BVH4."""
    n = np.ascontiguousarray(nodes, np.float32).reshape(-1, 8)
    ni = n.view(np.int32)
    out, todo, seen = {}, [1], set()
    while todo:
        q = todo.pop()
        if q in seen:
            continue
        seen.add(q)
        for i in range(4):
            lo, esc = int(ni[4*q + i, 3]) & 0xFFFFFFFF, int(ni[4*q + i, 7]) & 0xFFFFFFFF
            if lo == 0xFFFFFFFF and esc == 0xFFFFFFFF:
                continue
            if lo & 0x80000000:
                rec = lo & 0x7FFFFFFF
                out[int(ni[4*rec + 3, 0])] = int(ni[4*rec + 3, 1])
            else:
                todo.append(lo)
    return out


def test_two_trees_and_alpha_table(built):
    from hydracore_b200.layout import C
    scn = scenes.cornell_with_cutout(32, 32)
    assert scn.bvh1 is not None
    # instances 0 (room), 2 (light quad) are opaque; 1, 3, 4 use meshes with opacity-mapped materials (scene order = instance id)
    assert sorted(_instance_records(scn.bvh["nodes"])) == [0, 2] and sorted(_instance_records(scn.bvh1["nodes"])) == [1, 3, 4]
    assert scn.bvh["inv_matrices"].shape == (5, 16) and np.isfinite(scn.bvh["inv_matrices"]).all()
    tris, alpha = scn.bvh1["tris"], scn.bvh1["alpha"]
    n = tris.shape[0]
    ti = tris.view(np.int32)
    assert alpha.dtype == np.uint32 and alpha.shape == (n + 2*6, 2)                      # two materials with opacity maps -> two samplers
    k, seen_samplers = 0, set()
    while k < n:
        if ti[k, 2] == -1 and ti[k, 3] == -1:                                              # leaf header
            assert tuple(alpha[k]) == (0xFFFFFFFF, 0xFFFFFFFF)
            k += 1
            continue
        off = int(alpha[k, 0])
        assert n <= off < n + 12 and (off - n) % 6 == 0
        seen_samplers.add(off)
        sampler = alpha[off:off + 6].reshape(-1).view(np.float32)
        assert sampler.view(np.int32)[2] == 1 and sampler[1] == 1.0                        # texture id 1, gamma 1
        prim, geom = int(ti[k, 3]), int(ti[k + 1, 3])
        mesh = scn.meshes[geom]
        for j in range(3):                                                                  # packed uv decodes to the (wrapped) vertex uv
            p = int(alpha[k + j, 1])
            u, v = 2.0*(p & 0xFFFF)/65535.0 - 1.0, 2.0*(p >> 16)/65535.0 - 1.0
            uv = mesh.uv[int(mesh.idx[prim, j])]
            wrap = lambda x: x - int(x) if x > 1.0 else (int(x) - x if x < -1.0 else x)
            assert abs(u - wrap(float(uv[0]))) < 1e-4 and abs(v - wrap(float(uv[1]))) < 1e-4
        k += 3
    assert len(seen_samplers) == 2
    # a scene without opacity maps has no second tree, and its instance ids are the builder's own
    plain = scenes.cornell(32, 32)
    assert plain.bvh1 is None and sorted(_instance_records(plain.bvh["nodes"])) == list(range(plain.bvh["inv_matrices"].shape[0]))
    assert C["OPACITY_SAMPLER_OFFSET"] % 4 == 0


def test_light_pick_probabilities_follow_the_driver():
    from hydracore_b200 import materials as M
    from hydracore_b200.scene import light_pick_probs
    from hydracore_b200.layout import C
    L = [M.area_light((0, 1, 0), (1, 1), (5, 5, 5)), M.point_light((0, 2, 0), (3, 3, 3)), M.area_light((0, 3, 0), (1, 1), (0.001, 0.001, 0.001)),
         M.sky_light((1, 1, 1), 0), M.sphere_light((1, 1, 1), 0.5, (2, 2, 2)), M.point_light((0, 2, 0), (3, 3, 3))]
    L = np.stack(L)
    li = L.view(np.int32)
    li[:, C["PLIGHT_GROUP_ID"]] = -1
    li[4, C["PLIGHT_GROUP_ID"]] = 7
    li[5, C["PLIGHT_GROUP_ID"]] = 7                                   # lights 4 and 5 share a group: 5 groups in all
    li[1, C["PLIGHT_FLAGS"]] |= C["LIGHT_DO_NOT_SAMPLE_ME"]
    L[0, C["PLIGHT_PROB_MULT"]] = 2.0
    rev, fwd = light_pick_probs(L, False), light_pick_probs(L, True)
    g = np.float32(1.0)/np.float32(5)
    assert np.array_equal(rev, np.array([g*np.float32(2), 0, 0, g, g/np.float32(2), g/np.float32(2)], np.float32))     # do-not-sample and black lights: 0
    assert fwd[3] == 0 and np.array_equal(np.delete(fwd, 3), np.delete(rev, 3))                                        # sky domes are never picked forward


def test_normal_maps_go_to_the_aux_storage(built):
    from hydracore_b200.layout import C
    scn = scenes.cornell_normal_mapped(32, 32)
    blob = scn.globals_blob
    aux_off, aux_size = blob[C["EG_texturesAuxTableOffset"]//4], blob[C["EG_texturesAuxTableSize"]//4]
    table = blob[aux_off:aux_off + aux_size]
    assert list(table[:2]) == [-1, 1]                                 # texture id 1 at float4 offset 1 of "textures_aux"
    hdr = scn.storages["textures_aux"][16:32].view(np.int32)
    assert list(hdr) == [64, 64, 4, 4]
    mats = scn.storages["materials"].view(np.float32).reshape(-1, 192)
    with_map = [int(m.view(np.int32)[C["NORMAL_TEX_OFFSET"]]) for m in mats]
    assert with_map.count(1) == 6                                     # floor, GGX, blend head + 2 children, glass
