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


def test_scene_reader_lights_and_materials_from_xml(tmp_path, built):
    """hydra_scene.parse_library / build_scene on a hand-written statex library: spot and directional lights, a black sky dome, an opacity
    map, a torranse_sparrow reflection layer and a mirror (glossiness 1) - the converter rules cited in hydra_scene.py."""
    import struct
    from hydracore_b200 import hydra_scene as HS
    from hydracore_b200.layout import C
    d = tmp_path / "lib"
    (d / "data").mkdir(parents=True)
    # one quad mesh in the vsgf layout (24-byte header, pos4f, norm4f, uv2f, indices, material indices)
    pos = np.array([[-1, 0, -1, 1], [1, 0, -1, 1], [1, 0, 1, 1], [-1, 0, 1, 1]], np.float32)
    nrm = np.array([[0, 1, 0, 0]]*4, np.float32)
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    idx = np.array([0, 2, 1, 0, 3, 2], np.int32)
    mat = np.array([1, 1], np.int32)
    body = pos.tobytes() + nrm.tobytes() + uv.tobytes() + idx.tobytes() + mat.tobytes()
    (d / "data" / "chunk_00001.vsgf").write_bytes(struct.pack("<QIIII", 24 + len(body), 4, 6, 2, 0) + body)
    body2 = (pos*np.float32([4, 1, 4, 1]) - np.float32([0, 1, 0, 0])).astype(np.float32).tobytes() + nrm.tobytes() + uv.tobytes() + idx.tobytes() + np.array([0, 2], np.int32).tobytes()
    (d / "data" / "chunk_00002.vsgf").write_bytes(struct.pack("<QIIII", 24 + len(body2), 4, 6, 2, 0) + body2)      # an opaque floor: grey + mirror
    img = np.full((4, 4, 4), 255, np.uint8)
    (d / "data" / "chunk_00000.image4ub").write_bytes(struct.pack("<II", 4, 4) + img.tobytes())
    xml = """<?xml version="1.0"?>
<textures_lib><texture id="0" name="white" /><texture id="1" name="mask" loc="data/chunk_00000.image4ub" width="4" height="4" bytesize="64" /></textures_lib>
<materials_lib>
  <material id="0" name="a" type="hydra_material"><diffuse brdf_type="lambert"><color val="0.5 0.5 0.5" /></diffuse></material>
  <material id="1" name="b" type="hydra_material"><diffuse brdf_type="lambert"><color val="0.2 0.6 0.2" /></diffuse>
     <reflectivity brdf_type="torranse_sparrow"><color val="0.4 0.4 0.4" /><glossiness val="0.7" /><fresnel val="1" /><fresnel_ior val="1.5" /></reflectivity>
     <opacity smooth="0"><texture id="1" type="texref" /></opacity></material>
  <material id="2" name="m" type="hydra_material"><diffuse brdf_type="lambert"><color val="0 0 0" /></diffuse>
     <reflectivity brdf_type="phong"><color val="0.8 0.8 0.8" /><glossiness val="1" /></reflectivity></material>
</materials_lib>
<lights_lib>
  <light id="0" name="env" type="sky" shape="point" distribution="uniform"><intensity><color val="0 0 0" /><multiplier val="1" /></intensity></light>
  <light id="1" name="spot" type="point" shape="point" distribution="spot"><intensity><color val="1 1 1" /><multiplier val="10 10 10" /></intensity>
     <falloff_angle val="100" /><falloff_angle2 val="60" /></light>
  <light id="2" name="sun" type="directional" shape="point" distribution="directional"><intensity><color val="1 0.9 0.8" /><multiplier val="2" /></intensity>
     <size inner_radius="3" outer_radius="5" /><shadow_softness val="2" /></light>
</lights_lib>
<cam_lib><camera id="0" name="c" type="uvn"><fov>45</fov><nearClipPlane>0.01</nearClipPlane><farClipPlane>100</farClipPlane>
   <up>0 1 0</up><position>0 3 6</position><look_at>0 0 0</look_at></camera></cam_lib>
<geometry_lib><mesh id="0" name="q" type="vsgf" bytesize="1" loc="data/chunk_00001.vsgf" vertNum="4" triNum="2" />
   <mesh id="1" name="floor" type="vsgf" bytesize="1" loc="data/chunk_00002.vsgf" vertNum="4" triNum="2" /></geometry_lib>
<render_lib><render_settings type="HydraModern" id="0"><width>64</width><height>48</height><trace_depth>4</trace_depth><diff_trace_depth>2</diff_trace_depth></render_settings></render_lib>
<scenes><scene id="0" name="s">
  <instance id="0" mesh_id="0" mmat_id="-1" matrix="1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1 " light_id="-1" />
  <instance id="1" mesh_id="1" mmat_id="-1" matrix="1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1 " light_id="-1" />
  <instance_light id="0" light_id="0" matrix="1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1 " lgroup_id="-1" />
  <instance_light id="1" light_id="1" matrix="1 0 0 0 0 1 0 4 0 0 1 0 0 0 0 1 " lgroup_id="-1" />
  <instance_light id="2" light_id="2" matrix="1 0 0 5 0 0 1 5 0 -1 0 0 0 0 0 1 " lgroup_id="-1" />
</scene></scenes>
"""
    (d / "statex_00001.xml").write_text(xml)
    lib = HS.parse_library(str(d / "statex_00001.xml"))
    assert lib["lights"][1]["type"] == "spot" and lib["lights"][1]["color"] == [10.0, 10.0, 10.0] and lib["lights"][1]["params"][:2] == [100.0, 60.0]
    assert lib["lights"][2]["type"] == "directional" and lib["lights"][2]["params"] == [3.0, 5.0, 0.5]      # angle_radius = 0.25 * shadow_softness
    assert lib["materials"][1]["opacity_tex"] == 1 and lib["materials"][1]["reflect"]["brdf"] == "torranse_sparrow"
    scn = HS.build_scene(lib, 64, 48)
    L = np.stack(scn.lights)
    types = [int(x) for x in L.view(np.int32)[:, C["PLIGHT_TYPE"]]]
    assert types == [C["PLAIN_LIGHT_TYPE_SKY_DOME"], C["PLAIN_LIGHT_TYPE_POINT_SPOT"], C["PLAIN_LIGHT_TYPE_DIRECT"]]
    assert np.allclose(L[1, 2:5], (0, 4, 0)) and np.allclose(L[1, 5:8], (0, -1, 0))                         # spot at the instance position, axis -Y
    assert np.allclose(L[2, 5:8], (0, 0, 1), atol=1e-6) or np.allclose(L[2, 5:8], (0, 0, -1), atol=1e-6)    # rotated axis of the directional light
    assert abs(L[1, 14] - np.cos(np.radians(30.0))) < 1e-6 and abs(L[1, 15] - np.cos(np.radians(50.0))) < 1e-6
    blob = scn.globals_blob
    assert blob[C["EG_skyLightId"]//4] == 0 and blob[C["EG_sunNumber"]//4] == 8                             # a soft directional light fills the sun slots
    sel = blob[blob[C["EG_lightSelectorTableOffsetRev"]//4]:][:4].view(np.float32)
    assert sel[0] == 0.0 and sel[1] == 0.0 and abs(sel[3] - 2.0/3.0) < 1e-6                                 # the black sky is never picked
    assert scn.bvh1 is not None                                                                             # the opacity-mapped quad went to the alpha-tested tree ...
    mats = scn.storages["materials"].view(np.int32).reshape(-1, 192)
    assert (mats[:, C["PLAIN_MAT_TYPE_OFFSET"]] == C["PLAIN_MAT_CLASS_BLINN_SPECULAR"]).sum() == 1          # ... with a Blinn layer
    assert (mats[:, C["PLAIN_MAT_TYPE_OFFSET"]] == C["PLAIN_MAT_CLASS_PERFECT_MIRROR"]).sum() == 1          # glossiness 1 -> mirror
