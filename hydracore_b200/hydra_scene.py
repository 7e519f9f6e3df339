"""Reader for Hydra scene libraries (statex_*.xml + data/chunk_*.vsgf / *.image4ub), enough for the reference's test scenes of the
hydra_app/tests/test_42 kind: Lambert (+texture), Phong, diffuse+reflection blends, emissive area lights, rect area lights, DOF camera.

It produces a hydracore_b200.scene.Scene, i.e. the blobs RenderDriverRTE would hand to an IHWLayer.  What it mirrors:
  * .vsgf header {u64 fileBytes; u32 vertNum; u32 indNum; u32 matNum; u32 flags} then pos4f, norm4f, (tan4f when flags & 1), uv2f, indices,
    matIndices (HydraAPI HydraVSGFExport; byte offsets are also listed in the XML, statex_00001.xml:112-119);
  * .image4ub = {u32 w; u32 h} + RGBA8 (statex_00001.xml:3-5);
  * material conversion to first order (PlainMaterialConverter.cpp:1502-1602): diffuse -> Lambert, diffuse + reflectivity ->
    BlendMask(S = Phong/GGX, D = Lambert, mask = reflection colour) without fresnel (REFLECTION_WEIGHT_IS_ONE), emission -> emissive;
  * lights: rect area light, intensity = colour * multiplier, size = half extents, instance_light matrix = position / rotation
    (PlainLightConverter.cpp:130-300); scene matrices are row-major with the translation in [3], [7], [11].
Both the CUDA layer and the oracle consume the SAME packed blobs, so parity does not depend on how faithful this conversion is; it
matters only for "the same scene as the reference renders"."""
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

from . import materials as M
from . import scene as S


def read_vsgf(path, offset=0):
    d = open(path, "rb").read()[offset:]
    _bytes, vert_num, ind_num, _mat_num, flags = struct.unpack("<QIIII", d[:24])
    o = 24
    pos = np.frombuffer(d, np.float32, vert_num*4, o).reshape(-1, 4); o += vert_num*16
    nrm = np.frombuffer(d, np.float32, vert_num*4, o).reshape(-1, 4); o += vert_num*16
    tan = None
    if flags & 1:
        tan = np.frombuffer(d, np.float32, vert_num*4, o).reshape(-1, 4); o += vert_num*16
    uv = np.frombuffer(d, np.float32, vert_num*2, o).reshape(-1, 2); o += vert_num*8
    idx = np.frombuffer(d, np.int32, ind_num, o).reshape(-1, 3); o += ind_num*4
    mat = np.frombuffer(d, np.int32, ind_num//3, o)
    return dict(pos=pos[:, :3].copy(), norm=nrm[:, :3].copy(), tan=None if tan is None else tan.copy(), uv=uv.copy(), idx=idx.copy(), mat=mat.copy())


def read_image4ub(path, offset=8):
    d = open(path, "rb").read()
    w, h = struct.unpack("<II", d[:8])
    return np.frombuffer(d, np.uint8, w*h*4, offset).reshape(h, w, 4).copy()


def _floats(text):
    return [float(x.rstrip("f")) for x in text.replace(",", " ").split()]


def _val(node, default=None):
    """Hydra XML writes scalars and colours either as element text or as a val="..." attribute."""
    if node is None:
        return default
    v = node.get("val")
    if v is None:
        v = (node.text or "").strip()
    return v if v != "" else default


def parse_library(xml_path, mesh_fallback_dirs=()):
    """-> plain dict description of the scene library (no packing yet): textures, materials, lights, camera, meshes, instances, settings."""
    base = os.path.dirname(os.path.abspath(xml_path))
    root = ET.fromstring("<root>" + open(xml_path, encoding="utf-8").read().split("?>", 1)[-1] + "</root>")
    out = dict(textures={}, materials={}, lights={}, meshes={}, instances=[], light_instances=[], settings={})
    for t in root.find("textures_lib"):
        loc = t.get("loc")
        if loc and int(t.get("bytesize", "0")) > 16 and os.path.exists(os.path.join(base, loc)):
            out["textures"][int(t.get("id"))] = read_image4ub(os.path.join(base, loc), int(t.get("offset", "8")))
    for m in root.find("materials_lib"):
        d = dict(name=m.get("name"), light_id=int(m.get("light_id", "-1")))
        dif, ref, emi = m.find("diffuse"), m.find("reflectivity"), m.find("emission")
        if dif is not None:
            tex = dif.find("texture")
            tex = dif.find("color/texture") if tex is None else tex
            d["diffuse"] = dict(color=_floats(_val(dif.find("color"), "0 0 0")), tex=int(tex.get("id")) if tex is not None else 0,
                                brdf=dif.get("brdf_type", "lambert"), roughness=float(_val(dif.find("roughness"), "0")))
        if ref is not None:
            col = ref.find("color")
            d["reflect"] = dict(color=_floats(_val(col, "0 0 0")), gloss=float(_val(ref.find("glossiness"), "1")),
                                brdf=ref.get("brdf_type", "phong"), fresnel=_val(ref.find("fresnel"), "0") == "1", ior=float(_val(ref.find("fresnel_ior"), "1.5")))
            ani = ref.find("anisotropy")                                          # ReflectiveMaterialFromHydraMtl, PlainMaterialConverter.cpp:1063-1083
            if ani is not None:
                d["reflect"].update(aniso=float(_val(ani, "0")), aniso_rot=float(ani.get("rot", "0")), aniso_flip=int(ani.get("flip_axis", "0")) == 1)
            d["reflect"]["brdf"] = {"GGX": "ggx", "TRGGX": "trggx"}.get(d["reflect"]["brdf"], d["reflect"]["brdf"])
        if emi is not None:
            col = emi.find("color")
            d["emission"] = _floats(_val(col, "0 0 0"))
        opa = m.find("opacity")
        if opa is not None and opa.find("texture") is not None:                      # cut-out map: the mesh goes into the alpha-tested BVH tree
            d["opacity_tex"] = int(opa.find("texture").get("id"))
        out["materials"][int(m.get("id"))] = d
    for l in root.find("lights_lib"):
        size, inten = l.find("size"), l.find("intensity")
        mult = float(_val(inten.find("multiplier"), "1").split()[0])          # pugixml's as_float reads the leading number of "30 30 30"
        col = inten.find("color")
        half = (float(size.get("half_length", "0")), float(size.get("half_width", "0"))) if size is not None else (0.0, 0.0)
        out["lights"][int(l.get("id"))] = dict(type=l.get("type"), shape=l.get("shape", "point"), half=half, radius=float(size.get("radius", "0")) if size is not None else 0.0,
                                                color=[c*mult for c in _floats(_val(col, "1 1 1"))], mat_id=int(l.get("mat_id", "-1")),
                                                textured=col is not None and col.find("texture") is not None, perez=l.find("perez") is not None)
        d = out["lights"][int(l.get("id"))]
        if l.get("type") == "directional" or l.get("distribution") == "directional":        # CreateDirectLightFromXmlNode, PlainLightConverter.cpp:840-855
            soft = float(_val(l.find("shadow_softness"), "0"))
            d.update(type="directional", shape="point", params=[float(size.get("inner_radius", "0")) if size is not None else 0.0,
                                                                 float(size.get("outer_radius", "0")) if size is not None else 0.0,
                                                                 float(_val(l.find("angle_radius"), "0")) if l.find("angle_radius") is not None else 0.25*soft])
        elif l.get("shape", "point") == "point" and l.get("distribution") == "spot":           # CreatePointSpotLightFromXmlNode, :895-906
            d.update(type="spot", params=[float(_val(l.find("falloff_angle"), "0")), float(_val(l.find("falloff_angle2"), "0")), 0.0])
    cam = root.find("cam_lib")[0]
    g = lambda n, dflt: _val(cam.find(n), dflt)
    out["camera"] = dict(fov=float(g("fov", "45")), near=float(g("nearClipPlane", "0.01")), far=float(g("farClipPlane", "100")), up=_floats(g("up", "0 1 0")),
                         pos=_floats(g("position", "0 0 1")), look_at=_floats(g("look_at", "0 0 0")), dof=int(g("enable_dof", "0")) == 1,
                         lens_radius=_floats(g("dof_lens_radius", "0"))[0])
    for m in root.find("geometry_lib"):
        p = os.path.join(base, m.get("loc"))
        if not os.path.exists(p):                         # chunk stripped from the repository: look for the source mesh by name
            for d in mesh_fallback_dirs:
                q = os.path.join(d, os.path.basename(m.get("name")))
                if os.path.exists(q):
                    p = q
                    break
        out["meshes"][int(m.get("id"))] = read_vsgf(p, 0)
    rs = root.find("render_lib")[0]
    for k in ("width", "height", "trace_depth", "diff_trace_depth", "qmc_variant"):
        if rs.find(k) is not None:
            out["settings"][k] = int(_val(rs.find(k), "0"))
    scn = root.find("scenes")[0]
    for e in scn:
        mtx = np.array(_floats(e.get("matrix")), np.float32).reshape(4, 4)
        if e.tag == "instance":
            out["instances"].append(dict(mesh_id=int(e.get("mesh_id")), matrix=mtx, light_id=int(e.get("light_id", "-1")), linst_id=int(e.get("linst_id", "-1"))))
        elif e.tag == "instance_light":
            out["light_instances"].append(dict(id=int(e.get("id")), light_id=int(e.get("light_id")), matrix=mtx))
    out["light_instances"].sort(key=lambda d: d["id"])
    return out


def build_scene(lib, width, height):
    """plain description -> packed Scene (hydracore_b200.scene.Scene)."""
    c = lib["camera"]
    scn = S.Scene(width, height, S.Camera(pos=tuple(c["pos"]), look_at=tuple(c["look_at"]), up=tuple(c["up"]), fov=c["fov"], near=c["near"], far=c["far"],
                                          dof=c["dof"], lens_radius=c["lens_radius"]))
    scn.set_trace_depth(lib["settings"].get("trace_depth", 5), lib["settings"].get("diff_trace_depth", 3))
    tex_map = {0: 0}
    for tid in sorted(lib["textures"]):
        tex_map[tid] = scn.add_texture_rgba8(lib["textures"][tid])
    # one PlainLight per <instance_light>, in instance order (the driver instantiates lights per scene instance, RenderDriverRTE.cpp:1499-1521)
    linst_map, light_map = {}, {}
    for li in lib["light_instances"]:
        l, mtx = lib["lights"][li["light_id"]], li["matrix"]
        if l["type"] == "area" and l["shape"] == "rect":
            idx = scn.add_light(M.area_light(tuple(mtx[:3, 3]), l["half"], tuple(l["color"]), rotation=mtx[:3, :3]))
        elif l["type"] == "area" and l["shape"] == "sphere":
            scale = float(np.linalg.norm(mtx[:3, :3] @ (np.ones(3, np.float32)/np.float32(np.sqrt(3.0)))))       # SphereLight::Transform, PlainLightConverter.cpp:480-488
            idx = scn.add_light(M.sphere_light(tuple(mtx[:3, 3]), l["radius"]*scale, tuple(l["color"])))
        elif l["type"] == "point" and l["shape"] == "point":
            idx = scn.add_light(M.point_light(tuple(mtx[:3, 3]), tuple(l["color"])))
        elif l["type"] == "spot":
            L = M.spot_light(tuple(mtx[:3, 3]), (0.0, -1.0, 0.0), tuple(l["color"]), l["params"][0], l["params"][1])
            L[5:8] = mtx[:3, :3] @ np.array([0.0, -1.0, 0.0], np.float32)            # Transform() rotates the axis without re-normalising it (:606-618)
            idx = scn.add_light(L)
        elif l["type"] == "directional":
            L = M.direct_light(tuple(mtx[:3, 3]), (0.0, -1.0, 0.0), tuple(l["color"]), l["params"][0], l["params"][1], l["params"][2])
            L[5:8] = mtx[:3, :3] @ np.array([0.0, -1.0, 0.0], np.float32)
            idx = scn.add_light(L)
        elif l["type"] == "sky" and not l.get("textured") and not l.get("perez"):
            # constant-colour environment: 2x2 uniform pdf table (RenderDriverRTE_PdfTables.cpp:534-544); a black one is never picked
            idx = scn.add_light(M.sky_light(tuple(l["color"]), scn.add_sky_pdf_table()))
        else:
            raise ValueError("light %d: type %s / shape %s is not supported yet (rect and sphere area lights, omni / spot point lights, directional lights, untextured sky domes are)" % (li["light_id"], l["type"], l["shape"]))
        linst_map[li["id"]] = idx
        light_map.setdefault(li["light_id"], idx)
    mat_map = {}
    for mid in range(max(lib["materials"]) + 1):
        d = lib["materials"].get(mid, dict(diffuse=dict(color=[0.5, 0.5, 0.5], tex=0)))
        if "emission" in d and (max(d["emission"]) > 1e-5 or d.get("light_id", -1) >= 0 or ("diffuse" not in d and "reflect" not in d)):
            nodes = M.emissive(tuple(d["emission"]), light_map.get(d.get("light_id", -1), -1))
        else:
            dif = d.get("diffuse", dict(color=[0, 0, 0], tex=0))
            if dif.get("brdf", "lambert") == "orennayar":
                lam = M.orennayar(tuple(dif["color"]), dif.get("roughness", 0.0), tex_id=tex_map.get(dif["tex"], 0))
            else:
                lam = M.lambert(tuple(dif["color"]), tex_id=tex_map.get(dif["tex"], 0))  # textures that were not shipped (dl="1") read as white
            if "reflect" in d and max(d["reflect"]["color"]) > 1e-5:
                r = d["reflect"]
                if r["brdf"] not in _BRDF_CODE:
                    raise ValueError("material %d: reflectivity brdf_type '%s' is not supported yet (phong, ggx, torranse_sparrow, beckmann, trggx are)" % (mid, r["brdf"]))
                if r["gloss"] >= 0.995:                            # an untextured glossiness this high becomes a perfect mirror (PlainMaterialConverter.cpp:1128)
                    top = M.mirror(tuple(r["color"]))
                elif r["brdf"] in ("beckmann", "trggx"):
                    top = {"beckmann": M.beckmann, "trggx": M.trggx}[r["brdf"]](tuple(r["color"]), r["gloss"], aniso=r.get("aniso", 0.0), rot=r.get("aniso_rot", 0.0),
                                                                                  flip=r.get("aniso_flip", False))
                else:
                    top = {"ggx": M.ggx, "torranse_sparrow": M.blinn, "phong": M.phong}[r["brdf"]](tuple(r["color"]), r["gloss"])
                nodes = M.blend(tuple(r["color"]), top, lam, fresnel=r["fresnel"], ior=r["ior"])
            else:
                nodes = lam
        if d.get("opacity_tex", 0) > 0 and tex_map.get(d["opacity_tex"], 0) > 0:
            nodes = M.with_opacity(nodes, tex_map[d["opacity_tex"]])
        mat_map[mid] = scn.add_material(nodes)
    mesh_map = {}
    for inst in lib["instances"]:
        k = inst["mesh_id"]
        if k not in mesh_map:
            m = lib["meshes"][k]
            mesh_map[k] = scn.add_mesh(S.Mesh(m["pos"], m["idx"], norm=m["norm"], uv=m["uv"], mat=np.array([mat_map.get(int(x), 0) for x in m["mat"]], np.int32)))
        lidx = linst_map.get(inst.get("linst_id", -1), light_map.get(inst["light_id"], -1))
        scn.add_instance(mesh_map[k], inst["matrix"], light_id=lidx)
    return scn.build()


_BRDF_CODE = {"phong": 0, "ggx": 1, "torranse_sparrow": 2, "beckmann": 3, "trggx": 4}     # column 10 of the fixture's material rows
_BRDF_NAME = {v: k for k, v in _BRDF_CODE.items()}


def _fixture_arrays(lib):
    a = {"camera": np.array([lib["camera"][k] if not isinstance(lib["camera"][k], (list, tuple)) else 0 for k in ("fov", "near", "far", "lens_radius")], np.float64),
         "camera_vec": np.array([lib["camera"]["pos"], lib["camera"]["look_at"], lib["camera"]["up"]], np.float64), "camera_dof": np.array([int(lib["camera"]["dof"])]),
         "settings": np.array([lib["settings"].get(k, -1) for k in ("width", "height", "trace_depth", "diff_trace_depth", "qmc_variant")], np.int64)}
    used_meshes = sorted({i["mesh_id"] for i in lib["instances"]})
    a["mesh_ids"] = np.array(used_meshes, np.int64)
    for k in used_meshes:
        m = lib["meshes"][k]
        for f in ("pos", "norm", "uv", "idx", "mat"):
            a["mesh%d_%s" % (k, f)] = m[f]
    a["tex_ids"] = np.array(sorted(lib["textures"]), np.int64)
    for t in lib["textures"]:
        a["tex%d" % t] = lib["textures"][t]
    mats = []
    for mid in sorted(lib["materials"]):
        d = lib["materials"][mid]
        dif, ref = d.get("diffuse"), d.get("reflect")
        mats.append([mid, d.get("light_id", -1)] + (dif["color"] + [dif["tex"]] if dif else [0, 0, 0, -1]) +
                    (ref["color"] + [ref["gloss"], float(_BRDF_CODE[ref["brdf"]]), 1.0 if ref["fresnel"] else 0.0, ref["ior"]] if ref else [0, 0, 0, -1, 0, 0, 0]) +
                    (d["emission"] if "emission" in d else [-1, -1, -1]) + [d.get("opacity_tex", -1)] +
                    ([ref.get("aniso", 0.0), ref.get("aniso_rot", 0.0), 1.0 if ref.get("aniso_flip", False) else 0.0] if ref else [0.0, 0.0, 0.0]))
    a["materials"] = np.array(mats, np.float64)
    shapes = {"rect": 0, "sphere": 1, "point": 2, "sky": 3, "spot": 4, "directional": 5}
    a["lights"] = np.array([[lid] + list(l["half"]) + l["color"] + [l["mat_id"], shapes[l["type"] if l["type"] in ("sky", "spot", "directional") else l["shape"]], l["radius"]] + list(l.get("params", [0.0, 0.0, 0.0])) for lid, l in sorted(lib["lights"].items())], np.float64)
    a["light_instances"] = np.array([[li["id"], li["light_id"]] for li in lib["light_instances"]], np.int64).reshape(-1, 2)
    a["light_matrices"] = np.array([li["matrix"] for li in lib["light_instances"]], np.float32).reshape(-1, 4, 4)
    a["instances"] = np.array([[i["mesh_id"], i["light_id"], i.get("linst_id", -1)] for i in lib["instances"]], np.int64)
    a["instance_matrices"] = np.array([i["matrix"] for i in lib["instances"]], np.float32)
    return a


def save_fixtures(libs, path):
    """Store parsed libraries {name: lib} in ONE compressed .npz, so that tests and the bench can rebuild the scenes where the reference tree
    does not exist.  Keys are "<name>::<field>"; arrays shared between scenes (the teapot, textures) are stored once and aliased."""
    import hashlib
    out, seen = {}, {}
    for name, lib in libs.items():
        for k, v in _fixture_arrays(lib).items():
            v = np.ascontiguousarray(v)
            key = name + "::" + k
            h = (v.dtype.str, v.shape, hashlib.sha1(v.tobytes()).hexdigest())
            if v.nbytes > 4096 and h in seen:
                out[key + "@"] = np.array(seen[h])
            else:
                out[key] = v
                seen[h] = key
    np.savez_compressed(path, **out)


class _Prefixed:
    def __init__(self, z, prefix):
        self.z, self.p = z, prefix

    def __getitem__(self, k):
        if self.p + k + "@" in self.z.files:
            return self.z[str(self.z[self.p + k + "@"])]
        return self.z[self.p + k]


def fixture_scenes(path):
    return sorted({k.split("::")[0] for k in np.load(path).files})


def load_fixture(path, scene):
    z = _Prefixed(np.load(path), scene + "::")
    cam = dict(zip(("fov", "near", "far", "lens_radius"), [float(x) for x in z["camera"]]))
    cam.update(pos=list(z["camera_vec"][0]), look_at=list(z["camera_vec"][1]), up=list(z["camera_vec"][2]), dof=bool(z["camera_dof"][0]))
    lib = dict(camera=cam, textures={int(t): z["tex%d" % t] for t in z["tex_ids"]}, materials={}, lights={}, meshes={}, instances=[], light_instances=[],
               settings={k: int(v) for k, v in zip(("width", "height", "trace_depth", "diff_trace_depth", "qmc_variant"), z["settings"]) if v >= 0})
    for k in z["mesh_ids"]:
        lib["meshes"][int(k)] = {f: z["mesh%d_%s" % (k, f)] for f in ("pos", "norm", "uv", "idx", "mat")}
    for r in z["materials"]:
        d = dict(light_id=int(r[1]))
        if r[5] >= 0:
            d["diffuse"] = dict(color=list(r[2:5]), tex=int(r[5]))
        if r[9] >= 0:
            d["reflect"] = dict(color=list(r[6:9]), gloss=float(r[9]), brdf=_BRDF_NAME[int(round(r[10]))], fresnel=r[11] > 0.5, ior=float(r[12]))
            if len(r) > 19:                                                       # fixtures written before the anisotropic lobes have 17 columns
                d["reflect"].update(aniso=float(r[17]), aniso_rot=float(r[18]), aniso_flip=r[19] > 0.5)
        if r[13] >= 0:
            d["emission"] = list(r[13:16])
        if len(r) > 16 and r[16] > 0:
            d["opacity_tex"] = int(r[16])
        lib["materials"][int(r[0])] = d
    shapes = {0: ("area", "rect"), 1: ("area", "sphere"), 2: ("point", "point"), 3: ("sky", "point"), 4: ("spot", "point"), 5: ("directional", "point")}
    for r in z["lights"]:
        ty, sh = shapes[int(r[7])]
        lib["lights"][int(r[0])] = dict(type=ty, shape=sh, half=(float(r[1]), float(r[2])), color=list(r[3:6]), mat_id=int(r[6]), radius=float(r[8]),
                                        params=[float(x) for x in r[9:12]] if len(r) >= 12 else [0.0, 0.0, 0.0])
    for r, mtx in zip(z["light_instances"], z["light_matrices"]):
        lib["light_instances"].append(dict(id=int(r[0]), light_id=int(r[1]), matrix=mtx))
    for r, mtx in zip(z["instances"], z["instance_matrices"]):
        lib["instances"].append(dict(mesh_id=int(r[0]), light_id=int(r[1]), linst_id=int(r[2]), matrix=mtx))
    return lib
