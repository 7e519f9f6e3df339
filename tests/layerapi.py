"""ctypes binding of the IHWLayer test harness (hydracore_b200/cpp/layer_harness.cpp): every call goes through the reference's
IHWLayer / IMemoryStorage virtual interface of the C++ GPUCUDALayer, the way RenderDriverRTE drives a layer."""
import ctypes as ct
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hydracore_b200", "cpp", "_build", "libhydra_cuda_layer.so")
HRT_UNIFIED_IMAGE_SAMPLING = None     # filled from tests/golden/ref_consts.json by callers


def P(a):
    return a.ctypes.data_as(ct.c_void_p)


class LayerError(RuntimeError):
    pass


class CppLayer:
    """One GPUCUDALayer behind an IHWLayer*; method names are the reference's."""

    @staticmethod
    def available():
        return os.path.exists(LIB)

    def __init__(self, w, h, flags=0, device=0):
        self._L = ct.CDLL(LIB)
        L = self._L
        L.hl_last_error.restype = ct.c_char_p
        L.hl_create.restype = ct.c_void_p
        L.hl_create.argtypes = [ct.c_int]*4
        L.hl_get_spp.restype = ct.c_float
        L.hl_get_spp.argtypes = [ct.c_void_p]
        L.hl_get_spp_contrib.restype = ct.c_float
        L.hl_get_spp_contrib.argtypes = [ct.c_void_p]
        L.hl_shared_image_create.restype = ct.c_void_p
        L.hl_shared_image_create.argtypes = [ct.c_int, ct.c_int]
        L.hl_shared_image_destroy.argtypes = [ct.c_void_p]
        L.hl_shared_image_read.argtypes = [ct.c_void_p, ct.c_void_p, ct.POINTER(ct.c_float), ct.POINTER(ct.c_int)]
        L.hl_contribute.argtypes = [ct.c_void_p, ct.c_void_p]
        L.hl_available_memory.restype = ct.c_uint64
        L.hl_available_memory.argtypes = [ct.c_void_p, ct.c_int]
        for name in ("hl_destroy", "hl_create_storage", "hl_storage_update", "hl_storage_info", "hl_resize_tables", "hl_set_bvh", "hl_set_bvh2", "hl_set_remap", "hl_set_instances",
                     "hl_set_lights", "hl_set_camera", "hl_get_vars", "hl_set_vars", "hl_prepare", "hl_globals_blob", "hl_call", "hl_init_path_tracing",
                     "hl_passes", "hl_clear_accumulated", "hl_get_hdr", "hl_get_ldr", "hl_device_name", "hl_device_count", "hl_rays_stat",
                     "hl_store_cpu_data"):
            getattr(L, name).restype = None if name == "hl_destroy" else ct.c_int
        self._s = None
        s = L.hl_create(int(w), int(h), int(flags), int(device))
        if not s:
            raise LayerError("CreateCudaImpl failed: " + L.hl_last_error().decode(errors="replace"))
        self._s = ct.c_void_p(s)
        self.width, self.height = int(w), int(h)

    def close(self):
        if getattr(self, "_s", None):
            self._L.hl_destroy(self._s)
            self._s = None

    __del__ = close

    def _ck(self, rc, what):
        if rc != 0:
            raise LayerError(what + ": " + self._L.hl_last_error().decode(errors="replace"))

    def CreateMemStorage(self, name, max_bytes):
        self._ck(self._L.hl_create_storage(self._s, name.encode(), ct.c_uint64(int(max_bytes))), "CreateMemStorage")

    def StorageUpdate(self, name, obj_id, blob):
        b = np.ascontiguousarray(blob).view(np.uint8).reshape(-1)
        off = ct.c_int(-2)
        self._ck(self._L.hl_storage_update(self._s, name.encode(), int(obj_id), P(b), ct.c_uint64(b.size), ct.byref(off)), "IMemoryStorage::Update")
        return off.value

    def StorageInfo(self, name):
        sz, cap, mir = ct.c_uint64(), ct.c_uint64(), ct.c_int()
        self._ck(self._L.hl_storage_info(self._s, name.encode(), ct.byref(sz), ct.byref(cap), ct.byref(mir)), "storage info")
        return sz.value, cap.value, bool(mir.value)

    def ResizeTablesForEngineGlobals(self, geom, img, mat, light):
        self._ck(self._L.hl_resize_tables(self._s, geom, img, mat, light), "ResizeTablesForEngineGlobals")

    def SetAllBVH4(self, nodes, tris, bvh_type=b"object"):
        n = np.ascontiguousarray(nodes, np.float32).reshape(-1, 8)
        t = np.ascontiguousarray(tris, np.float32).reshape(-1, 4)
        self._keep_type = ct.c_char_p(bvh_type)
        self._ck(self._L.hl_set_bvh(self._s, P(n), n.shape[0], P(t), t.shape[0], self._keep_type), "SetAllBVH4")

    def SetAllBVH4TwoTrees(self, nodes, tris, nodes1, tris1, alpha1, bvh_type=b"object"):
        """ConvertionResult with treesNum = 2: tree 1 holds the meshes with opacity maps and carries pTriangleAlpha."""
        a = [np.ascontiguousarray(nodes, np.float32).reshape(-1, 8), np.ascontiguousarray(tris, np.float32).reshape(-1, 4),
             np.ascontiguousarray(nodes1, np.float32).reshape(-1, 8), np.ascontiguousarray(tris1, np.float32).reshape(-1, 4),
             np.ascontiguousarray(alpha1, np.uint32).reshape(-1, 2)]
        self._keep_type = ct.c_char_p(bvh_type)
        self._ck(self._L.hl_set_bvh2(self._s, P(a[0]), a[0].shape[0], P(a[1]), a[1].shape[0], P(a[2]), a[2].shape[0], P(a[3]), a[3].shape[0],
                                     P(a[4]), a[4].shape[0], self._keep_type), "SetAllBVH4")

    def SetAllRemapLists(self, all_lists, table, inst_remap_ids):
        a = [np.ascontiguousarray(all_lists, np.int32).reshape(-1), np.ascontiguousarray(table, np.int32).reshape(-1, 2),
             np.ascontiguousarray(inst_remap_ids, np.int32).reshape(-1)]
        self._ck(self._L.hl_set_remap(self._s, P(a[0]), a[0].size, P(a[1]), a[1].shape[0], P(a[2]), a[2].size), "SetAllRemapLists")

    def SetAllInstances(self, inv_matrices, light_ids):
        m = np.ascontiguousarray(inv_matrices, np.float32).reshape(-1, 16)
        li = np.ascontiguousarray(light_ids, np.int32).reshape(-1)
        self._ck(self._L.hl_set_instances(self._s, P(m), P(li), m.shape[0]), "SetAllInstMatrices / SetAllInstLightInstId")

    def SetAllLights(self, lights, select_table):
        l = np.ascontiguousarray(lights, np.float32).reshape(-1, 128)
        t = np.ascontiguousarray(select_table, np.float32).reshape(-1)
        self._ck(self._L.hl_set_lights(self._s, P(l), l.shape[0], P(t), t.size), "SetAllPODLights / SetAllLightsSelectTable")

    def SetCamMatrices(self, proj_inv, wv_inv, proj, wv, aspect, fov_x, look_at):
        a = [np.ascontiguousarray(x, np.float32).reshape(16) for x in (proj_inv, wv_inv, proj, wv)]
        la = np.ascontiguousarray(look_at, np.float32).reshape(3)
        self._ck(self._L.hl_set_camera(self._s, P(a[0]), P(a[1]), P(a[2]), P(a[3]), ct.c_float(aspect), ct.c_float(fov_x), P(la)), "SetCamMatrices")

    def GetAllFlagsAndVars(self):
        vi, vf, fl = np.zeros(64, np.int32), np.zeros(64, np.float32), ct.c_uint()
        self._ck(self._L.hl_get_vars(self._s, P(vi), P(vf), ct.byref(fl)), "GetAllFlagsAndVars")
        return vi, vf, fl.value

    def SetAllFlagsAndVars(self, varsI, varsF, flags):
        vi, vf = np.ascontiguousarray(varsI, np.int32), np.ascontiguousarray(varsF, np.float32)
        self._ck(self._L.hl_set_vars(self._s, P(vi), P(vf), ct.c_uint(int(flags))), "SetAllFlagsAndVars")

    def PrepareEngineGlobalsAndTables(self):
        self._ck(self._L.hl_prepare(self._s), "PrepareEngineGlobals / PrepareEngineTables")

    def EngineGlobalsBlob(self):
        n = ct.c_int()
        self._ck(self._L.hl_globals_blob(self._s, None, 0, ct.byref(n)), "GetEngineGlobals")
        out = np.zeros(n.value, np.int32)
        self._ck(self._L.hl_globals_blob(self._s, P(out), n.value, ct.byref(n)), "GetEngineGlobals")
        return out

    def CallNamedFunc(self, name, args):
        self._ck(self._L.hl_call(self._s, name.encode(), args.encode()), "CallNamedFunc")

    def CommIdHex(self):
        """rank 0 of a multi-process render: the NCCL unique id as 256 hex digits (CallNamedFunc("comm_id"))."""
        buf = ct.create_string_buffer(257)
        self._L.hl_comm_id_hex.restype = ct.c_int
        self._ck(self._L.hl_comm_id_hex(self._s, buf), "CallNamedFunc(comm_id)")
        return buf.value.decode()

    def InitPathTracing(self, seed):
        self._ck(self._L.hl_init_path_tracing(self._s, int(seed)), "InitPathTracing")

    def TracingPasses(self, n=1):
        self._ck(self._L.hl_passes(self._s, int(n)), "BeginTracingPass / EndTracingPass")

    def ClearAccumulatedColor(self):
        self._ck(self._L.hl_clear_accumulated(self._s), "ClearAccumulatedColor")

    def GetHDRImage(self):
        out = np.empty((self.height, self.width, 4), np.float32)
        self._ck(self._L.hl_get_hdr(self._s, P(out), self.width, self.height), "GetHDRImage")
        return out

    def GetLDRImage(self):
        out = np.empty((self.height, self.width), np.uint32)
        self._ck(self._L.hl_get_ldr(self._s, P(out), self.width, self.height), "GetLDRImage")
        return out

    def GetSPP(self):
        return float(self._L.hl_get_spp(self._s))

    def GetSPPContrib(self):
        return float(self._L.hl_get_spp_contrib(self._s))

    def ContribToExternalImageAccumulator(self, shared):
        self._ck(self._L.hl_contribute(self._s, shared.h), "ContribToExternalImageAccumulator")

    def GetDeviceName(self):
        buf = ct.create_string_buffer(256)
        self._ck(self._L.hl_device_name(self._s, buf, 256), "GetDeviceName")
        return buf.value.decode()

    def DeviceCount(self):
        return int(self._L.hl_device_count(self._s))

    def GetRaysStat(self):
        a, b, c = ct.c_float(), ct.c_float(), ct.c_int()
        self._ck(self._L.hl_rays_stat(self._s, ct.byref(a), ct.byref(b), ct.byref(c)), "GetRaysStat")
        return dict(raysPerSec=a.value, samplesPerSec=b.value, traceTimePerCent=c.value)

    def GetAvaliableMemoryAmount(self, all_mem=False):
        return int(self._L.hl_available_memory(self._s, 1 if all_mem else 0))

    def StoreCPUData(self):
        return bool(self._L.hl_store_cpu_data(self._s))


def load_scene_like_render_driver(lay, scn, consts):
    """Feed a hydracore_b200.scene.Scene to an IHWLayer the way RenderDriverRTE does: per-object IMemoryStorage::Update calls (ids -> tables),
    SetAllBVH4, instances, lights + selection table, camera, vars, PrepareEngineGlobals/Tables."""
    from hydracore_b200 import scene as S
    C = consts
    for name in ("textures", "textures_aux", "geom", "materials", "pdfs"):
        lay.CreateMemStorage(name, max(4096, 2*scn.storages[name].size + 4096))
    nl = len(scn.lights)
    lay.ResizeTablesForEngineGlobals(len(scn.meshes), len(scn.textures) + 1, len(scn.material_ids), max(nl, 1))
    for gid, m in enumerate(scn.meshes):
        lay.StorageUpdate("geom", gid, m.pack())
    heads = list(scn.material_ids) + [len(scn.materials)]
    for mid in range(len(scn.material_ids)):
        nodes = np.stack(scn.materials[heads[mid]:heads[mid + 1]]).astype(np.float32)
        lay.StorageUpdate("materials", mid, nodes)
    for k, t in enumerate(scn.textures):
        h, w = t.shape[0], t.shape[1]
        body = t.reshape(-1)
        chunk = np.concatenate([np.array([w, h, 4, 4], np.int32).view(np.uint8), body])
        lay.StorageUpdate("textures", k + 1, chunk)
    for tid in sorted(getattr(scn, "aux_ids", ())):           # normal maps live in "textures_aux" under the same id (RenderDriverRTE::UpdateImageAux)
        t = scn.textures[tid - 1]
        lay.StorageUpdate("textures_aux", tid, np.concatenate([np.array([t.shape[1], t.shape[0], 4, 4], np.int32).view(np.uint8), t.reshape(-1)]))
    for k, t in enumerate(getattr(scn, "pdf_tables", [])):
        lay.StorageUpdate("pdfs", k, t)
    if getattr(scn, "bvh1", None) is not None:
        lay.SetAllBVH4TwoTrees(scn.bvh["nodes"], scn.bvh["tris"], scn.bvh1["nodes"], scn.bvh1["tris"], scn.bvh1["alpha"])
    else:
        lay.SetAllBVH4(scn.bvh["nodes"], scn.bvh["tris"])
    if getattr(scn, "remap_lists", None):            # RenderDriverRTE::BeginScene sends the lists, EndScene the per-instance ids (RenderDriverRTE.cpp:1340-1376, 1478)
        lay.SetAllRemapLists(*scn.remap_arrays())
    lay.SetAllInstances(scn.bvh["inv_matrices"], scn.inst_light_ids)
    cam = scn.camera
    W, H = scn.width, scn.height
    aspect = float(W)/float(H)
    proj = S.perspective(cam.fov, aspect, cam.near, cam.far)
    view = S.look_at(cam.pos, cam.look_at, cam.up)
    fov_rad = np.float32(np.pi/180.0)*np.float32(cam.fov)
    varsI, varsF = scn.varsI.copy(), scn.varsF.copy()
    varsF[C["HRT_CAM_FOV"]] = fov_rad
    varsF[C["HRT_DOF_FOCAL_PLANE_DIST"]] = np.linalg.norm(np.asarray(cam.pos, np.float64) - np.asarray(cam.look_at, np.float64))
    varsI[C["HRT_ENABLE_DOF"]] = 1 if cam.dof else 0
    varsF[C["HRT_DOF_LENS_RADIUS"]] = cam.lens_radius if cam.dof else 0.0
    varsF[18:22] = scn.bsphere                       # HRT_BSPHERE_* (RenderDriverRTE.cpp:1483-1486)
    lay.SetAllFlagsAndVars(varsI, varsF, scn.flags | C["HRT_UNIFIED_IMAGE_SAMPLING"])
    lay.SetCamMatrices(S._cols(np.linalg.inv(proj)), S._cols(np.linalg.inv(view)), S._cols(proj), S._cols(view), aspect, float(fov_rad), cam.look_at)
    if nl > 0:
        lights = np.stack(scn.lights).astype(np.float32)
        lights[:, C["PLIGHT_PICK_PROB_FWD"]] = np.float32(1.0)/np.float32(nl)
        lights[:, C["PLIGHT_PICK_PROB_REV"]] = np.float32(1.0)/np.float32(nl)
        pref = np.zeros(nl + 1, np.float32)
        acc = np.float32(0)
        for i in range(nl):
            pref[i] = acc
            acc = np.float32(acc + np.float32(1.0)/np.float32(nl))
        pref[nl] = acc
        lay.SetAllLights(lights, pref)
    lay.PrepareEngineGlobalsAndTables()


class SharedImage:
    """In-process stand-in for HydraAPI's IHRSharedAccumImage (hydracore_b200/cpp/layer_harness.cpp)."""

    def __init__(self, w, h):
        self._L = ct.CDLL(LIB)
        self._L.hl_shared_image_create.restype = ct.c_void_p
        self._L.hl_shared_image_create.argtypes = [ct.c_int, ct.c_int]
        self._L.hl_shared_image_destroy.argtypes = [ct.c_void_p]
        self._L.hl_shared_image_read.argtypes = [ct.c_void_p, ct.c_void_p, ct.POINTER(ct.c_float), ct.POINTER(ct.c_int)]
        self.w, self.h_ = w, h
        self.h = ct.c_void_p(self._L.hl_shared_image_create(w, h))

    def read(self):
        out = np.zeros((self.h_, self.w, 4), np.float32)
        spp, cnt = ct.c_float(), ct.c_int()
        locked = self._L.hl_shared_image_read(self.h, P(out), ct.byref(spp), ct.byref(cnt))
        return out, spp.value, cnt.value, bool(locked)

    def close(self):
        if self.h:
            self._L.hl_shared_image_destroy(self.h)
            self.h = None
