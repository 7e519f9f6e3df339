"""Thin Python mirror of the reference's IHWLayer / IBVHBuilder2 method names over the C ABI (tests and bench.py only).

Method names, argument meaning and error behaviour follow hydra_drv/IHWLayer.h:97-246 and IBVHBuilderAPI.h:35-68 so that the
parity tests read like a RenderDriverRTE session: CreateMemStorage/Update -> SetAllBVH4 -> SetAllInstMatrices ->
PrepareEngineGlobals -> InitPathTracing -> BeginTracingPass/EndTracingPass -> GetHDRImage.
"""
import ctypes as ct
import numpy as np

from ._lib import load, check, hc_stats, HC_HOST, HC_DEVICE, HcError

STORAGE_SLOTS = {"textures": 0, "textures_aux": 1, "geom": 2, "materials": 3, "pdfs": 4}   # RenderDriverRTE.cpp:705-709
INTEGRATOR_PT, INTEGRATOR_MISPT, INTEGRATOR_MISPT_QMC = 0, 2, 3


def _ptr(a):
    return a.ctypes.data_as(ct.c_void_p)


class BvhBuilder:
    """IBVHBuilder2 stand-in (InstanceTriangleMeshes / CommitScene / ConvertMap)."""

    def __init__(self):
        self._h = None
        self._L = load()
        h = ct.c_void_p()
        check(self._L.hc_bvh_create(ct.byref(h)), "hc_bvh_create")
        self._h = h

    def close(self):
        if self._h:
            self._L.hc_bvh_destroy(self._h)
            self._h = None

    __del__ = close

    def add_mesh(self, vert4f, indices):
        v = np.ascontiguousarray(vert4f, dtype=np.float32).reshape(-1, 4)
        i = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1)
        out = ct.c_int()
        check(self._L.hc_bvh_add_mesh(self._h, _ptr(v), v.shape[0], _ptr(i), i.size, ct.byref(out)), "hc_bvh_add_mesh")
        return out.value

    def add_instance(self, mesh_id, matrix_row_major, real_id=None):
        """real_id: the scene-wide instance id written into the instance record (one builder per BVH tree); default = index in this builder."""
        m = np.ascontiguousarray(matrix_row_major, dtype=np.float32).reshape(16)
        if real_id is not None:
            check(self._L.hc_bvh_add_instance_id(self._h, int(mesh_id), _ptr(m), int(real_id)), "hc_bvh_add_instance_id")
            return int(real_id)
        out = ct.c_int()
        check(self._L.hc_bvh_add_instance(self._h, int(mesh_id), _ptr(m), ct.byref(out)), "hc_bvh_add_instance")
        return out.value

    def commit(self):
        """CommitScene + ConvertMap: returns dict(nodes u8[n*32], tris f32[m,4], inv_matrices f32[k,16], max_stack)."""
        check(self._L.hc_bvh_commit(self._h), "hc_bvh_commit")
        nodes, tris, inv = ct.c_void_p(), ct.c_void_p(), ct.c_void_p()
        nn, nt, ni, ms = ct.c_int(), ct.c_int(), ct.c_int(), ct.c_int()
        check(self._L.hc_bvh_result(self._h, ct.byref(nodes), ct.byref(nn), ct.byref(tris), ct.byref(nt), ct.byref(inv), ct.byref(ni),
                                    ct.byref(ms)), "hc_bvh_result")
        nodes_np = np.ctypeslib.as_array(ct.cast(nodes, ct.POINTER(ct.c_float)), shape=(nn.value, 8)).copy()
        tris_np = np.ctypeslib.as_array(ct.cast(tris, ct.POINTER(ct.c_float)), shape=(nt.value, 4)).copy()
        inv_np = np.ctypeslib.as_array(ct.cast(inv, ct.POINTER(ct.c_float)), shape=(ni.value, 16)).copy()
        return dict(nodes=nodes_np, tris=tris_np, inv_matrices=inv_np, max_stack=ms.value)

    def bounds(self):
        lo = (ct.c_float*3)()
        hi = (ct.c_float*3)()
        check(self._L.hc_bvh_bounds(self._h, lo, hi), "hc_bvh_bounds")
        return np.array(lo[:], np.float32), np.array(hi[:], np.float32)


class CudaLayer:
    """The CUDA IHWLayer: one per device."""

    def __init__(self, w=0, h=0, device=0):
        self._c = None
        self._L = load()
        c = ct.c_void_p()
        check(self._L.hc_ctx_create(int(device), ct.byref(c)), "hc_ctx_create")
        self._c = c
        self.width, self.height = w, h
        if w > 0 and h > 0:
            self.ResizeScreen(w, h)

    def close(self):
        if getattr(self, "_c", None):
            self._L.hc_ctx_destroy(self._c)
            self._c = None

    __del__ = close

    # ---- device
    def GetDeviceName(self):
        buf = ct.create_string_buffer(256)
        check(self._L.hc_device_name(self._c, buf, 256), "hc_device_name")
        return buf.value.decode()

    def GetAvaliableMemoryAmount(self, allMem=False):
        f, t = ct.c_size_t(), ct.c_size_t()
        check(self._L.hc_mem_info(self._c, ct.byref(f), ct.byref(t)), "hc_mem_info")
        return t.value if allMem else f.value

    def FinishAll(self):
        check(self._L.hc_sync(self._c), "hc_sync")

    def stream(self):
        s = ct.c_void_p()
        check(self._L.hc_stream(self._c, ct.byref(s)), "hc_stream")
        return s.value or 0

    # ---- storages (MemoryStorageCUDA): whole-blob upload, the id->offset tables live in the globals blob
    def UploadStorage(self, name, blob):
        slot = STORAGE_SLOTS[name]
        b = np.ascontiguousarray(blob).view(np.uint8).reshape(-1)
        check(self._L.hc_storage_reserve(self._c, slot, max(b.size, 16)), "hc_storage_reserve")
        if b.size:
            check(self._L.hc_storage_write(self._c, slot, 0, _ptr(b), b.size), "hc_storage_write")

    # ---- scene
    def PrepareEngineGlobals(self, blob_i32):
        b = np.ascontiguousarray(blob_i32).view(np.uint8).reshape(-1)
        check(self._L.hc_set_globals(self._c, _ptr(b), b.size), "hc_set_globals")

    def SetAllBVH4(self, nodes, tris, have_inst=True, tree=0, alpha=None):
        n = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1, 8)
        t = np.ascontiguousarray(tris, dtype=np.float32).reshape(-1, 4)
        if alpha is not None:
            a = np.ascontiguousarray(alpha, dtype=np.uint32).reshape(-1, 2)
            check(self._L.hc_set_bvh_alpha(self._c, tree, _ptr(n), n.shape[0], _ptr(t), t.shape[0], _ptr(a), a.shape[0], 1 if have_inst else 0), "hc_set_bvh_alpha")
            return
        check(self._L.hc_set_bvh(self._c, tree, _ptr(n), n.shape[0], _ptr(t), t.shape[0], 1 if have_inst else 0), "hc_set_bvh")

    def SetAllInstMatrices(self, inv_matrices):
        m = np.ascontiguousarray(inv_matrices, dtype=np.float32).reshape(-1, 16)
        check(self._L.hc_set_inst_matrices(self._c, _ptr(m), m.shape[0]), "hc_set_inst_matrices")

    def SetAllRemapLists(self, all_lists, table, inst_remap_ids):
        """SetAllRemapLists + SetAllInstIdToRemapId (IHWLayer.h:122-123); None clears."""
        if all_lists is None or len(all_lists) == 0:
            check(self._L.hc_set_remap_lists(self._c, None, None, 0, 0), "hc_set_remap_lists")
            check(self._L.hc_set_inst_remap_ids(self._c, None, 0), "hc_set_inst_remap_ids")
            return
        a = np.ascontiguousarray(all_lists, dtype=np.int32).reshape(-1)
        t = np.ascontiguousarray(table, dtype=np.int32).reshape(-1, 2)
        r = np.ascontiguousarray(inst_remap_ids, dtype=np.int32).reshape(-1)
        check(self._L.hc_set_remap_lists(self._c, _ptr(a), _ptr(t), a.size, t.shape[0]), "hc_set_remap_lists")
        check(self._L.hc_set_inst_remap_ids(self._c, _ptr(r), r.size), "hc_set_inst_remap_ids")

    def SetAllInstLightInstId(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.int32).reshape(-1)
        check(self._L.hc_set_inst_light_ids(self._c, _ptr(a), a.size), "hc_set_inst_light_ids")

    def ResizeScreen(self, w, h):
        check(self._L.hc_resize(self._c, int(w), int(h)), "hc_resize")
        self.width, self.height = int(w), int(h)

    # ---- ray casting
    def MakeEyeRays(self, w, h, offsets4=None):
        rays = np.empty((w*h, 8), np.float32)
        o = None if offsets4 is None else np.ascontiguousarray(offsets4, dtype=np.float32).reshape(-1, 4)
        check(self._L.hc_make_eye_rays(self._c, w, h, None if o is None else _ptr(o), _ptr(rays), HC_HOST), "hc_make_eye_rays")
        return rays

    def TraceClosest(self, rays8):
        r = np.ascontiguousarray(rays8, dtype=np.float32).reshape(-1, 8)
        hits = np.empty(r.shape[0], dtype=HIT_DTYPE)
        check(self._L.hc_trace_closest(self._c, _ptr(r), r.shape[0], _ptr(hits), HC_HOST), "hc_trace_closest")
        return hits

    def TraceShadow(self, rays8):
        r = np.ascontiguousarray(rays8, dtype=np.float32).reshape(-1, 8)
        vis = np.empty(r.shape[0], dtype=np.uint8)
        check(self._L.hc_trace_shadow(self._c, _ptr(r), r.shape[0], _ptr(vis), HC_HOST), "hc_trace_shadow")
        return vis

    def trace_closest_device(self, rays_ptr, n, hits_ptr):
        check(self._L.hc_trace_closest(self._c, ct.c_void_p(rays_ptr), n, ct.c_void_p(hits_ptr), HC_DEVICE), "hc_trace_closest")

    def trace_shadow_device(self, rays_ptr, n, vis_ptr):
        check(self._L.hc_trace_shadow(self._c, ct.c_void_p(rays_ptr), n, ct.c_void_p(vis_ptr), HC_DEVICE), "hc_trace_shadow")

    def make_eye_rays_device(self, w, h, rays_ptr, offsets_ptr=None):
        check(self._L.hc_make_eye_rays(self._c, w, h, ct.c_void_p(offsets_ptr) if offsets_ptr else None, ct.c_void_p(rays_ptr), HC_DEVICE),
              "hc_make_eye_rays")

    def MakeShadowRays(self, rays8, hits, light_pos):
        r = np.ascontiguousarray(rays8, dtype=np.float32).reshape(-1, 8)
        out = np.empty_like(r)
        lp = (ct.c_float*3)(*[float(v) for v in light_pos])
        check(self._L.hc_make_shadow_rays(self._c, _ptr(r), _ptr(np.ascontiguousarray(hits)), r.shape[0], lp, _ptr(out), HC_HOST), "hc_make_shadow_rays")
        return out

    def make_shadow_rays_device(self, rays_ptr, hits_ptr, n, light_pos, out_ptr):
        lp = (ct.c_float*3)(*[float(v) for v in light_pos])
        check(self._L.hc_make_shadow_rays(self._c, ct.c_void_p(rays_ptr), ct.c_void_p(hits_ptr), n, lp, ct.c_void_p(out_ptr), HC_DEVICE), "hc_make_shadow_rays")

    def RaycastPass(self, light_pos, hits_ptr=None, vis_ptr=None, space=HC_HOST):
        """K1 -> K2 -> shadow rays -> K2s over the whole screen; optional result pointers (host or device per `space`)."""
        lp = (ct.c_float*3)(*[float(v) for v in light_pos])
        check(self._L.hc_raycast_pass(self._c, lp, ct.c_void_p(hits_ptr) if hits_ptr else None, ct.c_void_p(vis_ptr) if vis_ptr else None, space),
              "hc_raycast_pass")

    def measure_read_bandwidth(self, nbytes, repeats):
        g = ct.c_float()
        check(self._L.hc_measure_read_bandwidth(self._c, int(nbytes), int(repeats), ct.byref(g)), "hc_measure_read_bandwidth")
        return g.value

    def last_trace_ms(self):
        ms = ct.c_float()
        check(self._L.hc_trace_last_ms(self._c, ct.byref(ms)), "hc_trace_last_ms")
        return ms.value

    # ---- path tracing
    def InitPathTracing(self, seed):
        check(self._L.hc_pt_init(self._c, int(seed)), "hc_pt_init")

    def SetTiles(self, tile, rank, world):
        check(self._L.hc_pt_set_tiles(self._c, int(tile), int(rank), int(world)), "hc_pt_set_tiles")

    def SetSampleStreams(self, streams, max_paths_in_flight=0):
        """S generators per pixel: pass p draws from stream p mod S, so up to S passes share one wavefront (hc_pt_set_sample_streams).  Call before InitPathTracing."""
        check(self._L.hc_pt_set_sample_streams(self._c, int(streams), int(max_paths_in_flight)), "hc_pt_set_sample_streams")

    def GroupPasses(self):
        """Passes one wavefront carries with the current streams / tiles / limit (after InitPathTracing)."""
        m = ct.c_int(0)
        check(self._L.hc_pt_group_passes(self._c, ct.byref(m)), "hc_pt_group_passes")
        return int(m.value)

    def SetShadowTrees(self, mode):
        """1 (library default): shadow rays walk every BVH tree, cut-outs occlude (GPUOCLLayer); 0: first tree only (the CPU integrators' shadowTrace)."""
        check(self._L.hc_pt_set_shadow_trees(self._c, int(mode)), "hc_pt_set_shadow_trees")

    def SetMaterialSort(self, enable=True, from_bounce=1):
        """enable: False / 0 = off, True / 1 = on, "auto" / 2 = the default (on with >= 3 materials and >= 384k paths per pass)."""
        mode = 2 if enable in ("auto", 2) else (1 if enable else 0)
        check(self._L.hc_pt_set_material_sort(self._c, mode, int(from_bounce)), "hc_pt_set_material_sort")

    def TracingPass(self, integrator=INTEGRATOR_MISPT, passes=1):
        """BeginTracingPass + EndTracingPass, `passes` times."""
        check(self._L.hc_pt_pass(self._c, int(integrator), int(passes)), "hc_pt_pass")

    def ClearAccumulatedColor(self):
        check(self._L.hc_fb_clear(self._c), "hc_fb_clear")

    def GetHDRImage(self):
        out = np.empty((self.height, self.width, 4), np.float32)
        check(self._L.hc_fb_read_hdr(self._c, _ptr(out), self.width, self.height), "hc_fb_read_hdr")
        return out

    def GetLDRImage(self):
        out = np.empty((self.height, self.width), np.uint32)
        check(self._L.hc_fb_read_ldr(self._c, _ptr(out), self.width, self.height), "hc_fb_read_ldr")
        return out

    # ---- multi-GPU exchange inside the library (hc_comm.cu): one process per GPU, NCCL over NVLink
    @staticmethod
    def CommUniqueId():
        """rank 0: 128 bytes of ncclGetUniqueId; the host distributes them (bench.py broadcasts them with torch.distributed)."""
        buf = (ct.c_ubyte*128)()
        check(load().hc_comm_unique_id(buf), "hc_comm_unique_id")
        return bytes(buf)

    def CommInit(self, unique_id, rank, world):
        buf = (ct.c_ubyte*128).from_buffer_copy(bytes(unique_id))
        check(self._L.hc_comm_init(self._c, buf, int(rank), int(world)), "hc_comm_init")

    def ReduceFramebuffer(self, dst=0, mode=0):
        """Combine the per-rank SUM buffers on `dst` (mode 0: owned tiles travel; mode 1: full-size sum).  Returns the device time in ms."""
        ms = ct.c_float()
        check(self._L.hc_fb_reduce(self._c, int(dst), int(mode), ct.byref(ms)), "hc_fb_reduce")
        return ms.value

    def GetSumImage(self):
        out = np.empty((self.height, self.width, 4), np.float32)
        check(self._L.hc_fb_read_sum(self._c, _ptr(out), self.width, self.height), "hc_fb_read_sum")
        return out

    def fb_device_ptr(self):
        p, n = ct.c_void_p(), ct.c_int64()
        check(self._L.hc_fb_device_ptr(self._c, ct.byref(p), ct.byref(n)), "hc_fb_device_ptr")
        return p.value, n.value

    def GetSPP(self):
        s = ct.c_float()
        check(self._L.hc_get_spp(self._c, ct.byref(s)), "hc_get_spp")
        return s.value

    def GetRaysStat(self):
        s = hc_stats()
        check(self._L.hc_get_stats(self._c, ct.byref(s)), "hc_get_stats")
        return {k: getattr(s, k) for k, _ in hc_stats._fields_}

    def ResetPerfCounters(self):
        check(self._L.hc_reset_stats(self._c), "hc_reset_stats")

    # ---- convenience: upload a hydracore_b200.scene.Scene
    def LoadScene(self, scn):
        for name in ("textures", "textures_aux", "geom", "materials", "pdfs"):
            self.UploadStorage(name, scn.storages[name])
        self.SetAllBVH4(scn.bvh["nodes"], scn.bvh["tris"])
        if getattr(scn, "bvh1", None) is not None:
            self.SetAllBVH4(scn.bvh1["nodes"], scn.bvh1["tris"], tree=1, alpha=scn.bvh1["alpha"])
        self.SetAllInstMatrices(scn.bvh["inv_matrices"])
        self.SetAllInstLightInstId(scn.inst_light_ids)
        if getattr(scn, "remap_lists", None):
            self.SetAllRemapLists(*scn.remap_arrays())
        else:
            self.SetAllRemapLists(None, None, None)
        self.ResizeScreen(scn.width, scn.height)
        self.PrepareEngineGlobals(scn.globals_blob)


HIT_DTYPE = np.dtype([("t", np.float32), ("primId", np.int32), ("instId", np.int32), ("geomId", np.int32)])
