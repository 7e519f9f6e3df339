"""ctypes access to the two checkers (TEST INFRASTRUCTURE): oracle/libhydra_oracle.so (our restatement) and
oracle/_ref/libhydra_ref.so (the reference's own headers + CPU integrators compiled in place, see oracle/Makefile)."""
import ctypes as ct
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HIT_DTYPE = np.dtype([("t", np.float32), ("primId", np.int32), ("instId", np.int32), ("geomId", np.int32)])


def P(a):
    return a.ctypes.data_as(ct.c_void_p)


def rays_from_pos_dir(pos_dir6, tfar=3.402823466e+38):
    n = pos_dir6.shape[0]
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = pos_dir6[:, 0:3]
    rays[:, 4:7] = pos_dir6[:, 3:6]
    rays[:, 7] = tfar
    return rays


class Oracle:
    def __init__(self):
        self.L = ct.CDLL(os.path.join(ROOT, "oracle", "libhydra_oracle.so"))
        self.L.orc_qmc_sobol.restype = ct.c_float
        self.L.orc_qmc_sobol.argtypes = [ct.c_uint32, ct.c_int, ct.c_void_p]

    def rng_init(self, seed):
        s = np.zeros(2, np.uint32)
        self.L.orc_rng_init(ct.c_int(seed), P(s))
        return s

    def rng_float4(self, state, n):
        out = np.zeros((n, 4), np.float32)
        self.L.orc_rng_float4(P(state), n, P(out))
        return out

    def rng_float1(self, state, n):
        out = np.zeros(n, np.float32)
        self.L.orc_rng_float1(P(state), n, P(out))
        return out

    def qmc_table(self):
        t = np.zeros((11, 31), np.uint32)
        self.L.orc_qmc_table(P(t))
        return t

    def qmc_sobol(self, pos, dim, table):
        return self.L.orc_qmc_sobol(int(pos), int(dim), P(table))

    def make_rand_eye_rays(self, globals_blob, w, h, xy, offsets):
        n = xy.shape[0]
        out = np.zeros((n, 6), np.float32)
        self.L.orc_make_rand_eye_rays(P(globals_blob), w, h, P(np.ascontiguousarray(xy, np.int32)), P(np.ascontiguousarray(offsets, np.float32)), n, P(out))
        return out

    def make_eye_rays_f4(self, globals_blob, lens):
        n = lens.shape[0]
        out, xy = np.zeros((n, 6), np.float32), np.zeros((n, 2), np.float32)
        self.L.orc_make_eye_rays_f4(P(globals_blob), P(np.ascontiguousarray(lens, np.float32)), n, P(out), P(xy))
        return out, xy

    def trace_closest(self, nodes, tris, rays8, count=False):
        n = rays8.shape[0]
        hits = np.zeros(n, HIT_DTYPE)
        cnt = np.zeros(3, np.uint64)
        self.L.orc_trace_closest(P(nodes), P(tris), P(rays8), ct.c_longlong(n), P(hits), P(cnt) if count else None)
        return (hits, cnt) if count else hits

    def trace_shadow(self, nodes, tris, rays8, count=False):
        n = rays8.shape[0]
        vis = np.zeros(n, np.uint8)
        cnt = np.zeros(3, np.uint64)
        self.L.orc_trace_shadow(P(nodes), P(tris), P(rays8), ct.c_longlong(n), P(vis), P(cnt) if count else None)
        return (vis, cnt) if count else vis


class Ref:
    PATH = os.path.join(ROOT, "oracle", "_ref", "libhydra_ref.so")

    @classmethod
    def try_load(cls):
        return cls() if os.path.exists(cls.PATH) else None

    def __init__(self):
        self.L = ct.CDLL(self.PATH)
        L = self.L
        L.ref_const_name.restype = ct.c_char_p
        L.ref_const_value.restype = ct.c_longlong
        L.ref_qmc_sobol.restype = ct.c_float
        L.ref_qmc_sobol.argtypes = [ct.c_uint32, ct.c_int, ct.c_void_p]
        L.ref_scene_create.restype = ct.c_void_p
        L.ref_scene_create.argtypes = [ct.c_void_p, ct.c_longlong] + [ct.c_void_p]*7 + [ct.c_int, ct.c_void_p, ct.c_int, ct.c_void_p, ct.c_int, ct.c_int]
        L.ref_scene_destroy.argtypes = [ct.c_void_p]
        if hasattr(L, "ref_scene_set_remap"):
            L.ref_scene_set_remap.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_void_p, ct.c_int, ct.c_void_p, ct.c_int]
        if hasattr(L, "ref_scene_set_tree1"):
            L.ref_scene_set_tree1.argtypes = [ct.c_void_p]*4
        L.ref_render_create.restype = ct.c_void_p
        L.ref_render_create.argtypes = [ct.c_void_p, ct.c_int, ct.c_int]
        L.ref_render_destroy.argtypes = [ct.c_void_p]
        L.ref_render_pass.argtypes = [ct.c_void_p] + [ct.c_int]*4
        L.ref_render_set_streams.argtypes = [ct.c_void_p, ct.c_int]; L.ref_render_set_streams.restype = None
        L.ref_render_pass_qmc_range.argtypes = [ct.c_void_p, ct.c_int, ct.c_int]
        L.ref_render_get_sum.argtypes = [ct.c_void_p, ct.c_void_p]
        for f in (L.ref_surface_eval, L.ref_material_sample, L.ref_material_eval, L.ref_light_sample, L.ref_emission_eval):
            f.restype = None

    def consts(self):
        return {self.L.ref_const_name(i).decode(): self.L.ref_const_value(i) for i in range(self.L.ref_num_consts())}

    def ms_tables(self):
        ggx, tr = np.zeros(64*64, np.uint16), np.zeros(64*64*64, np.uint16)
        self.L.ref_ms_tables(P(ggx), P(tr))
        return ggx, tr

    def rng_init(self, seed):
        s = np.zeros(2, np.uint32)
        self.L.ref_rng_init(ct.c_int(seed), P(s))
        return s

    def rng_float4(self, state, n):
        out = np.zeros((n, 4), np.float32)
        self.L.ref_rng_float4(P(state), n, P(out))
        return out

    def rng_float1(self, state, n):
        out = np.zeros(n, np.float32)
        self.L.ref_rng_float1(P(state), n, P(out))
        return out

    def qmc_table(self):
        t = np.zeros((11, 31), np.uint32)
        self.L.ref_qmc_table(P(t))
        return t

    def qmc_sobol(self, pos, dim, table):
        return self.L.ref_qmc_sobol(int(pos), int(dim), P(table))

    def make_rand_eye_rays(self, globals_blob, w, h, xy, offsets):
        n = xy.shape[0]
        out = np.zeros((n, 6), np.float32)
        self.L.ref_make_rand_eye_rays(P(globals_blob), w, h, P(np.ascontiguousarray(xy, np.int32)), P(np.ascontiguousarray(offsets, np.float32)), n, P(out))
        return out

    def make_eye_rays_f4(self, globals_blob, lens):
        n = lens.shape[0]
        out, xy = np.zeros((n, 6), np.float32), np.zeros((n, 2), np.float32)
        self.L.ref_make_eye_rays_f4(P(globals_blob), P(np.ascontiguousarray(lens, np.float32)), n, P(out), P(xy))
        return out, xy

    def trace_closest(self, nodes, tris, rays8):
        n = rays8.shape[0]
        hits = np.zeros(n, HIT_DTYPE)
        self.L.ref_trace_closest(P(nodes), P(tris), 1, P(rays8), ct.c_longlong(n), P(hits))
        return hits

    def trace_shadow(self, nodes, tris, rays8):
        n = rays8.shape[0]
        vis = np.zeros(n, np.uint8)
        self.L.ref_trace_shadow(P(nodes), P(tris), 1, P(rays8), ct.c_longlong(n), P(vis))
        return vis

    def omp_threads(self, n=None):
        """Set (n > 0) and return the number of OpenMP threads the reference code runs on."""
        if n:
            self.L.ref_omp_set_threads(int(n))
        return int(self.L.ref_omp_max_threads())

    def raycast_step(self, nodes, tris, rays8, light):
        """closest hit + one shadow ray per hit towards `light`, all inside the reference library; returns (hits, visible, rays traced)."""
        n = rays8.shape[0]
        hits, vis = np.zeros(n, HIT_DTYPE), np.zeros(n, np.uint8)
        self.L.ref_raycast_step.restype = ct.c_longlong
        lt = np.asarray(light, np.float32)
        traced = self.L.ref_raycast_step(P(nodes), P(tris), P(rays8), ct.c_longlong(n), P(lt), P(hits), P(vis))
        return hits, vis, int(traced)

    def trace_shadow_anyhit(self, nodes, tris, rays8):
        n = rays8.shape[0]
        vis = np.zeros(n, np.uint8)
        self.L.ref_trace_shadow_anyhit(P(nodes), P(tris), P(rays8), ct.c_longlong(n), P(vis))
        return vis

    # ---- scenes / integrators
    def scene(self, scn):
        return RefScene(self, scn)


class RefScene:
    """Keeps the numpy blobs alive for as long as the reference integrators point into them."""

    def __init__(self, ref, scn):
        self.ref, self.scn = ref, scn
        self._keep = [np.ascontiguousarray(scn.globals_blob, np.int32)] + [np.ascontiguousarray(scn.storages[k]) for k in
                     ("geom", "materials", "textures", "textures_aux", "pdfs")] + [np.ascontiguousarray(scn.bvh["nodes"], np.float32),
                     np.ascontiguousarray(scn.bvh["tris"], np.float32), np.ascontiguousarray(scn.bvh["inv_matrices"], np.float32),
                     np.ascontiguousarray(scn.inst_light_ids, np.int32)]
        g, geom, mats, tex, texa, pdfs, nodes, tris, inv, lids = self._keep
        self.h = ref.L.ref_scene_create(P(g), g.size, P(geom), P(mats), P(tex), P(texa), P(pdfs), P(nodes), P(tris), 1,
                                         P(inv), inv.shape[0], P(lids), scn.width, scn.height)
        if getattr(scn, "bvh1", None) is not None:          # meshes with opacity maps: second tree + alpha table, walked by BVH4InstTraverseAlpha
            t1 = [np.ascontiguousarray(scn.bvh1["nodes"], np.float32), np.ascontiguousarray(scn.bvh1["tris"], np.float32),
                  np.ascontiguousarray(scn.bvh1["alpha"], np.uint32)]
            self._keep += t1
            ref.L.ref_scene_set_tree1(self.h, P(t1[0]), P(t1[1]), P(t1[2]))
        if getattr(scn, "remap_lists", None):
            al, tb, ri = scn.remap_arrays()
            ref.L.ref_scene_set_remap(self.h, P(al), al.size, P(tb), tb.shape[0], P(ri), ri.size)
        self._renders = []

    def close(self):
        for r in self._renders:
            self.ref.L.ref_render_destroy(r)
        self._renders = []
        if self.h:
            self.ref.L.ref_scene_destroy(self.h)
            self.h = None

    def render(self, kind, seed, passes, window=None, streams=1):
        """kind: 0 PT (IntegratorStupidPT), 1 MISPT recursive, 2 MISPTLoop2, 3 MISPT+QMC.  Returns per-pixel SUM image and pass count.
        window = (x0, y0, x1, y1): only these pixels are rendered (per-pixel generators make them equal to the same pixels of a full frame)."""
        r = self.ref.L.ref_render_create(self.h, kind, seed)
        self._renders.append(r)
        if streams != 1:
            self.ref.L.ref_render_set_streams(r, int(streams))
        W, H = self.scn.width, self.scn.height
        x0, y0, x1, y1 = window if window is not None else (0, 0, W, H)
        for _ in range(passes):
            self.ref.L.ref_render_pass(r, int(x0), int(y0), int(x1), int(y1))
        out = np.zeros((H, W, 4), np.float32)
        n = self.ref.L.ref_render_get_sum(r, P(out))
        return out, n

    def surface_eval(self, rays8, hits):
        n = rays8.shape[0]
        out = np.zeros((n, 24), np.float32)
        self.ref.L.ref_surface_eval(self.h, P(rays8), P(hits), n, P(out))
        return out

    def material_sample(self, surf24, ray_dir3, rands10, flags):
        n = surf24.shape[0]
        out, mo = np.zeros((n, 8), np.float32), np.zeros(n, np.int32)
        self.ref.L.ref_material_sample(self.h, P(surf24), P(np.ascontiguousarray(ray_dir3, np.float32)), P(np.ascontiguousarray(rands10, np.float32)),
                                       P(np.ascontiguousarray(flags, np.uint32)), n, P(out), P(mo))
        return out, mo

    def material_eval(self, surf24, l3, v3):
        n = surf24.shape[0]
        out = np.zeros((n, 8), np.float32)
        self.ref.L.ref_material_eval(self.h, P(surf24), P(np.ascontiguousarray(l3, np.float32)), P(np.ascontiguousarray(v3, np.float32)), n, P(out))
        return out

    def light_sample(self, pos3, rnd4):
        n = pos3.shape[0]
        out = np.zeros((n, 12), np.float32)
        self.ref.L.ref_light_sample(self.h, P(np.ascontiguousarray(pos3, np.float32)), P(np.ascontiguousarray(rnd4, np.float32)), n, P(out))
        return out
