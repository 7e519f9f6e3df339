"""Where does the 512x512 / 64 spp PT frame of C1 differ from the reference integrator?  (pixels, magnitudes, pass of first difference)"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hydracore_b200 as hc
from hydracore_b200 import hydra_scene as HS
from tests import refapi
scn = HS.build_scene(HS.load_fixture(os.path.join(ROOT, "tests", "golden", "hydra_scenes.npz"), "test_42"), 512, 512)
ref = refapi.Ref.try_load()
ref.omp_threads(min(len(os.sched_getaffinity(0)), 32))
lay = hc.CudaLayer()
if hasattr(hc.load(), 'hc_pt_set_shadow_trees'):
    lay.SetShadowTrees(0)
lay.LoadScene(scn)
lay.InitPathTracing(777)
rs = ref.scene(scn)
r = ref.L.ref_render_create(rs.h, 0, 777)
rs._renders.append(r)
out = []
prev_bad = 0
for p in range(1, 65):
    lay.TracingPass(0, 1)
    ref.L.ref_render_pass(r, 0, 0, 512, 512)
    if p in (1, 2, 4, 8, 16, 32, 64):
        got = lay.GetHDRImage()[..., :3]*np.float32(p)
        want = np.zeros((512, 512, 4), np.float32)
        ref.L.ref_render_get_sum(r, want.ctypes.data_as(__import__("ctypes").c_void_p))
        want = want[..., :3]
        d = np.abs(got - want)
        rel = d/np.maximum(np.abs(want), 1e-3)
        bad = (rel > 1e-4).any(-1)
        out.append({"passes": p, "rel_rmse": float(np.sqrt((d**2).mean())/np.sqrt((want**2).mean())), "pixels_rel_gt_1e-4": int(bad.sum()),
                    "max_abs_diff": float(d.max()), "bitexact_pixels": int((d == 0).all(-1).sum())})
        print(out[-1], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "c1_diff.json"), "w"), indent=1)
