"""Wider parity sweep than the unit tests: every test scene at 256x256, 8 passes, seeds 1 and 4242, MISPT / PT / QMC against the reference's
CPU integrators compiled in place (oracle/_ref travels with the snapshot).  python scripts/gpu_parity_sweep.py  (under gpurun)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hydracore_b200 as hc  # noqa: E402
from hydracore_b200 import hydra_scene as HS  # noqa: E402
from tests import refapi, scenes  # noqa: E402

ref = refapi.Ref.try_load()
assert ref is not None, "oracle/_ref/libhydra_ref.so is missing"
W = H = int(os.environ.get("RES", "256"))
P = int(os.environ.get("PASSES", "8"))
mk_all = {"cornell": lambda: scenes.cornell(W, H, two_lights=True, dof=True), "orennayar": lambda: scenes.cornell_orennayar(W, H),
      "sphere_point": lambda: scenes.cornell_sphere_and_point_lights(W, H), "spot_direct": lambda: scenes.cornell_spot_and_direct_lights(W, H, True),
      "translucent_thin_glass": lambda: scenes.cornell_translucent(W, H), "normal_maps": lambda: scenes.cornell_normal_mapped(W, H),
      "remap": lambda: scenes.cornell_remap_lists(W, H), "mesh_light": lambda: scenes.cornell_mesh_light(W, H), "cylinder_light": lambda: scenes.cornell_cylinder_light(W, H, True), "cutout": lambda: scenes.cornell_with_cutout(W, H),
      "sky": lambda: scenes.open_box_under_sky(W, H, True), "sky_env": lambda: scenes.open_box_under_sky(W, H, False, env_map=True),
      }          # (tests/scenes.instanced_geometry is a ray-casting scene with a placeholder material: not path traced)
mk = {k: v for k, v in mk_all.items() if not os.environ.get("ONLY") or k in os.environ["ONLY"].split(",")}
fx = os.path.join(ROOT, "tests", "golden", "hydra_scenes.npz")
for n in (HS.fixture_scenes(fx) if not os.environ.get("ONLY") else []):
    mk[n] = (lambda n=n: HS.build_scene(HS.load_fixture(fx, n), W, H))
lay = hc.CudaLayer()
worst = 0.0
for name, f in mk.items():
    try:
        scn = f()
    except Exception as e:                      # ray-casting scenes without materials etc.
        print(name, "skipped:", str(e)[:80])
        continue
    rs = ref.scene(scn)
    for seed in (1, 4242):
        for integ, kind in ((2, 2), (0, 0), (3, 3)):
            lay.LoadScene(scn)
            try:
                lay.InitPathTracing(seed)
            except hc.HcError as e:
                print(name, "rejected:", str(e)[:100])
                break
            lay.TracingPass(integ, P)
            got = lay.GetHDRImage()[..., :3]*np.float32(lay.GetSPP())
            t0 = time.time()
            want, n = rs.render(kind, seed, P)
            want = want[..., :3]
            den = float(np.sqrt((want.astype(np.float64)**2).mean())) or 1.0
            rel = float(np.sqrt(((got.astype(np.float64) - want)**2).mean()))/den
            nd = int((np.abs(got - want).max(-1) > 1e-6*np.maximum(np.abs(want).max(-1), 1e-3)).sum())
            worst = max(worst, rel)
            print("%-28s seed %4d integ %d  relRMSE %.3e  pixels differing %6d / %d   (ref %.1f s)" % (name, seed, integ, rel, nd, W*H, time.time() - t0), flush=True)
    rs.close()
print("worst relRMSE", worst)
