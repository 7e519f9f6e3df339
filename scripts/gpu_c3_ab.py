"""C3 / C4 pass time A/B between library builds on the same box: python scripts/gpu_c3_ab.py  (HC_LIB selects the build)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hydracore_b200 as hc
from hydracore_b200 import scene as S
out = {}
from hydracore_b200 import hydra_scene as HS
W = int(os.environ.get("AB_W", "1920")); H = int(os.environ.get("AB_H", "1080"))
for key, build, integ in (("c1", lambda: HS.build_scene(HS.load_fixture(os.path.join(ROOT, "tests", "golden", "hydra_scenes.npz"), "test_42"), 512, 512), 0),
                          ("c3", lambda: S.scene_c3(W, H), 2), ("c4", lambda: S.scene_c4(W, H), 2)):
    scn = build()
    lay = hc.CudaLayer()
    lay.LoadScene(scn)
    lay.InitPathTracing(777)
    lay.TracingPass(integ, 2)
    best = None
    for rep in range(3):
        lay.ResetPerfCounters()
        t0 = time.perf_counter()
        NP = 64 if key == 'c1' else 8
        lay.TracingPass(integ, NP)
        dt = (time.perf_counter() - t0)/NP*1e3
        st = lay.GetRaysStat()
        row = {"ms_per_pass": round(dt, 3), "closest": round(st["msClosest"]/NP, 3), "shadow_added": round(st["msShadow"]/NP, 3), "shade": round(st["msShade"]/NP, 3), "other": round(st["msOther"]/NP, 3), "rays_closest_per_pass": st["raysClosest"]//NP}
        if best is None or row["ms_per_pass"] < best["ms_per_pass"]:
            best = row
    out[key] = best
    lay.close()
print(json.dumps(out))
