"""C3 / C4 pass time A/B between library builds on the same box: python scripts/gpu_c3_ab.py  (HC_LIB selects the build)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hydracore_b200 as hc
from hydracore_b200 import scene as S
out = {}
for key, build, integ in (("c3", lambda: S.scene_c3(1920, 1080), 2), ("c4", lambda: S.scene_c4(1920, 1080), 2)):
    scn = build()
    lay = hc.CudaLayer()
    lay.LoadScene(scn)
    lay.InitPathTracing(777)
    lay.TracingPass(integ, 2)
    best = None
    for rep in range(3):
        lay.ResetPerfCounters()
        t0 = time.perf_counter()
        lay.TracingPass(integ, 8)
        dt = (time.perf_counter() - t0)/8*1e3
        st = lay.GetRaysStat()
        row = {"ms_per_pass": round(dt, 3), "closest": round(st["msClosest"]/8, 3), "shadow_added": round(st["msShadow"]/8, 3), "shade": round(st["msShade"]/8, 3), "other": round(st["msOther"]/8, 3)}
        if best is None or row["ms_per_pass"] < best["ms_per_pass"]:
            best = row
    out[key] = best
    lay.close()
print(json.dumps(out))
