"""Does a single GPU gain from MORE than one 1080p pass in flight?  C3 / C4 with 8 sample streams and in-flight limits of 1, 2, 4, 8 frames
(late bounces then launch over 2x / 4x / 8x as many live paths).  python scripts/gpu_streams_cap.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hydracore_b200 as hc
from hydracore_b200 import scene as S
out = {}
for key, build in (("c3", lambda: S.scene_c3(1920, 1080)), ("c4", lambda: S.scene_c4(1920, 1080))):
    scn = build()
    for frames in (1, 2, 4, 8):
        lay = hc.CudaLayer()
        lay.SetSampleStreams(8, frames*1920*1080)
        lay.LoadScene(scn)
        lay.InitPathTracing(777)
        lay.TracingPass(2, 8)
        best = None
        for rep in range(3):
            lay.ResetPerfCounters()
            t0 = time.perf_counter()
            lay.TracingPass(2, 16)
            dt = (time.perf_counter() - t0)/16*1e3
            st = lay.GetRaysStat()
            row = {"ms_per_pass": round(dt, 3), "group": lay.GroupPasses(), "closest": round(st["msClosest"]/16, 3), "shadow_added": round(st["msShadow"]/16, 3),
                   "shade": round(st["msShade"]/16, 3), "other": round(st["msOther"]/16, 3)}
            if best is None or row["ms_per_pass"] < best["ms_per_pass"]:
                best = row
        out["%s_frames%d" % (key, frames)] = best
        print(key, frames, best, flush=True)
        lay.close()
print(json.dumps(out))
