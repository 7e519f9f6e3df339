"""Sample streams A/B on one GPU: ms per pass with S = 1 and S = 8 for C1 (512 x 512) and for C3 / C4 at 1080p when this GPU owns 1/G of the
tiles (G = 1, 2, 4, 8 emulated with hc_pt_set_tiles; no exchange here).  python scripts/gpu_streams_ab.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hydracore_b200 as hc
from hydracore_b200 import scene as S
from hydracore_b200 import hydra_scene as HS
out = {}
cases = [("c1", lambda: HS.build_scene(HS.load_fixture(os.path.join(ROOT, "tests", "golden", "hydra_scenes.npz"), "test_42"), 512, 512), 0, 64, (1,)),
         ("c3", lambda: S.scene_c3(1920, 1080), 2, 16, (1, 2, 4, 8)), ("c4", lambda: S.scene_c4(1920, 1080), 2, 16, (1, 2, 4, 8))]
for key, build, integ, NP, worlds in cases:
    scn = build()
    for G in worlds:
        for streams in (1, 8):
            lay = hc.CudaLayer()
            lay.SetSampleStreams(streams)
            lay.SetTiles(32, 0, G)
            lay.LoadScene(scn)
            lay.InitPathTracing(777)
            lay.TracingPass(integ, 8)
            best = None
            for rep in range(3):
                lay.ResetPerfCounters()
                t0 = time.perf_counter()
                lay.TracingPass(integ, NP)
                dt = (time.perf_counter() - t0)/NP*1e3
                st = lay.GetRaysStat()
                row = {"ms_per_pass": round(dt, 3), "group": lay.GroupPasses(), "closest": round(st["msClosest"]/NP, 3), "shadow_added": round(st["msShadow"]/NP, 3),
                       "shade": round(st["msShade"]/NP, 3), "other": round(st["msOther"]/NP, 3)}
                if best is None or row["ms_per_pass"] < best["ms_per_pass"]:
                    best = row
            out["%s_G%d_S%d" % (key, G, streams)] = best
            print(key, G, streams, best, flush=True)
            lay.close()
print(json.dumps(out))
