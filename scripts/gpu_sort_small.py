"""Material sort on / off against the number of live paths: C1 (512x512) and C3 at several resolutions, wall ms per pass over 64 passes.
python scripts/gpu_sort_small.py   (under gpurun)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hydracore_b200 as hc  # noqa: E402
from hydracore_b200 import scene as S, hydra_scene as HS  # noqa: E402

cases = [("c1_512", HS.build_scene(HS.load_fixture(os.path.join(ROOT, "tests", "golden", "hydra_scenes.npz"), "test_42"), 512, 512), 0)]
for w, h in ((480, 270), (960, 540), (1920, 1080)):
    cases.append(("c3_%dx%d" % (w, h), S.scene_c3(w, h), 2))
for name, scn, integ in cases:
    lay = hc.CudaLayer()
    lay.LoadScene(scn)
    res = {}
    for sort in (1, 0):
        lay.SetMaterialSort(sort, 1)
        lay.InitPathTracing(777)
        lay.TracingPass(integ, 2)
        best = 1e9
        for _ in range(3):
            lay.FinishAll()
            t0 = time.perf_counter()
            lay.TracingPass(integ, 64)
            lay.FinishAll()
            best = min(best, (time.perf_counter() - t0)/64)
        res["sort" if sort else "nosort"] = round(1e3*best, 4)
    print(name, scn.width*scn.height, res, flush=True)
    lay.close()
