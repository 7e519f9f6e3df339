"""Wall time per pass with / without CUDA-graph replay of the untimed passes (HC_PT_NO_GRAPH=1): python scripts/gpu_graph_ab.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hydracore_b200 as hc  # noqa: E402
from hydracore_b200 import scene as S, hydra_scene as HS  # noqa: E402

for name, scn, integ in (("c1", HS.build_scene(HS.load_fixture(os.path.join(ROOT, "tests", "golden", "hydra_scenes.npz"), "test_42"), 512, 512), 0),
                         ("c3", S.scene_c3(1920, 1080), 2), ("c3_480x270", S.scene_c3(480, 270), 2)):
    lay = hc.CudaLayer()
    lay.LoadScene(scn)
    lay.InitPathTracing(777)
    lay.TracingPass(integ, 2)
    for passes in (4, 8, 16, 64):
        res = {}
        for mode in ("graph", "direct"):
            if mode == "direct":
                os.environ["HC_PT_NO_GRAPH"] = "1"
            else:
                os.environ.pop("HC_PT_NO_GRAPH", None)
            best = 1e9
            for _ in range(3):
                lay.FinishAll()
                t0 = time.perf_counter()
                lay.TracingPass(integ, passes)
                lay.FinishAll()
                best = min(best, (time.perf_counter() - t0)/passes)
            res[mode] = 1e3*best
        print(name, passes, {k: round(v, 4) for k, v in res.items()}, flush=True)
    lay.close()
