"""A/B of the material sort on C3 / C4: python scripts/gpu_sort_ab.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hydracore_b200 as hc  # noqa: E402
from hydracore_b200 import scene as S  # noqa: E402

for which in ("c3", "c4"):
    scn = S.scene_c3(1920, 1080) if which == "c3" else S.scene_c4(1920, 1080)
    lay = hc.CudaLayer()
    lay.LoadScene(scn)
    imgs = {}
    for mode in ((0, 0), (1, 1), (1, 2), (1, 0)):
        lay.SetMaterialSort(*mode)
        lay.InitPathTracing(777)
        lay.TracingPass(2, 1)
        lay.ResetPerfCounters()
        lay.TracingPass(2, 3)
        st = lay.GetRaysStat()
        imgs[mode] = lay.GetHDRImage().copy()
        print(which, "sort", mode, "ms/pass %.3f" % lay.last_trace_ms(), {k: round(v/3, 3) for k, v in st.items() if k.startswith("ms")})
    ref = imgs[(0, 0)]
    for m, im in imgs.items():
        print("   identical to unsorted:", m, bool(np.array_equal(im, ref)))
    lay.close()
