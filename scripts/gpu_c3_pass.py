"""A few MISPT passes of the C3 scene (for ncu): python scripts/gpu_c3_pass.py [passes] [scene=c3|c4]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hydracore_b200 as hc  # noqa: E402
from hydracore_b200 import scene as S  # noqa: E402

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
which = sys.argv[2] if len(sys.argv) > 2 else "c3"
scn = S.scene_c3(1920, 1080) if which == "c3" else S.scene_c4(1920, 1080)
lay = hc.CudaLayer()
lay.LoadScene(scn)
lay.InitPathTracing(777)
lay.TracingPass(2, 1)
lay.ResetPerfCounters()
lay.TracingPass(2, passes)
st = lay.GetRaysStat()
print({k: (v/passes if k.startswith("ms") else v) for k, v in st.items()}, "ms/pass", lay.last_trace_ms())
