"""Scheduling parameters of k_trace under the path-tracing load with four passes in flight (8.3 M paths): C3 ms per pass for the refill threshold
(HC_TRACE_REFILL, default 24) and the vote bias (HC_TRACE_QBIAS, default 2), set per process.  python scripts/gpu_refill_c3.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import time
    import hydracore_b200 as hc
    from hydracore_b200 import scene as S
    scn = S.scene_c3(1920, 1080)
    lay = hc.CudaLayer()
    lay.SetSampleStreams(8)
    lay.LoadScene(scn)
    lay.InitPathTracing(777)
    lay.TracingPass(2, 8)
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        lay.TracingPass(2, 16)
        best = min(best, (time.perf_counter() - t0)/16*1e3)
    st = lay.GetRaysStat()
    print(json.dumps({"cfg": sys.argv[1], "ms_per_pass": round(best, 3)}), flush=True)
    sys.exit(0)
for refill, qbias in ((24, 2), (16, 2), (20, 2), (28, 2), (24, 3), (24, 4)):
    env = dict(os.environ, HC_TRACE_REFILL=str(refill), HC_TRACE_QBIAS=str(qbias))
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "refill %d qbias %d" % (refill, qbias)], env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=120)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "failed", flush=True)
