"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0 checked against the oracle."""
import numpy as np


def run():
    import hydracore_b200 as hc
    from tests import refapi, scenes
    scn = scenes.instanced_geometry(160, 120)
    lay = hc.CudaLayer()
    print("device:", lay.GetDeviceName())
    lay.LoadScene(scn)
    rays = lay.MakeEyeRays(160, 120, None)
    hits = lay.TraceClosest(rays)
    orc = refapi.Oracle()
    want = orc.trace_closest(scn.bvh["nodes"], scn.bvh["tris"], rays)
    bad = (hits["primId"] != want["primId"]) | (hits["instId"] != want["instId"]) | (hits["t"] != want["t"])
    assert bad.sum() <= 2, f"closest-hit mismatch on {bad.sum()} rays"
    sh = rays.copy()
    sh[:, 7] = 8.0
    vis = lay.TraceShadow(sh)
    assert (vis != orc.trace_shadow(scn.bvh["nodes"], scn.bvh["tris"], sh)).sum() <= 2
    # one small path-traced image (MISPT, 3 passes) against the golden image of the reference CPU integrator
    import os
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "images.npz"))
    cs = scenes.cornell(64, 64)
    lay.LoadScene(cs)
    lay.InitPathTracing(777)
    lay.TracingPass(2, 3)
    got = lay.GetHDRImage()[..., :3]*np.float32(3)
    want_img = golden["cornell_mispt_sum3"]
    rel = float(np.sqrt(((got - want_img)**2).mean())/np.sqrt((want_img**2).mean()))
    assert rel <= 1e-4, f"MISPT image differs from the reference integrator: relRMSE {rel}"
    print("smoke ok: %d rays, %.1f%% hit, launches=%d" % (rays.shape[0], 100.0*(hits["primId"] >= 0).mean(), lay.GetRaysStat()["kernelLaunches"]))
    lay.close()
