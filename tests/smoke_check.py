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
    print("smoke ok: %d rays, %.1f%% hit, launches=%d" % (rays.shape[0], 100.0*(hits["primId"] >= 0).mean(), lay.GetRaysStat()["kernelLaunches"]))
    lay.close()
