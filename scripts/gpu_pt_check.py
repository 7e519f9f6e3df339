import sys, time
sys.path.insert(0, '.')
import numpy as np
import hydracore_b200 as hc
from tests import refapi, scenes

ref = refapi.Ref.try_load()
lay = hc.CudaLayer()
for name, kw in (("cornell", dict()), ("cornell2L", dict(two_lights=True)), ("cornell_dof", dict(dof=True))):
    scn = scenes.cornell(128, 128, **kw)
    lay.LoadScene(scn)
    rs = ref.scene(scn)
    for kind, integ in ((2, hc.layer.INTEGRATOR_MISPT), (0, hc.layer.INTEGRATOR_PT)):
        for passes in (1, 4):
            lay.InitPathTracing(777)
            t = time.time(); lay.TracingPass(integ, passes); tg = time.time() - t
            got = lay.GetHDRImage()*lay.GetSPP()
            t = time.time(); want, n = rs.render(kind, 777, passes); tr = time.time() - t
            d = np.abs(got[..., :3] - want[..., :3])
            exact = (got[..., :3] == want[..., :3]).all(-1).mean()
            close = (d <= 1e-4*np.maximum(np.abs(want[..., :3]), 1e-3)).all(-1).mean()
            rel = np.sqrt((d**2).mean())/max(np.sqrt((want[..., :3]**2).mean()), 1e-9)
            print(f"{name} kind={kind} passes={passes}: mean got {got[...,:3].mean():.5f} want {want[...,:3].mean():.5f} exact {exact:.4f} close {close:.4f} relRMSE {rel:.3e} nan {np.isnan(got).sum()}  gpu {tg*1e3:.1f} ms ref {tr*1e3:.1f} ms")
    rs.close()
