"""Where does the end-to-end time of hc_raycast_pass go?  (wall clock around synchronous calls, pinned host buffers)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hydracore_b200 as hc  # noqa: E402
from hydracore_b200 import scene as S  # noqa: E402
from hydracore_b200._lib import HC_HOST, HC_DEVICE  # noqa: E402

scn = S.scene_c2(1920, 1080)
lay = hc.CudaLayer()
lay.LoadScene(scn)
n = 1920*1080
hits_h = torch.empty(n*4, dtype=torch.int32, pin_memory=True)
vis_h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
hits_d = torch.empty(n*4, dtype=torch.int32, device="cuda")
blob = torch.from_numpy(scn.globals_blob.copy()).pin_memory().numpy()
light = S.C2_LIGHT_POS


def t(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e3*(time.perf_counter() - t0)/reps


print("device only            %.3f ms" % t(lambda: lay.RaycastPass(light, None, None, HC_DEVICE)))
print("hits -> host           %.3f ms" % t(lambda: lay.RaycastPass(light, hits_h.data_ptr(), None, HC_HOST)))
print("vis -> host            %.3f ms" % t(lambda: lay.RaycastPass(light, None, vis_h.data_ptr(), HC_HOST)))
print("hits + vis -> host     %.3f ms" % t(lambda: lay.RaycastPass(light, hits_h.data_ptr(), vis_h.data_ptr(), HC_HOST)))
print("globals upload         %.3f ms" % t(lambda: lay.PrepareEngineGlobals(blob)))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    print("plain D2H 33 MB (torch) %.3f ms" % t(lambda: hits_h.copy_(hits_d, non_blocking=True)))
