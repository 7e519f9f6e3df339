"""BASELINE config C1: the reference's own test scene hydra_app/tests/test_42 (teapot 25,600 triangles + box + light quad, Lambert /
Phong blend / emissive, one rect area light, DOF camera).  Run where /root/reference exists:
    python tests/golden/make_test42.py
writes tests/golden/test_42_scene.npz (the parsed scene library: meshes, texture, material / light / camera parameters — so that tests
and bench.py can rebuild the scene on the GPU box, where the reference tree does not exist) and tests/golden/test_42_images.npz
(HDR sums of 4 passes at 128x128 from the reference's CPU integrators compiled in place, oracle/_ref: IntegratorStupidPT — the
"unidirectional PT" the config names — and IntegratorMISPTLoop2, per-pixel seeding of SURVEY.md 8c, seed 777)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from hydracore_b200 import hydra_scene as HS  # noqa: E402
from tests import refapi  # noqa: E402

REF = os.environ.get("HYDRA_REFERENCE", "/root/reference")


def main():
    lib = HS.parse_library(os.path.join(REF, "hydra_app/tests/test_42/statex_00001.xml"), mesh_fallback_dirs=[os.path.join(REF, "hydra_app/data/meshes")])
    HS.save_fixture(lib, os.path.join(HERE, "test_42_scene.npz"))
    scn = HS.build_scene(HS.load_fixture(os.path.join(HERE, "test_42_scene.npz")), 128, 128)
    ref = refapi.Ref.try_load()
    assert ref is not None, "oracle/_ref/libhydra_ref.so missing: run __graft_entry__.build() where the reference tree exists"
    rs = ref.scene(scn)
    out = {}
    for kind, tag in ((0, "pt"), (2, "mispt")):
        img, n = rs.render(kind, 777, 4)
        out["test_42_%s_sum4" % tag] = (img[..., :3]*np.float32(1.0)).astype(np.float32)
    rs.close()
    np.savez_compressed(os.path.join(HERE, "test_42_images.npz"), **out)
    print({k: float(v.mean()) for k, v in out.items()}, os.path.getsize(os.path.join(HERE, "test_42_scene.npz")))


if __name__ == "__main__":
    main()
