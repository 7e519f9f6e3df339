"""Scene libraries of the reference (hydra_app/tests/*) as test fixtures.  Run where /root/reference exists:
    python tests/golden/make_hydra_scenes.py
writes tests/golden/hydra_scenes.npz — the PARSED libraries (meshes, textures, material / light / camera parameters; shared arrays stored
once) so that tests and bench.py can rebuild the scenes on the GPU box, where the reference tree does not exist — and
tests/golden/hydra_scenes_images.npz — HDR sums of 4 passes at 128x128 from the reference's CPU integrators compiled in place
(oracle/_ref; IntegratorStupidPT = the "unidirectional PT" BASELINE config C1 names, and IntegratorMISPTLoop2; per-pixel seeding of
SURVEY.md 8c, seed 777).
  test_42          BASELINE config C1: teapot 25,600 triangles + box + light quad; Lambert (+texture), Phong blend, emissive; rect light; DOF
  test_42_ggx      the same with a GGX reflection layer
  test_224_sphere  another teapot, a rect and a SPHERE area light
  test_42_beckmann, test_224_sphere_microfacet   the same two with brdf_type="torranse_sparrow" (Blinn distribution) reflection layers
  test_42_with_mirror  mirror material (glossiness 1), a black sky-dome "environment" light beside the rect light, "30 30 30" multiplier
  test_224, test_223_small   eleven materials (Lambert, two Phong), textures, rect area light (test_224: two lights)
  demo_06          two triangles under a uniform sky light (its environment texture chunk is not in the reference tree: constant colour)
The other seven libraries of hydra_app/tests (014_Bump_height, Benchmark_Scene03, demo_05, teapot_cylinder, test_aniso, test_aniso2, test_pool) cannot be
loaded by anybody from this tree: their mesh chunks are listed in the reference's own .MISSING_LARGE_BLOBS."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from hydracore_b200 import hydra_scene as HS  # noqa: E402
from tests import refapi  # noqa: E402

REF = os.environ.get("HYDRA_REFERENCE", "/root/reference")
SCENES = ("test_42", "test_42_ggx", "test_224_sphere", "test_42_beckmann", "test_224_sphere_microfacet", "test_42_with_mirror",
          "test_224", "test_223_small", "demo_06")


def main():
    libs = {n: HS.parse_library(os.path.join(REF, "hydra_app/tests", n, "statex_00001.xml"), mesh_fallback_dirs=[os.path.join(REF, "hydra_app/data/meshes")])
            for n in SCENES}
    path = os.path.join(HERE, "hydra_scenes.npz")
    HS.save_fixtures(libs, path)
    ref = refapi.Ref.try_load()
    assert ref is not None, "oracle/_ref/libhydra_ref.so missing: run __graft_entry__.build() where the reference tree exists"
    out = {}
    for n in SCENES:
        scn = HS.build_scene(HS.load_fixture(path, n), 128, 128)
        rs = ref.scene(scn)
        for kind, tag in ((0, "pt"), (2, "mispt")):
            img, _n = rs.render(kind, 777, 4)
            out["%s_%s_sum4" % (n, tag)] = img[..., :3].astype(np.float32)
        rs.close()
    np.savez_compressed(os.path.join(HERE, "hydra_scenes_images.npz"), **out)
    print({k: float(v.mean()) for k, v in out.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
