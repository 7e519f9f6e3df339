"""Golden HDR images (MISPT and PT, 48x48, sums of 3 passes, seed 777) of the feature scenes in tests/scenes.py from the reference's own CPU
integrators (oracle/_ref), so that the GPU parity of every light / BSDF / map feature is also checked where _ref is not available.
    python tests/golden/make_feature_images.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import refapi, scenes  # noqa: E402

W = H = 48
FEATURES = {"orennayar": lambda: scenes.cornell_orennayar(W, H), "sphere_point": lambda: scenes.cornell_sphere_and_point_lights(W, H),
            "spot_direct": lambda: scenes.cornell_spot_and_direct_lights(W, H, True), "translucent_thin_glass": lambda: scenes.cornell_translucent(W, H),
            "normal_maps": lambda: scenes.cornell_normal_mapped(W, H), "remap": lambda: scenes.cornell_remap_lists(W, H),
            "cutout": lambda: scenes.cornell_with_cutout(W, H), "mesh_light": lambda: scenes.cornell_mesh_light(W, H),
            "cylinder_light": lambda: scenes.cornell_cylinder_light(W, H, True), "sky": lambda: scenes.open_box_under_sky(W, H, True),
            "sky_env": lambda: scenes.open_box_under_sky(W, H, False, env_map=True), "anisotropic": lambda: scenes.cornell_anisotropic(W, H),
            "mesh_light_tex": lambda: scenes.cornell_mesh_light(W, H, True), "cylinder_light_tex": lambda: scenes.cornell_cylinder_light(W, H, True, True),
            "area_spot": lambda: scenes.cornell_area_spot(W, H), "perez_sky": lambda: scenes.open_box_under_sky(W, H, False, perez=True),
            "ies": lambda: scenes.cornell_ies(W, H)}


def main():
    ref = refapi.Ref.try_load()
    assert ref is not None
    out = {}
    for name, f in FEATURES.items():
        rs = ref.scene(f())
        for kind, tag in ((2, "mispt"), (0, "pt")):
            img, _n = rs.render(kind, 777, 3)
            out[f"{name}_{tag}_sum3"] = img[..., :3].astype(np.float32)
        rs.close()
    p = os.path.join(HERE, "feature_images.npz")
    np.savez_compressed(p, **out)
    print("written", len(out), "images,", os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
