"""Golden HDR images from the reference's own CPU integrators (oracle/_ref), per-pixel seeding rule of SURVEY.md 8c.
    python tests/golden/make_golden_images.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import refapi, scenes  # noqa: E402


def main():
    ref = refapi.Ref.try_load()
    assert ref is not None
    out = {}
    for name, kw in (("cornell", dict()), ("cornell_two_lights_dof", dict(two_lights=True, dof=True))):
        scn = scenes.cornell(64, 64, **kw)
        rs = ref.scene(scn)
        for kind, tag in ((2, "mispt"), (0, "pt"), (3, "qmc")):
            img, n = rs.render(kind, 777, 3)
            out[f"{name}_{tag}_sum3"] = img[..., :3].astype(np.float32)
        rs.close()
    np.savez_compressed(os.path.join(HERE, "images.npz"), **out)
    print("written", {k: float(v.mean()) for k, v in out.items()})


if __name__ == "__main__":
    main()
