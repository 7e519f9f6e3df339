"""Golden vectors of the reference's microfacet lobes (hydra_drv/cmatpbrt.h, Beckmann and Trowbridge-Reitz in the local frame) and of its
erf / erf^-1, produced by oracle/_ref (the reference headers compiled in place).  python tests/golden/make_microfacet_golden.py"""
import ctypes as ct
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.test_microfacet import inputs, perez_inputs, _perez, P          # noqa: E402

R = ct.CDLL(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libhydra_ref.so"))
wo, wi, u, al, x = inputs(8192, 11)
out = {}
for kind in (0, 1):
    a = np.zeros((wo.shape[0], 8), np.float32)
    R.ref_pbrt_microfacet(kind, P(wo), P(wi), P(u), P(al), wo.shape[0], P(a))
    out["lobe%d" % kind] = a
e, ie = np.zeros_like(x), np.zeros_like(x)
R.ref_pbrt_erf(P(x), x.size, P(e), P(ie))
out["erf"], out["erfinv"] = e, ie
out["perez"] = _perez(R, "ref", perez_inputs(2048, 3))
np.savez_compressed(os.path.join(HERE, "microfacet.npz"), **out)
print({k: v.shape for k, v in out.items()})
