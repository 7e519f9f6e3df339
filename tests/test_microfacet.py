"""CPU: hydracore_b200/csrc/hc_microfacet.cuh (the anisotropic Beckmann / Trowbridge-Reitz lobes behind the reference's
PLAIN_MAT_CLASS_BECKMANN / PLAIN_MAT_CLASS_TRGGX materials) compiled as plain C++ and compared bit for bit with the reference's own
functions (hydra_drv/cmatpbrt.h:103-524): against committed golden vectors made by oracle/_ref, and against _ref itself where present.
The header is pure arithmetic, so the host build differs from the sm_100a build only in libm vs libdevice for the double transcendentals;
the GPU tests (test_path_gpu.py) cover the device side through whole materials."""
import ctypes as ct
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def P(a):
    return a.ctypes.data_as(ct.c_void_p)


def inputs(n, seed):
    """Directions on both hemispheres (normal and grazing incidence included), anisotropic and isotropic alphas over the range
    BeckmannRoughnessToAlpha produces, uniform samples including both ends."""
    rs = np.random.RandomState(seed)

    def dirs():
        v = rs.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        v[:, 2] = np.abs(v[:, 2])
        v[rs.rand(n) < 0.2, 2] *= -1
        return np.ascontiguousarray(v.astype(np.float32))
    wo, wi = dirs(), dirs()
    h = rs.normal(size=(n, 3))*0.3 + np.array([0.0, 0.0, 1.0])                            # half of the pairs reflect about a half vector near the normal,
    h /= np.linalg.norm(h, axis=1, keepdims=True)                                        # where the lobes are not negligible
    refl = (2.0*(wo*h).sum(1, keepdims=True)*h - wo).astype(np.float32)
    near = rs.rand(n) < 0.5
    wi[near] = refl[near]
    k = n//50
    wo[:k] = np.array([0, 0, 1], np.float32)
    wo[k:2*k, 2] = 1e-4
    u = rs.rand(n, 2).astype(np.float32)
    u[:k//4] = 0.0
    u[k//4:k//2] = np.float32(0.99999994)
    al = np.exp(rs.uniform(np.log(1e-3), np.log(1.7), size=(n, 2))).astype(np.float32)
    al[near] = np.maximum(al[near], np.float32(0.05))
    iso = rs.rand(n) < 0.4
    al[iso, 1] = al[iso, 0]
    x = np.concatenate([rs.uniform(-1.2, 1.2, n), rs.uniform(-6, 6, n)]).astype(np.float32)
    return wo, wi, u, al, x


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("mf")/"libhost_microfacet.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-msse4.2", "-ffp-contract=off", "-fPIC", "-shared",
                           os.path.join(ROOT, "tests", "host_microfacet.cpp"), "-o", so])
    return ct.CDLL(so)


def _run(lib, prefix, wo, wi, u, al, x):
    out = {}
    for kind in (0, 1):
        a = np.zeros((wo.shape[0], 8), np.float32)
        getattr(lib, prefix + "_pbrt_microfacet")(kind, P(wo), P(wi), P(u), P(al), wo.shape[0], P(a))
        out["lobe%d" % kind] = a
    e, ie = np.zeros_like(x), np.zeros_like(x)
    getattr(lib, prefix + "_pbrt_erf")(P(x), x.size, P(e), P(ie))
    out["erf"], out["erfinv"] = e, ie
    return out


def perez_inputs(n, seed):
    rs = np.random.RandomState(seed)
    d = rs.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    cases = []
    for sun, turb in (((0.3, -0.8, 0.52), 2.5), ((0.0, -1.0, 0.0), 2.0), ((0.7, -0.1, 0.7), 6.0), ((0.5, 0.2, 0.84), 3.0)):
        s = np.array(sun, np.float64)
        s = (s/np.linalg.norm(s)).astype(np.float32)
        dd = d.copy()
        k = n//20
        dd[:k] = -s + rs.normal(size=(k, 3))*0.02                          # directions into the sun disk
        dd /= np.linalg.norm(dd, axis=1, keepdims=True)
        cases.append((s, np.float32(turb), np.array([1.0, 0.9, 0.7], np.float32), np.ascontiguousarray(dd.astype(np.float32))))
    return cases


def _perez(lib, prefix, cases):
    f = getattr(lib, prefix + "_perez_sky")
    f.argtypes = [ct.c_void_p, ct.c_float, ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_void_p]
    out = []
    for s, turb, col, dd in cases:
        a = np.zeros_like(dd)
        f(P(s), float(turb), P(col), P(dd), dd.shape[0], P(a))
        out.append(a)
    return np.concatenate(out)


def _same(a, b):
    return bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


def test_microfacet_header_matches_golden_vectors(host):
    want = np.load(os.path.join(G, "microfacet.npz"))
    got = _run(host, "host", *inputs(8192, 11))
    for k in ("lobe0", "lobe1", "erf", "erfinv"):
        assert _same(got[k], want[k]), k
    got_sky = _perez(host, "host", perez_inputs(2048, 3))
    assert _same(got_sky, want["perez"]) and (want["perez"].max(1) > 1.5).sum() > 50 and want["perez"].mean() > 0.02      # sky colours, some inside the sun disk
    for k in ("lobe0", "lobe1"):                          # the vectors are not trivial: most BRDF / pdf values are positive, half vectors are unit
        assert (want[k][:, 0] > 0).mean() > 0.4 and (want[k][:, 1] > 0).mean() > 0.4
        assert np.allclose(np.linalg.norm(want[k][:, 2:5], axis=1), 1.0, atol=1e-5)


def test_microfacet_header_matches_live_reference(host, ref):
    args = inputs(200000, 5)
    got, want = _run(host, "host", *args), _run(ref.L, "ref", *args)
    for k in ("lobe0", "lobe1", "erf", "erfinv"):
        assert _same(got[k], want[k]), k
    cases = perez_inputs(100000, 2)
    assert _same(_perez(host, "host", cases), _perez(ref.L, "ref", cases))
