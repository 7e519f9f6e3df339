"""CPU: the C-ABI library builds, loads and exports every symbol include/hydracore_cuda.h declares; the ctypes table mirrors it.
No compute call is made here (no GPU in the build container); without a device hc_ctx_create must fail loudly."""
import ctypes as ct
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hydracore_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hc_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(built):
    from hydracore_b200 import _lib
    names = _declared()
    assert len(names) >= 35
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hydracore_cuda.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.hc_abi_version() == 2


def test_no_device_means_loud_failure(built):
    import hydracore_b200 as hc
    lib = hc.load()
    n = ct.c_int(-1)
    rc = lib.hc_device_count(ct.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(hc.HcError):
        hc.CudaLayer()


def test_bvh_builder_layout(built):
    """Host-side builder: Appendix-A invariants of the flattened tree (quad 0 root record, quad alignment, leaf headers)."""
    import numpy as np
    from tests import scenes
    scn = scenes.instanced_geometry()
    nodes, tris = scn.bvh["nodes"], scn.bvh["tris"]
    u = nodes.view(np.uint32)
    assert nodes.shape[0] % 4 == 0
    assert u[0, 3] == 1 and u[0, 7] == 0                       # root record: leftOffset = 1, not a leaf
    assert np.array_equal(nodes[1:3].reshape(16), np.eye(4, dtype=np.float32).reshape(16))   # identity matrix in nodes 1..2
    # every triangle leaf header: {first = hdr+1, count in 1..4, -1, -1}
    ti = tris.view(np.int32)
    hdr = 0
    ntri = 0
    while hdr < tris.shape[0]:
        assert ti[hdr, 0] == hdr + 1 and 1 <= ti[hdr, 1] <= 4 and ti[hdr, 2] == -1 and ti[hdr, 3] == -1
        ntri += ti[hdr, 1]
        hdr += 1 + 3*ti[hdr, 1]
    assert hdr == tris.shape[0]
    assert ntri == sum(m.tri_count for m in scn.meshes)       # mesh sub-trees are shared between instances
    assert scn.bvh["inv_matrices"].shape == (8, 16)
    assert scn.bvh["max_stack"] <= 64


def test_bvh_builder_against_brute_force(built, oracle):
    """The builder must not lose triangles: reference-style traversal of its tree == brute-force Moeller-Trumbore over all triangles."""
    import numpy as np
    from hydracore_b200 import scene as S
    from tests import refapi, scenes
    scn = S.Scene(64, 36, S.Camera(pos=(0.0, 9.0, 16.0), look_at=(0.0, 0.0, 1.0), fov=45.0))
    scn.add_instance(scn.add_mesh(S.grid_mesh(150, 149, size=60.0, amplitude=1.6)))
    scn.add_material(np.zeros(192, np.float32))
    scn.build()
    rays = refapi.rays_from_pos_dir(oracle.make_rand_eye_rays(scn.globals_blob, 64, 36, scenes.pixel_grid(64, 36), np.zeros((64*36, 4), np.float32)))
    h = oracle.trace_closest(scn.bvh["nodes"], scn.bvh["tris"], rays)
    assert (h["primId"] >= 0).mean() > 0.8
    m = scn.meshes[0]
    A, B, Cc = (m.pos[m.idx[:, k]].astype(np.float64) for k in range(3))
    e1, e2 = B - A, Cc - A
    for i in range(0, rays.shape[0], 23):
        o, d = rays[i, 0:3].astype(np.float64), rays[i, 4:7].astype(np.float64)
        pv = np.cross(d, e2)
        det = (e1*pv).sum(1)
        tv = o - A
        u = (tv*pv).sum(1)/det
        qv = np.cross(tv, e1)
        v = (qv*d).sum(1)/det
        t = (e2*qv).sum(1)/det
        ok = (u >= 0) & (v >= 0) & (u + v <= 1) & (t > 0)
        if ok.any():
            assert h["primId"][i] >= 0 and abs(t[ok].min() - h["t"][i]) <= 1e-4*t[ok].min()
        else:
            assert h["primId"][i] < 0


def test_packed_fp32_is_not_contracted(built):
    """Bit-parity guard for the traversal kernels: ptxas contracts mul.rn.f32x2 + add/sub.rn.f32x2 into FFMA2, which rounds once
    instead of twice.  The triangle test (hc_trace.cuh / hc_trace2.cuh) writes every difference of products as fma(b, -1, a); an FFMA2
    whose multiplier is not the immediate -1 means a packed multiply-add was fused behind our back."""
    import re
    import shutil
    import subprocess
    from hydracore_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.lib_path()], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    funcs = {}
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(ln)
    trace = {k: v for k, v in funcs.items() if "k_trace" in k}
    assert len(trace) >= 6, "closest / any-hit / second-tree / ray-generating instantiations of k_trace expected in the library"
    for name, lines in trace.items():
        text = "\n".join(lines)
        assert "FMUL2" in text and "FADD2" in text and "FMNMX3" in text, "the traversal kernel is expected to use packed FP32 and 3-input min/max: " + name
        ffma2 = [ln.strip() for ln in lines if " FFMA2 " in ln]
        assert ffma2, "expected FFMA2 (a - b as fma(b, -1, a)) in " + name
        other = [ln for ln in ffma2 if ", -1, " not in ln]
        allowed = 0
        assert len(other) == allowed, "contracted packed multiply-add in %s (%d FFMA2 without the -1 multiplier, %d expected):\n%s" % (
            name, len(other), allowed, "\n".join(other[:5]))
        assert text.count(".256") >= 3, "k_trace is expected to fetch triangle pair records by 256-bit loads: " + name


def test_cpp_layer_library_loads_and_fails_loudly_without_device(built):
    """The C++ drop-in (GPUCUDALayer : IHWLayer) is built where the reference headers exist; without a CUDA device CreateCudaImpl throws."""
    from tests import layerapi
    if not layerapi.CppLayer.available():
        pytest.skip("hydracore_b200/cpp/_build/libhydra_cuda_layer.so not present")
    lib = ct.CDLL(layerapi.LIB)
    for name in ("hl_create", "hl_create_storage", "hl_storage_update", "hl_set_bvh", "hl_prepare", "hl_init_path_tracing", "hl_passes", "hl_get_hdr"):
        assert hasattr(lib, name)
    import hydracore_b200 as hc
    n = ct.c_int(-1)
    if hc.load().hc_device_count(ct.byref(n)) == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(layerapi.LayerError, match="no CUDA device"):
        layerapi.CppLayer(16, 16)
