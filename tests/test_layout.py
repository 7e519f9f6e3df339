"""CPU: every layout constant the kernels and packers use (csrc/hc_layout.h) must equal the reference's own value
(tests/golden/ref_consts.json dumped from the reference headers through oracle/_ref; re-checked live when _ref is present)."""
import json
import os

from hydracore_b200 import layout

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _expected(name, ref):
    if name.startswith("EG_"):
        f = name[3:]
        if f == "sizeof":
            return ref["sizeof_EngineGlobals"]
        if f == "HEAD_BYTES":
            return ref["offsetof_EngineGlobals_suns"]
        return ref["offsetof_EngineGlobals_" + f]
    return ref.get(name)


def _check(ref):
    missing, wrong = [], []
    for name, val in layout.C.items():
        exp = _expected(name, ref)
        if exp is None:
            missing.append(name)
        elif exp != val:
            wrong.append((name, val, exp))
    assert not wrong, wrong
    assert not missing, f"constants with no reference counterpart: {missing}"


def test_layout_matches_golden():
    assert len(layout.C) > 80
    _check(json.load(open(os.path.join(G, "ref_consts.json"))))


def test_layout_matches_live_reference(ref):
    _check(ref.consts())
    assert ref.consts() == json.load(open(os.path.join(G, "ref_consts.json")))


def test_device_bvh_layout_invariants(built):
    """hc_bvh_device_layout (host side of hc_set_bvh): SoA quads keep their index, every triangle lands in exactly one pair record with
    edges B-A, C-A evaluated in float, odd leaves are padded with a zero triangle, empty child slots get an infinite box."""
    import ctypes as ct
    import numpy as np
    import hydracore_b200 as hc
    from tests import scenes
    scn = scenes.instanced_geometry()
    nodes = np.ascontiguousarray(scn.bvh["nodes"], np.float32)
    tris = np.ascontiguousarray(scn.bvh["tris"], np.float32)
    lib = hc.load()
    npairs, bound = ct.c_int64(), ct.c_int()
    P = lambda a: a.ctypes.data_as(ct.c_void_p)
    assert lib.hc_bvh_device_layout(P(nodes), nodes.shape[0], P(tris), tris.shape[0], None, None, 0, ct.byref(npairs), ct.byref(bound)) == 0
    assert 0 < bound.value <= 64 and npairs.value % 24 == 0
    dn = np.zeros((nodes.shape[0]//4, 32), np.float32)
    dp = np.zeros(npairs.value, np.float32)
    assert lib.hc_bvh_device_layout(P(nodes), nodes.shape[0], P(tris), tris.shape[0], P(dn), P(dp), dp.size, ct.byref(npairs), ct.byref(bound)) == 0
    pairs = dp.reshape(-1, 24)
    pi = pairs.view(np.int32)
    un = nodes.view(np.uint32).reshape(-1, 4, 8)                     # [quad][child][8 words]
    du = dn.view(np.uint32)
    # walk the top level from quad 1 and the mesh sub-trees behind the instance records
    ti = tris.view(np.int32)
    seen_tris, todo, visited = 0, [(1, False)], set()
    while todo:
        q, in_mesh = todo.pop()
        if (q, in_mesh) in visited:
            continue
        visited.add((q, in_mesh))
        for c in range(4):
            lo, esc = un[q, c, 3], un[q, c, 7]
            word = du[q, 24 + c]
            if lo == 0xFFFFFFFF and esc == 0xFFFFFFFF:
                assert word == 0xFFFFFFFF and np.isposinf(dn[q, 0 + c]) and np.isposinf(dn[q, 4 + c])
                continue
            assert np.array_equal(dn[q, [0 + c, 8 + c, 16 + c]], nodes[4*q + c, 0:3]) and np.array_equal(dn[q, [4 + c, 12 + c, 20 + c]], nodes[4*q + c, 4:7])
            off = int(lo & 0x7FFFFFFF)
            if not (lo & 0x80000000):
                assert word == off
                todo.append((off, in_mesh))
            elif not in_mesh:                                       # instance leaf -> record quad `off`
                assert word == (0x80000000 | off)
                rec = nodes[4*off:4*off + 4].reshape(-1)
                assert np.array_equal(dn[off, 0:16], rec[8:24])                       # inverse matrix
                assert du[off, 17] == rec.view(np.uint32)[24]                         # realInstId
                sub = int(un[off, 0, 3])
                if sub & 0x80000000:
                    leaf_words = [(int(du[off, 16]), sub & 0x7FFFFFFF)]
                else:
                    assert du[off, 16] == sub
                    todo.append((sub, True))
                    leaf_words = []
                for w, hdr in leaf_words:
                    seen_tris += _check_leaf(w, hdr, ti, tris, pairs, pi)
            else:
                seen_tris += _check_leaf(int(word), off, ti, tris, pairs, pi)
    assert seen_tris == sum(m.tri_count for m in scn.meshes)
    # error paths: out-of-range child offset
    bad = nodes.copy()
    bad.view(np.uint32)[4, 3] = 0x00FFFFFF
    assert lib.hc_bvh_device_layout(P(bad), bad.shape[0], P(tris), tris.shape[0], None, None, 0, ct.byref(npairs), ct.byref(bound)) != 0


def _check_leaf(word, hdr, ti, tris, pairs, pi):
    import numpy as np
    assert word & 0x80000000
    npair = ((word >> 25) & 63) + 1
    first = word & 0x01FFFFFF
    count = int(ti[hdr, 1])
    assert npair == (count + 1)//2 and ti[hdr, 0] == hdr + 1
    for k in range(count):
        A, B, Cc = tris[hdr + 1 + 3*k], tris[hdr + 2 + 3*k], tris[hdr + 3 + 3*k]
        R, s = pairs[first + k//2], k & 1
        assert np.array_equal(R[[0 + s, 2 + s, 4 + s]], A[:3])
        assert np.array_equal(R[[6 + s, 8 + s, 10 + s]], (B[:3] - A[:3]).astype(np.float32))
        assert np.array_equal(R[[12 + s, 14 + s, 16 + s]], (Cc[:3] - A[:3]).astype(np.float32))
        assert pi[first + k//2, 18 + s] == ti[hdr + 1 + 3*k, 3] and pi[first + k//2, 20 + s] == ti[hdr + 2 + 3*k, 3]
    if count & 1:
        R = pairs[first + count//2]
        assert not R[[1, 3, 5, 7, 9, 11, 13, 15, 17]].any() and pi[first + count//2, 19] == -1
    return count
