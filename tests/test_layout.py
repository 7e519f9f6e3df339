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
