"""Layout constants of the reference's blobs, parsed from csrc/hc_layout.h (single source of truth shared with the kernels).

``C['HRT_TRACE_DEPTH']`` etc.  Names drop the ``HC_`` prefix so they read like the reference's own
(hydra_drv/cglobals.h, cfetch.h, cmaterial.h, clight.h)."""
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "csrc", "hc_layout.h")


def _parse(path):
    out = {}
    rx = re.compile(r"^\s*#define\s+HC_([A-Za-z0-9_]+)\s+(\(?-?(?:0x[0-9A-Fa-f]+|\d+)\)?)\s*(?://.*)?$")
    with open(path) as f:
        for line in f:
            m = rx.match(line)
            if m:
                out[m.group(1)] = int(m.group(2).strip("()"), 0)
    return out


C = _parse(HEADER)
globals().update({k: v for k, v in C.items() if k.isidentifier()})
