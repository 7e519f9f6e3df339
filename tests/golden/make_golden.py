"""Generate the golden vectors under tests/golden/ from the REFERENCE'S OWN CODE (oracle/_ref/libhydra_ref.so = the reference
headers and CPU integrators compiled in place from /root/reference, see oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures are committed; the GPU box never needs /root/reference."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import refapi, scenes  # noqa: E402


def main():
    ref = refapi.Ref.try_load()
    assert ref is not None, "build oracle/_ref first (python -c 'import __graft_entry__ as g; g.build()')"

    json.dump(ref.consts(), open(os.path.join(HERE, "ref_consts.json"), "w"), indent=0, sort_keys=True)

    # a1: RandomGen known answers
    seeds = np.array([0, 1, 2, 5, 6, 7, 13, 777, 778, 65536, 123456789, 2147483647, -1, -3, -2147483648], np.int64)
    st0, f4, f1, st1 = [], [], [], []
    for s in seeds:
        st = ref.rng_init(int(s))
        st0.append(st.copy())
        f4.append(ref.rng_float4(st, 8))
        f1.append(ref.rng_float1(st, 8))
        st1.append(st.copy())
    # a2: Niederreiter table + samples
    table = ref.qmc_table()
    pos = np.array([0, 1, 2, 3, 7, 100, 65535, 1 << 20, (1 << 31) - 1], np.uint32)
    sob = np.array([[ref.qmc_sobol(p, d, table) for d in range(11)] for p in pos], np.float32)
    np.savez_compressed(os.path.join(HERE, "samplers.npz"), seeds=seeds, state0=np.array(st0), float4=np.array(f4), float1=np.array(f1),
                        state1=np.array(st1), qmc_table=table, qmc_pos=pos, qmc_val=sob)

    # a4: eye rays (pinhole and thin lens) + a7/a9: hits and visibility
    out = {}
    for name, dof in (("pinhole", False), ("dof", True)):
        scn = scenes.instanced_geometry(dof=dof)
        W, H = scn.width, scn.height
        xy = scenes.pixel_grid(W, H)
        rng = np.random.RandomState(7)
        offs = (rng.rand(W*H, 4)*2 - 1).astype(np.float32)
        out["eye_" + name] = ref.make_rand_eye_rays(scn.globals_blob, W, H, xy, offs)
        lens = rng.rand(W*H, 4).astype(np.float32)
        r, fxy = ref.make_eye_rays_f4(scn.globals_blob, lens)
        out["eyef4_" + name] = r
        out["eyef4_xy_" + name] = fxy
        out["offs_" + name] = offs
        out["lens_" + name] = lens
    scn = scenes.instanced_geometry()
    W, H = scn.width, scn.height
    prim = refapi.rays_from_pos_dir(ref.make_rand_eye_rays(scn.globals_blob, W, H, scenes.pixel_grid(W, H), np.zeros((W*H, 4), np.float32)))
    inco = scenes.incoherent_rays(6000, 11)
    out["hits_primary"] = ref.trace_closest(scn.bvh["nodes"], scn.bvh["tris"], prim)
    out["hits_incoherent"] = ref.trace_closest(scn.bvh["nodes"], scn.bvh["tris"], inco)
    sh = inco.copy()
    sh[:, 7] = np.random.RandomState(3).uniform(0.5, 14.0, sh.shape[0]).astype(np.float32)
    out["shadow_tfar"] = sh[:, 7].copy()
    out["vis_incoherent"] = ref.trace_shadow(scn.bvh["nodes"], scn.bvh["tris"], sh)
    out["vis_incoherent_anyhit"] = ref.trace_shadow_anyhit(scn.bvh["nodes"], scn.bvh["tris"], sh)
    s1 = scenes.single_triangle_leaf()
    p1 = refapi.rays_from_pos_dir(ref.make_rand_eye_rays(s1.globals_blob, 32, 32, scenes.pixel_grid(32, 32), np.zeros((1024, 4), np.float32)))
    out["hits_single_leaf"] = ref.trace_closest(s1.bvh["nodes"], s1.bvh["tris"], p1)
    # the BVH itself is part of the fixture: the builder may evolve, the reference traversal's answer on THIS tree may not
    out["bvh_nodes"] = scn.bvh["nodes"]
    out["bvh_tris"] = scn.bvh["tris"]
    out["bvh1_nodes"] = s1.bvh["nodes"]
    out["bvh1_tris"] = s1.bvh["tris"]
    out["rays_primary"] = prim
    out["rays_single_leaf"] = p1
    out["globals_pinhole"] = scenes.instanced_geometry(dof=False).globals_blob[:243]      # EngineGlobals head (972 bytes)
    out["globals_dof"] = scenes.instanced_geometry(dof=True).globals_blob[:243]
    np.savez_compressed(os.path.join(HERE, "raycast.npz"), **out)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
