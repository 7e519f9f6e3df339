"""Seeded scenes shared by the golden-vector generator, the CPU tests and the GPU parity tests."""
import numpy as np

from hydracore_b200 import scene as S


def instanced_geometry(width=96, height=72, dof=False):
    """Two instances of a displaced grid + six scaled spheres: exercises instance transforms (non-uniform scale), shared
    mesh sub-trees and top-level BVH descent.  Geometry only (one dummy material)."""
    scn = S.Scene(width, height, S.Camera(pos=(0, 6, 12), look_at=(0, 0, 0), fov=45, dof=dof, lens_radius=0.1))
    m = scn.add_mesh(S.grid_mesh(60, 60))
    m2 = scn.add_mesh(S.sphere_mesh(1.0, 24, 12))
    scn.add_instance(m, S.translate(0, 0, 0))
    scn.add_instance(m, S.translate(3, 2, -4) @ S.rotate_y(0.5))
    for i in range(6):
        scn.add_instance(m2, S.translate(-4 + 1.7*i, 1.5, 2 - i) @ S.scale(0.5 + 0.1*i, 0.8, 0.6))
    scn.add_material(np.zeros(192, np.float32))
    return scn.build()


def single_triangle_leaf():
    """One instance of a two-triangle mesh: the mesh root is itself a leaf (instance record points straight at triangles)."""
    scn = S.Scene(32, 32, S.Camera(pos=(0, 5, 0.01), look_at=(0, 0, 0), fov=60))
    m = scn.add_mesh(S.quad_mesh(2.0, 2.0))
    scn.add_instance(m, S.translate(0, 0, 0))
    scn.add_material(np.zeros(192, np.float32))
    return scn.build()


def pixel_grid(w, h):
    return np.stack(np.meshgrid(np.arange(w), np.arange(h)), -1).reshape(-1, 2).astype(np.int32)


def incoherent_rays(n, seed, radius=9.0):
    """Rays with random origins on a sphere shell aimed at random points near the centre: fully incoherent."""
    rng = np.random.RandomState(seed)
    o = rng.standard_normal((n, 3))
    o = radius*o/np.linalg.norm(o, axis=1, keepdims=True)
    t = rng.uniform(-3, 3, (n, 3))*np.array([1.5, 0.5, 1.5])
    d = t - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = o
    rays[:, 4:7] = d
    rays[:, 7] = 3.402823466e+38
    return rays
