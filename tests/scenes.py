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


def cornell(width=128, height=128, trace_depth=5, two_lights=False, dof=False, textured=True):
    """Cornell-box style scene in the spirit of hydra_app/tests/test_42: Lambert room (one wall textured), a mirror ball, a glass
    ball, a GGX box, a Phong-over-Lambert fresnel blend, one (or two) rectangular area lights with emissive meshes."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45, dof=dof, lens_radius=0.15))
    scn.set_trace_depth(trace_depth, 3)
    tex = 0
    if textured:
        rng = np.random.RandomState(5)
        img = np.zeros((32, 32, 4), np.uint8)
        yy, xx = np.mgrid[0:32, 0:32]
        chk = (((xx//4) + (yy//4)) % 2).astype(np.uint8)
        img[..., 0] = 60 + 180*chk
        img[..., 1] = 200 - 120*chk
        img[..., 2] = (rng.rand(32, 32)*255).astype(np.uint8)
        img[..., 3] = 255
        tex = scn.add_texture_rgba8(img)
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    floor = scn.add_material(M.lambert((0.9, 0.9, 0.9), tex_id=tex))
    mir = scn.add_material(M.mirror((0.9, 0.9, 0.9)))
    gls = scn.add_material(M.glass((0.95, 0.98, 0.95), ior=1.5, gloss=1.0))
    ggxm = scn.add_material(M.ggx((0.8, 0.6, 0.2), 0.7))
    bl = scn.add_material(M.blend((0.8, 0.8, 0.8), M.phong((0.9, 0.9, 0.9), 0.85), M.lambert((0.2, 0.3, 0.8)), fresnel=True, ior=1.5))
    rgl = scn.add_material(M.glass((0.9, 0.9, 1.0), ior=1.33, gloss=0.8))
    emi = scn.add_material(M.emissive((17.0, 15.0, 12.0), 0))
    # room: faces +x, -x, +y, -y, +z, -z
    room = scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, floor, white, white), inward=True, skip_faces=(4,)))     # open towards the camera (+z)
    scn.add_instance(room)
    sph = S.sphere_mesh(1.0, 32, 16)
    for mat, mtx in ((mir, S.translate(-2.0, -2.8, -1.0) @ S.scale(1.2, 1.2, 1.2)), (gls, S.translate(1.8, -2.9, 1.2) @ S.scale(1.1, 1.1, 1.1)),
                     (rgl, S.translate(0.0, 0.5, -2.0) @ S.scale(0.8, 1.3, 0.8))):
        m = S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, mat, np.int32))
        scn.add_instance(scn.add_mesh(m), mtx)
    box = S.box_mesh(0.9, 1.6, 0.9, mat_ids=(ggxm,)*6, inward=False)
    scn.add_instance(scn.add_mesh(box), S.translate(2.2, -2.4, -1.8) @ S.rotate_y(0.4))
    box2 = S.box_mesh(0.8, 0.8, 0.8, mat_ids=(bl,)*6, inward=False)
    scn.add_instance(scn.add_mesh(box2), S.translate(-0.3, -3.2, 1.8) @ S.rotate_y(-0.5))
    from hydracore_b200 import materials as M2
    lq = S.quad_mesh(1.0, 1.0, y=0.0, mat_id=emi, flip=True)
    lmesh = scn.add_mesh(lq)
    l0 = scn.add_light(M2.area_light((0.0, 3.95, 0.0), (1.0, 1.0), (17.0, 15.0, 12.0)))
    scn.add_instance(lmesh, S.translate(0.0, 3.95, 0.0), light_id=l0)
    if two_lights:
        emi2 = scn.add_material(M.emissive((4.0, 6.0, 9.0), 1))
        lq2 = S.quad_mesh(0.6, 0.6, y=0.0, mat_id=emi2, flip=True)
        R = S.rotate_x(-1.2)[:3, :3]
        l1 = scn.add_light(M2.area_light((-2.5, 1.0, 3.0), (0.6, 0.6), (4.0, 6.0, 9.0), rotation=R))
        scn.add_instance(scn.add_mesh(lq2), S.translate(-2.5, 1.0, 3.0) @ S.rotate_x(-1.2), light_id=l1)
    return scn.build()


def cornell_orennayar(width=96, height=96):
    """Cornell room with Oren-Nayar walls / balls (roughness 0.3, 0.8, 1.0, one of them textured), a GGX box and one rect area light."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    img = np.zeros((16, 16, 4), np.uint8)
    yy, xx = np.mgrid[0:16, 0:16]
    img[..., 0] = 80 + 150*(((xx//2) + (yy//2)) % 2)
    img[..., 1] = 200
    img[..., 2] = 90
    img[..., 3] = 255
    tex = scn.add_texture_rgba8(img)
    white = scn.add_material(M.orennayar((0.73, 0.73, 0.73), 0.8))
    red = scn.add_material(M.orennayar((0.65, 0.05, 0.05), 0.3))
    green = scn.add_material(M.orennayar((0.12, 0.45, 0.15), 1.0))
    floor = scn.add_material(M.orennayar((0.9, 0.9, 0.9), 0.5, tex_id=tex))
    ball = scn.add_material(M.orennayar((0.3, 0.4, 0.9), 0.9))
    ggxm = scn.add_material(M.ggx((0.8, 0.6, 0.2), 0.7))
    emi = scn.add_material(M.emissive((17.0, 15.0, 12.0), 0))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, floor, white, white), inward=True, skip_faces=(4,))))
    sph = S.sphere_mesh(1.0, 32, 16)
    scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, ball, np.int32))), S.translate(-1.5, -2.6, 0.5) @ S.scale(1.4, 1.4, 1.4))
    scn.add_instance(scn.add_mesh(S.box_mesh(0.9, 1.6, 0.9, mat_ids=(ggxm,)*6, inward=False)), S.translate(2.0, -2.4, -1.5) @ S.rotate_y(0.4))
    l0 = scn.add_light(M.area_light((0.0, 3.95, 0.0), (1.0, 1.0), (17.0, 15.0, 12.0)))
    scn.add_instance(scn.add_mesh(S.quad_mesh(1.0, 1.0, y=0.0, mat_id=emi, flip=True)), S.translate(0.0, 3.95, 0.0), light_id=l0)
    return scn.build()


def cornell_anisotropic(width=96, height=96):
    """Cornell room with the anisotropic microfacet materials of the reference (PLAIN_MAT_CLASS_BECKMANN / _TRGGX): isotropic and anisotropic lobes,
    rotated and axis-flipped tangent frames, textured colour / glossiness / anisotropy / rotation, a near-specular lobe (gloss 0.99: RAY_EVENT_S),
    one of them under a fresnel blend over Lambert, lit by a rect area light (so that the eval / pdf side is used by MIS as well)."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    rng = np.random.RandomState(3)
    img = np.zeros((16, 16, 4), np.uint8)
    yy, xx = np.mgrid[0:16, 0:16]
    img[..., 0] = 90 + 160*(((xx//2) + (yy//2)) % 2)
    img[..., 1] = 120 + 8*xx
    img[..., 2] = (rng.rand(16, 16)*255).astype(np.uint8)
    img[..., 3] = 255
    tex = scn.add_texture_rgba8(img)
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    floor = scn.add_material(M.trggx((0.8, 0.8, 0.8), 0.75, aniso=0.7, rot=0.35, tex_id=tex, gloss_tex_id=tex, aniso_tex_id=tex, rot_tex_id=tex))
    back = scn.add_material(M.beckmann((0.7, 0.7, 0.75), 0.55, aniso=0.5, rot=0.1, tex_id=tex))
    b_iso = scn.add_material(M.beckmann((0.9, 0.6, 0.3), 0.6))
    b_ani = scn.add_material(M.beckmann((0.5, 0.8, 0.9), 0.7, aniso=0.8, rot=0.15))
    t_ani = scn.add_material(M.trggx((0.9, 0.9, 0.5), 0.8, aniso=0.6, rot=0.4, flip=True))
    t_bl = scn.add_material(M.blend((0.8, 0.8, 0.8), M.trggx((0.95, 0.95, 0.95), 0.85), M.lambert((0.2, 0.3, 0.8)), fresnel=True, ior=1.5))
    b_spec = scn.add_material(M.beckmann((0.9, 0.9, 0.9), 0.99, aniso=0.3))
    emi = scn.add_material(M.emissive((17.0, 15.0, 12.0), 0))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, floor, white, back), inward=True, skip_faces=(4,))))
    sph = S.sphere_mesh(1.0, 32, 16)
    for mat, mtx in ((b_iso, S.translate(-2.4, -2.9, -0.5) @ S.scale(1.1, 1.1, 1.1)), (b_ani, S.translate(0.0, -2.8, 0.8) @ S.scale(1.2, 1.2, 1.2)),
                     (t_ani, S.translate(2.4, -2.9, -0.3) @ S.scale(1.1, 1.1, 1.1)), (t_bl, S.translate(-1.2, 0.6, -2.2) @ S.scale(0.9, 0.9, 0.9))):
        scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, mat, np.int32))), mtx)
    scn.add_instance(scn.add_mesh(S.box_mesh(0.8, 1.4, 0.8, mat_ids=(b_spec,)*6, inward=False)), S.translate(1.6, 0.2, -2.4) @ S.rotate_y(0.5))
    l0 = scn.add_light(M.area_light((0.0, 3.95, 0.0), (1.0, 1.0), (17.0, 15.0, 12.0)))
    scn.add_instance(scn.add_mesh(S.quad_mesh(1.0, 1.0, y=0.0, mat_id=emi, flip=True)), S.translate(0.0, 3.95, 0.0), light_id=l0)
    return scn.build()


def cornell_multiscatter(ms_tables, width=96, height=96):
    """Cornell room with energy-compensated microfacet materials (PLAIN_MATERIAL_ENERGY_FIX_OR_MULTISCATTER): rough GGX reflectors read the baked
    64 x 64 table EngineGlobals::m_essGgx2017Table, rough glass the 64^3 table m_essTranspTable (ms_tables = both, as InitEngineGlobals copies them)."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(6, 3)
    scn.ms_tables = ms_tables
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    g1 = scn.add_material(M.ggx((0.9, 0.7, 0.3), 0.4, multiscatter=True))
    g2 = scn.add_material(M.ggx((0.8, 0.8, 0.9), 0.75, multiscatter=True))
    gl1 = scn.add_material(M.glass((0.9, 0.95, 0.9), ior=1.5, gloss=0.6, multiscatter=True))
    gl2 = scn.add_material(M.glass((0.95, 0.9, 1.0), ior=1.8, gloss=0.85, multiscatter=True))
    gl3 = scn.add_material(M.glass((1.0, 1.0, 1.0), ior=2.6, gloss=0.7, multiscatter=True))       # relative IOR outside the table: no compensation
    emi = scn.add_material(M.emissive((17.0, 15.0, 12.0), 0))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, g2, white, white), inward=True, skip_faces=(4,))))
    sph = S.sphere_mesh(1.0, 32, 16)
    for mat, mtx in ((g1, S.translate(-2.4, -2.9, -0.8) @ S.scale(1.1, 1.1, 1.1)), (gl1, S.translate(0.0, -2.8, 1.0) @ S.scale(1.2, 1.2, 1.2)),
                     (gl2, S.translate(2.4, -2.9, -0.3) @ S.scale(1.1, 1.1, 1.1)), (gl3, S.translate(-0.8, 0.4, -2.0) @ S.scale(0.9, 0.9, 0.9))):
        scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, mat, np.int32))), mtx)
    l0 = scn.add_light(M.area_light((0.0, 3.95, 0.0), (1.0, 1.0), (17.0, 15.0, 12.0)))
    scn.add_instance(scn.add_mesh(S.quad_mesh(1.0, 1.0, y=0.0, mat_id=emi, flip=True)), S.translate(0.0, 3.95, 0.0), light_id=l0)
    return scn.build()


def cornell_area_spot(width=96, height=96):
    """The Cornell room under a rect area light with a SPOT distribution (distribution="spot": smoothstep between two cone cosines about the light's
    normal, clight.h:532-539, 576-590) and a tilted disk light with another cone; both have emissive meshes, so that eye rays and GI rays hit them."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    ggxm = scn.add_material(M.ggx((0.8, 0.6, 0.2), 0.7))
    mir = scn.add_material(M.mirror((0.9, 0.9, 0.9)))
    emi0 = scn.add_material(M.emissive((30.0, 26.0, 20.0), 0))
    emi1 = scn.add_material(M.emissive((9.0, 12.0, 20.0), 1))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, white, white, white), inward=True, skip_faces=(4,))))
    sph = S.sphere_mesh(1.0, 32, 16)
    scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, ggxm, np.int32))), S.translate(-2.0, -2.8, -1.0) @ S.scale(1.2, 1.2, 1.2))
    scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, mir, np.int32))), S.translate(1.8, -2.9, 0.8) @ S.scale(1.1, 1.1, 1.1))
    l0 = scn.add_light(M.area_light((0.0, 3.95, 0.0), (1.0, 1.0), (30.0, 26.0, 20.0), spot_angles_deg=(40.0, 90.0)))
    scn.add_instance(scn.add_mesh(S.quad_mesh(1.0, 1.0, y=0.0, mat_id=emi0, flip=True)), S.translate(0.0, 3.95, 0.0), light_id=l0)
    R = S.rotate_x(0.6)[:3, :3]
    l1 = scn.add_light(M.area_light((-3.0, 2.0, 1.0), (0.6, 0.6), (9.0, 12.0, 20.0), rotation=R, disk=True, spot_angles_deg=(60.0, 120.0)))
    scn.add_instance(scn.add_mesh(S.quad_mesh(0.6, 0.6, y=0.0, mat_id=emi1, flip=True)), S.translate(-3.0, 2.0, 1.0) @ S.rotate_x(0.6), light_id=l1)
    return scn.build()


def cornell_ies(width=96, height=96):
    """The Cornell room lit through IES photometric webs (LIGHT_HAS_IES): an omni point light, a rect area light that looks the web up per sample
    and a tilted disk light that looks it up from its centre (LIGHT_IES_POINT_AREA); the area lights have emissive meshes (eye rays see them white)."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    yy, xx = np.mgrid[0:16, 0:32]
    web = (0.15 + np.cos(np.pi*yy/15.0)**2*(0.6 + 0.4*np.cos(2.0*np.pi*xx/32.0*3))).astype(np.float32)      # three lobes around the axis
    ies = scn.add_ies_table(web)
    ies2 = scn.add_ies_table(web[::-1, :]*np.float32(0.5) + np.float32(0.1))
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    ggxm = scn.add_material(M.ggx((0.8, 0.6, 0.2), 0.7))
    emi0 = scn.add_material(M.emissive((20.0, 18.0, 14.0), 1))
    emi1 = scn.add_material(M.emissive((8.0, 11.0, 18.0), 2))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, white, white, white), inward=True, skip_faces=(4,))))
    sph = S.sphere_mesh(1.0, 32, 16)
    scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, ggxm, np.int32))), S.translate(-2.0, -2.8, -1.0) @ S.scale(1.2, 1.2, 1.2))
    scn.add_light(M.with_ies(M.point_light((2.0, 1.0, 1.5), (60.0, 50.0, 40.0)), ies, matrix=S.rotate_x(0.5)[:3, :3]))
    l1 = scn.add_light(M.with_ies(M.area_light((0.0, 3.95, 0.0), (1.0, 1.0), (20.0, 18.0, 14.0)), ies2))
    scn.add_instance(scn.add_mesh(S.quad_mesh(1.0, 1.0, y=0.0, mat_id=emi0, flip=True)), S.translate(0.0, 3.95, 0.0), light_id=l1)
    R = S.rotate_x(0.6)[:3, :3]
    l2 = scn.add_light(M.with_ies(M.area_light((-3.0, 2.0, 1.0), (0.6, 0.6), (8.0, 11.0, 18.0), rotation=R, disk=True), ies, matrix=R, point_area=True))
    scn.add_instance(scn.add_mesh(S.quad_mesh(0.6, 0.6, y=0.0, mat_id=emi1, flip=True)), S.translate(-3.0, 2.0, 1.0) @ S.rotate_x(0.6), light_id=l2)
    return scn.build()


def cornell_sphere_and_point_lights(width=96, height=96):
    """The Cornell room lit by a sphere area light (with its emissive mesh, so that paths can hit it) and an omni point light."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    ggxm = scn.add_material(M.ggx((0.8, 0.6, 0.2), 0.7))
    gls = scn.add_material(M.glass((0.95, 0.98, 0.95), ior=1.5, gloss=1.0))
    emi = scn.add_material(M.emissive((30.0, 26.0, 20.0), 0))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, white, white, white), inward=True, skip_faces=(4,))))
    sph = S.sphere_mesh(1.0, 32, 16)
    for mat, mtx in ((ggxm, S.translate(-2.0, -2.8, -1.0) @ S.scale(1.2, 1.2, 1.2)), (gls, S.translate(1.8, -2.9, 1.2) @ S.scale(1.1, 1.1, 1.1))):
        scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, mat, np.int32))), mtx)
    l0 = scn.add_light(M.sphere_light((0.5, 2.4, 0.0), 0.6, (30.0, 26.0, 20.0)))
    lm = S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, emi, np.int32))
    scn.add_instance(scn.add_mesh(lm), S.translate(0.5, 2.4, 0.0) @ S.scale(0.6, 0.6, 0.6), light_id=l0)
    scn.add_light(M.point_light((-2.5, 1.0, 3.0), (25.0, 35.0, 50.0)))
    return scn.build()


def cornell_spot_and_direct_lights(width=96, height=96, soft_sun=True):
    """The Cornell room lit only by delta lights: a spot light from the ceiling and a directional light through the open front (a soft
    one - a "sun" with a 2-degree cone - or a perfectly parallel one)."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    ggxm = scn.add_material(M.ggx((0.8, 0.6, 0.2), 0.7))
    phg = scn.add_material(M.phong((0.7, 0.7, 0.8), 0.8))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, white, white, white), inward=True, skip_faces=(4,))))
    sph = S.sphere_mesh(1.0, 32, 16)
    for mat, mtx in ((ggxm, S.translate(-2.0, -2.8, -1.0) @ S.scale(1.2, 1.2, 1.2)), (phg, S.translate(1.8, -2.9, 1.2) @ S.scale(1.1, 1.1, 1.1))):
        scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, mat, np.int32))), mtx)
    scn.add_light(M.spot_light((0.5, 3.6, 0.5), (-0.2, -1.0, -0.1), (60.0, 55.0, 40.0), falloff_angle=100.0, falloff_angle2=60.0))
    scn.add_light(M.direct_light((6.0, 5.0, 12.0), (-0.45, -0.4, -0.8), (2.0, 2.2, 2.6), radius1=3.0, radius2=5.0, soft_angle_deg=2.0 if soft_sun else 0.0))
    return scn.build()


def cornell_translucent(width=96, height=96):
    """The Cornell room with a translucent curtain (a quad, translucent on its own and as one side of a blend with Lambert) between the
    light and the floor, a translucent sphere, and two thin-glass panes (clear and frosted) in front of the camera."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    tr = scn.add_material(M.translucent((0.8, 0.75, 0.5)))
    tg = scn.add_material(M.thin_glass((0.9, 0.7, 0.7), gloss=1.0))
    tgg = scn.add_material(M.thin_glass((0.7, 0.9, 0.7), gloss=0.8))
    trb = scn.add_material(M.blend((0.5, 0.5, 0.5), M.translucent((0.3, 0.6, 0.9)), M.lambert((0.6, 0.6, 0.6)), fresnel=False))
    emi = scn.add_material(M.emissive((17.0, 15.0, 12.0), 0))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, white, white, white), inward=True, skip_faces=(4,))))
    scn.add_instance(scn.add_mesh(S.quad_mesh(3.0, 3.0, mat_id=tr)), S.translate(-0.8, 1.5, 0.0))
    sph = S.sphere_mesh(1.0, 32, 16)
    scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, trb, np.int32))), S.translate(1.5, -2.6, 1.0) @ S.scale(1.4, 1.4, 1.4))
    scn.add_instance(scn.add_mesh(S.quad_mesh(1.2, 1.2, mat_id=tg)), S.translate(-2.2, -1.0, 2.0) @ S.rotate_x(np.pi/2))        # clear pane facing the camera
    scn.add_instance(scn.add_mesh(S.quad_mesh(1.2, 1.2, mat_id=tgg)), S.translate(2.4, 0.5, 1.5) @ S.rotate_x(np.pi/2))        # frosted pane
    l0 = scn.add_light(M.area_light((0.0, 3.95, 0.0), (1.0, 1.0), (17.0, 15.0, 12.0)))
    scn.add_instance(scn.add_mesh(S.quad_mesh(1.0, 1.0, y=0.0, mat_id=emi, flip=True)), S.translate(0.0, 3.95, 0.0), light_id=l0)
    return scn.build()


def cornell_normal_mapped(width=96, height=96):
    """The Cornell room with normal-mapped surfaces: Lambert floor, a GGX sphere (map with inverted Y), a Phong-over-Lambert blend box
    (X / Y swapped) and a glass sphere: BumpMapping in the BSDF sample and in the explicit-light evaluation, with the cosine corrections."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    yy, xx = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
    hgt = 0.5*np.sin(xx*np.pi/8.0)*np.cos(yy*np.pi/5.0) + 0.25*np.sin((xx + 2*yy)*np.pi/11.0)
    dx, dy = np.gradient(hgt, axis=1), np.gradient(hgt, axis=0)
    nrm = np.stack([-dx, -dy, np.ones_like(hgt)], -1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    img = np.zeros((64, 64, 4), np.uint8)
    img[..., 0] = np.clip((nrm[..., 0]*0.5 + 0.5)*255.0, 0, 255).astype(np.uint8)
    img[..., 1] = np.clip((nrm[..., 1]*0.5 + 0.5)*255.0, 0, 255).astype(np.uint8)
    img[..., 2] = np.clip(nrm[..., 2]*255.0, 0, 255).astype(np.uint8)          # z is used as stored (materialNormalMapFetch, cmaterial.h:2216)
    img[..., 3] = 255
    nm = scn.add_normal_map(img)
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    floor = scn.add_material(M.with_normal_map(M.lambert((0.7, 0.7, 0.7)), nm, row0=(3, 0, 0, 0), row1=(0, 3, 0, 0)))
    ggxm = scn.add_material(M.with_normal_map(M.ggx((0.8, 0.6, 0.2), 0.75), nm, invert_y=True))
    bl = scn.add_material(M.with_normal_map(M.blend((0.5, 0.5, 0.5), M.phong((0.7, 0.7, 0.7), 0.7), M.lambert((0.2, 0.3, 0.7)), fresnel=True), nm, swap_xy=True))
    gls = scn.add_material(M.with_normal_map(M.glass((0.95, 0.98, 0.95), ior=1.5, gloss=1.0), nm))
    emi = scn.add_material(M.emissive((17.0, 15.0, 12.0), 0))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, floor, white, white), inward=True, skip_faces=(4,))))
    sph = S.sphere_mesh(1.0, 32, 16)
    for mat, mtx in ((ggxm, S.translate(-2.0, -2.8, -1.0) @ S.scale(1.2, 1.2, 1.2)), (gls, S.translate(1.8, -2.9, 1.2) @ S.scale(1.1, 1.1, 1.1))):
        scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, mat, np.int32))), mtx)
    scn.add_instance(scn.add_mesh(S.box_mesh(0.9, 1.4, 0.9, mat_ids=(bl,)*6, inward=False)), S.translate(0.3, -2.6, -1.8) @ S.rotate_y(0.5))
    l0 = scn.add_light(M.area_light((0.0, 3.95, 0.0), (1.0, 1.0), (17.0, 15.0, 12.0)))
    scn.add_instance(scn.add_mesh(S.quad_mesh(1.0, 1.0, y=0.0, mat_id=emi, flip=True)), S.translate(0.0, 3.95, 0.0), light_id=l0)
    return scn.build()


def cornell_remap_lists(width=96, height=96):
    """Three instances of ONE two-material box mesh: as modelled, with a remap list that swaps its materials for a mirror and a GGX one, and
    with a list whose only entry does not apply (from-id absent) - per-instance material overrides (remapMaterialId)."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    blue = scn.add_material(M.lambert((0.1, 0.2, 0.7)))
    mir = scn.add_material(M.mirror((0.9, 0.9, 0.9)))
    ggxm = scn.add_material(M.ggx((0.8, 0.6, 0.2), 0.7))
    emi = scn.add_material(M.emissive((17.0, 15.0, 12.0), 0))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, white, white, white), inward=True, skip_faces=(4,))))
    box = scn.add_mesh(S.box_mesh(0.8, 1.2, 0.8, mat_ids=(blue, blue, white, white, red, red), inward=False))
    swap = scn.add_remap_list([(blue, mir), (red, ggxm)])
    noop = scn.add_remap_list([(green, mir)])
    scn.add_instance(box, S.translate(-2.3, -2.8, 0.0) @ S.rotate_y(0.4))
    scn.add_instance(box, S.translate(0.0, -2.8, -1.0) @ S.rotate_y(-0.3), remap_list=swap)
    scn.add_instance(box, S.translate(2.3, -2.8, 0.5) @ S.rotate_y(0.9), remap_list=noop)
    l0 = scn.add_light(M.area_light((0.0, 3.95, 0.0), (1.0, 1.0), (17.0, 15.0, 12.0)))
    scn.add_instance(scn.add_mesh(S.quad_mesh(1.0, 1.0, y=0.0, mat_id=emi, flip=True)), S.translate(0.0, 3.95, 0.0), light_id=l0)
    return scn.build()


def _light_texture(scn):
    """16 x 8 RGBA8 stripes-and-gradient image for textured lights; returns (texture id, luminance image for the light's pdf table)."""
    yy, xx = np.mgrid[0:8, 0:16]
    img = np.zeros((8, 16, 4), np.uint8)
    img[..., 0] = 40 + 200*((xx//2) % 2)
    img[..., 1] = 60 + 20*yy
    img[..., 2] = 250 - 12*xx
    img[..., 3] = 255
    lum = (0.2126*img[..., 0] + 0.7152*img[..., 1] + 0.0722*img[..., 2]).astype(np.float32)/np.float32(255.0)
    return scn.add_texture_rgba8(img), lum


def cornell_mesh_light(width=96, height=96, textured=False):
    """The Cornell room lit by a MESH light: an emissive, rotated and scaled sphere sampled by triangle area (plus a small rect light).
    textured: the light's colour is modulated by a texture looked up at the mesh's texture coordinates (meshLightGetIntensity)."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    ltex = _light_texture(scn)[0] if textured else 0
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    ggxm = scn.add_material(M.ggx((0.8, 0.6, 0.2), 0.7))
    emi0 = scn.add_material(M.emissive((12.0, 10.0, 6.0), 0))
    emi1 = scn.add_material(M.emissive((17.0, 15.0, 12.0), 1))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, white, white, white), inward=True, skip_faces=(4,))))
    sph = S.sphere_mesh(1.0, 16, 8)
    scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, ggxm, np.int32))), S.translate(-2.0, -2.8, -1.0) @ S.scale(1.2, 1.2, 1.2))
    lmesh = scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, emi0, np.int32)))
    mtx = S.translate(1.5, 0.5, 0.5) @ S.rotate_y(0.7) @ S.scale(0.8, 0.5, 0.6)
    l0 = scn.add_mesh_light(lmesh, mtx, (12.0, 10.0, 6.0), tex_id=ltex)
    scn.add_instance(lmesh, mtx, light_id=l0)
    l1 = scn.add_light(M.area_light((0.0, 3.95, 0.0), (0.5, 0.5), (17.0, 15.0, 12.0)))
    scn.add_instance(scn.add_mesh(S.quad_mesh(0.5, 0.5, y=0.0, mat_id=emi1, flip=True)), S.translate(0.0, 3.95, 0.0), light_id=l1)
    return scn.build()


def cornell_cylinder_light(width=96, height=96, second_table=False, textured=False):
    """The Cornell room lit by a CYLINDER light (a tilted three-quarter tube with its emissive mesh).  second_table: another pdf table comes
    first, so that the light's table id is positive and the 2D-table branch of CylinderLightSamplePos runs instead of the uniform one.
    textured: colour texture on the light (cylinderLightGetIntensity) and the pdf table of its luminance image, as the driver builds it."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    ltex, llum = _light_texture(scn) if textured else (0, None)
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    ggxm = scn.add_material(M.ggx((0.8, 0.6, 0.2), 0.7))
    emi0 = scn.add_material(M.emissive((14.0, 12.0, 9.0), 0))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, white, white, white), inward=True, skip_faces=(4,))))
    sph = S.sphere_mesh(1.0, 16, 8)
    scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, ggxm, np.int32))), S.translate(-2.0, -2.8, -1.0) @ S.scale(1.2, 1.2, 1.2))
    if second_table:
        scn.add_sky_pdf_table()
    mtx = S.translate(0.5, 1.5, 0.0) @ S.rotate_x(1.1) @ S.rotate_y(0.4)
    l0 = scn.add_cylinder_light(mtx, 0.4, 3.0, 270.0, (14.0, 12.0, 9.0), tex_id=ltex, tex_lum=llum)
    scn.add_instance(scn.add_mesh(S.cylinder_mesh(0.4, 3.0, 24, phi_max=np.radians(270.0), mat_id=emi0)), mtx, light_id=l0)
    return scn.build()


def cornell_with_cutout(width=96, height=96):
    """The Cornell room with two instances of a quad whose material has an opacity (cut-out) map - a checker of opaque and transparent
    cells, bilinear and point sampled - in front of the back wall and above the floor: the quads go into the alpha-tested tree 1, rays
    pass through the transparent cells (BVH4InstTraverseAlpha), and - as in the reference's CPU integrators - they cast no shadows."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 0, 14.0), look_at=(0, 0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    rng = np.random.default_rng(5)
    cells = (rng.random((8, 8)) > 0.45).astype(np.uint8)*255
    img = np.zeros((64, 64, 4), np.uint8)
    img[..., :3] = np.kron(cells, np.ones((8, 8), np.uint8))[..., None]
    img[..., 3] = 255
    tex = scn.add_texture_rgba8(img)
    white = scn.add_material(M.lambert((0.73, 0.73, 0.73)))
    red = scn.add_material(M.lambert((0.65, 0.05, 0.05)))
    green = scn.add_material(M.lambert((0.12, 0.45, 0.15)))
    leaf = scn.add_material(M.with_opacity(M.lambert((0.2, 0.6, 0.25)), tex, gamma=1.0))
    leaf2 = scn.add_material(M.with_opacity(M.blend((0.6, 0.6, 0.6), M.ggx((0.8, 0.8, 0.8), 0.8), M.lambert((0.7, 0.3, 0.1))), tex, gamma=1.0, flags=C_TEX_POINT_SAM()))
    emi = scn.add_material(M.emissive((17.0, 12.0, 4.0), 0))
    scn.add_instance(scn.add_mesh(S.box_mesh(4.0, 4.0, 4.0, mat_ids=(green, red, white, white, white, white), inward=True, skip_faces=(4,))))
    q = S.quad_mesh(2.5, 2.5)
    q1 = scn.add_mesh(S.Mesh(q.pos, q.idx, norm=q.norm, uv=q.uv, mat=np.full(q.tri_count, leaf, np.int32)))
    q2 = scn.add_mesh(S.Mesh(q.pos, q.idx, norm=q.norm, uv=q.uv*np.float32(1.7) - np.float32(0.3), mat=np.full(q.tri_count, leaf2, np.int32)))
    scn.add_instance(q1, S.translate(-1.0, 0.2, -1.5) @ S.rotate_x(np.pi/2))
    l0 = scn.add_light(M.area_light((0.0, 3.95, 0.0), (1.0, 1.0), (17.0, 12.0, 4.0)))
    scn.add_instance(scn.add_mesh(S.quad_mesh(1.0, 1.0, y=0.0, mat_id=emi, flip=True)), S.translate(0.0, 3.95, 0.0), light_id=l0)
    scn.add_instance(q2, S.translate(1.2, -2.0, 1.0) @ S.rotate_x(0.3))
    scn.add_instance(q1, S.translate(1.5, 1.0, 0.5) @ S.rotate_x(np.pi/2) @ S.scale(0.5, 0.5, 0.5))
    return scn.build()


def C_TEX_POINT_SAM():
    from hydracore_b200.layout import C
    return C["TEX_POINT_SAM"]


def open_box_under_sky(width=96, height=96, with_area_light=True, env_map=False, perez=False):
    """Objects on a floor under a uniform sky-dome light (plus, optionally, a rect area light): rays that leave the scene pick up the
    environment colour with MIS, the sky is sampled through its pdf table."""
    from hydracore_b200 import materials as M
    scn = S.Scene(width, height, S.Camera(pos=(0, 2.0, 12.0), look_at=(0, -1.0, 0), fov=45))
    scn.set_trace_depth(5, 3)
    grey = scn.add_material(M.lambert((0.6, 0.6, 0.6)))
    red = scn.add_material(M.lambert((0.65, 0.1, 0.1)))
    ggxm = scn.add_material(M.ggx((0.8, 0.6, 0.2), 0.7))
    gls = scn.add_material(M.glass((0.95, 0.98, 0.95), ior=1.5, gloss=1.0))
    mir = scn.add_material(M.mirror((0.9, 0.9, 0.9)))
    emi = scn.add_material(M.emissive((12.0, 11.0, 9.0), 1))
    scn.add_instance(scn.add_mesh(S.quad_mesh(8.0, 8.0, y=0.0, mat_id=grey)), S.translate(0.0, -4.0, 0.0))
    sph = S.sphere_mesh(1.0, 32, 16)
    for mat, mtx in ((ggxm, S.translate(-2.4, -2.8, 0.0) @ S.scale(1.2, 1.2, 1.2)), (gls, S.translate(0.0, -2.9, 1.5) @ S.scale(1.1, 1.1, 1.1)),
                     (mir, S.translate(2.5, -2.7, -0.5) @ S.scale(1.3, 1.3, 1.3)), (red, S.translate(0.3, -3.2, -2.5) @ S.scale(0.8, 0.8, 0.8))):
        scn.add_instance(scn.add_mesh(S.Mesh(sph.pos, sph.idx, norm=sph.norm, uv=sph.uv, mat=np.full(sph.tri_count, mat, np.int32))), mtx)
    if env_map:
        # 16 x 8 HDR map: bright patch ("sun") + gradient; pdf table = prefix sums of its luminance (RenderDriverRTE_PdfTables.cpp:520-566)
        yy, xx = np.mgrid[0:8, 0:16]
        env = np.zeros((8, 16, 4), np.float32)
        env[..., 0] = 0.4 + 0.05*xx
        env[..., 1] = 0.5 + 0.04*yy
        env[..., 2] = 0.9 - 0.03*yy
        env[2, 5, 0:3] = (40.0, 36.0, 30.0)
        env[..., 3] = 1.0
        tex = scn.add_texture_f4(env)
        lum = (0.2126*env[..., 0] + 0.7152*env[..., 1] + 0.0722*env[..., 2]).astype(np.float32)
        tab = scn.add_sky_pdf_table(lum)
        scn.add_light(M.sky_light((1.0, 1.0, 1.0), tab, tex_id=tex))
    else:
        tab = scn.add_sky_pdf_table()
        # perez: the analytic all-weather sky with a sun disk (SKY_LIGHT_USE_PEREZ_ENVIRONMENT) instead of a constant colour; uniform pdf table
        scn.add_light(M.sky_light((0.9, 1.0, 1.3), tab, perez=dict(sun_dir=(0.35, -0.75, -0.55), turbidity=2.8, sun_color=(6.0, 5.5, 4.5)) if perez else None))
    if with_area_light:
        l1 = scn.add_light(M.area_light((0.0, 3.5, 0.0), (1.0, 1.0), (12.0, 11.0, 9.0)))
        scn.add_instance(scn.add_mesh(S.quad_mesh(1.0, 1.0, y=0.0, mat_id=emi, flip=True)), S.translate(0.0, 3.5, 0.0), light_id=l1)
    return scn.build()
