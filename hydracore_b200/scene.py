"""Scene packing into the reference's blob formats and synthetic scene generators.

This plays the part of RenderDriverRTE::UpdateMesh/UpdateMaterial/UpdateLight/UpdateCamera + IHWLayerDataAssembler for
tests and benchmarks: it produces exactly what the CUDA layer (and the CPU oracle) receive through the IHWLayer boundary —
five storages ("textures", "textures_aux", "geom", "materials", "pdfs"), the EngineGlobals + tables blob, the flat BVH4,
inverse instance matrices and the instance -> light table.  Formats cited per function (reference file:line).
"""
import math
import numpy as np

from . import layout as L
from .layer import BvhBuilder

C = L.C
INVALID_TEXTURE = -2            # 0xFFFFFFFE as int32 (cglobals.h:16)


def _round_blocks(elems, block):            # roundBlocks, cglobals.h:628-634
    if elems < block:
        return block
    return ((elems + block - 1)//block)*block


def _as_f(i):
    return np.array([i], np.int32).view(np.float32)[0]


# ---------------------------------------------------------------------------------------------------------------- meshes
class Mesh:
    """Triangle mesh as HydraAPI hands it to UpdateMesh (HRMeshDriverInput): pos4f, norm4f, tan4f, texcoord2f, indices, matIndices."""

    def __init__(self, pos, idx, norm=None, uv=None, mat=None, tan=None):
        self.pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        self.idx = np.ascontiguousarray(idx, np.int32).reshape(-1, 3)
        n = self.pos.shape[0]
        self.norm = _vertex_normals(self.pos, self.idx) if norm is None else np.ascontiguousarray(norm, np.float32).reshape(n, 3)
        self.uv = np.zeros((n, 2), np.float32) if uv is None else np.ascontiguousarray(uv, np.float32).reshape(n, 2)
        self.mat = np.zeros(self.idx.shape[0], np.int32) if mat is None else np.ascontiguousarray(mat, np.int32).reshape(-1)
        self.tan = np.zeros((n, 4), np.float32) if tan is None else np.ascontiguousarray(tan, np.float32).reshape(n, 4)

    @property
    def tri_count(self):
        return self.idx.shape[0]

    def vert4f(self):
        v = np.zeros((self.pos.shape[0], 4), np.float32)
        v[:, :3] = self.pos
        v[:, 3] = 1.0
        return v

    def pack(self):
        """PlainMesh blob: 64-byte header + pos4f(w=u) + norm4f(w=v) + tan4f + indices + matIndices + shadowOffsets,
        every array padded to 16 bytes, offsets in float4 units from the header (RenderDriverRTE.cpp:1059-1159, cfetch.h:1038-1119)."""
        nv, nt = self.pos.shape[0], self.idx.shape[0]
        align = 16
        header_size = _round_blocks(64, align)
        pos_off = header_size
        pos_size = _round_blocks(16*nv, align)
        norm_off = pos_off + pos_size
        norm_size = _round_blocks(16*nv, align)
        texc_off = norm_off + norm_size
        tang_off = texc_off
        tang_size = _round_blocks(16*nv, align)
        ind_off = tang_off + tang_size
        ind_size = _round_blocks(nt*3*4, align)
        mind_off = ind_off + ind_size
        mind_size = _round_blocks(nt*4, align)
        soff_off = mind_off + mind_size
        soff_size = _round_blocks(nt*4, align)
        total = soff_off + soff_size
        blob = np.zeros(total, np.uint8)
        hdr = np.zeros(16, np.int32)
        hdr[0] = pos_off//16      # vPosOffset
        hdr[1] = norm_off//16     # vNormOffset
        hdr[2] = texc_off//16     # vTexCoordOffset
        hdr[3] = ind_off//16      # vIndicesOffset
        hdr[4] = nv               # vPosNum
        hdr[5] = nv               # vNormNum
        hdr[6] = nv               # vTexCoordNum
        hdr[7] = nt*3             # tIndicesNum
        hdr[8] = mind_off//16     # mIndicesOffset
        hdr[9] = nt               # mIndicesNum
        hdr[10] = tang_off//16    # vTangentOffset
        hdr[11] = nv              # vTangentNum
        hdr[12] = total           # totalBytesNum
        hdr[13] = soff_off//16    # polyShadowOffset
        blob[0:64] = hdr.view(np.uint8)
        p4 = np.zeros((nv, 4), np.float32)
        p4[:, :3] = self.pos
        p4[:, 3] = self.uv[:, 0]
        n4 = np.zeros((nv, 4), np.float32)
        n4[:, :3] = self.norm
        n4[:, 3] = self.uv[:, 1]
        blob[pos_off:pos_off + 16*nv] = p4.view(np.uint8).reshape(-1)
        blob[norm_off:norm_off + 16*nv] = n4.view(np.uint8).reshape(-1)
        blob[tang_off:tang_off + 16*nv] = self.tan.view(np.uint8).reshape(-1)
        blob[ind_off:ind_off + nt*12] = self.idx.view(np.uint8).reshape(-1)
        blob[mind_off:mind_off + nt*4] = self.mat.view(np.uint8).reshape(-1)
        blob[soff_off:soff_off + nt*4] = self.shadow_offsets().view(np.uint8).reshape(-1)
        return blob

    def shadow_offsets(self):
        """CalcAuxShadowRaysOffsets (RenderDriverRTE.cpp:990-1056): min(0.05*sqrt(area), 0.00025*bboxMax) for smooth-shaded tris, else 0."""
        A, B, Cc = self.pos[self.idx[:, 0]], self.pos[self.idx[:, 1]], self.pos[self.idx[:, 2]]
        ext = (self.pos[self.idx.reshape(-1)].max(0) - self.pos[self.idx.reshape(-1)].min(0)).astype(np.float32)
        mesh_max = np.float32(0.00025)*ext.max()
        crpd = np.cross(A - B, A - Cc).astype(np.float32)
        ln = np.sqrt((crpd*crpd).sum(1)).astype(np.float32)
        fn = crpd/np.maximum(ln, np.float32(1e-30))[:, None]
        diff = np.zeros(self.idx.shape[0], np.float32)
        for k in range(3):
            nk = self.norm[self.idx[:, k]]
            d = fn - nk
            diff += np.sqrt((d*d).sum(1)).astype(np.float32)
        off = np.minimum(np.float32(0.05)*np.sqrt(ln*np.float32(0.5)), mesh_max).astype(np.float32)
        return np.where(diff > np.float32(0.001), off, np.float32(0.0)).astype(np.float32)


def _vertex_normals(pos, idx):
    fn = np.cross(pos[idx[:, 1]] - pos[idx[:, 0]], pos[idx[:, 2]] - pos[idx[:, 0]])
    n = np.zeros_like(pos)
    for k in range(3):
        np.add.at(n, idx[:, k], fn)
    ln = np.sqrt((n*n).sum(1))
    ln[ln == 0] = 1.0
    return (n/ln[:, None]).astype(np.float32)


def grid_mesh(nx, ny, size=10.0, amplitude=0.35, seed=1234, mat_blocks=None, flat=False):
    """Displaced nx x ny-quad grid in the XZ plane (2*nx*ny triangles): the synthetic "1M-triangle mesh" is grid_mesh(708, 707)."""
    rng = np.random.RandomState(seed)
    xs = np.linspace(-0.5*size, 0.5*size, nx + 1, dtype=np.float32)
    zs = np.linspace(-0.5*size, 0.5*size, ny + 1, dtype=np.float32)
    X, Z = np.meshgrid(xs, zs)
    k = 2.0*math.pi/size
    Y = (amplitude*(np.sin(3.0*k*X)*np.cos(2.0*k*Z) + 0.5*np.sin(7.0*k*X + 1.3)*np.sin(5.0*k*Z + 0.7))
         + 0.15*amplitude*rng.standard_normal(X.shape)).astype(np.float32)
    pos = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float32)
    uv = np.stack([(X/size + 0.5), (Z/size + 0.5)], -1).reshape(-1, 2).astype(np.float32)
    i0 = (np.arange(ny)[:, None]*(nx + 1) + np.arange(nx)[None, :]).reshape(-1)
    tri = np.empty((nx*ny, 2, 3), np.int64)
    tri[:, 0] = np.stack([i0, i0 + nx + 1, i0 + 1], -1)
    tri[:, 1] = np.stack([i0 + 1, i0 + nx + 1, i0 + nx + 2], -1)
    idx = tri.reshape(-1, 3).astype(np.int32)
    mat = None
    if mat_blocks is not None:
        # material id by contiguous triangle block: mat_blocks = [(fraction, matId), ...]
        nt = idx.shape[0]
        mat = np.zeros(nt, np.int32)
        start = 0
        for frac, mid in mat_blocks:
            end = min(nt, start + int(round(frac*nt)))
            mat[start:end] = mid
            start = end
        mat[start:] = mat_blocks[-1][1]
    m = Mesh(pos, idx, uv=uv, mat=mat)
    if flat:
        m.norm[:] = np.array([0, 1, 0], np.float32)
    return m


def quad_mesh(sx, sz, y=0.0, mat_id=0, flip=False):
    """Rectangle in the XZ plane, half sizes (sx, sz), normal +Y (or -Y when flip)."""
    pos = np.array([[-sx, y, -sz], [sx, y, -sz], [sx, y, sz], [-sx, y, sz]], np.float32)
    idx = np.array([[0, 2, 1], [0, 3, 2]], np.int32) if not flip else np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    nrm = np.tile(np.array([[0, -1.0 if flip else 1.0, 0]], np.float32), (4, 1))
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    return Mesh(pos, idx, norm=nrm, uv=uv, mat=np.full(2, mat_id, np.int32))


def box_mesh(hx, hy, hz, mat_ids=(0, 0, 0, 0, 0, 0), inward=True, skip_faces=()):
    """Axis-aligned box of half sizes (hx, hy, hz), 12 triangles, flat normals; inward=True gives a room (Cornell-box style)."""
    faces = []
    s = -1.0 if inward else 1.0
    defs = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)),
            ((0, -1, 0), (1, 0, 0), (0, 0, 1)), ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0), (1, 0, 0))]
    h = np.array([hx, hy, hz], np.float32)
    pos, idx, nrm, uv, mat = [], [], [], [], []
    for f, (n, u, v) in enumerate(defs):
        if f in skip_faces:
            continue
        n, u, v = np.array(n, np.float32), np.array(u, np.float32), np.array(v, np.float32)
        c = n*h
        corners = [c - u*h - v*h, c + u*h - v*h, c + u*h + v*h, c - u*h + v*h]
        b = len(pos)
        pos += corners
        nrm += [s*n]*4
        uv += [[0, 0], [1, 0], [1, 1], [0, 1]]
        tri = [[b, b + 1, b + 2], [b, b + 2, b + 3]]
        if inward:
            tri = [[t[0], t[2], t[1]] for t in tri]
        idx += tri
        mat += [mat_ids[f]]*2
        faces.append(f)
    return Mesh(np.array(pos, np.float32), np.array(idx, np.int32), norm=np.array(nrm, np.float32), uv=np.array(uv, np.float32),
                mat=np.array(mat, np.int32))


def sphere_mesh(radius, nu, nv, mat_id=0):
    """UV sphere with smooth normals, 2*nu*(nv-1) triangles."""
    th = np.linspace(0, math.pi, nv + 1)
    ph = np.linspace(0, 2*math.pi, nu + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    n = np.stack([np.sin(T)*np.cos(P), np.cos(T), np.sin(T)*np.sin(P)], -1).reshape(-1, 3).astype(np.float32)
    pos = (radius*n).astype(np.float32)
    uv = np.stack([P/(2*math.pi), T/math.pi], -1).reshape(-1, 2).astype(np.float32)
    idx = []
    for i in range(nv):
        for j in range(nu):
            a = i*(nu + 1) + j
            b = a + nu + 1
            if i != 0:
                idx.append([a, a + 1, b])
            if i != nv - 1:
                idx.append([a + 1, b + 1, b])
    idx = np.array(idx, np.int32)
    return Mesh(pos, idx, norm=n, uv=uv, mat=np.full(idx.shape[0], mat_id, np.int32))


def cylinder_mesh(radius, height, nu, phi_max=2*math.pi, mat_id=0):
    """Open tube section around the local Z axis (the parameterisation of CylinderLightSamplePos, clight.h:785-799): z in [-h/2, h/2],
    phi in [0, phi_max]; outward normals, uv = (z fraction, phi fraction), 2*nu triangles."""
    ph = np.linspace(0, phi_max, nu + 1)
    pos, nrm, uv = [], [], []
    for k, z in enumerate((-0.5*height, 0.5*height)):
        for j, p in enumerate(ph):
            pos.append([radius*math.cos(p), radius*math.sin(p), z])
            nrm.append([math.cos(p), math.sin(p), 0.0])
            uv.append([float(k), j/float(nu)])
    idx = []
    for j in range(nu):
        a, b = j, j + nu + 1
        idx += [[a, a + 1, b], [a + 1, b + 1, b]]
    idx = np.array(idx, np.int32)
    return Mesh(np.array(pos, np.float32), idx, norm=np.array(nrm, np.float32), uv=np.array(uv, np.float32), mat=np.full(idx.shape[0], mat_id, np.int32))


# ---------------------------------------------------------------------------------------------------------------- matrices
def translate(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (x, y, z)
    return m


def scale(x, y, z):
    return np.diag(np.array([x, y, z, 1.0], np.float32))


def rotate_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], np.float32)


def rotate_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], np.float32)


def look_at(eye, center, up):                    # lookAt, cglobals.h:1006-1038 (row-major numpy; stored as columns below)
    eye, center, up = (np.asarray(v, np.float64) for v in (eye, center, up))
    z = eye - center
    z /= np.linalg.norm(z)
    x = np.cross(up, z)
    y = np.cross(z, x)
    x /= np.linalg.norm(x)
    y /= np.linalg.norm(y)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = x, y, z
    m[:3, 3] = (-x.dot(eye), -y.dot(eye), -z.dot(eye))
    return m


def perspective(fov_deg, aspect, z_near, z_far):  # OpenGL frustum (LiteMath perspectiveMatrix)
    ymax = z_near*math.tan(fov_deg*math.pi/360.0)
    xmax = ymax*aspect
    m = np.zeros((4, 4))
    m[0, 0] = z_near/xmax
    m[1, 1] = z_near/ymax
    m[2, 2] = -(z_far + z_near)/(z_far - z_near)
    m[2, 3] = -2.0*z_far*z_near/(z_far - z_near)
    m[3, 2] = -1.0
    return m


def _cols(m):
    """4x4 row-major numpy matrix -> 16 floats in column storage (float4x4::m_col, make_float4x4 cglobals.h:790-798)."""
    return np.ascontiguousarray(np.asarray(m, np.float64).T, np.float32).reshape(16)


# ---------------------------------------------------------------------------------------------------------------- scene
class Camera:
    def __init__(self, pos=(0, 0, 15), look_at=(0, 0, 0), up=(0, 1, 0), fov=45.0, near=0.1, far=1000.0, dof=False, lens_radius=0.0):
        self.pos, self.look_at, self.up, self.fov, self.near, self.far = pos, look_at, up, fov, near, far
        self.dof, self.lens_radius = dof, lens_radius


class Scene:
    """Host-side scene description -> blobs.  See materials.py for PlainMaterial / PlainLight packing."""

    def __init__(self, width, height, camera=None):
        self.width, self.height = int(width), int(height)
        self.camera = camera or Camera()
        self.meshes = []          # Mesh
        self.instances = []       # (meshId, 4x4 row-major matrix, lightId or -1)
        self.materials = []       # list of 192-float PlainMaterial nodes (a material may span several consecutive nodes)
        self.material_ids = []    # matId -> index of its head node in self.materials
        self.lights = []          # list of 128-float PlainLight
        self.textures = []        # list of (w, h, rgba8 ndarray) ; texture id 0 is reserved ("no texture")
        self.pdf_tables = []      # list of float32 arrays {as_float(w), as_float(h), as_float(1), as_float(n+1), prefix sums..., 1.0} (sky-dome lights)
        self.varsI = np.zeros(64, np.int32)
        self.varsF = np.zeros(64, np.float32)
        self.flags = 0
        self.ms_tables = None     # (ggx u16[64*64], transp u16[64*64*64]) baked tables (bakeBrdfEnergy/MSTables*.cpp), optional
        self.varsF[C["HRT_ABLOW_SCALE_X"]] = 1.0      # AllRenderVarialbes ctor, IHWLayer.h:25-35
        self.varsF[C["HRT_ABLOW_SCALE_Y"]] = 1.0
        self.varsI[C["HRT_SHADOW_MATTE_BACK"]] = INVALID_TEXTURE
        self.varsF[C["HRT_IMAGE_GAMMA"]] = 2.2
        self.varsF[C["HRT_TEXINPUT_GAMMA"]] = 2.2
        self.set_trace_depth(5, 3)

    def set_trace_depth(self, trace_depth, diff_trace_depth):
        """XML trace_depth / diff_trace_depth are stored +1 (RenderDriverRTE.cpp:317-321)."""
        self.varsI[C["HRT_TRACE_DEPTH"]] = trace_depth + 1
        self.varsI[C["HRT_DIFFUSE_TRACE_DEPTH"]] = diff_trace_depth + 1

    def add_mesh(self, mesh):
        self.meshes.append(mesh)
        return len(self.meshes) - 1

    def add_instance(self, mesh_id, matrix=None, light_id=-1, remap_list=-1):
        """remap_list: id from add_remap_list - this instance renders with some material ids replaced (HydraAPI remap lists)."""
        self.instances.append((mesh_id, np.eye(4, dtype=np.float32) if matrix is None else np.asarray(matrix, np.float32), light_id))
        self.inst_remap = getattr(self, "inst_remap", [])
        self.inst_remap += [-1]*(len(self.instances) - 1 - len(self.inst_remap)) + [int(remap_list)]
        return len(self.instances) - 1

    def add_remap_list(self, pairs):
        """pairs: [(from_material_id, to_material_id), ...]; stored sorted by `from` (remapMaterialId does a binary search, cglobals.h:2931-2983)."""
        self.remap_lists = getattr(self, "remap_lists", [])
        self.remap_lists.append(sorted((int(a), int(b)) for a, b in pairs))
        return len(self.remap_lists) - 1

    def remap_arrays(self):
        """What RenderDriverRTE hands to SetAllRemapLists / SetAllInstIdToRemapId (RenderDriverRTE.cpp:1340-1376, 1478): all lists back to back,
        {offset, size in INTS} per list, list id per instance."""
        lists = getattr(self, "remap_lists", [])
        flat, table = [], []
        for l in lists:
            table.append((len(flat), 2*len(l)))
            for a, b in l:
                flat += [a, b]
        inst = list(getattr(self, "inst_remap", [])) + [-1]*(len(self.instances) - len(getattr(self, "inst_remap", [])))
        return np.array(flat, np.int32), np.array(table, np.int32).reshape(-1, 2), np.array(inst, np.int32)

    def add_material(self, nodes):
        """nodes: one PlainMaterial (192 floats) or a list of consecutive nodes (blend trees reference children by relative offset)."""
        nodes = [nodes] if isinstance(nodes, np.ndarray) and nodes.ndim == 1 else list(nodes)
        head = len(self.materials)
        self.materials += [np.ascontiguousarray(n, np.float32).reshape(192) for n in nodes]
        self.material_ids.append(head)
        return len(self.material_ids) - 1

    def add_light(self, plain_light):
        self.lights.append(np.ascontiguousarray(plain_light, np.float32).reshape(128))
        return len(self.lights) - 1

    def add_cylinder_light(self, matrix, radius, height, angle_deg, intensity, tex_id=0, tex_lum=None, **sampler_kw):
        """Cylinder light (CylinderLight + CreateCylinderLightFromXmlNode + Transform, PlainLightConverter.cpp:353-413, 867-893) with the 2x2
        uniform pdf table RenderDriverRTE::UpdatePdfTablesForLight builds for an untextured one (RenderDriverRTE.cpp:940-941)."""
        from . import materials as M
        f = np.float32
        mtx = np.asarray(matrix, np.float32).reshape(4, 4)
        z_min, z_max = f(-0.5)*f(height), f(0.5)*f(height)
        phi_max = f(f(np.pi/180.0)*f(angle_deg))
        L = M.point_light(tuple(mtx[:3, 3]), intensity)
        Li = L.view(np.int32)
        Li[C["PLIGHT_TYPE"]] = C["PLAIN_LIGHT_TYPE_CYLINDER"]
        L[16:25] = mtx[:3, :3].reshape(9)                                  # CYLINDER_LIGHT_MATRIX_E00.., rows
        L[25], L[26], L[27], L[28] = radius, z_min, z_max, phi_max
        Li[29], Li[30] = INVALID_TEXTURE, INVALID_TEXTURE                  # CYLINDER_TEX_ID / TEXMATRIX_ID
        if tex_id and tex_id > 0:                                          # PutSamplerAt(..., CYLINDER_TEX_ID, CYLINDER_TEXMATRIX_ID, CYLINDER_TEX_SAMPLER = 32), PlainLightConverter.cpp:374
            Li[29] = tex_id
            Li[30] = M._sampler(L, 32, tex_id, **sampler_kw)
        # CYLINDER_PDF_TABLE_ID: the 2x2 table of 0.25 an untextured light gets, or the table of the texture's luminance image (tex_lum) the
        # driver builds for a textured one (UpdatePdfTablesForLight, RenderDriverRTE_PdfTables.cpp)
        Li[31] = self.add_sky_pdf_table(tex_lum)
        d = f(1.0)/np.sqrt(f(3.0), dtype=f)
        vert = mtx[:3, :3] @ np.array([d, d, d], f)                        # mul(mrot, normalize(float3(1,1,1)))
        mult = np.sqrt((vert*vert).sum(dtype=f), dtype=f)
        L[C["PLIGHT_SURFACE_AREA"]] = f(f(f(f(z_max - z_min)*mult)*f(f(radius)*mult))*phi_max)
        return self.add_light(L)

    def add_mesh_light(self, mesh_id, matrix, intensity, tex_id=0, **sampler_kw):
        """Mesh light (MeshLight + Transform, PlainLightConverter.cpp:724-830): the light samples the triangles of `mesh_id` by area.  A copy of
        the packed mesh and the prefix sums of its triangle areas (CalcTrianglePickProbTable, RenderDriverRTE_PdfTables.cpp:650-674) go into the
        "pdfs" storage; position, rotation sub-matrix and the transformed surface area into the light.  Returns the light id; instance the mesh
        with the same matrix, an emissive material and light_id = this id so that paths can hit it."""
        from . import materials as M
        f = np.float32
        mesh = self.meshes[mesh_id]
        mtx = np.asarray(matrix, np.float32).reshape(4, 4)
        P = mesh.pos.astype(np.float32)

        def areas(pts):
            A, B, Cc = pts[mesh.idx[:, 0]], pts[mesh.idx[:, 1]], pts[mesh.idx[:, 2]]
            e1, e2 = (B - A).astype(f), (Cc - A).astype(f)
            cx = (e1[:, 1]*e2[:, 2]).astype(f) - (e1[:, 2]*e2[:, 1]).astype(f)
            cy = (e1[:, 2]*e2[:, 0]).astype(f) - (e1[:, 0]*e2[:, 2]).astype(f)
            cz = (e1[:, 0]*e2[:, 1]).astype(f) - (e1[:, 1]*e2[:, 0]).astype(f)
            ln = np.sqrt(((cx*cx).astype(f) + (cy*cy).astype(f)).astype(f) + (cz*cz).astype(f), dtype=f)
            return (f(0.5)*ln).astype(f)

        tri = areas(P)
        table = np.zeros(tri.size + 1, f)
        acc = f(0)
        for i, a in enumerate(tri):
            table[i] = acc
            acc = f(acc + a)
        table[tri.size] = acc
        self.pdf_tables.append(mesh.pack())                      # the PlainMesh copy
        self.pdf_tables.append(table)
        # mul(a_matrix, p) of LiteMath: row-major rows dotted with (p, 1), sums left to right
        Pw = np.stack([((mtx[r, 0]*P[:, 0]).astype(f) + (mtx[r, 1]*P[:, 1]).astype(f)).astype(f) + (mtx[r, 2]*P[:, 2]).astype(f) + mtx[r, 3] for r in range(3)], 1).astype(f)
        total = 0.0
        for a in areas(Pw):
            total += float(a)
        L = M.point_light(tuple(mtx[:3, 3]), intensity)
        Li = L.view(np.int32)
        Li[C["PLIGHT_TYPE"]] = C["PLAIN_LIGHT_TYPE_MESH"]
        Li[14], Li[15], Li[16] = len(self.pdf_tables) - 2, len(self.pdf_tables) - 1, tri.size     # MESH_LIGHT_MESH_OFFSET_ID / TABLE_OFFSET_ID / TRI_NUM
        L[20:29] = mtx[:3, :3].reshape(9)                                                              # MESH_LIGHT_MATRIX_E00, rows
        Li[30], Li[31] = INVALID_TEXTURE, INVALID_TEXTURE                                              # MESH_LIGHT_TEX_ID / TEXMATRIX_ID
        if tex_id and tex_id > 0:                                                                      # PutSamplerAt(..., MESH_LIGHT_TEX_SAMPLER = 32), PlainLightConverter.cpp:782
            Li[30] = tex_id
            Li[31] = M._sampler(L, 32, tex_id, **sampler_kw)
        L[C["PLIGHT_SURFACE_AREA"]] = f(total)
        return self.add_light(L)

    def add_sky_pdf_table(self, lum=None):
        """Pdf table of a sky-dome light as RenderDriverRTE::UpdatePdfTablesForLight builds it (RenderDriverRTE_PdfTables.cpp:520-566): prefix sums of
        the luminance image (2x2 of 0.25 for an untextured sky); returns the table id to put into SKY_DOME_PDF_TABLE0."""
        lum = np.full((2, 2), 0.25, np.float32) if lum is None else np.ascontiguousarray(lum, np.float32)
        h, w = lum.shape
        pref = np.zeros(w*h + 1, np.float32)
        acc = np.float32(0)
        for i, v in enumerate(lum.reshape(-1)):
            pref[i] = acc
            acc = np.float32(acc + v)
        pref[w*h] = acc
        data = np.zeros(pref.size + 5, np.float32)
        data[0:4] = np.array([w, h, 1, pref.size + 1], np.int32).view(np.float32)
        data[4:4 + pref.size] = pref
        data[-1] = 1.0
        self.pdf_tables.append(data)
        return len(self.pdf_tables) - 1

    def add_ies_table(self, web):
        """IES photometric web as AddIesTexTableToStorage stores it (RenderDriverRTE_PdfTables.cpp:385-443): single-channel float image over the
        sphere, normalised to its maximum, behind an image header {w, h, 1, 4}; returns its table id in the "pdfs" storage."""
        web = np.ascontiguousarray(web, np.float32)
        h, w = web.shape
        web = (web*(np.float32(1.0)/web.max())).astype(np.float32)
        data = np.zeros(web.size + 5, np.float32)
        data[0:4] = np.array([w, h, 1, 4], np.int32).view(np.float32)
        data[4:4 + web.size] = web.reshape(-1)
        self.pdf_tables.append(data)
        return len(self.pdf_tables) - 1

    def add_texture_rgba8(self, rgba):
        """Texture ids start at 1; id 0 means "white" (sample2D, cfetch.h:654-655)."""
        rgba = np.ascontiguousarray(rgba, np.uint8)
        assert rgba.ndim == 3 and rgba.shape[2] == 4
        self.textures.append(rgba)
        return len(self.textures)

    def add_normal_map(self, rgba):
        """Normal map (RGBA8): the driver keeps such images in the "textures_aux" storage under the same texture id
        (RenderDriverRTE::UpdateImageAux; sample2DAux looks the id up in the aux texture table, cfetch.h:766-789)."""
        tid = self.add_texture_rgba8(rgba)
        self.aux_ids = getattr(self, "aux_ids", set()) | {tid}
        return tid

    def add_texture_f4(self, rgba):
        """HDR texture: float4 texels (bpp 16, cfetch.h:364-584), e.g. an environment map."""
        rgba = np.ascontiguousarray(rgba, np.float32)
        assert rgba.ndim == 3 and rgba.shape[2] == 4
        self.textures.append(rgba)
        return len(self.textures)

    # ---- build everything the layer receives
    def build(self):
        # geometry storage + geometry table (float4 offsets)
        geom_chunks, geom_table, off = [], [], 0
        for m in self.meshes:
            b = m.pack()
            geom_table.append(off//16)
            geom_chunks.append(b)
            off += b.size
        geom = np.concatenate(geom_chunks) if geom_chunks else np.zeros(16, np.uint8)

        # materials storage: node k lives at float4 offset 48*k ; materialsTable[matId] -> offset of the head node
        if not self.materials:
            raise ValueError("scene has no materials")
        mats = np.stack(self.materials).astype(np.float32)
        mat_table = [48*h for h in self.material_ids]

        # textures storage: slot 0 unused; each texture = int4{w,h,depth=4,bpp=4} + texels (cfetch.h:364-584)
        tex_chunks, tex_table, off = [np.zeros(16, np.uint8)], [-1], 16
        for t in self.textures:
            h, w = t.shape[0], t.shape[1]
            hdr = np.array([w, h, 4, 16 if t.dtype == np.float32 else 4], np.int32).view(np.uint8)
            body = t.reshape(-1).view(np.uint8)
            pad = (-body.size) % 16
            chunk = np.concatenate([hdr, body, np.zeros(pad, np.uint8)])
            tex_table.append(off//16)
            tex_chunks.append(chunk)
            off += chunk.size
        textures = np.concatenate(tex_chunks)
        aux_chunks, self._aux_table, off = [np.zeros(16, np.uint8)], [-1]*len(tex_table), 16
        for tid in sorted(getattr(self, "aux_ids", ())):
            t = self.textures[tid - 1]
            body = t.reshape(-1).view(np.uint8)
            chunk = np.concatenate([np.array([t.shape[1], t.shape[0], 4, 4], np.int32).view(np.uint8), body, np.zeros((-body.size) % 16, np.uint8)])
            self._aux_table[tid] = off//16
            aux_chunks.append(chunk)
            off += chunk.size
        textures_aux = np.concatenate(aux_chunks)

        # BVH: tree 0 = opaque meshes, tree 1 = meshes with at least one opacity-mapped material (MeshHaveOpacity, RenderDriverRTE.cpp:1989-1991);
        # instance ids run over both trees in scene order, the inverse matrices are indexed by them
        from . import materials as M
        mat_opacity = [M.opacity_tex_id(self.materials[h]) for h in self.material_ids]
        mesh_alpha = [bool(any(mat_opacity[int(k)] > 0 for k in np.unique(m.mat))) for m in self.meshes]   # > 0: zero-filled placeholder materials of the ray-casting scenes have no map
        trees, lo, hi = [], None, None
        inv_all = np.zeros((len(self.instances), 16), np.float32)
        for tree_id in (0, 1):
            ids = [i for i, (mid, _m, _l) in enumerate(self.instances) if mesh_alpha[mid] == bool(tree_id)]
            if not ids:
                trees.append(None)
                continue
            bb = BvhBuilder()
            for m in self.meshes:                  # every mesh, so that builder mesh ids = scene mesh ids = geomId of the triangles; unused ones are not built
                bb.add_mesh(m.vert4f(), m.idx)
            for i in ids:
                mid, mat, _l = self.instances[i]
                bb.add_instance(mid, mat, real_id=i)
            t = bb.commit()
            inv_all[ids] = t["inv_matrices"]
            blo, bhi = bb.bounds()
            lo, hi = (blo, bhi) if lo is None else (np.minimum(lo, blo), np.maximum(hi, bhi))
            bb.close()
            trees.append(t)
        if trees[0] is None:
            raise ValueError("the scene needs at least one instance of a mesh without opacity maps (tree 0)")
        self.bvh = trees[0]
        self.bvh["inv_matrices"] = inv_all
        self.bvh1 = trees[1]
        if self.bvh1 is not None:
            self.bvh1["alpha"] = self._alpha_table(self.bvh1["tris"], mat_opacity)
        half = np.float32(0.5)*(hi - lo)                     # scene bounding sphere, RenderDriverRTE.cpp:1461-1467
        self.bsphere = np.concatenate([np.float32(0.5)*(hi + lo), [np.sqrt((half*half).sum(dtype=np.float32), dtype=np.float32)]]).astype(np.float32)
        self.inst_light_ids = np.array([l for (_m, _x, l) in self.instances], np.int32)

        # pdfs storage: tables padded to 16 bytes, pdfTableTable[id] -> float4 offset
        pdf_chunks, self._pdf_table_offsets, off = [], [], 0
        for t in self.pdf_tables:
            b = t.view(np.uint8)
            pad = (-b.size) % 16
            self._pdf_table_offsets.append(off//16)
            pdf_chunks.append(np.concatenate([b, np.zeros(pad, np.uint8)]))
            off += b.size + pad
        pdfs = np.concatenate(pdf_chunks) if pdf_chunks else np.zeros(16, np.uint8)
        self.storages = dict(textures=textures, textures_aux=textures_aux, geom=geom, materials=mats.view(np.uint8).reshape(-1),
                             pdfs=pdfs)
        self.globals_blob = self._pack_globals(geom_table, mat_table, tex_table)
        return self

    def _alpha_table(self, tris, mat_opacity):
        """RenderDriverRTE::CreateAlphaTestTable (RenderDriverRTE_AlphaTestTable.cpp:66-221): one uint2 per float4 of the triangle list -
        {offset of the opacity sampler | smooth flag | skip-shadow flag, texture coordinate packed to 2 x 16 bits} for the three float4 of a
        triangle, (-1, -1) for leaf headers - followed by the SWTexSampler copies (6 uint2 each), one per material with an opacity map."""
        n = tris.shape[0]
        ti = tris.view(np.int32)
        with_alpha = [k for k, t in enumerate(mat_opacity) if t > 0]
        out = np.zeros((n + 6*len(with_alpha), 2), np.uint32)
        samplers, k = {}, 0

        def wrap(v):                                   # WrapVal, cglobals.h:3014-3022
            v = np.float32(v)
            if v > 1.0:
                return np.float32(v - np.float32(int(v)))
            if v < -1.0:
                return np.float32(np.float32(int(v)) - v)
            return v

        def pack_tc(u, v):                             # CompressTexCoord16, RenderDriverRTE_AlphaTestTable.cpp:25-38
            tx = min(max(np.float32(np.float32(0.5)*wrap(u) + np.float32(0.5)), np.float32(0)), np.float32(1))
            ty = min(max(np.float32(np.float32(0.5)*wrap(v) + np.float32(0.5)), np.float32(0)), np.float32(1))
            return (int(np.float32(ty*np.float32(65535.0))) << 16) | int(np.float32(tx*np.float32(65535.0)))

        while k < n:
            if ti[k, 2] == -1 and ti[k, 3] == -1:      # object-list header
                out[k] = (0xFFFFFFFF, 0xFFFFFFFF)
                k += 1
                continue
            prim, geom = int(ti[k, 3]), int(ti[k + 1, 3])
            mesh = self.meshes[geom]
            mid = int(mesh.mat[prim])
            tex = mat_opacity[mid]
            if tex > 0:
                if mid not in samplers:
                    samplers[mid] = len(samplers)
                    head = self.materials[self.material_ids[mid]]
                    s0 = C["OPACITY_SAMPLER_OFFSET"]
                    out[n + 6*samplers[mid]:n + 6*samplers[mid] + 6] = np.ascontiguousarray(head[s0:s0 + 12], np.float32).view(np.uint32).reshape(6, 2)
                out[k, 0] = n + 6*samplers[mid]
                out[k + 1, 0] = 0                      # smoothOpacity
                out[k + 2, 0] = 0                      # skipShadow
                for j in range(3):
                    vi = int(mesh.idx[prim, j])
                    out[k + j, 1] = pack_tc(mesh.uv[vi, 0], mesh.uv[vi, 1])
            else:
                out[k:k + 3] = (np.uint32(INVALID_TEXTURE & 0xFFFFFFFF), 0xFFFFFFFF)
            k += 3
        return out

    def _pack_globals(self, geom_table, mat_table, tex_table):
        """EngineGlobals + tables blob (cfetch.h:21-81; CalcConstGlobDataOffsets / PrepareEngineGlobals / PrepareEngineTables /
        SetAllPODLights / SetCamMatrices, IHWLayerDataAssembler.cpp:94-452)."""
        W, H = self.width, self.height
        cam = self.camera
        nl = len(self.lights)
        sizes = dict(materials=len(mat_table), geometry=len(geom_table), textures=len(tex_table), texturesAux=len(tex_table),
                     pdfTable=max(nl, 1, len(self.pdf_tables)), lselRev=(nl + 1 if nl > 0 else 0), lselFwd=(nl + 1 if nl > 0 else 0), floats=0, lights=nl*128)
        cur = _round_blocks(C["EG_sizeof"]//4, 16)
        offs = {}
        for k in ("materials", "geometry", "textures", "texturesAux", "pdfTable", "lselRev", "lselFwd", "floats", "lights"):
            offs[k] = cur
            cur += _round_blocks(sizes[k], 16)
        blob = np.zeros(cur, np.int32)
        b8 = blob.view(np.uint8)

        def put_i(byte_off, v):
            blob[byte_off//4] = v

        def put_f(byte_off, arr):
            a = np.ascontiguousarray(arr, np.float32).reshape(-1)
            blob[byte_off//4:byte_off//4 + a.size] = a.view(np.int32)

        aspect = float(W)/float(H)
        proj = perspective(cam.fov, aspect, cam.near, cam.far)
        view = look_at(cam.pos, cam.look_at, cam.up)
        put_f(C["EG_mProj"], _cols(proj))
        put_f(C["EG_mWorldView"], _cols(view))
        put_f(C["EG_mProjInverse"], _cols(np.linalg.inv(proj)))
        put_f(C["EG_mWorldViewInverse"], _cols(np.linalg.inv(view)))

        varsI, varsF = self.varsI.copy(), self.varsF.copy()
        fov_rad = np.float32(math.pi/180.0)*np.float32(cam.fov)
        varsF[C["HRT_CAM_FOV"]] = fov_rad                              # UpdateCamera, RenderDriverRTE.cpp:1243
        varsF[C["HRT_FOV_X"]] = fov_rad                                # SetCamMatrices, IHWLayerDataAssembler.cpp:105-109
        varsF[C["HRT_FOV_Y"]] = fov_rad/np.float32(aspect)
        varsF[C["HRT_WIDTH_F"]] = W
        varsF[C["HRT_HEIGHT_F"]] = H
        varsF[C["HRT_DOF_FOCAL_PLANE_DIST"]] = np.linalg.norm(np.asarray(cam.pos, np.float64) - np.asarray(cam.look_at, np.float64))
        varsI[C["HRT_ENABLE_DOF"]] = 1 if cam.dof else 0
        varsF[C["HRT_DOF_LENS_RADIUS"]] = cam.lens_radius if cam.dof else 0.0
        varsF[18:22] = self.bsphere                                     # HRT_BSPHERE_CENTER_X..Z, HRT_BSPHERE_RADIUS (RenderDriverRTE.cpp:1483-1486)
        blob[C["EG_varsI"]//4:C["EG_varsI"]//4 + 64] = varsI
        put_f(C["EG_varsF"], varsF)

        rm = np.full(16, -1, np.int32)                                 # SetQMCVarRemapTable, IHWLayerDataAssembler.cpp:211-323
        variant = int(varsI[C["HRT_QMC_VARIANT"]])
        if varsI[C["HRT_ENABLE_DOF"]] != 1 and (variant & 1):
            variant -= 1
        table = {0: dict(SCR_X=0, SCR_Y=1, DOF_X=2, DOF_Y=3), 1: dict(SCR_X=0, SCR_Y=1, DOF_X=2, DOF_Y=3),
                 2: dict(SCR_X=0, SCR_Y=1, MAT_L=2, MAT_0=3, MAT_1=4),
                 3: dict(SCR_X=0, SCR_Y=1, DOF_X=2, DOF_Y=3, MAT_L=4, MAT_0=5, MAT_1=6),
                 4: dict(SCR_X=0, SCR_Y=1, LGT_N=2, LGT_0=3, LGT_1=4, LGT_2=5),
                 5: dict(SCR_X=0, SCR_Y=1, DOF_X=2, DOF_Y=3, LGT_N=4, LGT_0=5, LGT_1=6, LGT_2=7),
                 6: dict(SCR_X=0, SCR_Y=1, MAT_L=2, MAT_0=3, MAT_1=4, LGT_N=5, LGT_0=6, LGT_1=7, LGT_2=8),
                 7: dict(SCR_X=0, SCR_Y=1, DOF_X=2, DOF_Y=3, MAT_L=4, MAT_0=5, MAT_1=6, LGT_N=7, LGT_0=8, LGT_1=9, LGT_2=10)}
        for k, v in table.get(variant, table[0]).items():
            rm[C["QMC_VAR_" + k]] = v
        blob[C["EG_rmQMC"]//4:C["EG_rmQMC"]//4 + 16] = rm

        fwd = np.asarray(cam.look_at, np.float64) - np.asarray(cam.pos, np.float64)
        fwd /= np.linalg.norm(fwd)
        put_f(C["EG_camForward"], fwd)                                 # camForward[3]
        upv = np.linalg.inv(view)[:3, 1]
        put_f(C["EG_camForward"] + 12, upv/np.linalg.norm(upv))        # camUpVector[3]
        put_f(C["EG_camForward"] + 24, np.asarray(cam.look_at, np.float32))
        put_f(C["EG_imagePlaneDist"], [W/(2.0*math.tan(0.5*float(fov_rad)))])

        for key, name in (("materials", "materialsTable"), ("geometry", "geometryTable"), ("textures", "texturesTable"),
                          ("texturesAux", "texturesAuxTable"), ("pdfTable", "pdfTableTable")):
            put_i(C["EG_" + name + "Offset"], offs[key])
            put_i(C["EG_" + name + "Size"], sizes[key])
        put_i(C["EG_lightSelectorTableOffsetRev"], offs["lselRev"])
        put_i(C["EG_lightSelectorTableSizeRev"], sizes["lselRev"])
        put_i(C["EG_lightSelectorTableOffsetFwd"], offs["lselFwd"])
        put_i(C["EG_lightSelectorTableSizeFwd"], sizes["lselFwd"])
        put_i(C["EG_floatArraysOffset"], offs["floats"])
        put_i(C["EG_floatsArraysSize"], 0)
        put_i(C["EG_g_flags"], self.flags)
        sky = [i for i, L in enumerate(self.lights) if int(L[C["PLIGHT_TYPE"]:C["PLIGHT_TYPE"] + 1].view(np.int32)[0]) == C["PLAIN_LIGHT_TYPE_SKY_DOME"]]
        put_i(C["EG_skyLightId"], sky[0] if sky else -1)    # SetAllPODLights, IHWLayerDataAssembler.cpp:404-416
        put_i(C["EG_lightsOffset"], offs["lights"])
        put_i(C["EG_lightsSize"], nl*128)
        put_i(C["EG_lightsNum"], nl)
        put_i(C["EG_sunNumber"], 0)
        put_i(C["EG_m_allTablesAreReady"], 1)
        if self.ms_tables is not None:
            ggx, transp = self.ms_tables
            b8[C["EG_m_essGgx2017Table"]:C["EG_m_essGgx2017Table"] + 64*64*2] = np.ascontiguousarray(ggx, np.uint16).view(np.uint8)
            b8[C["EG_m_essTranspTable"]:C["EG_m_essTranspTable"] + 64*64*64*2] = np.ascontiguousarray(transp, np.uint16).view(np.uint8)

        blob[offs["materials"]:offs["materials"] + len(mat_table)] = mat_table
        blob[offs["geometry"]:offs["geometry"] + len(geom_table)] = geom_table
        blob[offs["textures"]:offs["textures"] + len(tex_table)] = tex_table
        blob[offs["texturesAux"]:offs["texturesAux"] + len(tex_table)] = getattr(self, "_aux_table", [-1]*len(tex_table))
        blob[offs["pdfTable"]:offs["pdfTable"] + sizes["pdfTable"]] = -1
        for i, o in enumerate(getattr(self, "_pdf_table_offsets", [])):
            blob[offs["pdfTable"] + i] = o
        if nl > 0:
            # light selection tables = prefix sums of the pick probabilities, N = lights+1 entries (RenderDriverRTE.cpp:1499-1521, clight.h:1774-1793)
            lights = np.stack(self.lights).astype(np.float32)
            prefs = {}
            for fwd, slot in ((False, C["PLIGHT_PICK_PROB_REV"]), (True, C["PLIGHT_PICK_PROB_FWD"])):
                w = light_pick_probs(lights, fwd)
                pref = np.zeros(nl + 1, np.float32)
                acc = np.float32(0)
                for i in range(nl):
                    pref[i] = acc
                    acc = np.float32(acc + w[i])
                pref[nl] = acc
                with np.errstate(divide="ignore", invalid="ignore"):
                    lights[:, slot] = w*(np.float32(1.0)/acc)                                # stored per light, then normalised (RenderDriverRTE.cpp:1509-1516)
                prefs[fwd] = pref
            # EngineGlobals::suns: SetAllPODLights copies the FIRST soft directional light into every one of the MAX_SUN_NUM = 8 slots (its search
            # restarts from light 0 for each slot, IHWLayerDataAssembler.cpp:421-450), with the pick probabilities already normalised
            # (RenderDriverRTE.cpp:1503-1520)
            li = lights.view(np.int32)
            soft = [i for i in range(nl) if li[i, C["PLIGHT_TYPE"]] == C["PLAIN_LIGHT_TYPE_DIRECT"] and lights[i, 16] > np.float32(1e-6)]   # DIRECT_LIGHT_SSOFTNESS
            if soft:
                for k in range(8):
                    blob[C["EG_suns"]//4 + k*128:C["EG_suns"]//4 + (k + 1)*128] = li[soft[0]]
                blob[C["EG_sunNumber"]//4] = 8
            blob[offs["lselRev"]:offs["lselRev"] + nl + 1] = prefs[False].view(np.int32)
            blob[offs["lselFwd"]:offs["lselFwd"] + nl + 1] = prefs[True].view(np.int32)
            blob[offs["lights"]:offs["lights"] + nl*128] = lights.reshape(-1).view(np.int32)
        return blob


def light_pick_probs(lights, fwd):
    """RenderDriverRTE::CalcLightPickProbTable (RenderDriverRTE_PdfTables.cpp:575-647): uniform over light groups (a light with group id -1 is
    its own group), zero for lights flagged LIGHT_DO_NOT_SAMPLE_ME, for black lights (|colour| < 0.01), for sky domes that a sky portal
    replaces and for sky domes in the forward (light tracing) table; times PLIGHT_PROB_MULT when that is positive.  Not normalised."""
    li = lights.view(np.int32)
    gid = li[:, C["PLIGHT_GROUP_ID"]]
    members = {}
    for g in gid:
        if g != -1:
            members[int(g)] = members.get(int(g), 0) + 1
    n_groups = int((gid == -1).sum()) + len(members)
    pick_group = np.float32(1.0)/np.float32(n_groups)
    portal_sky = {int(li[i, C["AREA_LIGHT_SKY_SOURCE"]]) for i in range(len(lights))
                  if li[i, C["PLIGHT_TYPE"]] == C["PLAIN_LIGHT_TYPE_AREA"] and (li[i, C["PLIGHT_FLAGS"]] & C["AREA_LIGHT_SKY_PORTAL"])}
    out = np.zeros(len(lights), np.float32)
    for i in range(len(lights)):
        pp = pick_group if gid[i] == -1 else np.float32(pick_group/np.float32(members[int(gid[i])]))
        if li[i, C["PLIGHT_FLAGS"]] & C["LIGHT_DO_NOT_SAMPLE_ME"]:
            pp = np.float32(0)
        if li[i, C["PLIGHT_TYPE"]] == C["PLAIN_LIGHT_TYPE_SKY_DOME"] and (fwd or i in portal_sky):
            pp = np.float32(0)
        col = lights[i, C["PLIGHT_COLOR_X"]:C["PLIGHT_COLOR_X"] + 3]
        if np.sqrt(np.float32(col[0]*col[0] + col[1]*col[1] + col[2]*col[2])) < np.float32(0.01):
            pp = np.float32(0)
        if lights[i, C["PLIGHT_PROB_MULT"]] > 0:
            pp = np.float32(pp*lights[i, C["PLIGHT_PROB_MULT"]])
        out[i] = pp
    return out


# ---------------------------------------------------------------------------------------------------------------- BASELINE configs
C2_LIGHT_POS = (0.0, 25.0, 0.0)


def scene_c2(width=1920, height=1080):
    """BASELINE config C2: synthetic 1,001,112-triangle mesh (708 x 707 displaced grid, seed 1234), one instance, identity
    matrix, 1080p pinhole camera looking down at the terrain so that ~all primary rays hit; one point light for shadow rays."""
    scn = Scene(width, height, Camera(pos=(0.0, 9.0, 16.0), look_at=(0.0, 0.0, 1.0), fov=45.0))
    scn.add_instance(scn.add_mesh(grid_mesh(708, 707, size=60.0, amplitude=1.6, seed=1234)))
    scn.add_material(np.zeros(192, np.float32))
    return scn.build()


def scene_c3(width=1920, height=1080, trace_depth=8):
    """BASELINE config C3: the C2 1M-triangle terrain with materials by triangle block {Lambert 60 %, GGX gloss 0.7 20 %, glass IOR 1.5
    10 %, Lambert+GGX fresnel blend 10 %}, two rectangular area lights (with their emissive quads), MISPT with trace_depth 8
    (-> HRT_TRACE_DEPTH = 9; the oracle ignores diff_trace_depth)."""
    from . import materials as M
    scn = Scene(width, height, Camera(pos=(0.0, 9.0, 16.0), look_at=(0.0, 0.0, 1.0), fov=45.0))
    scn.set_trace_depth(trace_depth, 3)
    lam = scn.add_material(M.lambert((0.62, 0.58, 0.50)))
    ggx = scn.add_material(M.ggx((0.85, 0.70, 0.35), 0.7))
    gls = scn.add_material(M.glass((0.95, 0.98, 0.95), ior=1.5, gloss=1.0))
    bld = scn.add_material(M.blend((0.8, 0.8, 0.8), M.ggx((0.9, 0.9, 0.9), 0.85), M.lambert((0.2, 0.35, 0.7)), fresnel=True, ior=1.5))
    emi0 = scn.add_material(M.emissive((60.0, 54.0, 45.0), 0))
    emi1 = scn.add_material(M.emissive((20.0, 30.0, 45.0), 1))
    scn.add_instance(scn.add_mesh(grid_mesh(708, 707, size=60.0, amplitude=1.6, seed=1234,
                                            mat_blocks=[(0.6, lam), (0.2, ggx), (0.1, gls), (0.1, bld)])))
    l0 = scn.add_light(M.area_light((0.0, 14.0, 0.0), (6.0, 6.0), (60.0, 54.0, 45.0)))
    scn.add_instance(scn.add_mesh(quad_mesh(6.0, 6.0, y=0.0, mat_id=emi0, flip=True)), translate(0.0, 14.0, 0.0), light_id=l0)
    l1 = scn.add_light(M.area_light((-14.0, 9.0, 10.0), (3.0, 3.0), (20.0, 30.0, 45.0)))
    scn.add_instance(scn.add_mesh(quad_mesh(3.0, 3.0, y=0.0, mat_id=emi1, flip=True)), translate(-14.0, 9.0, 10.0), light_id=l1)
    return scn.build()


def scene_c4(width=1920, height=1080, instances=200, grid=(224, 224), seed=99):
    """BASELINE config C4: 20 M instanced triangles = `instances` rigid copies (random rotation about Y + translation, seed 99) of one
    ~100k-triangle displaced patch (224 x 224 quads = 100,352 triangles), five diffuse materials by triangle block (Lambert in four colours,
    Oren-Nayar) so that the material sort of the live-path queue has something to sort, one large area light."""
    from . import materials as M
    rng = np.random.RandomState(seed)
    scn = Scene(width, height, Camera(pos=(0.0, 55.0, 95.0), look_at=(0.0, 0.0, 5.0), fov=45.0))
    scn.set_trace_depth(5, 3)
    lam = scn.add_material(M.lambert((0.7, 0.7, 0.7)))
    emi = scn.add_material(M.emissive((40.0, 40.0, 40.0), 0))
    lam2 = scn.add_material(M.lambert((0.75, 0.45, 0.35)))
    oren = scn.add_material(M.orennayar((0.55, 0.65, 0.45), 0.6))
    lam3 = scn.add_material(M.lambert((0.40, 0.50, 0.75)))
    lam4 = scn.add_material(M.lambert((0.80, 0.78, 0.55)))
    patch = scn.add_mesh(grid_mesh(grid[0], grid[1], size=12.0, amplitude=0.9, seed=4321,
                                   mat_blocks=[(0.30, lam), (0.25, lam2), (0.20, oren), (0.15, lam3), (0.10, lam4)]))
    side = int(math.ceil(math.sqrt(instances)))
    for k in range(instances):
        gx, gz = k % side, k//side
        t = translate((gx - 0.5*(side - 1))*11.0 + rng.uniform(-1, 1), rng.uniform(-1.5, 1.5), (gz - 0.5*(side - 1))*11.0 + rng.uniform(-1, 1))
        scn.add_instance(patch, t @ rotate_y(rng.uniform(0, 2*math.pi)))
    l0 = scn.add_light(M.area_light((0.0, 60.0, 0.0), (40.0, 40.0), (40.0, 40.0, 40.0)))
    scn.add_instance(scn.add_mesh(quad_mesh(40.0, 40.0, y=0.0, mat_id=emi, flip=True)), translate(0.0, 60.0, 0.0), light_id=l0)
    return scn.build()
