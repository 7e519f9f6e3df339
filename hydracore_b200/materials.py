"""PlainMaterial (192 floats) and PlainLight (128 floats) packers for synthetic scenes.

Field meaning follows the reference's converters (hydra_drv/PlainMaterialConverter.cpp, PlainLightConverter.cpp) and the slot
maps in hydra_drv/cmaterial.h / clight.h; every offset comes from csrc/hc_layout.h (pinned against the reference in
tests/test_layout.py).  Only what the CUDA layer supports this round is offered."""
import math

import numpy as np

from .layout import C

INVALID_TEXTURE = -2


def _i2f(v):
    return np.array([v], np.int32).view(np.float32)[0]


def _node(mat_type, flags):
    m = np.zeros(192, np.float32)
    m[C["PLAIN_MAT_TYPE_OFFSET"]] = _i2f(mat_type)
    m[C["PLAIN_MAT_FLAGS_OFFSET"]] = _i2f(flags)
    for k in ("EMISSIVE_TEXID_OFFSET", "EMISSIVE_TEXMATRIXID_OFFSET", "OPACITY_TEX_OFFSET", "OPACITY_TEX_MATRIX", "NORMAL_TEX_OFFSET",
              "NORMAL_TEX_MATRIX", "PROC_TEX1_F4_HEAD_OFFSET"):
        m[C[k]] = _i2f(INVALID_TEXTURE)
    m[C["EMISSIVE_LIGHTID_OFFSET"]] = _i2f(-1)
    return m


def _sampler(m, float_offset, tex_id, flags=0, gamma=2.2, row0=(1, 0, 0, 0), row1=(0, 1, 0, 0)):
    """SWTexSampler (48 bytes: flags, gamma, texId, pad, row0, row1; cfetch.h:108-131) at `float_offset`; returns its float4 index."""
    assert float_offset % 4 == 0
    m[float_offset + 0] = _i2f(flags)
    m[float_offset + 1] = gamma
    m[float_offset + 2] = _i2f(tex_id)
    m[float_offset + 3] = _i2f(0)
    m[float_offset + 4:float_offset + 8] = row0
    m[float_offset + 8:float_offset + 12] = row1
    return float_offset//4


def _color_slot(m, color, tex_id, texid_off, texmat_off, sampler_off, **kw):
    m[10:13] = color
    if tex_id and tex_id > 0:
        m[texid_off] = _i2f(tex_id)
        m[texmat_off] = _i2f(_sampler(m, sampler_off, tex_id, **kw))
    else:
        m[texid_off] = _i2f(INVALID_TEXTURE)
        m[texmat_off] = _i2f(INVALID_TEXTURE)


def with_opacity(nodes, tex_id, **kw):
    """Opacity (cut-out) map on a material: IMaterial::SetOpacitySampler puts texture id, sampler index and the sampler into the HEAD node
    (PlainMaterialConverter.cpp:1428-1443).  Meshes that use such a material go into the alpha-tested BVH tree (RenderDriverRTE.cpp:1989);
    the texel's max(rgb) > 0.5 keeps the hit (ctrace.h:384-398).  Sampler keywords as for colour textures (flags, gamma, row0, row1)."""
    nodes = [nodes] if isinstance(nodes, np.ndarray) and nodes.ndim == 1 else list(nodes)
    head = nodes[0]
    head[C["OPACITY_TEX_OFFSET"]] = _i2f(tex_id)
    head[C["OPACITY_TEX_MATRIX"]] = _i2f(_sampler(head, C["OPACITY_SAMPLER_OFFSET"], tex_id, **kw))
    return nodes


def with_normal_map(nodes, tex_id, invert_x=False, invert_y=False, swap_xy=False, **kw):
    """Normal map on a material: IMaterial::SetNormalSampler + the "push down" to every node of a blend tree (PlainMaterialConverter.cpp:1395-1460):
    NORMAL_TEX_OFFSET = texture id (looked up in the AUX texture table), NORMAL_TEX_MATRIX = float4 index of the sampler at
    NORMAL_SAMPLER_OFFSET; flags PLAIN_MATERIAL_INVERT_NMAP_X / _Y / _SWAP_NMAP_XY.  The image comes from Scene.add_normal_map."""
    nodes = [nodes] if isinstance(nodes, np.ndarray) and nodes.ndim == 1 else list(nodes)
    kw.setdefault("gamma", 1.0)
    for n in nodes:
        n[C["NORMAL_TEX_OFFSET"]] = _i2f(tex_id)
        n[C["NORMAL_TEX_MATRIX"]] = _i2f(_sampler(n, C["NORMAL_SAMPLER_OFFSET"], tex_id, **kw))
        f = int(np.asarray(n[C["PLAIN_MAT_FLAGS_OFFSET"]:C["PLAIN_MAT_FLAGS_OFFSET"] + 1], np.float32).view(np.int32)[0])
        f |= (C["PLAIN_MATERIAL_INVERT_NMAP_X"] if invert_x else 0) | (C["PLAIN_MATERIAL_INVERT_NMAP_Y"] if invert_y else 0) | \
             (C["PLAIN_MATERIAL_INVERT_SWAP_NMAP_XY"] if swap_xy else 0)
        n[C["PLAIN_MAT_FLAGS_OFFSET"]] = _i2f(f)
    return nodes


def opacity_tex_id(head_node):
    return int(np.asarray(head_node[C["OPACITY_TEX_OFFSET"]:C["OPACITY_TEX_OFFSET"] + 1], np.float32).view(np.int32)[0])


def lambert(color, tex_id=0, **kw):
    m = _node(C["PLAIN_MAT_CLASS_LAMBERT"], C["PLAIN_MATERIAL_HAS_DIFFUSE"])
    _color_slot(m, color, tex_id, C["LAMBERT_TEXID_OFFSET"], C["LAMBERT_TEXMATRIXID_OFFSET"], C["LAMBERT_SAMPLER0"], **kw)
    return m


def thin_glass(color, gloss=1.0, tex_id=0, **kw):
    """Thin glass (ThinGlassMaterial, PlainMaterialConverter.cpp:254-285; cmaterial.h:471-555): light passes straight through (a glossy lobe about
    the ray direction when gloss < 1), tinted by `color`; Phong's slot layout."""
    return _glossy(C["PLAIN_MAT_CLASS_THIN_GLASS"], color, gloss, tex_id, 0, **kw)


def translucent(color, tex_id=0, **kw):
    """Translucent Lambert (TranslucentMaterial, PlainMaterialConverter.cpp; cmaterial.h:1850-1910): cosine-distributed transmission through
    the surface, Lambert's colour / texture slots; flag PLAIN_MATERIAL_HAS_DIFFUSE like the converter sets."""
    m = _node(C["PLAIN_MAT_CLASS_TRANSLUCENT"], C["PLAIN_MATERIAL_HAS_DIFFUSE"])
    _color_slot(m, color, tex_id, C["LAMBERT_TEXID_OFFSET"], C["LAMBERT_TEXMATRIXID_OFFSET"], C["LAMBERT_SAMPLER0"], **kw)
    return m


def orennayar(color, roughness, tex_id=0, **kw):
    """Oren-Nayar diffuse (OrenNayarMaterial, PlainMaterialConverter.cpp:142-170): sigma = roughness*pi/2, A and B precomputed in float."""
    m = _node(C["PLAIN_MAT_CLASS_OREN_NAYAR"], C["PLAIN_MATERIAL_HAS_DIFFUSE"])
    _color_slot(m, color, tex_id, C["LAMBERT_TEXID_OFFSET"], C["LAMBERT_TEXMATRIXID_OFFSET"], C["LAMBERT_SAMPLER0"], **kw)
    f = np.float32
    sigma = f(f(roughness)*f(np.float64(3.14159265358979323846)/2.0))        # M_PI is the <cmath> double in the reference's host build; /2.0f in double, product in float
    sigma2 = f(sigma*sigma)
    m[15] = f(roughness)                                                      # ORENNAYAR_ROUGHNESS
    m[16] = f(f(1.0) - f(sigma2/f(f(2.0)*f(sigma2 + f(0.33)))))               # ORENNAYAR_A
    m[17] = f(f(f(0.45)*sigma2)/f(sigma2 + f(0.09)))                          # ORENNAYAR_B
    return m


def _glossy(mat_type, color, gloss, tex_id, flags, **kw):
    m = _node(mat_type, flags)
    _color_slot(m, color, tex_id, C["PHONG_TEXID_OFFSET"], C["PHONG_TEXMATRIXID_OFFSET"], C["PHONG_SAMPLER0_OFFSET"], **kw)
    m[C["PHONG_GLOSINESS_OFFSET"]] = gloss
    m[C["PHONG_GLOSINESS_TEXID_OFFSET"]] = _i2f(INVALID_TEXTURE)
    m[C["PHONG_GLOSINESS_TEXMATRIXID_OFFSET"]] = _i2f(INVALID_TEXTURE)
    return m


def phong(color, gloss, tex_id=0, energy_fix=False, **kw):
    return _glossy(C["PLAIN_MAT_CLASS_PHONG_SPECULAR"], color, gloss, tex_id,
                   C["PLAIN_MATERIAL_ENERGY_FIX_OR_MULTISCATTER"] if energy_fix else 0, **kw)


def blinn(color, gloss, tex_id=0, **kw):
    """BlinnTorranceSrappowMaterial (PlainMaterialConverter.cpp:462-485; XML brdf_type="torranse_sparrow"): Phong's slots, always CAST_CAUSTICS."""
    return _glossy(C["PLAIN_MAT_CLASS_BLINN_SPECULAR"], color, gloss, tex_id, C["PLAIN_MATERIAL_CAST_CAUSTICS"], **kw)


def ggx(color, gloss, tex_id=0, multiscatter=False, **kw):
    return _glossy(C["PLAIN_MAT_CLASS_GGX"], color, gloss, tex_id,
                   C["PLAIN_MATERIAL_ENERGY_FIX_OR_MULTISCATTER"] if multiscatter else 0, **kw)


def _aniso(mat_type, color, gloss, aniso, rot, flip, tex_id, gloss_tex_id, aniso_tex_id, rot_tex_id, **kw):
    m = _node(mat_type, C["PLAIN_MATERIAL_CAST_CAUSTICS"] | (C["PLAIN_MATERIAL_FLIP_TANGENT"] if flip else 0))
    m[C["BECKMANN_COLORX_OFFSET"]:C["BECKMANN_COLORX_OFFSET"] + 3] = color
    m[C["BECKMANN_COSPOWER_OFFSET"]] = 0.0
    m[C["BECKMANN_GLOSINESS_OFFSET"]] = gloss
    m[C["BECKMANN_ANISOTROPY_OFFSET"]] = aniso
    m[C["BECKMANN_ANISO_ROT_OFFSET"]] = rot
    for tid, id_off, mat_off, samp_off in ((tex_id, "BECKMANN_TEXID_OFFSET", "BECKMANN_TEXMATRIXID_OFFSET", "BECKMANN_SAMPLER0_OFFSET"),
                                           (gloss_tex_id, "BECKMANN_GLOSINESS_TEXID_OFFSET", "BECKMANN_GLOSINESS_TEXMATRIXID_OFFSET", "BECKMANN_SAMPLER1_OFFSET"),
                                           (aniso_tex_id, "BECKMANN_ANISO_TEXID_OFFSET", "BECKMANN_ANISO_TEXMATRIXID_OFFSET", "BECKMANN_SAMPLER2_OFFSET"),
                                           (rot_tex_id, "BECKMANN_ROT_TEXID_OFFSET", "BECKMANN_ROT_TEXMATRIXID_OFFSET", "BECKMANN_SAMPLER3_OFFSET")):
        if tid and tid > 0:
            m[C[id_off]] = _i2f(tid)
            m[C[mat_off]] = _i2f(_sampler(m, C[samp_off], tid, **kw))
        else:
            m[C[id_off]] = _i2f(INVALID_TEXTURE)
            m[C[mat_off]] = _i2f(INVALID_TEXTURE)
            _sampler(m, C[samp_off], INVALID_TEXTURE)                # DummySampler(): the converter always writes the four samplers
    return m


def beckmann(color, gloss, aniso=0.0, rot=0.0, flip=False, tex_id=0, gloss_tex_id=0, aniso_tex_id=0, rot_tex_id=0, **kw):
    """BeckmannMaterial (PlainMaterialConverter.cpp:501-565; XML brdf_type="beckmann"): anisotropic Beckmann lobe, alpha_x from the glossiness,
    alpha_y narrowed by `aniso` in [0, 1], tangent frame turned by `rot` (fraction of a full turn) about the normal; four samplers (colour,
    glossiness, anisotropy, rotation); always CAST_CAUSTICS, FLIP_TANGENT swaps the frame axes."""
    return _aniso(C["PLAIN_MAT_CLASS_BECKMANN"], color, gloss, aniso, rot, flip, tex_id, gloss_tex_id, aniso_tex_id, rot_tex_id, **kw)


def trggx(color, gloss, aniso=0.0, rot=0.0, flip=False, tex_id=0, gloss_tex_id=0, aniso_tex_id=0, rot_tex_id=0, **kw):
    """TRGGXMaterial (PlainMaterialConverter.cpp:568-632; XML brdf_type="trggx"): the Trowbridge-Reitz counterpart in Beckmann's slots."""
    return _aniso(C["PLAIN_MAT_CLASS_TRGGX"], color, gloss, aniso, rot, flip, tex_id, gloss_tex_id, aniso_tex_id, rot_tex_id, **kw)


def mirror(color):
    m = _node(C["PLAIN_MAT_CLASS_PERFECT_MIRROR"], 0)
    m[10:13] = color
    m[C["MIRROR_TEXID_OFFSET"]] = _i2f(INVALID_TEXTURE)
    m[C["MIRROR_TEXMATRIXID_OFFSET"]] = _i2f(INVALID_TEXTURE)
    return m


def glass(color, ior=1.5, gloss=1.0, multiscatter=False):
    """multiscatter: PLAIN_MATERIAL_ENERGY_FIX_OR_MULTISCATTER - rough glass takes its energy compensation from the baked 64^3 table
    EngineGlobals::m_essTranspTable (Scene.ms_tables must hold the tables then)."""
    m = _node(C["PLAIN_MAT_CLASS_GLASS"], C["PLAIN_MATERIAL_HAS_TRANSPARENCY"] | C["PLAIN_MATERIAL_HAVE_BTDF"] |
              (C["PLAIN_MATERIAL_ENERGY_FIX_OR_MULTISCATTER"] if multiscatter else 0))
    m[10:13] = color
    m[C["GLASS_TEXID_OFFSET"]] = _i2f(INVALID_TEXTURE)
    m[C["GLASS_TEXMATRIXID_OFFSET"]] = _i2f(INVALID_TEXTURE)
    m[C["GLASS_IOR_OFFSET"]] = ior
    m[C["GLASS_FOG_COLORX_OFFSET"]:C["GLASS_FOG_COLORX_OFFSET"] + 3] = 1.0
    m[C["GLASS_GLOSINESS"]] = gloss
    m[C["GLASS_GLOSINESS_TEXID_OFFSET"]] = _i2f(INVALID_TEXTURE)
    m[C["GLASS_GLOSINESS_TEXMATRIXID_OFFSET"]] = _i2f(INVALID_TEXTURE)
    return m


def emissive(color, light_id=-1):
    """Material of a light's mesh (EmissiveMaterial, PlainMaterialConverter.cpp:31-47)."""
    m = _node(C["PLAIN_MAT_CLASS_EMISSIVE"], C["PLAIN_MATERIAL_IS_LIGHT"])
    m[C["EMISSIVE_COLORX_OFFSET"]:C["EMISSIVE_COLORX_OFFSET"] + 3] = color
    m[C["EMISSIVE_LIGHTID_OFFSET"]] = _i2f(light_id)
    return m


def blend(mask_color, top, bottom, fresnel=True, ior=1.5, sigmoid_exp=None):
    """BlendMask node followed by its children: `top` (material 1, chosen with probability alpha) at relative offset +1 and `bottom`
    (material 2) right after it — the layout the converter emits for diffuse+reflect materials (PlainMaterialConverter.cpp:793-813);
    non-fresnel blends get BLEND_MASK_REFLECTION_WEIGHT_IS_ONE (:787-788).  `top` / `bottom` are nodes or node lists."""
    top = [top] if isinstance(top, np.ndarray) and top.ndim == 1 else list(top)
    bottom = [bottom] if isinstance(bottom, np.ndarray) and bottom.ndim == 1 else list(bottom)
    m = _node(C["PLAIN_MAT_CLASS_BLEND_MASK"], C["PLAIN_MATERIAL_SURFACE_BLEND"])
    m[10:13] = mask_color
    m[C["BLEND_MASK_TEXID_OFFSET"]] = _i2f(INVALID_TEXTURE)
    m[C["BLEND_MASK_TEXMATRIXID_OFFSET"]] = _i2f(INVALID_TEXTURE)
    flags = C["BLEND_MASK_FRESNEL"] if fresnel else C["BLEND_MASK_REFLECTION_WEIGHT_IS_ONE"]
    m[C["BLEND_MASK_FLAGS_OFFSET"]] = _i2f(flags)
    m[C["BLEND_MASK_MATERIAL1_OFFSET"]] = _i2f(1)
    m[C["BLEND_MASK_MATERIAL2_OFFSET"]] = _i2f(1 + len(top))
    m[C["BLEND_MASK_FRESNEL_IOR"]] = ior
    btype = C["BLEND_FRESNEL"] if fresnel else C["BLEND_SIMPLE"]
    if sigmoid_exp is not None:
        btype = C["BLEND_SIGMOID"]
        m[C["BLEND_SIGMOID_EXP"]] = sigmoid_exp
    m[C["BLEND_TYPE"]] = _i2f(btype)
    m[C["BLEND_FLAGS"]] = _i2f(0)
    return [m] + top + bottom


def area_light(pos, half_size, intensity, rotation=None, disk=False, pick_prob=1.0, spot_angles_deg=None):
    """Rectangular / disk area light (AreaDiffuseLight, PlainLightConverter.cpp:130-300): local normal (0,-1,0), sample position
    R*(+-sx, 0, +-sy) + pos, surface area 4*sx*sy (pi*r^2 for disks); field map hydra_drv/clight.h:15-64, 493-521."""
    L = np.zeros(128, np.float32)
    R = np.eye(3, dtype=np.float32) if rotation is None else np.asarray(rotation, np.float32).reshape(3, 3)
    L[C["PLIGHT_TYPE"]] = _i2f(C["PLAIN_LIGHT_TYPE_AREA"])
    L[C["PLIGHT_FLAGS"]] = _i2f(0)
    L[C["PLIGHT_POS_X"]:C["PLIGHT_POS_X"] + 3] = pos
    n = R @ np.array([0, -1, 0], np.float32)
    L[C["PLIGHT_NORM_X"]:C["PLIGHT_NORM_X"] + 3] = n/np.linalg.norm(n)
    L[C["PLIGHT_COLOR_X"]:C["PLIGHT_COLOR_X"] + 3] = intensity
    L[C["PLIGHT_COLOR_TEX"]] = _i2f(INVALID_TEXTURE)
    L[C["PLIGHT_COLOR_TEX_MATRIX"]] = _i2f(INVALID_TEXTURE)
    sx, sy = float(half_size[0]), float(half_size[1])
    L[C["PLIGHT_SURFACE_AREA"]] = (math.pi*sx*sx) if disk else (4.0*sx*sy)
    L[C["AREA_LIGHT_SIZE_X"]] = sx
    L[C["AREA_LIGHT_SIZE_Y"]] = sy
    L[C["AREA_LIGHT_MATRIX_E00"]:C["AREA_LIGHT_MATRIX_E00"] + 9] = R.reshape(9)
    L[C["AREA_LIGHT_IS_DISK"]] = _i2f(1 if disk else 0)
    L[C["AREA_LIGHT_SPOT_DISTR"]] = _i2f(0)
    if spot_angles_deg is not None:                                   # distribution="spot": cosines of the half angles of the inner and outer cone (PlainLightConverter.cpp:241-245)
        a1, a2 = spot_angles_deg
        L[C["AREA_LIGHT_SPOT_DISTR"]] = _i2f(1)
        L[C["AREA_LIGHT_SPOT_COS1"]] = np.float32(math.cos(math.radians(0.5*a1)))
        L[C["AREA_LIGHT_SPOT_COS2"]] = np.float32(math.cos(math.radians(0.5*a2)))
    L[C["PLIGHT_PROB_MULT"]] = 1.0
    L[C["PLIGHT_PICK_PROB_FWD"]] = pick_prob
    L[C["PLIGHT_PICK_PROB_REV"]] = pick_prob
    return L


def sphere_light(pos, radius, intensity, pick_prob=1.0):
    """Sphere area light (SphereLight, PlainLightConverter.cpp:445-492): centre, radius, surface area 4*pi*r^2, uniform emission."""
    L = np.zeros(128, np.float32)
    L[C["PLIGHT_TYPE"]] = _i2f(C["PLAIN_LIGHT_TYPE_SPHERE"])
    L[C["PLIGHT_FLAGS"]] = _i2f(0)
    L[C["PLIGHT_POS_X"]:C["PLIGHT_POS_X"] + 3] = pos
    L[C["PLIGHT_COLOR_X"]:C["PLIGHT_COLOR_X"] + 3] = intensity
    L[C["PLIGHT_COLOR_TEX"]] = _i2f(INVALID_TEXTURE)
    L[C["PLIGHT_COLOR_TEX_MATRIX"]] = _i2f(INVALID_TEXTURE)
    r = np.float32(radius)
    L[14] = r                                                              # SPHERE_LIGHT_RADIUS, clight.h:33
    L[C["PLIGHT_SURFACE_AREA"]] = np.float32(4.0)*np.float32(3.1415926535)*r*r
    L[C["PLIGHT_PROB_MULT"]] = 1.0
    L[C["PLIGHT_PICK_PROB_FWD"]] = pick_prob
    L[C["PLIGHT_PICK_PROB_REV"]] = pick_prob
    return L


def point_light(pos, intensity, pick_prob=1.0):
    """Omni point light without IES (PointLight, PlainLightConverter.cpp:628-690): surface area 1e-10."""
    L = np.zeros(128, np.float32)
    L[C["PLIGHT_TYPE"]] = _i2f(C["PLAIN_LIGHT_TYPE_POINT_OMNI"])
    L[C["PLIGHT_FLAGS"]] = _i2f(0)
    L[C["PLIGHT_POS_X"]:C["PLIGHT_POS_X"] + 3] = pos
    L[C["PLIGHT_COLOR_X"]:C["PLIGHT_COLOR_X"] + 3] = intensity
    L[C["PLIGHT_COLOR_TEX"]] = _i2f(INVALID_TEXTURE)
    L[C["PLIGHT_COLOR_TEX_MATRIX"]] = _i2f(INVALID_TEXTURE)
    L[C["PLIGHT_SURFACE_AREA"]] = 1e-10
    L[C["PLIGHT_PROB_MULT"]] = 1.0
    L[C["PLIGHT_PICK_PROB_FWD"]] = pick_prob
    L[C["PLIGHT_PICK_PROB_REV"]] = pick_prob
    return L


def with_ies(L, tex_table_id, pdf_table_id=None, matrix=None, point_area=False):
    """IES photometric web on a point or area light (LIGHT_HAS_IES; PlainLightConverter.cpp:255-270, 676-694): ids of the web image and of its pdf
    table in the "pdfs" storage (Scene.add_ies_table), the light's rotation for the look-up (IES_LIGHT_MATRIX, rows) and its inverse;
    point_area: LIGHT_IES_POINT_AREA - an area light looks the web up from its centre."""
    Li = L.view(np.int32)
    Li[C["PLIGHT_FLAGS"]] |= C["LIGHT_HAS_IES"] | (C["LIGHT_IES_POINT_AREA"] if point_area else 0)
    Li[C["IES_SPHERE_TEX_ID"]] = tex_table_id
    Li[C["IES_SPHERE_PDF_ID"]] = tex_table_id if pdf_table_id is None else pdf_table_id
    R = np.eye(3, dtype=np.float32) if matrix is None else np.asarray(matrix, np.float32).reshape(3, 3)
    L[C["IES_LIGHT_MATRIX_E00"]:C["IES_LIGHT_MATRIX_E00"] + 9] = R.reshape(9)
    L[C["IES_INV_MATRIX_E00"]:C["IES_INV_MATRIX_E00"] + 9] = np.linalg.inv(R.astype(np.float64)).astype(np.float32).reshape(9)
    return L


def spot_light(pos, direction, intensity, falloff_angle, falloff_angle2):
    """Spot light (SpotLight + CreatePointSpotLightFromXmlNode, PlainLightConverter.cpp:568-626, 895-906): cone angles in degrees, full
    intensity inside falloff_angle2 (cos1), smooth fall-off to zero at falloff_angle (cos2).  `direction` is the light's axis."""
    L = point_light(pos, intensity)
    L[C["PLIGHT_TYPE"]] = _i2f(C["PLAIN_LIGHT_TYPE_POINT_SPOT"])
    d = np.asarray(direction, np.float32)
    L[C["PLIGHT_NORM_X"]:C["PLIGHT_NORM_X"] + 3] = d/np.float32(np.linalg.norm(d))
    deg = np.float32(np.pi/180.0)
    L[14] = np.float32(np.cos(np.float32(0.5)*deg*np.float32(falloff_angle2)))       # POINT_LIGHT_SPOT_COS1
    L[15] = np.float32(np.cos(np.float32(0.5)*deg*np.float32(falloff_angle)))        # POINT_LIGHT_SPOT_COS2
    return L


def direct_light(pos, direction, intensity, radius1, radius2, soft_angle_deg=0.0):
    """Directional light (DirectLight, PlainLightConverter.cpp:500-566): parallel rays along `direction` inside a cylinder of radius1..radius2
    (smooth fall-off between them) around the axis through `pos`; soft_angle_deg > 0 jitters the direction inside a cone (a "sun")."""
    L = point_light(pos, intensity)
    L[C["PLIGHT_TYPE"]] = _i2f(C["PLAIN_LIGHT_TYPE_DIRECT"])
    d = np.asarray(direction, np.float32)
    L[C["PLIGHT_NORM_X"]:C["PLIGHT_NORM_X"] + 3] = d/np.float32(np.linalg.norm(d))
    alpha = np.float32(np.pi/180.0)*np.float32(soft_angle_deg)
    L[14], L[15] = radius1, radius2                                                     # DIRECT_LIGHT_RADIUS1 / 2
    L[16] = np.float32(soft_angle_deg)/np.float32(0.25)                                 # DIRECT_LIGHT_SSOFTNESS
    L[17], L[18] = np.float32(np.tan(alpha)), np.float32(np.cos(alpha))                 # DIRECT_LIGHT_ALPHA_TAN / _COS
    L[C["PLIGHT_SURFACE_AREA"]] = np.float32(np.pi)*np.float32(radius2)*np.float32(radius2)
    return L


def sky_light(color, pdf_table_id, pick_prob=1.0, tex_id=None, gamma=1.0, perez=None):
    """Sky-dome light, optionally textured (environment map), without the Perez model (SkyDomeLight, PlainLightConverter.cpp:909-1060):
    identity sampler matrices, the pdf table built by Scene.add_sky_pdf_table (from the map's luminance when textured)."""
    L = np.zeros(128, np.float32)
    Li = L.view(np.int32)
    Li[C["PLIGHT_TYPE"]] = C["PLAIN_LIGHT_TYPE_SKY_DOME"]
    Li[C["PLIGHT_FLAGS"]] = 0
    L[C["PLIGHT_COLOR_X"]:C["PLIGHT_COLOR_X"] + 3] = color
    Li[C["PLIGHT_COLOR_TEX"]] = INVALID_TEXTURE
    Li[C["PLIGHT_COLOR_TEX_MATRIX"]] = INVALID_TEXTURE
    L[17:20] = color                                  # SKY_DOME_COLOR_AUX_X..Z
    Li[20] = INVALID_TEXTURE                          # SKY_DOME_COLOR_TEX_AUX
    Li[21] = INVALID_TEXTURE                          # SKY_DOME_COLOR_TEX_MATRIX_AUX
    Li[22] = INVALID_TEXTURE                          # SKY_DOME_AUX_TEX_MATRIX_INV
    L[23:26] = (0.0, -1.0, 0.0)                       # SKY_DOME_SUN_DIR
    Li[30] = pdf_table_id                             # SKY_DOME_PDF_TABLE0
    Li[31] = pdf_table_id                             # SKY_DOME_PDF_TABLE1
    for base in (32, 44):                             # SKY_DOME_SAMPLER0 / SAMPLER1: {flags, gamma, texId, dummy} + row0 + row1
        Li[base + 0] = 0
        L[base + 1] = 1.0
        Li[base + 2] = INVALID_TEXTURE
        L[base + 4:base + 8] = (1, 0, 0, 0)
        L[base + 8:base + 12] = (0, 1, 0, 0)
    for base in (56, 72):                             # SKY_DOME_INV_MATRIX0 / 1: identity float4x4 (columns)
        L[base:base + 16] = np.eye(4, dtype=np.float32).reshape(16)
    Li[88] = -1                                       # SKY_DOME_SUN_DIR_ID
    if perez is not None:                             # <perez turbidity=...>: analytic sky, dict(sun_dir=(x, y, z) pointing AWAY from the sun, turbidity=t, sun_color=(r, g, b))
        Li[C["PLIGHT_FLAGS"]] |= C["SKY_LIGHT_USE_PEREZ_ENVIRONMENT"]          # PlainLightConverter.cpp:920-923; sun direction / colour are filled by the driver from the sun light
        d = np.asarray(perez["sun_dir"], np.float32)
        L[C["SKY_DOME_SUN_DIR_X"]:C["SKY_DOME_SUN_DIR_X"] + 3] = d/np.linalg.norm(d)
        L[C["SKY_DOME_TURBIDITY"]] = perez["turbidity"]
        L[C["SKY_SUN_COLOR_X"]:C["SKY_SUN_COLOR_X"] + 3] = perez.get("sun_color", (1.0, 1.0, 1.0))
    if tex_id is not None:                            # PlainLightConverter.cpp:969-978: texture id for the pdf table builder, sampler at offset 0
        Li[C["PLIGHT_COLOR_TEX"]] = tex_id
        Li[C["PLIGHT_COLOR_TEX_MATRIX"]] = 0
        Li[20], Li[21] = tex_id, 0
        for base in (32, 44):
            L[base + 1] = gamma
            Li[base + 2] = tex_id
    L[C["PLIGHT_PROB_MULT"]] = 1.0
    L[C["PLIGHT_PICK_PROB_FWD"]] = pick_prob
    L[C["PLIGHT_PICK_PROB_REV"]] = pick_prob
    return L
