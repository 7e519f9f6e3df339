// hc_shade.cuh — device functions of the shading half of the path: surface evaluation, software
// textures, BSDF evaluation / sampling over the blend tree, area-light sampling, emission and MIS, path-flag update.
//
// These restate, in our own code, the reference's device math (hydra_drv/{ctrace,cfetch,cmaterial,clight,cbidir,cglobals}.h;
// file:line cited per function).  Formulas, operation order and — where it changes rounding — the C++ promotion of the
// reference's unqualified math calls to double (see hc_math.cuh) are kept, so that the same random numbers give the same
// path on the GPU and in the CPU oracle.  Supported: Lambert, Oren-Nayar, translucent Lambert, Phong, Blinn (Torrance-Sparrow), GGX (Heitz VNDF sampling +
// multiscattering table), anisotropic Beckmann / TRGGX (hc_microfacet.cuh), perfect mirror, GGX glass (+ multiscattering table), thin glass, blend masks
// (simple / fresnel / sigmoid), emissive materials, normal maps; rectangular / disk area lights (spot cone, IES web), sphere lights, omni (IES web) / spot point
// lights, directional lights, mesh and cylinder lights (also textured), sky domes (constant, textured, Perez: hc_perez.cuh); RGBA8, float4 and single-channel textures.
// Everything else is rejected with an error at hc_pt_init (no silent fallback).
#pragma once
#include "hc_math.cuh"
#include "hc_layout.h"
#include "hc_trace.cuh"
#include "hc_microfacet.cuh"
#include "hc_perez.cuh"

#define HC_GEPSILON   5e-6f      // cglobals.h:68-70
#define HC_DEPSILON   1e-20f
#define HC_DEPSILON2  1e-30f
#define HC_INV_PI     0.31830988618379067154f
#define HC_INV_TWOPI  0.15915494309189533577f
#define HC_M_TWOPI    6.28318530717958647692f
#define HC_M_PI_D     3.14159265358979323846       // M_PI of <cmath> (a double) where the reference's host build uses it
#define HC_PI_D       3.14159265358979323846     // M_PI comes from <math.h> as a DOUBLE in the oracle build (cglobals.h:43 is skipped)
#define HC_INVALID_TEXTURE ((int)0xFFFFFFFE)

HC_DEV double D(float x) { return (double)x; }

// device view of everything the shading kernels read; filled on the host from the uploaded blobs
struct HcScene
{
  const int*    __restrict__ globals;       // EngineGlobals + tables blob, indexed in ints from its base (cfetch.h:135-213)
  const float4* __restrict__ geom;          // "geom" storage
  const float4* __restrict__ materials;     // "materials" storage
  const int4*   __restrict__ textures;      // "textures" storage
  const int4*   __restrict__ texturesAux;   // "textures_aux" storage (normal maps)
  const float4* __restrict__ pdfs;          // "pdfs" storage (sky-dome pdf tables)
  const float4* __restrict__ instMatrices;  // inverse instance matrices, 4 float4 each
  const int*    __restrict__ instLightIds;  // instance -> light index or -1
  const int*    __restrict__ remapLists;    // material remap lists {from, to} pairs back to back, or nullptr (SetAllRemapLists, IHWLayer.h:122)
  const int2*   __restrict__ remapTable;    // per list {offset into remapLists, size}
  const int*    __restrict__ remapInst;     // instance -> remap list id or -1 (SetAllInstIdToRemapId, IHWLayer.h:123)
  int remapListsSize, remapTableSize, remapInstSize;
  int materialsTableOffset, geometryTableOffset, texturesTableOffset, texturesAuxTableOffset, pdfTableTableOffset;
  int lightSelTableOffsetRev, lightSelTableSizeRev, lightsOffset, lightsNum, skyLightId;
  int gflags, traceDepth, diffTraceDepth;
  int essGgxTableOffsetBytes;
};

struct HcSurfaceHit       // SurfaceHit, cglobals.h:2514-2528
{
  float3 pos, normal, flatNormal, tangent, biTangent;
  float2 texCoord;
  int    matId;
  float  t, sRayOff;
  bool   hfi;
};

struct HcMatSample { float3 color, direction; float pdf; int flags; };           // MatSample, cglobals.h:388-396
struct HcBxDF { float3 brdf, btdf; float pdfFwd, pdfRev; bool diffuse; };       // BxDFResult, cmaterial.h:2373-2384
struct HcShadowSample { float3 pos, color; float pdf, maxDist, cosAtLight; bool isPoint; };   // ShadowSample, cglobals.h:2448-2456

// ------------------------------------------------------------------------------------------------------------------ a1: RandomGen
struct HcRng { unsigned x, y; };
HC_DEV unsigned NextState(HcRng& g)                       // crandom.h:19-25
{
  const unsigned x = g.x*17u + g.y*13123u;
  g.x = (x << 13) ^ x;
  g.y ^= (x << 7);
  return x;
}
HC_DEV HcRng RandomGenInit(int seed)                      // crandom.h:27-43
{
  const unsigned s = (unsigned)seed;
  HcRng g;
  g.x = s*(s*s*15731u + 74323u) + 871483u;
  g.y = s*(s*s*13734u + 37828u) + 234234u;
  const int n = seed % 7;
  for (int i = 0; i < n; i++) NextState(g);
  return g;
}
HC_DEV float4 rndFloat4_Pseudo(HcRng& g)                  // crandom.h:51-63
{
  const unsigned x = NextState(g);
  const unsigned x1 = x*(x*x*15731u + 74323u) + 871483u;
  const unsigned y1 = x*(x*x*13734u + 37828u) + 234234u;
  const unsigned z1 = x*(x*x*11687u + 26461u) + 137589u;
  const unsigned w1 = x*(x*x*15707u + 789221u) + 1376312589u;
  const float scale = 1.0f/4294967296.0f;
  return make_float4((float)x1*scale, (float)y1*scale, (float)z1*scale, (float)w1*scale);
}
HC_DEV float rndFloat1_Pseudo(HcRng& g)                   // crandom.h:77-83
{
  const unsigned x = NextState(g);
  const unsigned t = x*(x*x*15731u + 74323u) + 871483u;
  return (float)t*(1.0f/4294967296.0f);
}
HC_DEV float rndQmcSobolN(unsigned pos, int dim, const unsigned* __restrict__ table)    // crandom.h:228-236
{
  unsigned result = 0, data = pos;
  for (int bit = 0; bit < HC_QRNG_RESOLUTION_K; bit++, data >>= 1)
    if (data & 1u) result ^= table[bit + dim*HC_QRNG_RESOLUTION_K];
  return (float)(result + 1u)*(1.0f/(float)0x80000001U);
}
HC_DEV float rndQmcTab(HcRng& g, const int* __restrict__ tab, unsigned pos, int var, const unsigned* __restrict__ table)   // crandom.h:250-258
{
  const int dim = tab[var];
  return (dim < 0) ? rndFloat1_Pseudo(g) : rndQmcSobolN(pos, dim, table);
}

// ------------------------------------------------------------------------------------------------------------------ small helpers
HC_DEV float epsilonOfPos(float3 p)                       // cglobals.h:737 (one double product of two floats == the float product)
{
  return fmaxf(fmaxf(fabsf(p.x), fmaxf(fabsf(p.y), fabsf(p.z))), 2.0f*HC_GEPSILON)*HC_GEPSILON;
}
HC_DEV float misHeuristicPower1(float p) { return isfinite(p) ? fabsf(p) : 0.0f; }          // cglobals.h:738
HC_DEV float misWeightHeuristic(float a, float b)          // cglobals.h:741-745 (balance heuristic on |p|)
{
  const float w = misHeuristicPower1(a)/fmaxf(misHeuristicPower1(a) + misHeuristicPower1(b), HC_DEPSILON2);
  return isfinite(w) ? w : 0.0f;
}
HC_DEV float3 OffsRayPos(float3 hitPos, float3 n, float3 dir)                                 // cglobals.h:764-769
{
  const float s = dot(dir, n) < 0.0f ? -1.0f : 1.0f;
  const float e = epsilonOfPos(hitPos);
  return hitPos + s*e*n;
}
HC_DEV float3 OffsShadowRayPos(float3 hitPos, float3 n, float3 dir, float aux)                // cglobals.h:779-784
{
  const float s = dot(dir, n) < 0.0f ? -1.0f : 1.0f;
  const float e = epsilonOfPos(hitPos);
  return hitPos + s*(e + aux)*n;
}
HC_DEV float3 reflect3(float3 dir, float3 n) { return normalize((n*dot(dir, n)*(-2.0f)) + dir); }   // cglobals.h:688-693
HC_DEV float3 clamp3(float3 u, float a, float b) { return f3(clampf(u.x, a, b), clampf(u.y, a, b), clampf(u.z, a, b)); }

HC_DEV void CoordinateSystem(float3 v1, float3& v2, float3& v3)                               // cglobals.h:1502-1519
{
  if (fabsf(v1.x) > fabsf(v1.y))
  {
    const float invLen = (float)(1.0/sqrt(D(v1.x*v1.x + v1.z*v1.z)));
    v2 = f3((-1.0f)*v1.z*invLen, 0.0f, v1.x*invLen);
  }
  else
  {
    const float invLen = (float)(1.0/sqrt(D(v1.y*v1.y + v1.z*v1.z)));
    v2 = f3(0.0f, v1.z*invLen, (-1.0f)*v1.y*invLen);
  }
  v3 = cross(v1, v2);
}

HC_DEV float3 MapSampleToCosineDistribution(float r1, float r2, float3 direction, float3 hitNorm, float power)   // cglobals.h:1521-1561
{
  if (power >= 1e6f) return direction;
  const float sinPhi = hc_sin(2.0f*r1*3.141592654f);
  const float cosPhi = hc_cos(2.0f*r1*3.141592654f);
  const float cosTheta = hc_pow(1.0f - r2, 1.0f/(power + 1.0f));
  const float sinTheta = sqrtf(1.0f - cosTheta*cosTheta);
  const float3 dev = f3(sinTheta*cosPhi, sinTheta*sinPhi, cosTheta);
  float3 nx, nzT;
  CoordinateSystem(direction, nx, nzT);
  const float3 ny = nzT, nz = direction;              // the reference swaps ny and nz after CoordinateSystem
  float3 res = nx*dev.x + ny*dev.y + nz*dev.z;
  const float invSign = dot(direction, hitNorm) > 0.0f ? 1.0f : -1.0f;
  if (invSign*dot(res, hitNorm) < 0.0f)
    res = (-1.0f)*nx*dev.x + ny*dev.y - nz*dev.z;
  return res;
}

HC_DEV float3 MapSampleToModifiedCosineDistribution(float r1, float r2, float3 direction, float3 hitNorm, float power, bool& under)   // cglobals.h:1565-1601
{
  if (power >= 1e6f) return direction;        // NB: leaves `under` untouched, as the reference does
  const float sinPhi = hc_sin(2.0f*r1*3.141592654f);
  const float cosPhi = hc_cos(2.0f*r1*3.141592654f);
  const float sinTheta = (float)sqrt(1.0 - hc_d_pow(D(r2), D(2.0f/(power + 1.0f))));
  float3 dev;
  dev.x = sinTheta*cosPhi; dev.y = sinTheta*sinPhi;
  dev.z = sqrtf(1.0f - dev.x*dev.x - dev.y*dev.y);
  float3 nx, nzT;
  CoordinateSystem(direction, nx, nzT);
  const float3 ny = nzT, nz = direction;
  float3 res = nx*dev.x + ny*dev.y + nz*dev.z;
  under = false;
  const float invSign = dot(direction, hitNorm) >= 0.0f ? 1.0f : -1.0f;
  if (invSign*dot(res, hitNorm) < 0.0f)
  {
    res = (-1.0f)*nx*dev.x - ny*dev.y + nz*dev.z;
    under = true;
  }
  return res;
}

// ------------------------------------------------------------------------------------------------------------------ table access
HC_DEV const float* MaterialAt(const HcScene& s, int matId)            // materialAt, cfetch.h:201-211
{
  const int off = s.globals[s.materialsTableOffset + matId];
  return reinterpret_cast<const float*>(s.materials + off);
}
HC_DEV int   MatI(const float* m, int i) { return __float_as_int(m[i]); }
HC_DEV const float* LightAt(const HcScene& s, int id)                  // lightAt, clight.h:1739-1749
{
  if (id < 0) return nullptr;
  return reinterpret_cast<const float*>(s.globals + s.lightsOffset) + id*HC_LIGHT_DATA_SIZE;
}

// software textures: hc_texture.cuh (shared with the alpha-tested traversal)

// sample2DExt (cfetch.h:681-713) without procedural textures (none are supported: readProcTex then returns w = -1 and the image wins).
// samplerOffset counts float4 from the start of the material node (ReadSampler, cfetch.h:588-612).
// the fetch itself is out of line: 17 call sites would otherwise each inline the bilinear / sRGB code (untextured slots return early, inline)
static __device__ __noinline__ float3 Sample2DFetch(int samplerOffset, float2 tc, const float* mat, const int4* __restrict__ textures, const int* __restrict__ texTable)
{
  const int4*   mi = reinterpret_cast<const int4*>(mat);
  const float4* mf = reinterpret_cast<const float4*>(mat);
  const int4 header = mi[samplerOffset];
  const int flags = header.x; const float gamma = __int_as_float(header.y); const int texId = header.z;
  if (texId <= 0) return f3(1, 1, 1);
  const float4 row0 = mf[samplerOffset + 1], row1 = mf[samplerOffset + 2];
  const float2 tct = f2(row0.x*tc.x + row0.y*tc.y + row0.w, row1.x*tc.x + row1.y*tc.y + row1.w);     // mul2x4, cfetch.h:640-646
  const int offset = texTable[texId];
  float4 c = (offset >= 0) ? ReadImageSw4(textures + offset, tct, flags, (gamma != 1.0f)) : make_float4(1, 1, 1, 1);
  if (flags & HC_TEX_ALPHASRC_W) { c.x = c.w; c.y = c.w; c.z = c.w; }
  return f3(c.x, c.y, c.z);
}
HC_DEV float3 Sample2D(int samplerOffset, float2 tc, const float* mat, const HcScene& s)
{
  if (samplerOffset == HC_INVALID_TEXTURE || samplerOffset < 0) return f3(1, 1, 1);
  return Sample2DFetch(samplerOffset, tc, mat, s.textures, s.globals + s.texturesTableOffset);
}

// remapMaterialId (cglobals.h:2931-2983): per-instance material override, binary search for the first "from" id >= matId in the instance's list.
// The reference searches offsAndSize.y PAIRS although the driver stores the number of INTS there (RenderDriverRTE.cpp:1348-1365), so the search
// can run past the list; reads are clamped to the array here (the reference reads whatever follows).
HC_DEV int RemapMaterialId(const HcScene& s, int mId, int instId)
{
  if (mId < 0 || instId < 0 || instId >= s.remapInstSize || s.remapInst == nullptr || s.remapLists == nullptr || s.remapTable == nullptr) return mId;
  const int listId = s.remapInst[instId];
  if (listId < 0 || listId >= s.remapTableSize) return mId;
  const int2 os = s.remapTable[listId];
  int low = 0, high = os.y - 1;
  while (low <= high)
  {
    const int mid = low + ((high - low)/2);
    const int at = os.x + mid*2;
    const int from = (at >= 0 && at < s.remapListsSize) ? s.remapLists[at] : 0x7fffffff;
    if (from >= mId) high = mid - 1; else low = mid + 1;
  }
  if (high + 1 < os.y)
  {
    const int at = os.x + (high + 1)*2;
    if (at < 0 || at + 1 >= s.remapListsSize) return mId;
    return (s.remapLists[at] == mId) ? s.remapLists[at + 1] : mId;
  }
  return mId;
}

// ------------------------------------------------------------------------------------------------------------------ a10: surface evaluation
// surfaceEvalLS (ctrace.h:2005-2109) + the world transform of IntegratorCommon::surfaceEval / kernel_EvalSurface
// (CPUExp_Integrators_Common.cpp:193-242, CPUExp_Integrators_PT_Loop.cpp:35-84) and ComputeHit (shaders/trace.cl:130-227)
HC_DEV HcSurfaceHit SurfaceEval(const HcScene& s, float3 rpos, float3 rdir, const HcHit hit)
{
  const HcMat4 inv = loadMat4(s.instMatrices + hit.instId*4);
  const float3 o = mul4x3(inv, rpos), d = mul3x3(inv, rdir);

  const int meshOff = s.globals[s.geometryTableOffset + hit.geomId];
  const float4* mesh = s.geom + meshOff;
  const int4 h0 = reinterpret_cast<const int4*>(mesh)[0];      // vPosOffset vNormOffset vTexCoordOffset vIndicesOffset
  const int4 h2 = reinterpret_cast<const int4*>(mesh)[2];      // mIndicesOffset mIndicesNum vTangentOffset vTangentNum
  const int4 h3 = reinterpret_cast<const int4*>(mesh)[3];      // totalBytesNum polyShadowOffset ...
  const float4* vPos = mesh + h0.x; const float4* vNorm = mesh + h0.y; const float4* vTan = mesh + h2.z;
  const int* vIdx = reinterpret_cast<const int*>(mesh + h0.w);
  const int* mIdx = reinterpret_cast<const int*>(mesh + h2.x);
  const float* sOff = reinterpret_cast<const float*>(mesh + h3.y);

  HcSurfaceHit sh;
  sh.matId = mIdx[hit.primId];
  const int iA = vIdx[hit.primId*3 + 0], iB = vIdx[hit.primId*3 + 1], iC = vIdx[hit.primId*3 + 2];
  const float4 A1 = vPos[iA], B1 = vPos[iB], C1 = vPos[iC];
  const float4 A2 = vNorm[iA], B2 = vNorm[iB], C2 = vNorm[iC];
  const float3 A = f3(A1), B = f3(B1), C = f3(C1);
  const float3 An = f3(A2), Bn = f3(B2), Cn = f3(C2);
  const float2 At = f2(A1.w, A2.w), Bt = f2(B1.w, B2.w), Ct = f2(C1.w, C2.w);

  // triBaricentrics, ctrace.h:1986-2003
  const float3 e1 = B - A, e2 = C - A;
  const float3 pvec = cross(d, e2);
  const float det = dot(e1, pvec);
  const float invDet = 1.0f/det;
  const float3 tvec = o - A;
  const float v = dot(tvec, pvec)*invDet;
  const float3 qvec = cross(tvec, e1);
  const float u = dot(d, qvec)*invDet;
  const float2 uv = f2(u, v);

  const float w0 = 1.0f - uv.x - uv.y;
  sh.pos      = w0*A + uv.y*B + uv.x*C;
  sh.texCoord = w0*At + uv.y*Bt + uv.x*Ct;
  sh.normal   = w0*An + uv.y*Bn + uv.x*Cn;
  sh.t        = hit.t;
  sh.sRayOff  = sOff[hit.primId];

  const float4 At4 = vTan[iA], Bt4 = vTan[iB], Ct4 = vTan[iC];
  sh.flatNormal = normalize(cross(A - B, A - C));
  if (dot(d, sh.flatNormal) > 0.025f) sh.flatNormal = sh.flatNormal*(-1.0f);
  const float maxEdge = fmaxf(fmaxf(length(A - B), length(A - C)), length(B - C));
  if (sh.sRayOff > 1e-5f*maxEdge)
  {
    if (dot(d, sh.normal) > 0.120f)      { sh.normal = sh.normal*(-1.0f); sh.hfi = true; }
    else if (dot(d, sh.normal) > 0.0f)   { sh.normal = sh.flatNormal; sh.hfi = false; }
    else sh.hfi = false;
  }
  else
  {
    if (dot(d, sh.normal) > 0.0f) { sh.normal = sh.normal*(-1.0f); sh.hfi = true; }
    else sh.hfi = false;
  }
  const float handedness = (At4.w < 0.0f || Bt4.w < 0.0f || Ct4.w < 0.0f) ? -1.0f : 1.0f;
  sh.tangent   = normalize(w0*f3(At4) + uv.y*f3(Bt4) + uv.x*f3(Ct4));
  sh.biTangent = normalize(handedness > 0.0f ? cross(sh.normal, sh.tangent) : cross(sh.tangent, sh.normal));
  const bool badTangent = (!isfinite(sh.biTangent.x) || !isfinite(sh.biTangent.y) || !isfinite(sh.biTangent.z));
  if (fabsf(fabsf(dot(sh.normal, sh.tangent)) - 1.0f) < 1e-4f || badTangent)
    CoordinateSystem(sh.normal, sh.tangent, sh.biTangent);

  // to world space
  const HcMat4 m = inverse4x4(inv);
  const float multInv = (float)(1.0/sqrt(3.0));
  const float so = multInv*sh.sRayOff;
  const float3 shadowStart = mul3x3(m, f3(so, so, so));
  const HcMat4 nm = transpose(inv);
  HcSurfaceHit ws = sh;
  ws.pos        = mul4x3(m, sh.pos);
  ws.normal     = normalize(mul3x3(nm, sh.normal));
  ws.flatNormal = normalize(mul3x3(nm, sh.flatNormal));
  ws.tangent    = normalize(mul3x3(nm, sh.tangent));
  ws.biTangent  = normalize(mul3x3(nm, sh.biTangent));
  ws.t          = length(ws.pos - rpos);
  ws.sRayOff    = length(shadowStart);
  ws.matId      = RemapMaterialId(s, ws.matId, hit.instId);
  return ws;
}

// ------------------------------------------------------------------------------------------------------------------ a13/a14: BSDFs
HC_DEV float3 Mat3(const float* m, int i) { return f3(m[i], m[i + 1], m[i + 2]); }

// cubic spline glossiness -> Phong exponent (cosPowerFromGlosiness + its coefficient table, cmaterial.h:425-466; table = fitted data)
__device__ static const float kGlossCoeff[10][4] = {
  { 8.88178419700125e-14f, -1.77635683940025e-14f, 5.0f, 1.0f },        { 357.142857142857f, -35.7142857142857f, 5.0f, 1.5f },
  { -2142.85714285714f, 428.571428571429f, 8.57142857142857f, 2.0f },   { 428.571428571431f, -42.8571428571432f, 30.0f, 5.0f },
  { 2095.23809523810f, -152.380952380952f, 34.2857142857143f, 8.0f },   { -4761.90476190476f, 1809.52380952381f, 66.6666666666667f, 12.0f },
  { 9914.71215351811f, 1151.38592750533f, 285.714285714286f, 32.0f },   { 45037.7068059246f, 9161.90096119855f, 813.432835820895f, 82.0f },
  { 167903.678757035f, 183240.189801913f, 3996.94423223835f, 300.0f },  { -20281790.7444668f, 6301358.14889336f, 45682.0925553320f, 2700.0f } };
HC_DEV float cosPowerFromGlosiness(float g)
{
  const float cMax = 1000000.0f;
  const float x = g;
  const int k = (fabsf(x - 1.0f) < 1e-5f) ? 10 : (int)(x*10.0f);
  const float x1 = (x - (float)k*0.1f);
  if (k == 10 || x >= 0.99f) return cMax;
  return kGlossCoeff[k][3] + kGlossCoeff[k][2]*x1 + kGlossCoeff[k][1]*x1*x1 + kGlossCoeff[k][0]*x1*x1*x1;
}

// glossiness slot shared by Phong / GGX (phongGlosiness, ggxGlosiness: cmaterial.h:918-931, 1197-1210; same offsets 16/17/18)
HC_DEV float Glosiness(const float* m, float2 tc, const HcScene& s)
{
  if (MatI(m, HC_PHONG_GLOSINESS_TEXID_OFFSET) != HC_INVALID_TEXTURE)
  {
    const float3 gc = Sample2D(MatI(m, HC_PHONG_GLOSINESS_TEXMATRIXID_OFFSET), tc, m, s);
    return clampf(m[HC_PHONG_GLOSINESS_OFFSET]*maxcomp(gc), 0.0f, 0.99f);
  }
  return m[HC_PHONG_GLOSINESS_OFFSET];
}

// ---- Lambert (cmaterial.h:212-257)
HC_DEV float3 LambertColor(const float* m, float2 tc, const HcScene& s)
{
  const float3 tex = Sample2D(MatI(m, HC_LAMBERT_TEXMATRIXID_OFFSET), tc, m, s);
  return clamp3(tex*Mat3(m, HC_LAMBERT_COLORX_OFFSET), 0.0f, 1.0f);
}
HC_DEV void LambertSample(const float* m, float r1, float r2, float3 n, float2 tc, const HcScene& s, HcMatSample& out)
{
  const float3 color = LambertColor(m, tc, s);
  const float3 newDir = MapSampleToCosineDistribution(r1, r2, n, n, 1.0f);
  const float cosTheta = dot(newDir, n);
  out.direction = newDir;
  out.pdf = cosTheta*HC_INV_PI;
  out.color = color*HC_INV_PI;
  if (cosTheta <= HC_DEPSILON) out.color = f3(0, 0, 0);
  out.flags = HC_RAY_EVENT_D;
}

// ---- Oren-Nayar (cmaterial.h:264-370; helpers cmatpbrt.h:17-31).  Colour / texture slots coincide with Lambert's (cmaterial.h:1912).
#define HC_ORENNAYAR_A 16
#define HC_ORENNAYAR_B 17
HC_DEV float OrennayarFunc(float3 l, float3 v, float3 n, float A, float B)                 // cmaterial.h:288-334
{
  const float cosTheta_wi = dot(l, n), cosTheta_wo = dot(v, n);
  const float sinTheta_wi = sqrtf(fmaxf(0.0f, 1.0f - cosTheta_wi*cosTheta_wi));
  const float sinTheta_wo = sqrtf(fmaxf(0.0f, 1.0f - cosTheta_wo*cosTheta_wo));
  float3 nx, ny; const float3 nz = n;
  CoordinateSystem(nz, nx, ny);
  const float3 wo = f3(-dot(v, nx), -dot(v, ny), -dot(v, nz));
  const float3 wi = f3(-dot(l, nx), -dot(l, ny), -dot(l, nz));
  float maxcos = 0.0f;
  if ((double)sinTheta_wi > 1e-4 && (double)sinTheta_wo > 1e-4)                            // the reference compares against double literals
  {
    const float sinphii = (sinTheta_wi == 0.0f) ? 0.0f : clampf(wi.y/sinTheta_wi, -1.0f, 1.0f), cosphii = (sinTheta_wi == 0.0f) ? 1.0f : clampf(wi.x/sinTheta_wi, -1.0f, 1.0f);
    const float sinphio = (sinTheta_wo == 0.0f) ? 0.0f : clampf(wo.y/sinTheta_wo, -1.0f, 1.0f), cosphio = (sinTheta_wo == 0.0f) ? 1.0f : clampf(wo.x/sinTheta_wo, -1.0f, 1.0f);
    const float dcos = cosphii*cosphio + sinphii*sinphio;
    maxcos = fmaxf(0.0f, dcos);
  }
  float sinalpha, tanbeta;
  if (fabsf(cosTheta_wi) > fabsf(cosTheta_wo)) { sinalpha = sinTheta_wo; tanbeta = sinTheta_wi/fmaxf(fabsf(cosTheta_wi), HC_DEPSILON); }
  else                                         { sinalpha = sinTheta_wi; tanbeta = sinTheta_wo/fmaxf(fabsf(cosTheta_wo), HC_DEPSILON); }
  return A + B*maxcos*sinalpha*tanbeta;
}
HC_DEV float3 OrennayarEvalBxDF(const float* m, float3 l, float3 v, float3 n, float2 tc, const HcScene& s)
{
  return LambertColor(m, tc, s)*HC_INV_PI*OrennayarFunc(l, v, n, m[HC_ORENNAYAR_A], m[HC_ORENNAYAR_B]);
}
HC_DEV void OrennayarSample(const float* m, float r1, float r2, float3 rayDir, float3 n, float2 tc, const HcScene& s, HcMatSample& out)   // cmaterial.h:351-370
{
  const float3 color = LambertColor(m, tc, s);
  const float3 newDir = MapSampleToCosineDistribution(r1, r2, n, n, 1.0f);
  const float cosTheta = dot(newDir, n);
  out.direction = newDir;
  out.pdf = cosTheta*HC_INV_PI;
  out.color = color*HC_INV_PI*OrennayarFunc(newDir, (-1.0f)*rayDir, n, m[HC_ORENNAYAR_A], m[HC_ORENNAYAR_B]);
  if (cosTheta <= HC_DEPSILON) out.color = f3(0, 0, 0);
  out.flags = HC_RAY_EVENT_D;
}

// ---- translucent Lambert (cmaterial.h:1850-1910): diffuse transmission through a thin sheet.  Colour / texture slots coincide with Lambert's.
HC_DEV float TranslucentEvalPDF(float3 l, float3 v, float3 n)
{
  const float sign1 = dot(l, n) > 0 ? 1.0f : -1.0f, sign2 = dot(v, n) > 0 ? 1.0f : -1.0f;
  const float coeff = (sign1*sign2 < 0.0f) ? 1.0f : 0.0f;
  return fabsf(dot(l, n))*HC_INV_PI*coeff;
}
HC_DEV float3 TranslucentEvalBxDF(const float* m, float3 l, float3 v, float3 n, float2 tc, const HcScene& s)
{
  const float sign1 = dot(l, n) > 0 ? 1.0f : -1.0f, sign2 = dot(v, n) > 0 ? 1.0f : -1.0f;
  const float coeff = (sign1*sign2 < 0.0f) ? 1.0f : 0.0f;
  return LambertColor(m, tc, s)*coeff*HC_INV_PI;
}
HC_DEV void TranslucentSample(const float* m, float r1, float r2, float3 n, float2 tc, const HcScene& s, HcMatSample& out)
{
  const float3 kd = LambertColor(m, tc, s);
  const float3 nn = (-1.0f)*n;
  const float3 newDir = MapSampleToCosineDistribution(r1, r2, nn, nn, 1.0f);
  const float cosTheta = dot(newDir, nn);
  out.direction = newDir;
  out.pdf = cosTheta*HC_INV_PI;
  out.color = kd*HC_INV_PI;
  if (cosTheta <= 1e-6f) out.color = f3(0, 0, 0);
  out.flags = (HC_RAY_EVENT_D | HC_RAY_EVENT_T);
}

// ---- thin glass (cmaterial.h:471-555): straight-through (optionally glossy) transmission, invisible to explicit light sampling
HC_DEV void ThinglassSample(const float* m, float r1, float r2, float3 rayDir, float3 n, float2 tc, const HcScene& s, HcMatSample& out)
{
  const float3 tex = Sample2D(MatI(m, HC_PHONG_TEXMATRIXID_OFFSET), tc, m, s);
  const float3 gc = Sample2D(MatI(m, HC_PHONG_GLOSINESS_TEXMATRIXID_OFFSET), tc, m, s);           // thinglassCosPower: the slots coincide with Phong's
  const float cosPower = cosPowerFromGlosiness(clampf(m[HC_PHONG_GLOSINESS_OFFSET]*maxcomp(gc), 0.0f, 1.0f));
  float pdf = 1.0f, fVal = 1.0f;
  if (cosPower < 1e6f)
  {
    bool under = false;
    const float3 oldDir = rayDir;
    rayDir = MapSampleToModifiedCosineDistribution(r1, r2, rayDir, (-1.0f)*n, cosPower, under);
    const float cosTheta = clampf(dot(oldDir, rayDir), 0.0f, (float)(HC_M_PI_D*D(0.499995f)));
    fVal = (float)(D((cosPower + 2.0f)*HC_INV_TWOPI)*hc_d_pow(D(cosTheta), D(cosPower)));
    if (under) fVal = 0.0f;
    pdf = (float)(hc_d_pow(D(cosTheta), D(cosPower))*D(cosPower + 1.0f)*D(0.5f*HC_INV_PI));
  }
  const float cosThetaOut = dot(rayDir, n);
  const float cosMult = 1.0f/fmaxf(fabsf(cosThetaOut), 1e-6f);
  out.direction = rayDir;
  out.pdf = pdf;
  out.color = fVal*Mat3(m, HC_PHONG_COLORX_OFFSET)*tex*cosMult;
  if (cosThetaOut >= -1e-6f) out.color = f3(0, 0, 0);
  out.flags = (HC_RAY_EVENT_S | HC_RAY_EVENT_T | HC_RAY_EVENT_TNINGLASS);
}

// ---- perfect mirror (cmaterial.h:385-421)
HC_DEV void MirrorSample(const float* m, float3 rayDir, float3 n, float2 tc, const HcScene& s, HcMatSample& out)
{
  const float3 tex = Sample2D(MatI(m, HC_MIRROR_TEXMATRIXID_OFFSET), tc, m, s);
  float3 newDir = reflect3(rayDir, n);
  if (dot(rayDir, n) > 0.0f) newDir = rayDir;
  const float cosOut = dot(newDir, n);
  out.direction = newDir;
  out.pdf = 1.0f;
  out.color = Mat3(m, HC_MIRROR_COLORX_OFFSET)*tex*(1.0f/fmaxf(cosOut, 1e-6f));
  if (cosOut <= 1e-6f) out.color = f3(0, 0, 0);
  out.flags = HC_RAY_EVENT_S;
}

// ---- modified Phong (cmaterial.h:908-1017)
HC_DEV float PhongEnergyFix(float dotRL, float3 l, float3 n) { return dotRL/fmaxf(dot(n, l), 1e-6f); }
HC_DEV float PhongEvalPDF(const float* m, float3 l, float3 v, float3 n, float2 tc, const HcScene& s)
{
  if (dot(n, v) < 1e-6f || dot(n, l) < 1e-6f) return 1.0f;
  const float cosPower = cosPowerFromGlosiness(Glosiness(m, tc, s));
  const float3 r = reflect3((-1.0f)*v, n);
  const float cosTheta = clampf(fabsf(dot(l, r)), 0.0f, 1.0f);
  return (float)(hc_d_pow(D(cosTheta), D(cosPower))*D(cosPower + 1.0f)*D(HC_INV_TWOPI));
}
HC_DEV float3 PhongEvalBxDF(const float* m, float3 l, float3 v, float3 n, float2 tc, const HcScene& s)
{
  if (dot(n, v) < 1e-6f || dot(n, l) < 1e-6f) return f3(0, 0, 0);
  const float3 tex = Sample2D(MatI(m, HC_PHONG_TEXMATRIXID_OFFSET), tc, m, s);
  const float3 color = clamp3(Mat3(m, HC_PHONG_COLORX_OFFSET)*tex, 0.0f, 1.0f);
  const float cosPower = cosPowerFromGlosiness(Glosiness(m, tc, s));
  const float3 r = reflect3((-1.0f)*v, n);
  const float cosAlpha = clampf(dot(l, r), 0.0f, 1.0f);
  const bool fix = (MatI(m, HC_PLAIN_MAT_FLAGS_OFFSET) & HC_PLAIN_MATERIAL_ENERGY_FIX_OR_MULTISCATTER) != 0;
  const float energyFix = fix ? PhongEnergyFix(cosAlpha, l, n) : 1.0f;
  return color*(cosPower + 2.0f)*HC_INV_TWOPI*hc_pow(cosAlpha, cosPower)*energyFix;
}
HC_DEV void PhongSample(const float* m, float r1, float r2, float3 rayDir, float3 n, float2 tc, const HcScene& s, HcMatSample& out)
{
  const float3 tex = Sample2D(MatI(m, HC_PHONG_TEXMATRIXID_OFFSET), tc, m, s);
  const float3 color = clamp3(Mat3(m, HC_PHONG_COLORX_OFFSET)*tex, 0.0f, 1.0f);
  const float gloss = Glosiness(m, tc, s);
  const float cosPower = cosPowerFromGlosiness(gloss);
  bool under = false;
  const float3 r = reflect3(rayDir, n);
  const float3 newDir = MapSampleToModifiedCosineDistribution(r1, r2, r, n, cosPower, under);
  const float3 v = rayDir*(-1.0f), l = newDir;
  if (dot(n, v) < 1e-6f || dot(n, l) < 1e-6f || under) { out.color = f3(0, 0, 0); out.pdf = 1.0f; }
  else
  {
    const float cosAlpha = clampf(dot(newDir, r), 0.0f, 1.0f);
    const float eqTemp = (float)(hc_d_pow(D(cosAlpha), D(cosPower))*D(HC_INV_TWOPI));
    const bool fix = (MatI(m, HC_PLAIN_MAT_FLAGS_OFFSET) & HC_PLAIN_MATERIAL_ENERGY_FIX_OR_MULTISCATTER) != 0;
    const float energyFix = fix ? PhongEnergyFix(cosAlpha, newDir, n) : 1.0f;
    out.pdf = eqTemp*(cosPower + 1.0f);
    out.color = eqTemp*(cosPower + 2.0f)*color*energyFix;
  }
  out.direction = newDir;
  out.flags = (gloss >= 0.99f) ? HC_RAY_EVENT_S : HC_RAY_EVENT_G;
}

// ---- Blinn distribution + Torrance-Sparrow geometry term (cmaterial.h:1020-1160, cmatpbrt.h:33-105).  Slots coincide with Phong's.
HC_DEV float TorranceSparrowG(float NdotWh, float NdotWo, float NdotWi, float WOdotWhAbs)          // cmatpbrt.h:44-64
{
  const float WOdotWh = fmaxf(WOdotWhAbs, HC_DEPSILON);
  return fminf(1.0f, fminf(2.0f*NdotWh*NdotWo/WOdotWh, 2.0f*NdotWh*NdotWi/WOdotWh));
}
HC_DEV float TorranceSparrowGF1(float3 wo, float3 wi)                                              // local frame, cmatpbrt.h:66-84
{
  const float cosThetaO = fabsf(wo.z), cosThetaI = fabsf(wi.z);
  if (cosThetaI == 0.0f || cosThetaO == 0.0f) return 0.0f;
  float3 wh = wi + wo;
  if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return 0.0f;
  wh = normalize(wh);
  const float G = TorranceSparrowG(fabsf(wh.z), fabsf(wo.z), fabsf(wi.z), fabsf(dot(wo, wh)));
  return fminf(G*1.0f/fmaxf(4.0f*cosThetaI*cosThetaO, HC_DEPSILON), 250.0f);
}
HC_DEV float TorranceSparrowGF2(float3 wo, float3 wi, float3 n)                                    // world frame, cmatpbrt.h:86-105
{
  const float cosThetaO = fabsf(dot(wo, n)), cosThetaI = fabsf(dot(wi, n));
  if (cosThetaI == 0.0f || cosThetaO == 0.0f) return 0.0f;
  float3 wh = wi + wo;
  if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return 0.0f;
  wh = normalize(wh);
  const float G = TorranceSparrowG(fabsf(dot(wh, n)), fabsf(dot(wo, n)), fabsf(dot(wi, n)), fabsf(dot(wo, wh)));
  return fminf(G*1.0f/fmaxf(4.0f*cosThetaI*cosThetaO, HC_DEPSILON), 10.0f);          // 10 here, 250 in the local-frame variant
}
HC_DEV float BlinnEvalPDF(const float* m, float3 l, float3 v, float3 n, float2 tc, const HcScene& s)
{
  if (dot(n, v) < 1e-6f || dot(n, l) < 1e-6f) return 1.0f;
  const float exponent = cosPowerFromGlosiness(Glosiness(m, tc, s));
  const float3 wh = normalize(l + v);
  const float costheta = fabsf(dot(wh, n));
  return (float)((D(exponent + 1.0f)*hc_d_pow(D(costheta), D(exponent)))/D(HC_M_TWOPI*4.0f*dot(l, wh)));
}
HC_DEV float3 BlinnEvalBxDF(const float* m, float3 l, float3 v, float3 n, float2 tc, const HcScene& s)
{
  if (dot(n, v) < 1e-6f || dot(n, l) < 1e-6f) return f3(0, 0, 0);
  const float3 tex = Sample2D(MatI(m, HC_PHONG_TEXMATRIXID_OFFSET), tc, m, s);
  const float3 color = clamp3(Mat3(m, HC_PHONG_COLORX_OFFSET)*tex, 0.0f, 1.0f);
  const float exponent = cosPowerFromGlosiness(Glosiness(m, tc, s));
  const float3 wh = normalize(l + v);
  const float cosThetaH = fabsf(dot(wh, n));
  const float Dh = (float)(D((exponent + 2.0f)*HC_INV_TWOPI)*hc_d_pow(D(cosThetaH), D(exponent)));
  return color*Dh*TorranceSparrowGF2(l, v, n);
}
HC_DEV void BlinnSample(const float* m, float r1, float r2, float3 rayDir, float3 n, float2 tc, const HcScene& s, HcMatSample& out)
{
  const float3 tex = Sample2D(MatI(m, HC_PHONG_TEXMATRIXID_OFFSET), tc, m, s);
  const float3 color = clamp3(Mat3(m, HC_PHONG_COLORX_OFFSET)*tex, 0.0f, 1.0f);
  const float gloss = Glosiness(m, tc, s);
  float3 nx, ny; const float3 nz = n;
  CoordinateSystem(nz, nx, ny);
  const float3 wo = f3(-dot(rayDir, nx), -dot(rayDir, ny), -dot(rayDir, nz));
  const float exponent = cosPowerFromGlosiness(gloss);
  const float costheta = hc_pow(r1, 1.0f/(exponent + 1.0f));
  const float sintheta = sqrtf(fmaxf(0.0f, 1.0f - costheta*costheta));
  const float phi = r2*HC_M_TWOPI;
  const float3 wh = f3((float)(D(sintheta)*hc_d_cos(D(phi))), (float)(D(sintheta)*hc_d_sin(D(phi))), costheta);
  const float3 wi = (2.0f*dot(wo, wh)*wh) - wo;
  const float3 newDir = normalize(wi.x*nx + wi.y*ny + wi.z*nz);
  const float3 v = rayDir*(-1.0f), l = newDir;
  if (dot(n, v) < 1e-6f || dot(n, l) < 1e-6f) { out.color = f3(0, 0, 0); out.pdf = 1.0f; }
  else
  {
    const float Dh = (float)(D((exponent + 2.0f)*HC_INV_TWOPI)*hc_d_pow(D(costheta), D(exponent)));
    out.color = color*Dh*TorranceSparrowGF1(wo, wi);
    out.pdf = (float)((D(exponent + 1.0f)*hc_d_pow(D(costheta), D(exponent)))/D(fmaxf(HC_M_TWOPI*4.0f*dot(wo, wh), HC_DEPSILON)));
  }
  out.direction = newDir;
  out.flags = (gloss >= 0.99f) ? HC_RAY_EVENT_S : HC_RAY_EVENT_G;
}

// ---- GGX (cmaterial.h:1212-1285, 1317-1381, 1454-1520)
HC_DEV float SmithGGXMasking(float dotNV, float roughSqr)
{
  const float denomC = (float)(sqrt(D(roughSqr + (1.0f - roughSqr)*dotNV*dotNV)) + D(dotNV));
  return 2.0f*dotNV/fmaxf(denomC, 1e-6f);
}
HC_DEV float SmithGGXMaskingShadowing(float dotNL, float dotNV, float roughSqr)
{
  const float denomA = (float)(D(dotNV)*sqrt(D(roughSqr + (1.0f - roughSqr)*dotNL*dotNL)));
  const float denomB = (float)(D(dotNL)*sqrt(D(roughSqr + (1.0f - roughSqr)*dotNV*dotNV)));
  return 2.0f*dotNL*dotNV/fmaxf(denomA + denomB, 1e-6f);
}
HC_DEV float GGX_Distribution(float cosNH, float alpha)
{
  const float alpha2 = alpha*alpha;
  const float nh2 = clampf(cosNH*cosNH, 0.0f, 1.0f);
  const float den = nh2*alpha2 + (1.0f - nh2);
  return (float)(D(alpha2)/fmax(HC_PI_D*D(den)*D(den), D(1e-6f)));
}
HC_DEV float3 GgxVndf(float3 wo, float roughness, float u1, float u2)
{
  const float3 v = normalize(f3(wo.x*roughness, wo.y*roughness, wo.z));
  const float3 XAxis = f3(1.0f, 0.0f, 0.0f), ZAxis = f3(0.0f, 0.0f, 1.0f);
  const float3 t1 = (v.z < 0.999f) ? normalize(cross(v, ZAxis)) : XAxis;
  const float3 t2 = cross(t1, v);
  const float a = 1.0f/(1.0f + v.z);
  const float r = sqrtf(u1);
  const float phi = (u2 < a) ? (float)(D(u2/a)*HC_PI_D) : (float)(HC_PI_D + D((u2 - a)/(1.0f - a))*HC_PI_D);
  const float p1 = (float)(D(r)*hc_d_cos(D(phi)));
  const float p2 = (float)(D(r)*hc_d_sin(D(phi))*D((u2 < a) ? 1.0f : v.z));
  const float3 n = p1*t1 + p2*t2 + sqrtf(fmaxf(0.0f, 1.0f - p1*p1 - p2*p2))*v;
  return normalize(f3(roughness*n.x, roughness*n.y, fmaxf(0.0f, n.z)));
}
// GetMultiscatteringFrom2dTable + BilinearFrom2dTable (cmaterial.h:61-93, 152-159) over EngineGlobals::m_essGgx2017Table
HC_DEV float3 GgxMultiscatter(const HcScene& s, float roughness, float dotNV, float3 color)
{
  const unsigned short* tab = reinterpret_cast<const unsigned short*>(reinterpret_cast<const char*>(s.globals) + s.essGgxTableOffsetBytes);
  const int W = 64, H = 64;
  float x = clampf(dotNV*(float)W, 0.0f, W - 1.0001f), y = clampf(roughness*(float)H, 0.0f, H - 1.0001f);
  const int fy = (int)floorf(y), fx = (int)floorf(x);
  const int d1 = fy*W + fx, d2 = d1 + 1, d3 = (fy + 1)*W + fx, d4 = d3 + 1;
  const float dx = x - fx, dy = y - fy;
  const float m1 = (1.0f - dx)*(1.0f - dy), m2 = dx*(1.0f - dy), m3 = dy*(1.0f - dx), m4 = dx*dy;
  float val = 1.0f;
  if (fy >= 0 && fx >= 0 && fy <= H - 2 && fx <= W - 2) val = tab[d1]*m1 + tab[d2]*m2 + tab[d3]*m3 + tab[d4]*m4;
  const float Ess = val*(1.0f/65535.0f);
  const float3 t = color*(1.0f - Ess)/fmaxf(Ess, 1e-6f);
  return f3(1.0f + t.x, 1.0f + t.y, 1.0f + t.z);
}
// GetMultiscatteringFrom3dTable + BilinearFrom3dTable (cmaterial.h:96-149, 161-196) over EngineGlobals::m_essTranspTable (64 x 64 x 64: cosine, roughness,
// relative IOR in [0.4166, 2.4]); the X / Y range test and the 1/65536 scale are the reference's
HC_DEV float3 GlassMultiscatter(const HcScene& s, float roughness, float dotNV, float ior, float3 color)
{
  if (!(ior >= 0.4166f && ior <= 2.4f)) return f3(1.0f, 1.0f, 1.0f);
  const unsigned short* tab = reinterpret_cast<const unsigned short*>(reinterpret_cast<const char*>(s.globals) + HC_EG_m_essTranspTable);
  const int W = 64, H = 64, Dp = 64, plane = W*H;
  const float iorNormal = (ior - 0.4166f)/(2.4f - 0.4166f);
  const float x = clampf(dotNV*(float)W, 0.0f, W - 1.0001f), y = clampf(roughness*(float)H, 0.0f, H - 1.0001f), z = clampf(iorNormal*(float)Dp, 0.0f, Dp - 1.0001f);
  const int fx = (int)floorf(x), fy = (int)floorf(y), fz = (int)floorf(z);
  const int z0 = fz*plane, z1 = (fz + 1)*plane;
  const int d1 = z0 + fy*W + fx, d3 = z0 + (fy + 1)*W + fx, d2 = d1 + 1, d4 = d3 + 1;
  const int d5 = z1 + fy*W + fx, d7 = z1 + (fy + 1)*W + fx, d6 = d5 + 1, d8 = d7 + 1;
  const float dx = x - fx, dy = y - fy, dz = z - fz;
  const float m1 = (1.0f - dx)*(1.0f - dy), m2 = dx*(1.0f - dy), m3 = dy*(1.0f - dx), m4 = dx*dy;
  float val = 1.0f;
  if (fy >= 0 && fx >= 0 && fy <= H - 2 && fx <= W - 2)
  {
    const float p1 = tab[d1]*m1 + tab[d2]*m2 + tab[d3]*m3 + tab[d4]*m4;
    const float p2 = tab[d5]*m1 + tab[d6]*m2 + tab[d7]*m3 + tab[d8]*m4;
    val = p1 + dz*(p2 - p1);
  }
  const float Ess = val*(1.0f/65536.0f);
  const float3 t = color*(1.0f - Ess)/fmaxf(Ess, 1e-6f);
  return f3(1.0f + t.x, 1.0f + t.y, 1.0f + t.z);
}
HC_DEV float3 GgxColor(const float* m, float2 tc, const HcScene& s)
{
  const float3 tex = Sample2D(MatI(m, HC_GGX_TEXMATRIXID_OFFSET), tc, m, s);
  return clamp3(Mat3(m, HC_GGX_COLORX_OFFSET)*tex, 0.0f, 1.0f);
}
HC_DEV float Ggx2EvalPDF(const float* m, float3 l, float3 v, float3 n, float2 tc, const HcScene& s)
{
  const float dotNV = dot(n, v), dotNL = dot(n, l);
  if (dotNV < 1e-6f || dotNL < 1e-6f) return 1.0f;
  const float gloss = Glosiness(m, tc, s);
  const float roughness = 1.0f - gloss, roughSqr = roughness*roughness;
  const float3 h = normalize(v + l);
  const float dotNH = dot(n, h), dotHV = dot(h, v);
  const float G1 = SmithGGXMasking(dotNV, roughSqr);
  const float Dd = GGX_Distribution(dotNH, roughSqr);
  const float Dv = Dd*G1*dotHV/fmaxf(dotNV, 1e-6f);
  const float jacob = 1.0f/fmaxf(4.0f*dotHV, 1e-6f);
  return Dv*jacob;
}
HC_DEV float3 GgxEvalBxDF(const float* m, float3 l, float3 v, float3 n, float2 tc, const HcScene& s)
{
  const float dotNV = dot(n, v), dotNL = dot(n, l);
  if (dotNV < 1e-6f || dotNL < 1e-6f) return f3(0, 0, 0);
  const float3 color = GgxColor(m, tc, s);
  const float gloss = Glosiness(m, tc, s);
  const float roughness = 1.0f - gloss, roughSqr = roughness*roughness;
  const float3 h = normalize(v + l);
  const float dotNH = dot(n, h);
  const float Dd = GGX_Distribution(dotNH, roughSqr);
  const float G = SmithGGXMaskingShadowing(dotNL, dotNV, roughSqr);
  const float Pss = Dd*G/fmaxf(4.0f*dotNV*dotNL, 1e-6f);
  float3 Pms = f3(1, 1, 1);
  if (MatI(m, HC_PLAIN_MAT_FLAGS_OFFSET) & HC_PLAIN_MATERIAL_ENERGY_FIX_OR_MULTISCATTER) Pms = GgxMultiscatter(s, roughness, dotNV, color);
  return color*Pss*Pms;
}
HC_DEV void GgxSample2(const float* m, float r1, float r2, float3 rayDir, float3 nrm, float2 tc, const HcScene& s, HcMatSample& out)
{
  const float3 color = GgxColor(m, tc, s);
  const float gloss = Glosiness(m, tc, s);
  const float roughness = 1.0f - gloss, roughSqr = roughness*roughness;
  float3 nx, ny; const float3 nz = nrm;
  CoordinateSystem(nz, nx, ny);
  float Pss = 1.0f; float3 Pms = f3(1.0f, 1.0f, 1.0f);
  const float3 wo = f3(-dot(rayDir, nx), -dot(rayDir, ny), -dot(rayDir, nz));
  const float3 wh = GgxVndf(wo, roughSqr, r1, r2);
  const float3 wi = 2.0f*dot(wo, wh)*wh - wo;
  const float3 newDir = normalize(wi.x*nx + wi.y*ny + wi.z*nz);
  const float3 v = rayDir*(-1.0f), l = newDir;
  const float dotNV = dot(nrm, v), dotNL = dot(nrm, l);
  if (dotNV < 1e-6f || dotNL < 1e-6f) { Pss = 0.0f; out.pdf = 1.0f; }
  else
  {
    const float3 h = normalize(v + l);
    const float dotNH = dot(nrm, h), dotHV = dot(h, v);
    const float Dd = GGX_Distribution(dotNH, roughSqr);
    const float G1 = SmithGGXMasking(dotNV, roughSqr);
    const float G2 = SmithGGXMaskingShadowing(dotNL, dotNV, roughSqr);
    Pss = Dd*G2/fmaxf(4.0f*dotNV, 1e-6f);
    const float Dv = Dd*G1*dotHV/fmaxf(dotNV, 1e-6f);
    const float jacob = 1.0f/fmaxf(4.0f*dotHV, 1e-6f);
    out.pdf = Dv*jacob;
    if (MatI(m, HC_PLAIN_MAT_FLAGS_OFFSET) & HC_PLAIN_MATERIAL_ENERGY_FIX_OR_MULTISCATTER) Pms = GgxMultiscatter(s, roughness, dotNV, color);
  }
  out.direction = newDir;
  out.color = color*Pss*Pms/fmaxf(dotNL, 1e-6f);
  out.flags = (gloss >= 0.99f) ? HC_RAY_EVENT_S : HC_RAY_EVENT_G;
}

// ---- GGX glass (glassGloss cmaterial.h:610-618, myRefractGgx :674-706, GlassGGXSampleAndEvalBRDF :775-883); eval is zero (:620-629)
HC_DEV void GlassGgxSample(const float* m, float3 rands, float3 rayDir, float3 nrm, float2 tc, bool hfi, const HcScene& s, HcMatSample& out)
{
  const float3 tex = Sample2D(MatI(m, HC_GLASS_TEXMATRIXID_OFFSET), tc, m, s);
  const float3 color = clamp3(Mat3(m, HC_GLASS_COLORX_OFFSET)*tex, 0.0f, 1.0f);
  const float3 gc = Sample2D(MatI(m, HC_GLASS_GLOSINESS_TEXMATRIXID_OFFSET), tc, m, s);
  const float gloss = clampf(m[HC_GLASS_GLOSINESS]*maxcomp(gc), 0.0f, 1.0f);
  const float roughness = clampf(1.0f - gloss, 0.0f, 1.0f), roughSqr = roughness*roughness;
  const float IOR = m[HC_GLASS_IOR_OFFSET];
  const float3 normal2 = hfi ? (-1.0f)*nrm : nrm;
  bool spec = true; float Pss = 1.0f; float3 Pms = f3(1.0f, 1.0f, 1.0f);
  out.pdf = 1.0f;

  // myRefractGgx(ray_dir, normal2, IOR, 1.0f, rands.z)
  float3 rdir; bool success; float reta = 1.0f/IOR;
  {
    float3 nn = normal2;
    float cosTheta = dot(nn, rayDir)*(-1.0f);
    if (cosTheta < 0.0f) { cosTheta = cosTheta*(-1.0f); nn = nn*(-1.0f); reta = 1.0f/reta; }
    const float dotVN = cosTheta*(-1.0f);
    const float k = 1.0f - reta*reta*(1.0f - cosTheta*cosTheta);
    if (k > 0.0f) { rdir = normalize(reta*rayDir + (float)(D(reta*cosTheta) - sqrt(D(k)))*nn); success = true; }
    else          { rdir = normalize((nn*dotVN*(-2.0f)) + rayDir); success = false; reta = 1.0f; }
  }

  if (gloss < 0.999f)
  {
    spec = false;
    float eta = 1.0f/IOR;
    const float cosTheta = dot(normal2, rayDir)*(-1.0f);
    if (cosTheta < 0.0f) eta = 1.0f/eta;
    float3 nx, ny; const float3 nz = nrm;
    CoordinateSystem(nz, nx, ny);
    const float3 wo = f3(-dot(rayDir, nx), -dot(rayDir, ny), -dot(rayDir, nz));
    const float3 wh = GgxVndf(wo, roughSqr, rands.x, rands.y);
    const float dotWoWh = dot(wo, wh);
    float3 newDir;
    const float radicand = 1.0f + eta*eta*(dotWoWh*dotWoWh - 1.0f);
    if (radicand > 0.0f) { newDir = (float)(D(eta*dotWoWh) - sqrt(D(radicand)))*wh - eta*wo; success = true; reta = eta; }
    else                 { newDir = 2.0f*dotWoWh*wh - wo; success = false; reta = 1.0f; }
    rdir = normalize(newDir.x*nx + newDir.y*ny + newDir.z*nz);
    const float3 v = rayDir*(-1.0f), l = rdir;
    const float dotNV = fabsf(dot(nrm, v)), dotNL = fabsf(dot(nrm, l));
    const float G1 = SmithGGXMasking(dotNV, roughSqr);
    const float G2 = SmithGGXMaskingShadowing(dotNL, dotNV, roughSqr);
    Pss = G2/fmaxf(G1, 1e-6f);
    if (MatI(m, HC_PLAIN_MAT_FLAGS_OFFSET) & HC_PLAIN_MATERIAL_ENERGY_FIX_OR_MULTISCATTER) Pms = GlassMultiscatter(s, roughness, dotNV, 1.0f/eta, color);
  }

  const float cosOut = dot(rdir, nrm);
  const float cosMult = 1.0f/fmaxf(fabsf(cosOut), 1e-6f);
  out.direction = rdir;
  const float adjoint = reta*reta;                      // camera paths (a_isFwdDir == false)
  if (success) out.color = color*adjoint*Pss*Pms*cosMult;
  else         out.color = f3(1.0f, 1.0f, 1.0f)*Pss*Pms*cosMult;
  out.flags = spec ? (HC_RAY_EVENT_S | HC_RAY_EVENT_T) : (HC_RAY_EVENT_G | HC_RAY_EVENT_T);
  if (success && cosOut >= -1e-6f) out.color = f3(0.0f, 0.0f, 0.0f);
  else if (!success && cosOut < 1e-6f) out.color = f3(0.0f, 0.0f, 0.0f);
}

// ---- blend mask (fresnel helpers cglobals.h:1875-1926; blendMaskAlpha2 / blendSelectBRDF cmaterial.h:2034-2137)
HC_DEV float fresnelDielectric(float c1, float c2, float etaExt, float etaInt)
{
  const float Rs = (etaExt*c1 - etaInt*c2)/(etaExt*c1 + etaInt*c2);
  const float Rp = (etaInt*c1 - etaExt*c2)/(etaInt*c1 + etaExt*c2);
  return (Rs*Rs + Rp*Rp)/2.0f;
}
HC_DEV float fresnelReflectionCoeff(float cosTheta1, float etaExt, float etaInt)
{
  if (cosTheta1 < 0.0f) { const float t = etaInt; etaInt = etaExt; etaExt = t; }
  const float sinTheta2 = (float)(D(etaExt/etaInt)*sqrt(fmax(0.0, D(1.0f - cosTheta1*cosTheta1))));
  if (sinTheta2 > 1.0f) return 1.0f;
  const float cosTheta2 = sqrtf(fmaxf(0.0f, 1.0f - sinTheta2*sinTheta2));
  return fresnelDielectric(fabsf(cosTheta1), cosTheta2, etaInt, etaExt);
}
HC_DEV float BlendMaskAlpha(const float* m, float3 v, float3 n, float2 tc, const HcScene& s)
{
  const float3 tex = Sample2D(MatI(m, HC_BLEND_MASK_TEXMATRIXID_OFFSET), tc, m, s);
  const float3 lum1 = clamp3(tex*Mat3(m, HC_BLEND_MASK_COLORX_OFFSET), 0.0f, 1.0f);
  const int bflags = MatI(m, HC_BLEND_MASK_FLAGS_OFFSET);
  float lum;
  if (bflags & HC_BLEND_MASK_EXTRUSION_LUMINANCE) lum = dot(f3(0.2126f, 0.7152f, 0.0722f), lum1);
  else lum = fmaxf(lum1.x, fmaxf(lum1.y, lum1.z));
  const float normAngle = fabsf(dot(v, n));
  if (MatI(m, HC_BLEND_TYPE) == HC_BLEND_SIGMOID)
  {
    const float x2 = -5.0f + 10.0f*lum;                                                     // maxSigmoid, cmaterial.h:2026-2030
    lum = (float)(D(1.04f)/(D(1.0f) + hc_d_exp(D(-m[HC_BLEND_SIGMOID_EXP]*x2))) - D(0.02f));
  }
  if (bflags & HC_BLEND_MASK_FRESNEL)
    return clampf(lum*fresnelReflectionCoeff(fabsf(normAngle), 1.0f, m[HC_BLEND_MASK_FRESNEL_IOR]), 0.0f, 1.0f);
  return clampf(lum, 0.0f, 1.0f);
}
HC_DEV bool IsLeaf(const float* m) { return MatI(m, HC_PLAIN_MAT_TYPE_OFFSET) != HC_PLAIN_MAT_CLASS_BLEND_MASK; }

// ---- normal maps: sample2DAux (cfetch.h:766-789) from the "textures_aux" storage, materialNormalMapFetch / BumpMapping (cmaterial.h:2209-2243)
HC_DEV float3 NormalMapFetch(const float* m, float2 tc, const HcScene& s)
{
  float3 fromTex = f3(1, 1, 1);
  const int texId = MatI(m, HC_NORMAL_TEX_OFFSET), samplerOffset = MatI(m, HC_NORMAL_TEX_MATRIX);
  if (samplerOffset != HC_INVALID_TEXTURE)
  {
    const int4 header = reinterpret_cast<const int4*>(m)[samplerOffset];
    const float4 row0 = reinterpret_cast<const float4*>(m)[samplerOffset + 1], row1 = reinterpret_cast<const float4*>(m)[samplerOffset + 2];
    const int flags = header.x; const float gamma = __int_as_float(header.y);
    if (header.z != 0)
    {
      const float2 tct = f2(row0.x*tc.x + row0.y*tc.y + row0.w, row1.x*tc.x + row1.y*tc.y + row1.w);
      float4 c = ReadImageSw4(s.texturesAux + s.globals[s.texturesAuxTableOffset + texId], tct, flags, (gamma != 1.0f));
      if (flags & HC_TEX_ALPHASRC_W) { c.x = c.w; c.y = c.w; c.z = c.w; }
      fromTex = f3(c.x, c.y, c.z);
    }
  }
  float3 ts = f3(2.0f*fromTex.x - 1.0f, 2.0f*fromTex.y - 1.0f, fromTex.z);
  const int mflags = MatI(m, HC_PLAIN_MAT_FLAGS_OFFSET);
  if (mflags & HC_PLAIN_MATERIAL_INVERT_NMAP_Y) ts.y *= (-1.0f);
  if (mflags & HC_PLAIN_MATERIAL_INVERT_NMAP_X) ts.x *= (-1.0f);
  if (mflags & HC_PLAIN_MATERIAL_INVERT_SWAP_NMAP_XY) { const float t = ts.x; ts.x = ts.y; ts.y = t; }
  return normalize(ts);
}
HC_DEV float3 BumpMapping(float3 tangent, float3 bitangent, float3 normal, float2 tc, const float* m, const HcScene& s)
{
  const float3 ts = NormalMapFetch(m, tc, s);
  const float3 a0 = tangent, a1 = bitangent, a2 = normal;                                  // rows of the tangent transform; inverse(): cglobals.h:893-916
  const float det = a0.x*(a1.y*a2.z - a1.z*a2.y) - a0.y*(a1.x*a2.z - a1.z*a2.x) + a0.z*(a1.x*a2.y - a1.y*a2.x);
  float3 b0 = f3((a1.y*a2.z - a1.z*a2.y), -(a0.y*a2.z - a0.z*a2.y), (a0.y*a1.z - a0.z*a1.y));
  float3 b1 = f3(-(a1.x*a2.z - a1.z*a2.x), (a0.x*a2.z - a0.z*a2.x), -(a0.x*a1.z - a0.z*a1.x));
  float3 b2 = f3((a1.x*a2.y - a1.y*a2.x), -(a0.x*a2.y - a0.y*a2.x), (a0.x*a1.y - a0.y*a1.x));
  const float sc = 1.0f/det;
  b0 = b0*sc; b1 = b1*sc; b2 = b2*sc;
  return normalize(f3(b0.x*ts.x + b0.y*ts.y + b0.z*ts.z, b1.x*ts.x + b1.y*ts.y + b1.z*ts.z, b2.x*ts.x + b2.y*ts.y + b2.z*ts.z));
}
HC_DEV bool HasNormalMap(const float* m) { return MatI(m, HC_NORMAL_TEX_OFFSET) != HC_INVALID_TEXTURE; }

// ---- anisotropic Beckmann / Trowbridge-Reitz materials (PLAIN_MAT_CLASS_BECKMANN, _TRGGX; cmaterial.h:1529-1830).  Both share one slot map and
// one tangent-frame rule; the lobe itself (D, G, pdf, visible-normal sampling in the local frame) is hc_microfacet.cuh, KIND 0 / 1.
HC_DEV HcMf3 ToMf(float3 v) { return mf3(v.x, v.y, v.z); }
HC_DEV float AnisoGloss(const float* m, float2 tc, const HcScene& s)                        // beckmannGlosiness, cmaterial.h:1567-1580
{
  if (MatI(m, HC_BECKMANN_GLOSINESS_TEXID_OFFSET) != HC_INVALID_TEXTURE)
  {
    const float3 gc = Sample2D(MatI(m, HC_BECKMANN_GLOSINESS_TEXMATRIXID_OFFSET), tc, m, s);
    return clampf(m[HC_BECKMANN_GLOSINESS_OFFSET]*maxcomp(gc), 0.0f, 0.99f);
  }
  return m[HC_BECKMANN_GLOSINESS_OFFSET];
}
HC_DEV float2 AnisoAlphaXY(const float* m, float2 tc, const HcScene& s)                     // beckmannAnisotropy + beckmannAlphaXY, cmaterial.h:1582-1609
{
  const float3 ac = Sample2D(MatI(m, HC_BECKMANN_ANISO_TEXMATRIXID_OFFSET), tc, m, s);
  const float aniso = clampf(m[HC_BECKMANN_ANISOTROPY_OFFSET]*maxcomp(ac), 0.0f, 1.0f);
  const float roughness = 0.5f - 0.5f*AnisoGloss(m, tc, s);
  const float anisoMult = 1.0f - aniso;
  return f2(mfRoughnessToAlpha(roughness*roughness), mfRoughnessToAlpha(roughness*roughness*anisoMult*anisoMult));
}
// BeckmanTangentSpace (cmaterial.h:1611-1636): anisotropic lobes use the surface's (bitangent, tangent) turned about the normal by the
// rotation parameter (RotateAroundVector4x4, cglobals.h:1122-1150), isotropic ones any frame; FLIP_TANGENT swaps the two axes
HC_DEV void AnisoFrame(const float* m, float2 alpha, float3 nz, float3 tan, float3 bitan, float2 tc, const HcScene& s, float3& nx, float3& ny)
{
  if (fabsf(alpha.x - alpha.y) > 1e-5f)
  {
    const float3 rc = Sample2D(MatI(m, HC_BECKMANN_ROT_TEXMATRIXID_OFFSET), tc, m, s);
    const float rotVal = clampf(m[HC_BECKMANN_ANISO_ROT_OFFSET]*maxcomp(rc), 0.0f, 1.0f);
    const float ang = rotVal*HC_M_TWOPI;
    const float c = hc_cos(ang), sn = hc_sin(ang), omc = 1.0f - c;
    HcMat4 r;
    r.c0 = make_float4(omc*nz.x*nz.x + c,       omc*nz.y*nz.x + sn*nz.z, omc*nz.x*nz.z - sn*nz.y, 0.0f);
    r.c1 = make_float4(omc*nz.x*nz.y - sn*nz.z, omc*nz.y*nz.y + c,       omc*nz.z*nz.y + sn*nz.x, 0.0f);
    r.c2 = make_float4(omc*nz.x*nz.z + sn*nz.y, omc*nz.y*nz.z - sn*nz.x, omc*nz.z*nz.z + c,       0.0f);
    r.c3 = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    nx = mul3x3(r, bitan);
    ny = mul3x3(r, tan);
  }
  else
    CoordinateSystem(nz, nx, ny);
  if (MatI(m, HC_PLAIN_MAT_FLAGS_OFFSET) & HC_PLAIN_MATERIAL_FLIP_TANGENT) { const float3 t = nx; nx = ny; ny = t; }
}
HC_DEV float3 AnisoColor(const float* m, float2 tc, const HcScene& s)
{
  const float3 tex = Sample2D(MatI(m, HC_BECKMANN_TEXMATRIXID_OFFSET), tc, m, s);
  return clamp3(tex*Mat3(m, HC_BECKMANN_COLORX_OFFSET), 0.0f, 1.0f);
}
// beckmannEvalPDF / trggxEvalPDF (cmaterial.h:1638-1662, 1740-1765).  As in the reference, wo is the view direction's NEGATIVE in the local frame and
// the half vector normalize(l + v) is used in world coordinates.
template<int KIND>
HC_DEV float AnisoEvalPDF(const float* m, float3 l, float3 v, float3 n, float3 tan, float3 bitan, float2 tc, const HcScene& s)
{
  if (dot(n, v) < 1e-6f || dot(n, l) < 1e-6f) return 1.0f;
  const float2 alpha = AnisoAlphaXY(m, tc, s);
  float3 nx, ny;
  AnisoFrame(m, alpha, n, tan, bitan, tc, s, nx, ny);
  const HcMf3 wo = mf3(-dot(v, nx), -dot(v, ny), -dot(v, n));
  return mfPdf<KIND>(wo, ToMf(normalize(l + v)), alpha.x, alpha.y);
}
template<int KIND>
HC_DEV float3 AnisoEvalBxDF(const float* m, float3 l, float3 v, float3 n, float3 tan, float3 bitan, float2 tc, const HcScene& s)      // cmaterial.h:1664-1690, 1767-1792
{
  if (dot(n, v) < 1e-6f || dot(n, l) < 1e-6f) return f3(0.0f, 0.0f, 0.0f);
  const float3 color = AnisoColor(m, tc, s);
  const float2 alpha = AnisoAlphaXY(m, tc, s);
  float3 nx, ny;
  AnisoFrame(m, alpha, n, tan, bitan, tc, s, nx, ny);
  const HcMf3 wo = mf3(-dot(v, nx), -dot(v, ny), -dot(v, n)), wi = mf3(-dot(l, nx), -dot(l, ny), -dot(l, n));
  return color*mfBrdf<KIND>(wo, wi, alpha.x, alpha.y);
}
template<int KIND>
HC_DEV void AnisoSample(const float* m, float r1, float r2, float3 rayDir, float3 n, float2 tc, float3 tan, float3 bitan, const HcScene& s, HcMatSample& out)
{                                                                                                    // cmaterial.h:1692-1728, 1794-1830
  const float3 color = AnisoColor(m, tc, s);
  const float2 alpha = AnisoAlphaXY(m, tc, s);
  const float gloss = AnisoGloss(m, tc, s);
  float3 nx, ny;
  AnisoFrame(m, alpha, n, tan, bitan, tc, s, nx, ny);
  const HcMf3 wo = mf3(-dot(rayDir, nx), -dot(rayDir, ny), -dot(rayDir, n));
  const HcMf3 wh = mfSampleWh<KIND>(wo, r1, r2, alpha.x, alpha.y);
  const float k = 2.0f*mfDot(wo, wh);
  const HcMf3 wi = mf3(k*wh.x - wo.x, k*wh.y - wo.y, k*wh.z - wo.z);                                 // wo mirrored about wh
  const float3 newDir = normalize(wi.x*nx + wi.y*ny + wi.z*n);
  const float3 v = rayDir*(-1.0f);
  if (dot(n, v) < 1e-6f || dot(n, newDir) < 1e-6f) { out.color = f3(0.0f, 0.0f, 0.0f); out.pdf = 1.0f; }
  else { out.color = color*mfBrdf<KIND>(wo, wi, alpha.x, alpha.y); out.pdf = mfPdf<KIND>(wo, wh, alpha.x, alpha.y); }
  out.direction = newDir;
  out.flags = (gloss >= 0.99f) ? HC_RAY_EVENT_S : HC_RAY_EVENT_G;
}

// leaf dispatch: MaterialLeafSampleAndEvalBRDF (cmaterial.h:2245-2335)
// NMAP: the "extended" instantiation - the scene has at least one normal-mapped or anisotropic (Beckmann / TRGGX) material (found by
// ValidateScene); scenes without any run the instantiation without this code, which keeps the register-limited shade kernel as it was
template<bool NMAP>
HC_DEV void LeafSample(const float* m, const HcSurfaceHit& shIn, float3 rayDir, float3 rands, const HcScene& s, HcMatSample& out)
{
  out.color = f3(0.0f, 0.0f, 0.0f); out.direction = f3(0.0f, 1.0f, 0.0f); out.pdf = 1.0f; out.flags = 0;
  struct { float3 normal; float2 texCoord; bool hfi; } sh = { shIn.normal, shIn.texCoord, shIn.hfi };
  const bool hasNormalMap = NMAP && HasNormalMap(m);
  if (hasNormalMap)
  {
    const bool isGlass = (MatI(m, HC_PLAIN_MAT_TYPE_OFFSET) == HC_PLAIN_MAT_CLASS_GLASS);
    const float3 flatNorm = (shIn.hfi && !isGlass) ? (-1.0f)*shIn.flatNormal : shIn.flatNormal;
    sh.normal = BumpMapping(shIn.tangent, shIn.biTangent, flatNorm, shIn.texCoord, m, s);
  }
  switch (MatI(m, HC_PLAIN_MAT_TYPE_OFFSET))
  {
    case HC_PLAIN_MAT_CLASS_PHONG_SPECULAR: PhongSample(m, rands.x, rands.y, rayDir, sh.normal, sh.texCoord, s, out); break;
    case HC_PLAIN_MAT_CLASS_BLINN_SPECULAR: BlinnSample(m, rands.x, rands.y, rayDir, sh.normal, sh.texCoord, s, out); break;
    case HC_PLAIN_MAT_CLASS_GGX:            GgxSample2(m, rands.x, rands.y, rayDir, sh.normal, sh.texCoord, s, out); break;
    case HC_PLAIN_MAT_CLASS_PERFECT_MIRROR: MirrorSample(m, rayDir, sh.normal, sh.texCoord, s, out); break;
    case HC_PLAIN_MAT_CLASS_GLASS:          GlassGgxSample(m, rands, rayDir, sh.normal, sh.texCoord, sh.hfi, s, out); break;
    case HC_PLAIN_MAT_CLASS_LAMBERT:        LambertSample(m, rands.x, rands.y, sh.normal, sh.texCoord, s, out); break;
    case HC_PLAIN_MAT_CLASS_THIN_GLASS:     ThinglassSample(m, rands.x, rands.y, rayDir, sh.normal, sh.texCoord, s, out); break;
    case HC_PLAIN_MAT_CLASS_TRANSLUCENT:    TranslucentSample(m, rands.x, rands.y, sh.normal, sh.texCoord, s, out); break;
    case HC_PLAIN_MAT_CLASS_OREN_NAYAR:     OrennayarSample(m, rands.x, rands.y, rayDir, sh.normal, sh.texCoord, s, out); break;
    case HC_PLAIN_MAT_CLASS_BECKMANN:       if (NMAP) AnisoSample<0>(m, rands.x, rands.y, rayDir, sh.normal, sh.texCoord, shIn.tangent, shIn.biTangent, s, out); break;
    case HC_PLAIN_MAT_CLASS_TRGGX:          if (NMAP) AnisoSample<1>(m, rands.x, rands.y, rayDir, sh.normal, sh.texCoord, shIn.tangent, shIn.biTangent, s, out); break;
    default: break;
  }
  if (hasNormalMap)                          // the caller multiplies by the cosine to the UNBUMPED normal (cmaterial.h:2320-2330)
  {
    const float cosThetaOut1 = fabsf(dot(out.direction, shIn.normal)), cosThetaOut2 = fabsf(dot(out.direction, sh.normal));
    out.color *= (cosThetaOut2/fmaxf(cosThetaOut1, HC_DEPSILON2));
  }
  if (out.pdf <= 0.0f) out.color = f3(0, 0, 0);
}

// MaterialSampleAndEvalBxDF (cmaterial.h:2345-2371) with materialRandomWalkBRDF (:2180-2207); rands[0..2] direction, rands[3..9] layers
template<bool NMAP>
HC_DEV void MaterialSampleAndEval(const float* mat, const float* rands, const HcSurfaceHit& sh, float3 rayDir, unsigned rayFlags,
                                  const HcScene& s, HcMatSample& out)
{
  const unsigned other = (rayFlags & 0xFFFF0000u) >> 16;
  const bool canReflOnly = (MatI(mat, HC_PLAIN_MAT_FLAGS_OFFSET) & HC_PLAIN_MATERIAL_CAN_SAMPLE_REFL_ONLY) != 0;
  const bool reflOnly = ((other & HC_RAY_GRAMMAR_DIRECT_LIGHT) != 0) && canReflOnly;
  int localOffs = 0; float w = 1.0f;
  int selOffs = 0; float selW = 1.0f;
  const float* node = mat;
  int i = 0;
  while (!IsLeaf(node) && i < HC_MMLT_FLOATS_PER_MLAYER)
  {
    const float r3 = rands[HC_MMLT_FLOATS_PER_SAMPLE + i];
    {                                                                                    // blendSelectBRDF
      float alpha = BlendMaskAlpha(node, rayDir, sh.normal, sh.texCoord, s);
      const int o1 = MatI(node, HC_BLEND_MASK_MATERIAL1_OFFSET), o2 = MatI(node, HC_BLEND_MASK_MATERIAL2_OFFSET);
      float w1 = 1.0f; const float w2 = 1.0f;
      const int bflags = MatI(node, HC_BLEND_MASK_FLAGS_OFFSET);
      if ((bflags & HC_BLEND_MASK_REFLECTION_WEIGHT_IS_ONE) && IsLeaf(node + o1*HC_PLAIN_MATERIAL_DATA_SIZE)) w1 = alpha;
      if ((bflags & HC_BLEND_MASK_FRESNEL) != 0 && (reflOnly && i == 0)) { w1 = alpha; alpha = 1.0f; }
      if (r3 <= alpha) { selOffs = o1; selW = w1; } else { selOffs = o2; selW = w2; }
    }
    w = w*selW;
    localOffs += selOffs;
    node += selOffs*HC_PLAIN_MATERIAL_DATA_SIZE;
    i++;
  }
  const float* leaf = mat + localOffs*HC_PLAIN_MATERIAL_DATA_SIZE;
  LeafSample<NMAP>(leaf, sh, rayDir, f3(rands[0], rands[1], rands[2]), s, out);
  out.color *= 1.0f/fmaxf(w, 0.015625f);
  if ((MatI(leaf, HC_PLAIN_MAT_FLAGS_OFFSET) & HC_PLAIN_MATERIAL_SKIP_SKY_PORTAL))       // materialIsSkyPortal && isEyeRay, cmaterial.h:2366-2370
  {
    const bool nonSpec = (other & HC_RAY_EVENT_D) || (other & HC_RAY_EVENT_G);
    if ((((rayFlags & 0x0000FF00u) >> 8) == 0) || !nonSpec) { out.color = f3(1, 1, 1); out.pdf = 1.0f; }
  }
}

// materialLeafEval (cmaterial.h:2425-2551), camera direction (EVAL_FLAG_DEFAULT: no adjoint-BSDF fix, clamp 1e-6 in the normal-map cosine fix)
template<bool NMAP>
HC_DEV HcBxDF LeafEval(const float* m, float3 l, float3 v, const HcSurfaceHit& sh, const HcScene& s)
{
  HcBxDF r; r.brdf = f3(0, 0, 0); r.btdf = f3(0, 0, 0); r.pdfFwd = 0.0f; r.pdfRev = 0.0f; r.diffuse = false;
  float3 n = sh.normal; const float2 tc = sh.texCoord;
  float cosMult = 1.0f, cosMult2 = 1.0f;
  if (NMAP && HasNormalMap(m))
  {
    n = BumpMapping(sh.tangent, sh.biTangent, sh.flatNormal, tc, m, s);
    const float clampVal = 1e-6f;
    const float cosThetaOut1 = fmaxf(dot(l, sh.normal), 0.0f), cosThetaOut2 = fmaxf(dot(l, n), 0.0f);
    cosMult = cosThetaOut2/fmaxf(cosThetaOut1, clampVal);
    if (cosThetaOut1 <= 0.0f) cosMult = 0.0f;
    const float cosThetaOut3 = fmaxf(-dot(l, sh.normal), 0.0f), cosThetaOut4 = fmaxf(-dot(l, n), 0.0f);
    cosMult2 = cosThetaOut4/fmaxf(cosThetaOut3, clampVal);
    if (cosThetaOut3 <= 0.0f) cosMult2 = 0.0f;
  }
  switch (MatI(m, HC_PLAIN_MAT_TYPE_OFFSET))
  {
    case HC_PLAIN_MAT_CLASS_PHONG_SPECULAR:
      r.brdf = PhongEvalBxDF(m, l, v, n, tc, s)*cosMult; r.pdfFwd = PhongEvalPDF(m, l, v, n, tc, s); r.pdfRev = PhongEvalPDF(m, v, l, n, tc, s); break;
    case HC_PLAIN_MAT_CLASS_BLINN_SPECULAR:
      r.brdf = BlinnEvalBxDF(m, l, v, n, tc, s)*cosMult; r.pdfFwd = BlinnEvalPDF(m, l, v, n, tc, s); r.pdfRev = BlinnEvalPDF(m, v, l, n, tc, s); break;
    case HC_PLAIN_MAT_CLASS_GGX:
      r.brdf = GgxEvalBxDF(m, l, v, n, tc, s)*cosMult; r.pdfFwd = Ggx2EvalPDF(m, l, v, n, tc, s); r.pdfRev = Ggx2EvalPDF(m, v, l, n, tc, s); break;
    case HC_PLAIN_MAT_CLASS_LAMBERT:
      r.brdf = LambertColor(m, tc, s)*HC_INV_PI*cosMult; r.pdfFwd = fabsf(dot(l, n))*HC_INV_PI; r.pdfRev = fabsf(dot(v, n))*HC_INV_PI; r.diffuse = true; break;
    case HC_PLAIN_MAT_CLASS_OREN_NAYAR:
      r.brdf = OrennayarEvalBxDF(m, l, v, n, tc, s)*cosMult; r.pdfFwd = fabsf(dot(l, n))*HC_INV_PI; r.pdfRev = fabsf(dot(v, n))*HC_INV_PI; r.diffuse = true; break;
    case HC_PLAIN_MAT_CLASS_TRANSLUCENT:
      r.btdf = TranslucentEvalBxDF(m, l, v, n, tc, s)*cosMult2; r.pdfFwd = TranslucentEvalPDF(l, v, n); r.pdfRev = TranslucentEvalPDF(v, l, n); r.diffuse = true; break;
    case HC_PLAIN_MAT_CLASS_BECKMANN:
      if (NMAP) { r.brdf = AnisoEvalBxDF<0>(m, l, v, n, sh.tangent, sh.biTangent, tc, s)*cosMult; r.pdfFwd = AnisoEvalPDF<0>(m, l, v, n, sh.tangent, sh.biTangent, tc, s);
                  r.pdfRev = AnisoEvalPDF<0>(m, v, l, n, sh.tangent, sh.biTangent, tc, s); }
      break;
    case HC_PLAIN_MAT_CLASS_TRGGX:
      if (NMAP) { r.brdf = AnisoEvalBxDF<1>(m, l, v, n, sh.tangent, sh.biTangent, tc, s)*cosMult; r.pdfFwd = AnisoEvalPDF<1>(m, l, v, n, sh.tangent, sh.biTangent, tc, s);
                  r.pdfRev = AnisoEvalPDF<1>(m, v, l, n, sh.tangent, sh.biTangent, tc, s); }
      break;
    default: break;      // mirror, glass and thin glass evaluate to zero for explicit light (cmaterial.h:396-404, 620-629)
  }
  return r;
}

// materialEval (cmaterial.h:2554-2628): explicit-stack walk of the blend tree, same push/pop order (so sums round identically)
template<bool NMAP>
HC_DEV HcBxDF MaterialEval(const float* mat, float3 l, float3 v, const HcSurfaceHit& sh, const HcScene& s)
{
  const float3 n = sh.normal; const float2 tc = sh.texCoord;
  HcBxDF val; val.brdf = f3(0, 0, 0); val.btdf = f3(0, 0, 0); val.pdfFwd = 0.0f; val.pdfRev = 0.0f; val.diffuse = true;
  float stackW[HC_MIX_TREE_MAX_DEEP]; int stackO[HC_MIX_TREE_MAX_DEEP];
  int top = 0, cur = 0; float curW = 1.0f;
  do
  {
    if (top > 0) { top--; cur = stackO[top]; curW = stackW[top]; }
    const float* m = mat + cur*HC_PLAIN_MATERIAL_DATA_SIZE;
    if (!IsLeaf(m))
    {
      const float alpha = BlendMaskAlpha(m, v, n, tc, s);
      const int o1 = MatI(m, HC_BLEND_MASK_MATERIAL1_OFFSET), o2 = MatI(m, HC_BLEND_MASK_MATERIAL2_OFFSET);
      float w1 = alpha; const float w2 = 1.0f - alpha;
      if ((MatI(m, HC_BLEND_MASK_FLAGS_OFFSET) & HC_BLEND_MASK_REFLECTION_WEIGHT_IS_ONE) && IsLeaf(m + o1*HC_PLAIN_MATERIAL_DATA_SIZE)) w1 = 1.0f;
      if (top < HC_MIX_TREE_MAX_DEEP) { stackW[top] = curW*w1; stackO[top] = cur + o1; top++; }
      if (top < HC_MIX_TREE_MAX_DEEP) { stackW[top] = curW*w2; stackO[top] = cur + o2; top++; }
    }
    else
    {
      const HcBxDF b = LeafEval<NMAP>(m, l, v, sh, s);
      val.brdf += curW*b.brdf; val.btdf += curW*b.btdf; val.pdfFwd += curW*b.pdfFwd; val.pdfRev += curW*b.pdfRev;
      val.diffuse = val.diffuse && b.diffuse;
    }
  } while (top > 0);
  return val;
}

// materialEvalEmission (cmaterial.h:2918-2978): blend nodes carry emissive headers too, so every visited node contributes
HC_DEV float3 MaterialEvalEmission(const float* mat, float3 v, float3 n, float2 tc, const HcScene& s)
{
  float3 val = f3(0.0f, 0.0f, 0.0f);
  float stackW[HC_MIX_TREE_MAX_DEEP]; int stackO[HC_MIX_TREE_MAX_DEEP];
  int top = 0, cur = 0; float curW = 1.0f;
  do
  {
    if (top > 0) { top--; cur = stackO[top]; curW = stackW[top]; }
    const float* m = mat + cur*HC_PLAIN_MATERIAL_DATA_SIZE;
    if (!IsLeaf(m))
    {
      const float alpha = BlendMaskAlpha(m, v, n, tc, s);
      const int o1 = MatI(m, HC_BLEND_MASK_MATERIAL1_OFFSET), o2 = MatI(m, HC_BLEND_MASK_MATERIAL2_OFFSET);
      if (top < HC_MIX_TREE_MAX_DEEP) { stackW[top] = curW*alpha; stackO[top] = cur + o1; top++; }
      if (top < HC_MIX_TREE_MAX_DEEP) { stackW[top] = curW*(1.0f - alpha); stackO[top] = cur + o2; top++; }
    }
    const float3 tex = Sample2D(MatI(m, HC_EMISSIVE_TEXMATRIXID_OFFSET), tc, m, s);      // materialLeafEvalEmission, cmaterial.h:21-26
    val += curW*(Mat3(m, HC_EMISSIVE_COLORX_OFFSET)*tex);
  } while (top > 0);
  return val;
}

// ------------------------------------------------------------------------------------------------------------------ a11/a12: lights
// SelectRandomLightRev + SelectIndexPropToOpt (clight.h:1774-1793, cglobals.h:2806-2862)
HC_DEV int SelectIndexPropToOpt(float r, const float* __restrict__ acc, int N, float& pdf)          // cglobals.h:2806-2862
{
  int left = 0, right = N - 2, counter = 0, cur = -1;
  const float x = r*acc[N - 1];
  while (right - left > 1 && counter < 50)
  {
    const int size = right + left;
    const int p1 = (size % 2 == 0) ? (size + 1)/2 : (size + 0)/2;
    const float a = acc[p1], b = acc[p1 + 1];
    if (a < x && x <= b) { cur = p1; break; }
    else if (x <= a) right = p1;
    else if (x > b) left = p1;
    counter++;
  }
  if (cur < 0)
  {
    const float a1 = acc[left], b1 = acc[left + 1], a2 = acc[right], b2 = acc[right + 1];
    if (a1 < x && x <= b1) cur = left;
    if (a2 < x && x <= b2) cur = right;
  }
  if (x == 0.0f) cur = 0;
  else if (cur < 0) cur = (right + left + 1)/2;
  pdf = (acc[cur + 1] - acc[cur])/acc[N - 1];
  return cur;
}

HC_DEV int SelectRandomLightRev(float r, const HcScene& s, float& pickProb)
{
  const int N = s.lightSelTableSizeRev;
  if (N == 0) { pickProb = 1.0f; return -1; }
  if (N <= 2) { pickProb = 1.0f; return 0; }
  return SelectIndexPropToOpt(r, reinterpret_cast<const float*>(s.globals + s.lightSelTableOffsetRev), N, pickProb);
}

HC_DEV float AreaLightEvalPDF(const float* L, float3 rayDir, float hitDist)             // areaDiffuseLightEvalPDF, clight.h:524-530 ; PdfAtoW cglobals.h:1754
{
  const float3 ln = Mat3(L, HC_PLIGHT_NORM_X);
  const float pdfA = 1.0f/fmaxf(L[HC_PLIGHT_SURFACE_AREA], HC_DEPSILON);
  const float cosVal = (__float_as_int(L[HC_PLIGHT_FLAGS]) & HC_LIGHT_HAS_IES) ? fabsf(dot(rayDir, -1.0f*ln)) : fmaxf(dot(rayDir, -1.0f*ln), 0.0f);   // an IES light shines to both sides
  return (pdfA*hitDist*hitDist)/fmaxf(cosVal, HC_DEPSILON2);
}

// ---- IES distributions (LIGHT_HAS_IES): the photometric web as a single-channel float image over the sphere, kept in the "pdfs" storage
// (AddIesTexTableToStorage, RenderDriverRTE_PdfTables.cpp:385-475); lightDistributionMask (clight.h:465-483) looks it up along the light's own axes
HC_DEV float2 SphereMapTo2DTexCoord(float3 rayDir, float& sinTheta);
HC_DEV float3 Mat3x3MulVec(const float* M, float3 v);
HC_DEV float3 LightDistributionMask(const float* L, float3 rayDir, const HcScene& s)
{
  if (!(__float_as_int(L[HC_PLIGHT_FLAGS]) & HC_LIGHT_HAS_IES)) return f3(1.0f, 1.0f, 1.0f);
  rayDir = normalize(Mat3x3MulVec(L + HC_IES_LIGHT_MATRIX_E00, rayDir));
  float sintheta = 0.0f;
  const float2 tc = SphereMapTo2DTexCoord((-1.0f)*rayDir, sintheta);
  const int offset = s.globals[s.pdfTableTableOffset + __float_as_int(L[HC_IES_SPHERE_TEX_ID])];
  const float val = ReadImageSw1(reinterpret_cast<const int4*>(s.pdfs + offset), tc, HC_TEX_CLAMP_U | HC_TEX_CLAMP_V);
  return f3(val, val, val);
}

// areaSpotLightAttenuation (clight.h:7-12, 532-539): smoothstep between the two cone cosines about the light's normal
HC_DEV float AreaSpotAttenuation(const float* L, float3 shadowRayDir)
{
  const float cos1 = L[HC_AREA_LIGHT_SPOT_COS1], cos2 = L[HC_AREA_LIGHT_SPOT_COS2];
  const float cosTheta = fmaxf(dot(shadowRayDir, Mat3(L, HC_PLIGHT_NORM_X)), 0.0f);
  const float tVal = (cosTheta - cos2)/(cos1 - cos2);
  const float t = fminf(fmaxf(tVal, 0.0f), 1.0f);
  return t*t*(3.0f - 2.0f*t);
}
// areaDiffuseLightGetIntensity (clight.h:542-611) for untextured lights that are no sky portals: base colour, cut by the spot cone for light
// samples and GI hits; an EYE ray that hits a spot-distributed light sees it white (colour / its largest component), as the reference draws it
HC_DEV float3 AreaLightIntensity(const float* L, float3 rayDir, bool eyeRay, const HcScene& s)
{
  float3 color = Mat3(L, HC_PLIGHT_COLOR_X);
  if (__float_as_int(L[HC_PLIGHT_FLAGS]) & HC_LIGHT_HAS_IES)                              // takes precedence over the spot cone (clight.h:560-574)
  {
    if (!eyeRay) color *= LightDistributionMask(L, rayDir, s);
    else color *= (1.0f/fmaxf(color.x, fmaxf(color.y, color.z)));
  }
  else if (__float_as_int(L[HC_AREA_LIGHT_SPOT_DISTR]) != 0)
  {
    if (!eyeRay) color *= clampf(AreaSpotAttenuation(L, (-1.0f)*rayDir), 0.0f, 1.0f);
    else color *= (1.0f/fmaxf(color.x, fmaxf(color.y, color.z)));
  }
  return color;
}
// AreaLightSampleRev (clight.h:1180-1229), untextured, no sky portal (rejected at init)
HC_DEV void AreaLightSampleRev(const float* L, float3 rands, float3 illum, const HcScene& s, HcShadowSample& out)
{
  const float ox = rands.x*2.0f - 1.0f, oy = rands.y*2.0f - 1.0f;
  float3 sp = f3(ox*L[HC_AREA_LIGHT_SIZE_X], 0.0f, oy*L[HC_AREA_LIGHT_SIZE_Y]);
  if (__float_as_int(L[HC_AREA_LIGHT_IS_DISK]) != 0)
  {
    // MapSamplesToDisc (cglobals.h:1609-1655)
    const float x = ox, y = oy; float r = 0.0f, phi = 0.0f;
    if (x > y && x > -y)  { r = x;  phi = 0.25f*3.141592654f*(y/x); }
    if (x < y && x > -y)  { r = y;  phi = 0.25f*3.141592654f*(2.0f - x/y); }
    if (x < y && x < -y)  { r = -x; phi = 0.25f*3.141592654f*(4.0f + y/x); }
    if (x > y && x < -y)  { r = -y; phi = 0.25f*3.141592654f*(6 - x/y); }
    const float2 xz = f2(r*hc_sin(phi), r*hc_cos(phi))*L[HC_AREA_LIGHT_SIZE_X];
    sp = f3(xz.x, 0.0f, xz.y);
  }
  const float* M = L + HC_AREA_LIGHT_MATRIX_E00;                                         // matrix3x3f_mult_float3, cglobals.h:1091-1098
  sp = f3(M[0]*sp.x + M[1]*sp.y + M[2]*sp.z, M[3]*sp.x + M[4]*sp.y + M[5]*sp.z, M[6]*sp.x + M[7]*sp.y + M[8]*sp.z);
  sp = sp + Mat3(L, HC_PLIGHT_POS_X);
  const float3 rayDir = normalize(sp - illum);
  const float hitDist = length(sp - illum);
  const float3 ln = Mat3(L, HC_PLIGHT_NORM_X);
  out.isPoint = false;
  out.pos = sp + epsilonOfPos(sp)*ln;
  float3 customRayDir = rayDir;                                                          // LIGHT_IES_POINT_AREA: the web is looked up from the light's centre
  if (__float_as_int(L[HC_PLIGHT_FLAGS]) & HC_LIGHT_IES_POINT_AREA) customRayDir = normalize(Mat3(L, HC_PLIGHT_POS_X) - illum);
  out.color = AreaLightIntensity(L, customRayDir, false, s);
  out.pdf = AreaLightEvalPDF(L, rayDir, hitDist);
  out.maxDist = hitDist;
  out.cosAtLight = -dot(rayDir, ln);
}

// ---- sphere and omni point lights (clight.h:1287-1333, 1387-1407).  The reference's host build takes M_PI from <cmath> (a double) and
// resolves sin / cos / acos on float arguments to the double functions, so those expressions are evaluated in double and rounded once.
#define HC_SPHERE_LIGHT_RADIUS 14
HC_DEV float PdfAtoW(float pdfA, float dist, float cosThere) { return (pdfA*dist*dist)/fmaxf(cosThere, HC_DEPSILON2); }   // cglobals.h:1754-1757

HC_DEV float SphereLightEvalPDF(const float* L, float3 illum, float3 lpos, float3 lnorm)                                   // clight.h:1287-1302
{
  const float lradius = L[HC_SPHERE_LIGHT_RADIUS];
  const float3 lcenter = Mat3(L, HC_PLIGHT_POS_X);
  const float3 diff = lcenter - illum;                                                  // DistanceSquared(a, b) = dot(b - a, b - a), cglobals.h:1152
  if (dot(diff, diff) - lradius*lradius <= 0.0f) return 1.0f;
  const float pdfA = 1.0f/L[HC_PLIGHT_SURFACE_AREA];
  const float dist = length(lpos - illum);
  const float3 dirToV = normalize(lpos - illum);
  const float cosAtLight = fabsf(dot(dirToV, lnorm));
  return PdfAtoW(pdfA, dist, cosAtLight);
}

HC_DEV void SphereLightSampleRev(const float* L, float3 rands, float3 illum, HcShadowSample& out)                          // clight.h:1309-1333
{
  const float theta = (float)(2.0*3.14159265358979323846*(double)rands.x);
  const float phi   = (float)hc_d_acos((double)(1.0f - 2.0f*rands.y));
  const float x = (float)(hc_d_sin((double)phi)*hc_d_cos((double)theta));
  const float y = (float)(hc_d_sin((double)phi)*hc_d_sin((double)theta));
  const float z = (float)hc_d_cos((double)phi);
  const float3 lcenter = Mat3(L, HC_PLIGHT_POS_X);
  const float lradius = L[HC_SPHERE_LIGHT_RADIUS];
  const float3 samplePos = lcenter + lradius*f3(x, y, z);
  const float3 lightNorm = normalize(samplePos - lcenter);
  const float3 dirToV = normalize(samplePos - illum);
  out.isPoint = false;
  out.pos = samplePos;
  out.color = Mat3(L, HC_PLIGHT_COLOR_X);                                                // sphereLightGetIntensity = lightBaseColor
  out.pdf = SphereLightEvalPDF(L, illum, samplePos, lightNorm);
  out.maxDist = length(samplePos - illum);
  out.cosAtLight = fabsf(dot(lightNorm, dirToV));
}

HC_DEV void PointLightSampleRev(const float* L, float3 illum, const HcScene& s, HcShadowSample& out)                      // clight.h:1394-1407
{
  const float3 samplePos = Mat3(L, HC_PLIGHT_POS_X);
  const float hitDist = length(samplePos - illum);
  out.isPoint = true;
  out.pos = samplePos;
  out.color = LightDistributionMask(L, normalize(samplePos - illum), s)*Mat3(L, HC_PLIGHT_COLOR_X);      // pointLightGetIntensity: (1,1,1) without LIGHT_HAS_IES
  out.pdf = PdfAtoW(1.0f, hitDist, 1.0f);
  out.maxDist = hitDist;
  out.cosAtLight = 1.0f;
}

// ---- sky dome (clight.h:286-463, cfetch.h:258-296, cbidir.h:492-533): constant, textured or Perez (hc_perez.cuh); without a secondary (AUX) sky.
// M_PI is the <cmath> double in the reference's host build; sin / cos / acos / atan2 resolve to the double functions (hc_math.cuh).
#define HC_SKY_DOME_PDF_TABLE0  30
#define HC_SKY_DOME_SAMPLER0    32
#define HC_SKY_DOME_MATRIX0     36
#define HC_SKY_DOME_INV_MATRIX0 56
HC_DEV float2 SphereMapTo2DTexCoord(float3 rayDir, float& sinTheta)                                                          // cfetch.h:258-281
{
  const float x = rayDir.z, y = rayDir.x, z = -rayDir.y;
  const float theta = hc_acos(z);
  float phi = hc_atan2(y, x);
  if (phi < 0.0f) phi = (float)((double)phi + 2.0*HC_M_PI_D);                                // phi += 2.0f*M_PI
  const float texX = clampf(phi*0.5f*HC_INV_PI, 0.0f, 1.0f);
  const float texY = clampf(theta*HC_INV_PI, 0.0f, 1.0f);
  sinTheta = sqrtf(1.0f - rayDir.y*rayDir.y);
  return f2(texX, texY);
}
HC_DEV float3 TexCoord2DToSphereMap(float2 tc, float& sinThetaOut)                                                            // cfetch.h:283-296
{
  const float phi   = (float)((double)(tc.x*2.0f)*HC_M_PI_D);
  const float theta = (float)((double)tc.y*HC_M_PI_D);
  const float sinTheta = hc_sin(theta);
  const float x = (float)((double)sinTheta*hc_d_cos((double)phi));
  const float y = (float)((double)sinTheta*hc_d_sin((double)phi));
  const float z = hc_cos(theta);
  sinThetaOut = sinTheta;
  return f3(y, -z, x);
}
HC_DEV const float* PdfTableHeader(int tableId, const HcScene& s)                                                              // cfetch.h:153-163
{
  const int offset = s.globals[s.pdfTableTableOffset + tableId];
  return reinterpret_cast<const float*>(s.pdfs + offset);
}
HC_DEV float3 SkyLightIntensityTexturedEnv(const float* L, float3 dir, const HcScene& s)                                       // clight.h:292-307
{
  float sintheta = 0.0f;
  const float2 tc = SphereMapTo2DTexCoord(dir, sintheta);
  const float3 texColor = Sample2D(__float_as_int(L[HC_PLIGHT_COLOR_TEX_MATRIX]), tc, L + HC_SKY_DOME_SAMPLER0, s);
  if (__float_as_int(L[HC_PLIGHT_FLAGS]) & HC_SKY_LIGHT_USE_PEREZ_ENVIRONMENT)              // analytic sky + sun disk instead of the map (skyLightPerezColor, hc_perez.cuh)
  {
    const HcMf3 c = mfPerezSkyColor(ToMf(Mat3(L, HC_SKY_DOME_SUN_DIR_X)), L[HC_SKY_DOME_TURBIDITY], ToMf(Mat3(L, HC_SKY_SUN_COLOR_X)), ToMf(dir));
    return Mat3(L, HC_PLIGHT_COLOR_X)*f3(c.x, c.y, c.z);
  }
  return Mat3(L, HC_PLIGHT_COLOR_X)*texColor;
}
HC_DEV float EvalMap2DPdf(float2 t, const float* __restrict__ intervals, int sizeX, int sizeY)                                 // clight.h:309-337 (quirks included)
{
  const float fw = (float)sizeX, fh = (float)sizeY;
  if (t.x < 0.0f || t.x > 1.0f) t.x -= (float)((int)(t.x));
  if (t.y < 0.0f || t.x > 1.0f) t.y -= (float)((int)(t.y));
  int pixelX = (int)(fw*t.x - 0.5f), pixelY = (int)(fh*t.y - 0.5f);
  if (pixelX >= sizeX) pixelX = sizeX - 1;
  if (pixelY >= sizeY) pixelY = sizeY - 1;
  if (pixelX < 0) pixelX += sizeX;
  if (pixelY < 0) pixelY += sizeY;
  const int pixelOffset = pixelY*sizeX + pixelX, maxSize = sizeX*sizeY;
  const int offset0 = (pixelOffset + 0 < maxSize + 0) ? pixelOffset + 0 : maxSize - 1;
  const int offset1 = (pixelOffset + 1 < maxSize + 1) ? pixelOffset + 1 : maxSize;
  return (intervals[offset1] - intervals[offset0])*(fw*fh)/intervals[sizeX*sizeY];
}
HC_DEV float SkyLightEvalPDF(const float* L, float3 rayDir, const HcScene& s)                                                   // clight.h:339-363
{
  const float* hdr = PdfTableHeader(__float_as_int(L[HC_SKY_DOME_PDF_TABLE0]), s);
  const float* intervals = hdr + 4;
  const int sizeX = __float_as_int(hdr[0]), sizeY = __float_as_int(hdr[1]);
  float sintheta = 0.0f;
  const float2 tc = SphereMapTo2DTexCoord(rayDir, sintheta);
  if (sintheta == 0.0f) return 0.0f;
  const float4 row0 = *reinterpret_cast<const float4*>(L + HC_SKY_DOME_MATRIX0), row1 = *reinterpret_cast<const float4*>(L + HC_SKY_DOME_MATRIX0 + 4);
  const float2 tct = f2(row0.x*tc.x + row0.y*tc.y + row0.w, row1.x*tc.x + row1.y*tc.y + row1.w);
  const float mapPdf = EvalMap2DPdf(tct, intervals, sizeX, sizeY);
  return (float)((double)(mapPdf*1.0f)/((double)2.0f*HC_M_PI_D*HC_M_PI_D*(double)fmaxf(fabsf(sintheta), HC_DEPSILON)));
}
HC_DEV void SkyLightSampleRev(const float* L, float3 rands, float3 illum, const HcScene& s, HcShadowSample& out)                // clight.h:427-463
{
  const float* hdr = PdfTableHeader(__float_as_int(L[HC_SKY_DOME_PDF_TABLE0]), s);
  const float* intervals = hdr + 4;
  const int sizeX = __float_as_int(hdr[0]), sizeY = __float_as_int(hdr[1]);
  const float fw = (float)sizeX, fh = (float)sizeY, fN = fw*fh;
  float pdf = 1.0f;                                                                          // sampleMap2D, clight.h:375-396
  int pixelOffset = SelectIndexPropToOpt(rands.z, intervals, sizeX*sizeY + 1, pdf);
  if (pixelOffset >= sizeX*sizeY) pixelOffset = sizeX*sizeY - 1;
  const int yPos = pixelOffset/sizeX, xPos = pixelOffset - yPos*sizeX;
  const float texX = (1.0f/fw)*(((float)(xPos) + 0.5f) + (rands.x*2.0f - 1.0f)*0.5f);
  const float texY = (1.0f/fh)*(((float)(yPos) + 0.5f) + (rands.y*2.0f - 1.0f)*0.5f);
  const float mapPdf = pdf*fN;
  const HcMat4 m = loadMat4(L + HC_SKY_DOME_INV_MATRIX0);
  const float3 tct = mul4x3(m, f3(texX, texY, 0.0f));                                        // mul(float4x4, float3) = mul4x3
  float sintheta = 0.0f;
  const float3 sampleDir = TexCoord2DToSphereMap(f2(tct.x, tct.y), sintheta);
  const float radius = reinterpret_cast<const float*>(s.globals)[HC_EG_varsF/4 + 21];        // varsF[HRT_BSPHERE_RADIUS]
  const float3 samplePos = illum + sampleDir*radius;
  const float3 txClr = Sample2D(__float_as_int(L[HC_PLIGHT_COLOR_TEX_MATRIX]), f2(tct.x, tct.y), L + HC_SKY_DOME_SAMPLER0, s);
  out.isPoint = false;
  out.pos = samplePos;
  out.color = Mat3(L, HC_PLIGHT_COLOR_X)*txClr;
  out.pdf = (float)((double)(mapPdf*1.0f)/((double)2.0f*HC_M_PI_D*HC_M_PI_D*(double)fmaxf(fabsf(sintheta), HC_DEPSILON)));
  out.maxDist = length(illum - samplePos);
  out.cosAtLight = 1.0f;                                                                     // not written by the reference for sky lights; unused by the MISPT loop
}
// environmentColor (cbidir.h:492-533).  prevMaterialOffset: the MISPT integrators never set it (-1 from makeInitialMisData, cglobals.h:1391-1399);
// IntegratorStupidPT recurses with a value-initialised MisData() (CPUExp_Integrators_PT.cpp:37), i.e. offset 0 = the first material node.
HC_DEV float3 EnvironmentColor(const HcScene& s, float3 rayDir, float prevPdf, bool prevSpecular, int prevMaterialOffset, unsigned flags)
{
  if (s.skyLightId == -1) return f3(0, 0, 0);
  const unsigned rayBounceNum = (flags & 0x0000FF00u) >> 8, diffBounceNum = flags & 0xFFu;
  const float* L = LightAt(s, s.skyLightId);
  float3 envColor = SkyLightIntensityTexturedEnv(L, rayDir, s);
  if (rayBounceNum > 0 && !(s.gflags & HC_HRT_STUPID_PT_MODE) && !prevSpecular)
  {
    const float lgtPdf = L[HC_PLIGHT_PICK_PROB_REV]*SkyLightEvalPDF(L, rayDir, s);
    envColor *= misWeightHeuristic(prevPdf, lgtPdf);
  }
  if (prevMaterialOffset >= 0)
  {
    const float* prevMat = reinterpret_cast<const float*>(s.materials + prevMaterialOffset);
    const bool disableCaustics = (diffBounceNum > 0) && !(s.gflags & HC_HRT_ENABLE_PT_CAUSTICS) &&
                                 ((MatI(prevMat, HC_PLAIN_MAT_FLAGS_OFFSET) & HC_PLAIN_MATERIAL_CAST_CAUSTICS) != 0);
    if (disableCaustics) envColor = f3(0, 0, 0);
  }
  return envColor;
}

// LightSampleRev / lightEvalPDF dispatch (clight.h:1561-1633) over the light types hc_pt_init accepts
// ---- spot and directional lights (clight.h:892-912, 1416-1506; MapSamplesToCone cglobals.h:1655-1680).  Both are delta lights (isPoint).
#define HC_POINT_LIGHT_SPOT_COS1  14
#define HC_POINT_LIGHT_SPOT_COS2  15
#define HC_DIRECT_LIGHT_RADIUS1   14
#define HC_DIRECT_LIGHT_RADIUS2   15
#define HC_DIRECT_LIGHT_SSOFTNESS 16
#define HC_DIRECT_LIGHT_ALPHA_TAN 17
#define HC_DIRECT_LIGHT_ALPHA_COS 18
HC_DEV float LocalSmoothstep(float edge0, float edge1, float x)                                                             // clight.h:7-12
{
  const float tVal = (x - edge0)/(edge1 - edge0);
  const float t = fminf(fmaxf(tVal, 0.0f), 1.0f);
  return t*t*(3.0f - 2.0f*t);
}
HC_DEV void SpotLightSampleRev(const float* L, float3 illum, HcShadowSample& out)                                            // clight.h:1432-1450
{
  const float3 samplePos = Mat3(L, HC_PLIGHT_POS_X), norm = Mat3(L, HC_PLIGHT_NORM_X);
  const float hitDist = length(samplePos - illum);
  const float3 rayDir = normalize(samplePos - illum);
  const float cosT = fmaxf(dot((-1.0f)*rayDir, norm), 0.0f);                                // pointSpotLightAttenuation, clight.h:1416-1423
  const float3 color = Mat3(L, HC_PLIGHT_COLOR_X)*LocalSmoothstep(L[HC_POINT_LIGHT_SPOT_COS2], L[HC_POINT_LIGHT_SPOT_COS1], cosT);
  out.isPoint = true;
  out.pos = samplePos;
  out.color = color;
  out.pdf = PdfAtoW(1.0f, hitDist, 1.0f);
  out.maxDist = hitDist;
  out.cosAtLight = fmaxf(-dot(rayDir, norm), 0.0f);
}
HC_DEV float3 MapSamplesToCone(float cosCutoff, float sx, float sy, float3 direction)
{
  const float cosTheta = (1.0f - sx) + sx*cosCutoff;
  const float sinTheta = sqrtf(1.0f - cosTheta*cosTheta);
  const float sinPhi = (float)hc_d_sin(2.0*HC_M_PI_D*(double)sy);                                 // 2.0f*M_PI*sample.y with the <cmath> double M_PI
  const float cosPhi = (float)hc_d_cos(2.0*HC_M_PI_D*(double)sy);
  const float3 dev = f3(cosPhi*sinTheta, sinPhi*sinTheta, cosTheta);
  float3 nx, nzT;
  CoordinateSystem(direction, nx, nzT);
  const float3 ny = nzT, nz = direction;
  return nx*dev.x + ny*dev.y + nz*dev.z;
}
HC_DEV float DirectLightAttenuation(const float* L, float3 illum)                                                            // clight.h:892-912
{
  const float3 lpos = Mat3(L, HC_PLIGHT_POS_X), norm = Mat3(L, HC_PLIGHT_NORM_X);
  const float cosAlpha = dot(normalize(illum - lpos), norm);
  if (!(cosAlpha > 0.0f)) return 0.0f;
  const float sinAlpha = sqrtf(1.0f - cosAlpha*cosAlpha);
  const float d = length(illum - lpos)*sinAlpha;
  const float r1 = L[HC_DIRECT_LIGHT_RADIUS1], r2 = L[HC_DIRECT_LIGHT_RADIUS2];
  return LocalSmoothstep(fmaxf(r2, r1), fminf(r2, r1), d);
}
HC_DEV void DirectLightSampleRev(const float* L, float3 rands, float3 illum, HcShadowSample& out)                            // clight.h:1478-1506
{
  const float3 lpos = Mat3(L, HC_PLIGHT_POS_X);
  float3 norm = Mat3(L, HC_PLIGHT_NORM_X);
  const float pdfW = 1.0f;
  if (L[HC_DIRECT_LIGHT_SSOFTNESS] > 1e-5f) norm = MapSamplesToCone(L[HC_DIRECT_LIGHT_ALPHA_COS], rands.x, rands.y, norm);
  const float3 AC = illum - lpos;
  const float CBLen = dot(normalize(AC), norm)*length(AC);
  out.isPoint = true;
  out.pos = illum - norm*CBLen;
  out.color = Mat3(L, HC_PLIGHT_COLOR_X)*DirectLightAttenuation(L, illum)*pdfW;
  out.pdf = pdfW;
  out.maxDist = CBLen;
  out.cosAtLight = 1.0f;
}
HC_DEV float DirectLightEvalPDF(const float* L, float3 rayDir)                                                               // clight.h:1462-1476
{
  if (!(L[HC_DIRECT_LIGHT_SSOFTNESS] > 1e-5f)) return 1.0f;
  const float tanAlpha = L[HC_DIRECT_LIGHT_ALPHA_TAN];
  const float cosTheta = -dot(rayDir, Mat3(L, HC_PLIGHT_NORM_X));
  return (float)(HC_M_PI_D*(double)(tanAlpha*tanAlpha)*(double)(cosTheta*cosTheta*cosTheta));
}

// ---- mesh lights (clight.h:957-1029, 1513-1546): an emissive mesh sampled by area; the mesh (a PlainMesh copy) and the prefix sums of its
// triangle areas live in the "pdfs" storage (MeshLight, PlainLightConverter.cpp:724-783; CalcTrianglePickProbTable)
#define HC_MESH_LIGHT_MESH_OFFSET_ID  14
#define HC_MESH_LIGHT_TABLE_OFFSET_ID 15
#define HC_MESH_LIGHT_TRI_NUM         16
#define HC_MESH_LIGHT_MATRIX_E00      20
#define HC_MESH_LIGHT_TEX_ID          30
#define HC_MESH_LIGHT_TEXMATRIX_ID    31      // float4 index of the SWTexSampler inside the light record (MESH_LIGHT_TEX_SAMPLER / 4), clight.h:174-176
HC_DEV float3 Mat3x3MulVec(const float* M, float3 v)                                                                          // matrix3x3f_mult_float3, cglobals.h:1091-1098
{
  return f3(M[0]*v.x + M[1]*v.y + M[2]*v.z, M[3]*v.x + M[4]*v.y + M[5]*v.z, M[6]*v.x + M[7]*v.y + M[8]*v.z);
}
HC_DEV void MeshLightSampleRev(const float* L, float3 rands, float3 illum, const HcScene& s, HcShadowSample& out)
{
  const int meshId = __float_as_int(L[HC_MESH_LIGHT_MESH_OFFSET_ID]), pdftId = __float_as_int(L[HC_MESH_LIGHT_TABLE_OFFSET_ID]);
  const int triNum = __float_as_int(L[HC_MESH_LIGHT_TRI_NUM]);
  const float4* mesh = s.pdfs + s.globals[s.pdfTableTableOffset + meshId];
  const float* table = reinterpret_cast<const float*>(s.pdfs + s.globals[s.pdfTableTableOffset + pdftId]);
  const int4 h0 = reinterpret_cast<const int4*>(mesh)[0];                        // vPosOffset vNormOffset vTexCoordOffset vIndicesOffset
  const float4* vpos = mesh + h0.x; const float4* vnorm = mesh + h0.y;
  const int* indices = reinterpret_cast<const int*>(mesh + h0.w);
  float pickProb = 1.0f;
  const int tri = SelectIndexPropToOpt(rands.z, table, triNum + 1, pickProb);
  const int iA = indices[tri*3 + 0], iB = indices[tri*3 + 1], iC = indices[tri*3 + 2];
  const float4 dA = vpos[iA], dB = vpos[iB], dC = vpos[iC], eA = vnorm[iA], eB = vnorm[iB], eC = vnorm[iC];
  const float3 A = f3(dA), B = f3(dB), C = f3(dC);
  const float3 nA = f3(eA), nB = f3(eB), nC = f3(eC);
  float u = rands.x, v = rands.y;
  if (u + v > 1.0f) { u = 1.0f - u; v = 1.0f - v; }
  const float w = 1.0f - u - v;
  float3 samplePos = (A*u + B*v + C*w), sampleNorm = (nA*u + nB*v + nC*w);
  const float2 sampleTc = f2(dA.w, eA.w)*u + f2(dB.w, eB.w)*v + f2(dC.w, eC.w)*w;           // texture coordinates ride in pos.w / norm.w (clight.h:1007-1024)
  const float pdfA = 1.0f/L[HC_PLIGHT_SURFACE_AREA];
  const float* M = L + HC_MESH_LIGHT_MATRIX_E00;
  samplePos = Mat3x3MulVec(M, samplePos);
  sampleNorm = normalize(Mat3x3MulVec(M, sampleNorm));
  samplePos = samplePos + Mat3(L, HC_PLIGHT_POS_X);
  const float3 rayDir = normalize(samplePos - illum);
  const float hitDist = length(samplePos - illum);
  const float cosVal = fmaxf(-dot(rayDir, sampleNorm), 0.0f);
  out.isPoint = false;
  out.pos = samplePos + epsilonOfPos(samplePos)*sampleNorm;
  out.color = Sample2D(__float_as_int(L[HC_MESH_LIGHT_TEXMATRIX_ID]), sampleTc, L, s)*Mat3(L, HC_PLIGHT_COLOR_X);   // meshLightGetIntensity, clight.h:957-963
  out.pdf = PdfAtoW(pdfA, hitDist, cosVal);
  out.maxDist = hitDist;
  out.cosAtLight = cosVal;
}
HC_DEV float MeshLightEvalPDF(const float* L, float3 rayDir, float3 lnorm, float hitDist)                                     // clight.h:1541-1546
{
  const float pdfA = 1.0f/fmaxf(L[HC_PLIGHT_SURFACE_AREA], HC_DEPSILON);
  const float cosVal = fmaxf(dot(rayDir, (-1.0f)*lnorm), 0.0f);
  return PdfAtoW(pdfA, hitDist, cosVal);
}

// ---- cylinder lights (clight.h:753-812, 1338-1380): a tube section sampled uniformly or - when its pdf table id is positive - through the 2D
// table the driver builds for it (UpdatePdfTablesForLight: 2x2 uniform when the light has no texture).  sincos2f = double sin / cos.
#define HC_CYLINDER_LIGHT_MATRIX_E00 16
#define HC_CYLINDER_LIGHT_RADIUS     25
#define HC_CYLINDER_LIGHT_ZMIN       26
#define HC_CYLINDER_LIGHT_ZMAX       27
#define HC_CYLINDER_LIGHT_PHIMAX     28
#define HC_CYLINDER_TEX_ID           29
#define HC_CYLINDER_TEXMATRIX_ID     30      // float4 index of the sampler inside the light record (CYLINDER_TEX_SAMPLER / 4), clight.h:111-114
#define HC_CYLINDER_PDF_TABLE_ID     31
HC_DEV void CylinderLightSampleRev(const float* L, float3 rands, float3 illum, const HcScene& s, HcShadowSample& out)
{
  float2 tc = f2(rands.x, rands.y); float mapPdf = 1.0f;
  const int texId = __float_as_int(L[HC_CYLINDER_PDF_TABLE_ID]);
  if (texId > 0)
  {
    const float* hdr = PdfTableHeader(texId, s);
    const float* intervals = hdr + 4;
    const int sizeX = __float_as_int(hdr[0]), sizeY = __float_as_int(hdr[1]);
    const float fw = (float)sizeX, fh = (float)sizeY, fN = fw*fh;
    float pdf = 1.0f;                                                                        // sampleMap2D, clight.h:375-396
    int pixelOffset = SelectIndexPropToOpt(rands.z, intervals, sizeX*sizeY + 1, pdf);
    if (pixelOffset >= sizeX*sizeY) pixelOffset = sizeX*sizeY - 1;
    const int yPos = pixelOffset/sizeX, xPos = pixelOffset - yPos*sizeX;
    tc.x = (1.0f/fw)*(((float)(xPos) + 0.5f) + (rands.x*2.0f - 1.0f)*0.5f);
    tc.y = (1.0f/fh)*(((float)(yPos) + 0.5f) + (rands.y*2.0f - 1.0f)*0.5f);
    mapPdf = pdf*fN;
  }
  const float pdfA = mapPdf/L[HC_PLIGHT_SURFACE_AREA];
  const float zMin = L[HC_CYLINDER_LIGHT_ZMIN], zMax = L[HC_CYLINDER_LIGHT_ZMAX], radius = L[HC_CYLINDER_LIGHT_RADIUS], phiMax = L[HC_CYLINDER_LIGHT_PHIMAX];
  const float z = zMin + tc.x*(zMax - zMin);
  const float phi = tc.y*phiMax;
  const float sn = hc_sin(phi), cs = hc_cos(phi);
  float3 pObj = f3(radius*cs, radius*sn, z);
  float3 n = normalize(f3(pObj.x, pObj.y, 0.0f));
  const float hitRad = sqrtf(pObj.x*pObj.x + pObj.y*pObj.y);
  pObj.x *= radius/hitRad;
  pObj.y *= radius/hitRad;
  const float* M = L + HC_CYLINDER_LIGHT_MATRIX_E00;
  n = normalize(Mat3x3MulVec(M, n));
  const float3 center = Mat3(L, HC_PLIGHT_POS_X);
  const float3 samplePos = center + Mat3x3MulVec(M, pObj) + epsilonOfPos(center)*n;
  const float hitDist = length(samplePos - illum);
  const float3 rayDir = normalize(samplePos - illum);
  const float cosVal = fmaxf(dot(rayDir, (-1.0f)*n), 0.0f);
  out.isPoint = false;
  out.pos = samplePos;
  out.color = Sample2D(__float_as_int(L[HC_CYLINDER_TEXMATRIX_ID]), tc, L, s)*Mat3(L, HC_PLIGHT_COLOR_X);      // cylinderLightGetIntensity, clight.h:753-759
  out.pdf = PdfAtoW(pdfA, hitDist, cosVal);
  out.maxDist = hitDist;
  out.cosAtLight = cosVal;
}
HC_DEV float CylinderLightEvalPDF(const float* L, float3 illum, float3 lpos, float3 lnorm, float2 texCoord, const HcScene& s)  // clight.h:1338-1357
{
  float mapPdf = 1.0f;
  const int texId = __float_as_int(L[HC_CYLINDER_PDF_TABLE_ID]);
  if (texId)                                                                                 // sic: any non-zero id (ids < 0 are rejected at init)
  {
    const float* hdr = PdfTableHeader(texId, s);
    mapPdf = EvalMap2DPdf(texCoord, hdr + 4, __float_as_int(hdr[0]), __float_as_int(hdr[1]));
  }
  const float hitDist = length(lpos - illum);
  const float3 rayDir = normalize(lpos - illum);
  const float pdfA = mapPdf/fmaxf(L[HC_PLIGHT_SURFACE_AREA], HC_DEPSILON);
  const float cosVal = fmaxf(dot(rayDir, (-1.0f)*lnorm), 0.0f);
  return PdfAtoW(pdfA, hitDist, cosVal);
}

HC_DEV void LightSampleRev(const float* L, float3 rands, float3 illum, const HcScene& s, HcShadowSample& out)
{
  const int type = __float_as_int(L[HC_PLIGHT_TYPE]);
  if (type == HC_PLAIN_LIGHT_TYPE_SKY_DOME) SkyLightSampleRev(L, rands, illum, s, out);
  else if (type == HC_PLAIN_LIGHT_TYPE_SPHERE) SphereLightSampleRev(L, rands, illum, out);
  else if (type == HC_PLAIN_LIGHT_TYPE_POINT_OMNI) PointLightSampleRev(L, illum, s, out);
  else if (type == HC_PLAIN_LIGHT_TYPE_POINT_SPOT) SpotLightSampleRev(L, illum, out);
  else if (type == HC_PLAIN_LIGHT_TYPE_DIRECT) DirectLightSampleRev(L, rands, illum, out);
  else if (type == HC_PLAIN_LIGHT_TYPE_MESH) MeshLightSampleRev(L, rands, illum, s, out);
  else if (type == HC_PLAIN_LIGHT_TYPE_CYLINDER) CylinderLightSampleRev(L, rands, illum, s, out);
  else AreaLightSampleRev(L, rands, illum, s, out);
}

HC_DEV float LightEvalPDF(const float* L, float3 illum, float3 rayDir, float3 lpos, float3 lnorm, float2 texCoord, const HcScene& s)
{
  const float hitDist = length(illum - lpos);
  const int type = __float_as_int(L[HC_PLIGHT_TYPE]);
  if (type == HC_PLAIN_LIGHT_TYPE_SPHERE) return SphereLightEvalPDF(L, illum, lpos, lnorm);
  if (type == HC_PLAIN_LIGHT_TYPE_POINT_OMNI || type == HC_PLAIN_LIGHT_TYPE_POINT_SPOT)                                    // pointLightEvalPDF / spotLightEvalPDF,
    return PdfAtoW(1.0f, length(Mat3(L, HC_PLIGHT_POS_X) - illum), 1.0f);                                                  // clight.h:1387-1392, 1425-1430
  if (type == HC_PLAIN_LIGHT_TYPE_DIRECT) return DirectLightEvalPDF(L, rayDir);
  if (type == HC_PLAIN_LIGHT_TYPE_MESH) return MeshLightEvalPDF(L, rayDir, lnorm, hitDist);
  if (type == HC_PLAIN_LIGHT_TYPE_CYLINDER) return CylinderLightEvalPDF(L, illum, lpos, lnorm, texCoord, s);
  return AreaLightEvalPDF(L, rayDir, hitDist);
}

// emissionEval (cbidir.h:653-678) through IntegratorCommon::emissionEval (CPUExp_Integrators_Common.cpp:509-520)
HC_DEV float3 EmissionEval(const HcScene& s, float3 rayPos, float3 rayDir, const HcSurfaceHit& sh, unsigned flags, int instId)
{
  const float* L = (s.lightsNum > 0) ? LightAt(s, s.instLightIds[instId]) : nullptr;
  const float* mat = MaterialAt(s, sh.matId);
  const float3 normal = sh.hfi ? (-1.0f)*sh.normal : sh.normal;
  const bool hasIES = (L != nullptr) && (__float_as_int(L[HC_PLIGHT_FLAGS]) & HC_LIGHT_HAS_IES) != 0;     // a light with a photometric web emits from its back side too
  if (dot(rayDir, normal) >= 0.0f && !hasIES) return f3(0, 0, 0);
  float3 outColor = MaterialEvalEmission(mat, rayDir, normal, sh.texCoord, s);
  if ((MatI(mat, HC_PLAIN_MAT_FLAGS_OFFSET) & HC_PLAIN_MATERIAL_FORBID_EMISSIVE_GI) && (flags & 0xFFu) > 0) outColor = f3(0, 0, 0);
  if (s.lightsNum > 0 && L != nullptr)                                                  // lightGetIntensity (clight.h:1661-1706): base colour, times the light's texture
  {                                                                                     // at the hit's texture coordinates for cylinder and mesh lights
    outColor = Mat3(L, HC_PLIGHT_COLOR_X);
    const int type = __float_as_int(L[HC_PLIGHT_TYPE]);
    if (type == HC_PLAIN_LIGHT_TYPE_AREA)
    {
      float3 customDir = rayDir;
      if (__float_as_int(L[HC_PLIGHT_FLAGS]) & HC_LIGHT_IES_POINT_AREA) customDir = normalize(Mat3(L, HC_PLIGHT_POS_X) - rayPos);
      outColor = AreaLightIntensity(L, customDir, (flags & 0xFFu) == 0, s);                 // eyeRay: no diffuse bounce so far
    }
    else if (type == HC_PLAIN_LIGHT_TYPE_CYLINDER) outColor = Sample2D(__float_as_int(L[HC_CYLINDER_TEXMATRIX_ID]), sh.texCoord, L, s)*outColor;
    else if (type == HC_PLAIN_LIGHT_TYPE_MESH) outColor = Sample2D(__float_as_int(L[HC_MESH_LIGHT_TEXMATRIX_ID]), sh.texCoord, L, s)*outColor;
  }
  return outColor;
}

// flagsNextBounceLite (cmaterial.h:3262-3293)
HC_DEV unsigned FlagsNextBounceLite(unsigned flags, const HcMatSample& ms, const HcScene& s)
{
  const bool diffuse = (ms.flags & HC_RAY_EVENT_D) != 0;
  const unsigned bounce = (flags & 0x0000FF00u) >> 8, diffB = flags & 0xFFu;
  unsigned other = (flags & 0xFFFF0000u) >> 16;
  flags = (flags & 0xFFFF00FFu) | ((bounce + 1) << 8);
  if (diffuse) flags = (flags & 0xFFFFFF00u) | (diffB + 1);
  const unsigned bounce2 = bounce + 1, diff2 = flags & 0xFFu;
  if ((bounce2 >= (unsigned)s.traceDepth) || (diff2 >= (unsigned)s.diffTraceDepth + 1)) other |= HC_RAY_IS_DEAD;
  if (ms.flags & HC_RAY_EVENT_G) other |= HC_RAY_EVENT_G;
  if ((ms.flags & HC_RAY_EVENT_S) != 0 || (ms.flags & HC_RAY_EVENT_T) != 0) other |= HC_RAY_EVENT_S;
  if (ms.flags & HC_RAY_EVENT_D) other |= HC_RAY_EVENT_D;
  if (ms.flags & HC_RAY_EVENT_T) other |= HC_RAY_EVENT_T;
  return (flags & 0x0000FFFFu) | (other << 16);
}
