// TEST INFRASTRUCTURE ONLY — never imported, linked or called by the product path (hydracore_b200/).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it, as the checker.
//
// oracle/hydra_oracle.cpp — plain C++ RESTATEMENT (own words, no reference code) of the reference algorithms on the
// ray-casting part of the hot path, each function citing the reference file:line it follows (relative to the reference
// tree Ray-Tracing-Systems/HydraCore):
//   a1  RandomGen / NextState / RandomGenInit / rndFloat4_Pseudo / rndFloat1_Pseudo      hydra_drv/crandom.h:10-83
//   a2  Niederreiter base-2 table + rndQmcSobolN                                          hydra_drv/qmc_sobol_niederreiter.cpp:179-186, crandom.h:224-236
//   a4  MakeRandEyeRay / MakeEyeRayFromF4Rnd                                              hydra_drv/cfetch.h:877-968
//   a6  SafeInverse, RayBoxIntersectionLite2                                              hydra_drv/cglobals.h:726-735, ctrace.h:32-53
//   a7  BVH4InstTraverse (closest hit, two-level)                                         hydra_drv/ctrace.h:841-1062
//   a8  IntersectAllPrimitivesInLeaf (Moeller-Trumbore with the -1e-6 slack)              hydra_drv/ctrace.h:124-182
//   a9  shadow semantics of IntegratorCommon::shadowTrace                                 hydra_drv/CPUExp_Integrators_Common.cpp:156-180
// plus the traversal work counters Q (quads fetched), L (triangle leaves entered), T (triangles tested) that define the
// algorithmic bytes per ray of SURVEY.md 8(d):  B_ray = 32 + 4 + 16 + 128 Q + 16 L + 48 T.
//
// PINNING: tests/test_oracle.py checks every function here (a) against golden vectors generated from the reference's own
// headers compiled in place (oracle/_ref, script tests/golden/make_golden.py, fixtures tests/golden/*.npz) and (b) live
// against oracle/_ref/libhydra_ref.so whenever that library is present.  The shading half of the path (a10-a16) is checked
// directly against oracle/_ref (the reference's own code), not restated here.
// Built with -ffp-contract=off so that, like the reference's SSE4.2 build, no multiply-add is fused.
#include <cmath>
#include <cstdint>
#include <cstring>

namespace
{
struct V3 { float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator*(V3 a, float s) { return { a.x*s, a.y*s, a.z*s }; }
inline V3 operator*(float s, V3 a) { return { a.x*s, a.y*s, a.z*s }; }
inline V3 operator/(V3 a, float s) { return { a.x/s, a.y/s, a.z/s }; }
inline float dot(V3 a, V3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
inline V3 cross(V3 a, V3 b) { return { a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x }; }
inline V3 normalize(V3 a) { return a/std::sqrt(a.x*a.x + a.y*a.y + a.z*a.z); }
struct F4 { float x, y, z, w; };
struct M4 { F4 c[4]; };                                     // four columns (cglobals.h:206-209)
inline int32_t asInt(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }
inline uint32_t asUint(float f) { uint32_t i; std::memcpy(&i, &f, 4); return i; }

inline V3 mul4x3(const M4& m, V3 v)                          // cglobals.h:306-313
{
  return { v.x*m.c[0].x + v.y*m.c[1].x + v.z*m.c[2].x + m.c[3].x,
           v.x*m.c[0].y + v.y*m.c[1].y + v.z*m.c[2].y + m.c[3].y,
           v.x*m.c[0].z + v.y*m.c[1].z + v.z*m.c[2].z + m.c[3].z };
}
inline V3 mul3x3(const M4& m, V3 v)                          // cglobals.h:315-322
{
  return { v.x*m.c[0].x + v.y*m.c[1].x + v.z*m.c[2].x,
           v.x*m.c[0].y + v.y*m.c[1].y + v.z*m.c[2].y,
           v.x*m.c[0].z + v.y*m.c[1].z + v.z*m.c[2].z };
}

// ---------------------------------------------------------------------------------------------------- a1: RandomGen
struct Rng { uint32_t x, y; };
inline uint32_t NextState(Rng& g)                            // crandom.h:19-25
{
  const uint32_t x = g.x*17u + g.y*13123u;
  g.x = (x << 13) ^ x;
  g.y ^= (x << 7);
  return x;
}
inline Rng RngInit(int32_t seedSigned)                       // crandom.h:27-43 (signed int arithmetic that wraps)
{
  const uint32_t s = (uint32_t)seedSigned;
  Rng g;
  g.x = s*(s*s*15731u + 74323u) + 871483u;
  g.y = s*(s*s*13734u + 37828u) + 234234u;
  const int n = seedSigned % 7;                              // C remainder: negative seeds run zero iterations
  for (int i = 0; i < n; i++) NextState(g);
  return g;
}
inline void RngFloat4(Rng& g, float out[4])                  // crandom.h:51-63
{
  const uint32_t x = NextState(g);
  const uint32_t x1 = x*(x*x*15731u + 74323u) + 871483u;
  const uint32_t y1 = x*(x*x*13734u + 37828u) + 234234u;
  const uint32_t z1 = x*(x*x*11687u + 26461u) + 137589u;
  const uint32_t w1 = x*(x*x*15707u + 789221u) + 1376312589u;
  const float scale = 1.0f/4294967296.0f;
  out[0] = (float)x1*scale; out[1] = (float)y1*scale; out[2] = (float)z1*scale; out[3] = (float)w1*scale;
}
inline float RngFloat1(Rng& g)                               // crandom.h:77-83
{
  const uint32_t x = NextState(g);
  const uint32_t t = x*(x*x*15731u + 74323u) + 871483u;
  return (float)t*(1.0f/4294967296.0f);
}

// ---------------------------------------------------------------------------------------------------- a2: Niederreiter base 2
// Bratley, Fox, Niederreiter, "Implementation and test of low discrepancy sequences" (ACM TOMACS 2(3), 1992; TOMS 738),
// in the 63-bit form the reference embeds (qmc_sobol_niederreiter.cpp:75-167) and then truncates to the top 31 bits
// (initQuasirandomGenerator, :179-186).  Per dimension: an irreducible polynomial p over GF(2) of degree e; every e-th
// column the running power b = p^q is multiplied by p once more and the linear recurrence defined by b generates v[].
const int QDIM = 11, QRES = 31;

inline int PolyDeg(uint64_t p) { int d = -1; while (p) { d++; p >>= 1; } return d; }
inline uint64_t PolyMul(uint64_t a, uint64_t b) { uint64_t r = 0; while (b) { if (b & 1) r ^= a; a <<= 1; b >>= 1; } return r; }

void NiederreiterTable(uint32_t table[QDIM][QRES])
{
  // the first 11 irreducible polynomials over GF(2) by increasing bit pattern (what GeneratePolynomials(buffer, false)
  // finds, qmc_sobol_niederreiter.cpp:15-66): x, x+1, x^2+x+1, x^3+x+1, x^3+x^2+1, x^4+x+1, x^4+x^3+1, x^4+x^3+x^2+x+1,
  // x^5+x^2+1, x^5+x^3+1, x^5+x^3+x^2+x+1
  static const uint64_t irred[QDIM] = { 2, 3, 7, 11, 13, 19, 25, 31, 37, 41, 47 };
  const int NB = 63;
  for (int dim = 0; dim < QDIM; dim++)
  {
    const uint64_t p = irred[dim];
    const int e = PolyDeg(p);
    uint64_t cj[NB]; for (int i = 0; i < NB; i++) cj[i] = 0;
    uint64_t b = 1; int m = 0;                 // b(x) = p(x)^q, degree m
    int v[NB + 16];
    int u = e;
    for (int j = NB - 1; j >= 32; --j, ++u)     // only bits 62..32 survive the truncation to 31 bits
    {
      if (u == e)
      {
        u = 0;
        const int m1 = m;
        b = PolyMul(b, p); m += e;
        for (int i = 0; i < m1; i++) v[i] = 0;
        for (int i = m1; i < m; i++) v[i] = 1;
        for (int i = m; i <= NB + e - 2; i++)
        {
          int acc = 0;
          for (int k = 1; k <= m; k++) acc ^= v[i - k] & int((b >> (m - k)) & 1u);   // b[k] = coefficient of x^(m-k)
          v[i] = acc;
        }
      }
      for (int i = 0; i < NB; i++) cj[i] |= (uint64_t)v[i + u] << j;
    }
    for (int bit = 0; bit < QRES; bit++) table[dim][bit] = (uint32_t)((cj[bit] >> 32) & 0x7FFFFFFFu);
  }
}

inline float QmcSobolN(uint32_t pos, int dim, const uint32_t* table)   // crandom.h:228-236
{
  uint32_t result = 0, data = pos;
  for (int bit = 0; bit < QRES; bit++, data >>= 1)
    if (data & 1) result ^= table[bit + dim*QRES];
  return (float)(result + 1)*(1.0f/(float)0x80000001U);
}

// ---------------------------------------------------------------------------------------------------- a4: eye rays
struct Camera
{
  M4 projInv, worldViewInv;
  float fov, tiltX, tiltY, lensRadius, focalDist, fwidth, fheight;
  int enableDof;
};

Camera CameraFromGlobals(const unsigned char* g)
{
  // EngineGlobals field offsets (cfetch.h:21-81): mProjInverse @128, mWorldViewInverse @192, varsI @256, varsF @512
  Camera c;
  std::memcpy(&c.projInv, g + 128, 64);
  std::memcpy(&c.worldViewInv, g + 192, 64);
  const int* varsI = (const int*)(g + 256);
  const float* varsF = (const float*)(g + 512);
  c.fov = varsF[14];            // HRT_CAM_FOV
  c.tiltX = varsF[2]; c.tiltY = varsF[4];
  c.lensRadius = varsF[0]; c.focalDist = varsF[1];
  c.fwidth = varsF[25]; c.fheight = varsF[26];
  c.enableDof = (varsI[0] == 1);
  return c;
}

inline V3 EyeRayDirNormalized(float x, float y, const M4& m)           // cglobals.h:1069-1078
{
  const float px = 2.0f*x - 1.0f, py = 2.0f*y - 1.0f, pz = 0.0f, pw = 1.0f;
  F4 r;
  r.x = px*m.c[0].x + py*m.c[1].x + pz*m.c[2].x + pw*m.c[3].x;
  r.y = px*m.c[0].y + py*m.c[1].y + pz*m.c[2].y + pw*m.c[3].y;
  r.z = px*m.c[0].z + py*m.c[1].z + pz*m.c[2].z + pw*m.c[3].z;
  r.w = px*m.c[0].w + py*m.c[1].w + pz*m.c[2].w + pw*m.c[3].w;
  return normalize(V3{ r.x/r.w, r.y/r.w, r.z/r.w });
}

inline void MapSamplesToDisc(float x, float y, float& ox, float& oy)  // cglobals.h:1609-1655 ; sin/cos evaluated in double like the oracle build
{
  float r = 0.0f, phi = 0.0f;
  if (x > y && x > -y) { r = x;  phi = 0.25f*3.141592654f*(y/x); }
  if (x < y && x > -y) { r = y;  phi = 0.25f*3.141592654f*(2.0f - x/y); }
  if (x < y && x < -y) { r = -x; phi = 0.25f*3.141592654f*(4.0f + y/x); }
  if (x > y && x < -y) { r = -y; phi = 0.25f*3.141592654f*(6 - x/y); }
  ox = r*(float)std::sin((double)phi); oy = r*(float)std::cos((double)phi);
}

inline void MultRay3(const M4& m, V3& pos, V3& dir)                   // matrix4x4f_mult_ray3, cglobals.h:1080-1087
{
  const V3 p = mul4x3(m, pos), p2 = mul4x3(m, pos + 100.0f*dir);
  pos = p; dir = normalize(p2 - p);
}

void MakeRandEyeRay(int x, int y, int w, int h, const float offs[4], const Camera& cam, V3& rpos, V3& rdir)   // cfetch.h:877-931 (tilt = 0 only)
{
  V3 pos{ 0, 0, 0 };
  V3 dir = EyeRayDirNormalized(((float)x + 0.5f)/(float)w, ((float)y + 0.5f)/(float)h, cam.projInv);
  const float sinFov = (float)std::sin((double)(0.5f*cam.fov));
  const float pxX = sinFov*(1.0f/(float)w), pxY = sinFov*(1.0f/(float)h);
  dir.x += pxX*offs[0];
  dir.y += pxY*offs[1];
  dir.z = -std::sqrt(1.0f - (dir.x*dir.x + dir.y*dir.y));
  if (cam.enableDof)
  {
    const float tFocus = cam.focalDist/(-dir.z);
    const V3 focus = pos + dir*tFocus;
    float dx, dy; MapSamplesToDisc(1.0f*offs[2], 1.0f*offs[3], dx, dy);
    pos.x += cam.lensRadius*dx; pos.y += cam.lensRadius*dy;
    dir = normalize(focus - pos);
  }
  MultRay3(cam.worldViewInv, pos, dir);
  rpos = pos; rdir = dir;
}

void MakeEyeRayFromF4Rnd(const float lens[4], const Camera& cam, V3& rpos, V3& rdir, float& fx, float& fy)    // cfetch.h:933-968 (tilt = 0 only)
{
  const float x = cam.fwidth*lens[0], y = cam.fheight*lens[1];
  V3 pos{ 0, 0, 0 };
  V3 dir = EyeRayDirNormalized(x/cam.fwidth, y/cam.fheight, cam.projInv);
  if (cam.enableDof)
  {
    const float tFocus = cam.focalDist/(-dir.z);
    const V3 focus = pos + dir*tFocus;
    float dx, dy; MapSamplesToDisc(lens[2] - 0.5f, lens[3] - 0.5f, dx, dy);
    pos.x += cam.lensRadius*2.0f*dx; pos.y += cam.lensRadius*2.0f*dy;
    dir = normalize(focus - pos);
  }
  MultRay3(cam.worldViewInv, pos, dir);
  fx = lens[0]*cam.fwidth; fy = lens[1]*cam.fheight;
  rpos = pos; rdir = dir;
}

// ---------------------------------------------------------------------------------------------------- a5-a8: traversal
struct Hit { float t; int32_t primId, instId, geomId; };
struct Counters { uint64_t quads, leaves, tris; };
const float MAXF = 3.402823466e+38f;  // MAXFLOAT: <math.h> (and OpenCL) define it as FLT_MAX, so the 1e37f fallback of ctrace.h:665-667 is never taken
const int   STACK = 80;              // ctrace.h:576

inline V3 SafeInverse(V3 d)          // cglobals.h:726-735
{
  const float eps = 1.0e-36f;
  return { 1.0f/(std::fabs(d.x) > eps ? d.x : std::copysign(eps, d.x)),
           1.0f/(std::fabs(d.y) > eps ? d.y : std::copysign(eps, d.y)),
           1.0f/(std::fabs(d.z) > eps ? d.z : std::copysign(eps, d.z)) };
}

struct Slab { float tmin, tmax; };
inline Slab RayBox(V3 o, V3 inv, const float* lo, const float* hi)     // RayBoxIntersectionLite2, ctrace.h:32-53
{
  const float l0 = inv.x*(lo[0] - o.x), h0 = inv.x*(hi[0] - o.x);
  const float l1 = inv.y*(lo[1] - o.y), h1 = inv.y*(hi[1] - o.y);
  const float l2 = inv.z*(lo[2] - o.z), h2 = inv.z*(hi[2] - o.z);
  float tmin = std::fmin(l0, h0), tmax = std::fmax(l0, h0);
  tmin = std::fmax(tmin, std::fmin(l1, h1)); tmax = std::fmin(tmax, std::fmax(l1, h1));
  tmin = std::fmax(tmin, std::fmin(l2, h2)); tmax = std::fmin(tmax, std::fmax(l2, h2));
  return { tmin, tmax };
}

inline void LeafTriangles(V3 o, V3 d, int leafOffset, float tMin, Hit& hit, const F4* tris, int instId, Counters* cnt)   // ctrace.h:124-182
{
  const int first = asInt(tris[leafOffset].x), count = asInt(tris[leafOffset].y);
  if (cnt) { cnt->leaves++; cnt->tris += (uint64_t)count; }
  for (int a = first; a < first + 3*count; a += 3)
  {
    const F4 d1 = tris[a], d2 = tris[a + 1], d3 = tris[a + 2];
    const V3 A{ d1.x, d1.y, d1.z }, B{ d2.x, d2.y, d2.z }, C{ d3.x, d3.y, d3.z };
    const V3 e1 = B - A, e2 = C - A;
    const V3 pvec = cross(d, e2);
    const V3 tvec = o - A;
    const V3 qvec = cross(tvec, e1);
    const float invDet = 1.0f/dot(e1, pvec);
    const float v = dot(tvec, pvec)*invDet;
    const float u = dot(qvec, d)*invDet;
    const float t = dot(e2, qvec)*invDet;
    if (v > -1e-6f && u > -1e-6f && (u + v < 1.0f + 1e-6f) && t > tMin && t < hit.t)
    {
      hit.t = t; hit.primId = asInt(d1.w); hit.geomId = asInt(d2.w); hit.instId = instId;
    }
  }
}

// The reference's control flow, restated: inner loop descends to a leaf sorting the four children near-to-far, the stack holds
// child references only (no distances), a popped quad is always re-fetched and re-tested against the current hit.t.
Hit TraverseClosest(V3 o, V3 d, float tMin, Hit hit, const F4* bvh, const F4* tris, Counters* cnt)                     // ctrace.h:841-1062
{
  V3 inv = SafeInverse(d);
  int32_t stackData[STACK + 2]; int32_t* stack = stackData + 2; stackData[0] = stackData[1] = 0;
  int top = 0; int32_t node = 1; bool searching = true;
  int instDeep = 0, instTop = 0, instId = -1; V3 oldO{ 0, 0, 0 }, oldD{ 0, 0, 0 };

  while (top >= 0)
  {
    while (searching)
    {
      if (cnt) cnt->quads++;
      float tm[4]; int32_t ch[4];
      for (int i = 0; i < 4; i++)
      {
        const int n = 4*node + i;
        const F4 h1 = bvh[2*n], h2 = bvh[2*n + 1];
        const float lo[3] = { h1.x, h1.y, h1.z }, hi[3] = { h2.x, h2.y, h2.z };
        const bool valid = !(asUint(h1.w) == 0xffffffffu && asUint(h2.w) == 0xffffffffu);     // IsValidNode, cglobals.h:1321
        const Slab s = RayBox(o, inv, lo, hi);
        const bool hitChild = (s.tmin <= s.tmax) && (s.tmax >= tMin) && (s.tmin <= hit.t) && valid;
        tm[i] = hitChild ? s.tmin : MAXF; ch[i] = asInt(h1.w);
      }
      auto cswap = [&](int a, int b) { if (tm[b] < tm[a]) { float t = tm[a]; tm[a] = tm[b]; tm[b] = t; int32_t c = ch[a]; ch[a] = ch[b]; ch[b] = c; } };
      cswap(0, 1); cswap(2, 3); cswap(0, 2); cswap(1, 3); cswap(1, 2);                        // network of ctrace.h:900-962
      const bool space = (top < STACK);
      if (tm[3] < MAXF && space) stack[top++] = ch[3];
      if (tm[2] < MAXF && space) stack[top++] = ch[2];
      if (tm[1] < MAXF && space) stack[top++] = ch[1];
      if (tm[0] < MAXF) node = ch[0];
      else if (top >= 0) { top--; node = stack[top]; }
      searching = !(node & 0x80000000) && (top >= 0);
      node = node & 0x7fffffff;
      if (top < instTop && instDeep == 1) { o = oldO; d = oldD; inv = SafeInverse(d); instDeep = 0; }
    }

    if (top >= 0 && instDeep == 1)
    {
      LeafTriangles(o, d, node, tMin, hit, tris, instId, cnt);
      top--; node = stack[top];
    }
    else if (top >= 0 && instDeep == 0)
    {
      instDeep = 1; oldO = o; oldD = d;
      const int32_t next = asInt(bvh[node*8 + 0].w);
      M4 m; m.c[0] = bvh[node*8 + 2]; m.c[1] = bvh[node*8 + 3]; m.c[2] = bvh[node*8 + 4]; m.c[3] = bvh[node*8 + 5];
      instId = asInt(bvh[node*8 + 6].x);
      o = mul4x3(m, o); d = mul3x3(m, d); inv = SafeInverse(d);      // direction is NOT renormalised (ctrace.h:1040-1042)
      instTop = top;
      node = next;
    }
    searching = !(node & 0x80000000);
    node = node & 0x7fffffff;
    if (top < instTop && instDeep == 1) { o = oldO; d = oldD; inv = SafeInverse(d); instDeep = 0; }
  }
  return hit;
}

inline Hit MissHit(float t) { Hit h; h.t = t; h.primId = -1; h.instId = -1; h.geomId = (int32_t)0xC0000000u; return h; }   // Make_Lite_Hit(t, -1), cglobals.h:1258-1268
inline bool HitSome(const Hit& h) { return h.primId != -1 && std::isfinite(h.t); }                                          // cglobals.h:1271
} // namespace

extern "C"
{
void orc_rng_init(int seed, uint32_t* state2) { Rng g = RngInit(seed); state2[0] = g.x; state2[1] = g.y; }
void orc_rng_float4(uint32_t* state2, int n, float* out)
{
  Rng g{ state2[0], state2[1] };
  for (int i = 0; i < n; i++) RngFloat4(g, out + 4*i);
  state2[0] = g.x; state2[1] = g.y;
}
void orc_rng_float1(uint32_t* state2, int n, float* out)
{
  Rng g{ state2[0], state2[1] };
  for (int i = 0; i < n; i++) out[i] = RngFloat1(g);
  state2[0] = g.x; state2[1] = g.y;
}
void  orc_qmc_table(uint32_t* out) { NiederreiterTable((uint32_t (*)[QRES])out); }
float orc_qmc_sobol(uint32_t pos, int dim, const uint32_t* table) { return QmcSobolN(pos, dim, table); }

void orc_make_rand_eye_rays(const void* globals, int w, int h, const int* xy, const float* offsets4, int n, float* out6)
{
  const Camera cam = CameraFromGlobals((const unsigned char*)globals);
  for (int i = 0; i < n; i++)
  {
    V3 p, d; MakeRandEyeRay(xy[2*i], xy[2*i + 1], w, h, offsets4 + 4*i, cam, p, d);
    float* o = out6 + 6*i; o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = d.x; o[4] = d.y; o[5] = d.z;
  }
}
void orc_make_eye_rays_f4(const void* globals, const float* lens4, int n, float* out6, float* outXY)
{
  const Camera cam = CameraFromGlobals((const unsigned char*)globals);
  for (int i = 0; i < n; i++)
  {
    V3 p, d; float fx, fy; MakeEyeRayFromF4Rnd(lens4 + 4*i, cam, p, d, fx, fy);
    float* o = out6 + 6*i; o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = d.x; o[4] = d.y; o[5] = d.z;
    outXY[2*i] = fx; outXY[2*i + 1] = fy;
  }
}

// rays: n x 8 floats {pos.xyz, _, dir.xyz, tFar}; counters3 (may be NULL) accumulates {Q, L, T} over all rays
void orc_trace_closest(const void* nodes, const void* tris, const float* rays8, long long n, void* hitsOut, uint64_t* counters3)
{
  Hit* out = (Hit*)hitsOut;
  uint64_t q = 0, l = 0, t = 0;
  #pragma omp parallel for schedule(dynamic, 256) reduction(+:q, l, t)
  for (long long i = 0; i < n; i++)
  {
    const float* r = rays8 + 8*i;
    Counters c{ 0, 0, 0 };
    out[i] = TraverseClosest(V3{ r[0], r[1], r[2] }, V3{ r[4], r[5], r[6] }, 0.0f, MissHit(MAXF), (const F4*)nodes, (const F4*)tris, counters3 ? &c : nullptr);
    q += c.quads; l += c.leaves; t += c.tris;
  }
  if (counters3) { counters3[0] += q; counters3[1] += l; counters3[2] += t; }
}

// CPU shadow semantics: closest hit, then occluded iff 0 < t < tFar (CPUExp_Integrators_Common.cpp:156-180)
void orc_trace_shadow(const void* nodes, const void* tris, const float* rays8, long long n, unsigned char* visibleOut, uint64_t* counters3)
{
  uint64_t q = 0, l = 0, t = 0;
  #pragma omp parallel for schedule(dynamic, 256) reduction(+:q, l, t)
  for (long long i = 0; i < n; i++)
  {
    const float* r = rays8 + 8*i;
    Counters c{ 0, 0, 0 };
    const Hit h = TraverseClosest(V3{ r[0], r[1], r[2] }, V3{ r[4], r[5], r[6] }, 0.0f, MissHit(MAXF), (const F4*)nodes, (const F4*)tris, counters3 ? &c : nullptr);
    visibleOut[i] = (HitSome(h) && h.t > 0.0f && h.t < r[7]) ? 0 : 1;
    q += c.quads; l += c.leaves; t += c.tris;
  }
  if (counters3) { counters3[0] += q; counters3[1] += l; counters3[2] += t; }
}
} // extern "C"
