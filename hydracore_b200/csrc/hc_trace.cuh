// hc_trace.cuh — K2 / K2s: closest-hit and any-hit traversal of the reference's flat two-level BVH4.
//
// Replaces BVH4TraversalInstKernel / BVH4TraversalInstShadowKenrel (reference hydra_drv/shaders/trace.cl:50, 309), i.e. the
// device functions BVH4InstTraverse / BVH4InstTraverseShadow (hydra_drv/ctrace.h:841-1062, 1065-1294) with the triangle test
// IntersectAllPrimitivesInLeaf (ctrace.h:124-182), the slab test RayBoxIntersectionLite2 (ctrace.h:32-53) and SafeInverse
// (cglobals.h:726-735).  The triangle test uses the same IEEE float operations in the same order (no FMA contraction) and the same
// strict comparisons, children are visited near to far through the reference's sorting network, so the closest hit is the same hit.
// What differs is HOW the tree is stored and walked:
//
//   * hc_set_bvh re-lays the reference blobs out for the device (ConvertBvhForDevice, hc_api.cu).  A quad stays 128 bytes (one L1 line) and
//     keeps its index, but becomes SoA: six float4 rows {minx[4], maxx[4], miny[4], maxy[4], minz[4], maxz[4]}, one uint4 row of child words.
//     Two children share a 64-bit register pair, so one Blackwell packed-FP32 instruction (FADD2 / FMUL2: sub.rn.f32x2, mul.rn.f32x2,
//     each half rounded exactly like the scalar op) slab-tests two children; the near/far row of an axis is picked by the sign of
//     the ray direction (bit-identical to min(lo,hi) / max(lo,hi) for a finite ray) and the three axes are merged with the
//     three-input FMNMX3.  Invalid children carry an infinite box that can never pass, so IsValidNode costs nothing.
//     The slab test is the reference's arithmetic BIT FOR BIT.  Round 2 measured a cheaper alternative - centre / half-extent rows fetched
//     by three 256-bit loads at fixed offsets, conservative boxes with a 2^-19 margin - and rejected it: 35 % fewer L1 wavefronts bought
//     0-5 % of speed, and in the band where a ray passes a leaf box within rounding while the triangle test's 1e-6 slack still accepts it the
//     reference culls and a conservative box does not: 8 of the 16.8 M paths of the C1 frame (512 x 512, 64 spp) ended differently
//     (profiles/r02_k2_variants.md, gpurun_out/c1_diff.json).
//   * triangles are stored two to a record (96 B, three 256-bit loads) with precomputed edges {A, B-A, C-A}, SoA over the pair, so the
//     whole Moeller-Trumbore test runs in packed FP32 on two triangles at once; the leaf's triangle count lives in the child word,
//     which removes the dependent header fetch.
//   * stack entries carry the child's entry distance, and an entry whose distance exceeds the current hit is dropped at pop time
//     without fetching its quad (the reference re-fetches and re-tests all four children);
//   * the world-space ray is parked in local memory while the ray is inside an instance;
//   * "while-while" control flow (all lanes descend, then all lanes intersect) inside persistent warps that pull rays from a
//     global counter with ballot/popc aggregation (one atomic per refill).
// Variants measured and rejected in round 2 (postponed leaves, stack in shared memory, chunked ray supply, more resident warps) are in
// DESIGN.md section 3 with their numbers (profiles/r02_k2_variants.md, profiles/r02_k2_ncu_compare.md).
#pragma once
#include "hc_math.cuh"
#include "hc_texture.cuh"

#define HC_LEAF_BIT      0x80000000u
#define HC_NODE_SENTINEL 0xffffffffu   // "ray finished"; also the child word of an empty slot
#define HC_MAXFLOAT      3.402823466e+38f // MAXFLOAT = FLT_MAX from <math.h> / OpenCL (the 1e37f fallback of ctrace.h:665-667 is never taken)
#define HC_TRI_EPS       1e-6f            // barycentric slack, ctrace.h:111
#define HC_STACK_CAP     80               // entries per ray = the reference's STACK_SIZE (ctrace.h:576); hc_set_bvh checks the tree against it

// triangle-leaf child word: [31] leaf | [30:25] pair records - 1 | [24:0] index of the first pair record (96 bytes each)
#define HC_LEAF_PAIRS_SHIFT 25
#define HC_LEAF_PAIRS_MAX   64
#define HC_LEAF_INDEX_MASK  0x01ffffffu
#define HC_PAIR_F4          6             // float4 per pair record

struct HcHit { float t; int primId; int instId; int geomId; };   // == Lite_Hit (cglobals.h:1248-1256)

struct HcBvh
{
  const float4* __restrict__ nodes;   // device layout: 8 float4 per quad (SoA boxes + child words) or per instance record
  const float4* __restrict__ tris;    // device layout: pair records
  // alpha-tested tree only (BVH4InstTraverseAlpha, ctrace.h:1297-1520): per pair record two uint4 {sampler offset, packed uv of A, B, C}
  // taken from the reference's per-triangle alpha table, the table itself (its tail holds the opacity samplers), the texture storage
  const uint4*  __restrict__ alphaPairs    = nullptr;
  const uint2*  __restrict__ alphaTable    = nullptr;
  const int4*   __restrict__ textures      = nullptr;
  const int*    __restrict__ texturesTable = nullptr;   // EngineGlobals texture table: id -> float4 offset of the image header
  int singleLevel = 0;                                   // 1: no instance level (bvhType "triangle4v", BVH4Traverse ctrace.h:669-838): the leaves of the tree that
                                                         // starts at quad 1 hold world-space triangles, each with its own instance id
};
#define HC_INST_SINGLE 0x7ffffffe                        // HcRayTrav::instId of a ray in a single-level tree: "inside" from the start, never left

HC_DEV float3 SafeInverse(float3 d)
{
  const float ooeps = 1.0e-36f;
  float3 r;
  r.x = 1.0f/(fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
  r.y = 1.0f/(fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
  r.z = 1.0f/(fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
  return r;
}

// ------------------------------------------------------------------------------------------------ packed FP32 (sm_100a)
typedef unsigned long long hc_f2;      // two floats in one 64-bit register pair {lo, hi}

HC_DEV hc_f2 pk2(float lo, float hi)          { hc_f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
HC_DEV hc_f2 bc2(float v)                     { return pk2(v, v); }          // ptxas folds this into the .F32 broadcast operand form
HC_DEV void  upk2(hc_f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
HC_DEV hc_f2 sub2(hc_f2 a, hc_f2 b)           { hc_f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
HC_DEV hc_f2 mul2(hc_f2 a, hc_f2 b)           { hc_f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
HC_DEV float max3f(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
HC_DEV float min3f(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
HC_DEV hc_f2 lo2(float4 v) { return pk2(v.x, v.y); }
HC_DEV hc_f2 hi2(float4 v) { return pk2(v.z, v.w); }

// ptxas 12.9 contracts mul.rn.f32x2 + add/sub.rn.f32x2 into FFMA2 (it never does that to the scalar .rn forms), which would round
// once instead of twice and break bit-parity with the reference.  A difference of two PRODUCTS is therefore written as
// fma(b, -1, a): exact, one instruction, and not contractible.  Sums of products are avoided altogether: cross products are
// produced as (x, -y, -z) by swapping the operands of the y and z differences, so that every dot product
// (a.x*b.x + a.y*b.y) + a.z*b.z becomes (a.x*b.x - a.y*(-b.y)) - a.z*(-b.z) — the same roundings, negation being exact.
// tests/test_abi.py scans the SASS of the built library: an FFMA2 whose multiplier is not the immediate -1 fails the build check.
HC_DEV hc_f2 dif2(hc_f2 a, hc_f2 b)
{ hc_f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(b), "l"(0xbf800000bf800000ull), "l"(a)); return r; }   // a - b

struct HcVec2 { hc_f2 x, y, z; };
// cross(a, b) = (ay*bz - az*by, az*bx - ax*bz, ax*by - ay*bx), returned as (x, -y, -z)
HC_DEV HcVec2 cross2_xnynz(const HcVec2& a, const HcVec2& b)
{
  HcVec2 r;
  r.x = dif2(mul2(a.y, b.z), mul2(a.z, b.y));
  r.y = dif2(mul2(a.x, b.z), mul2(a.z, b.x));
  r.z = dif2(mul2(a.y, b.x), mul2(a.x, b.y));
  return r;
}
// dot(a, b) with b given as (x, -y, -z): (ax*bx + ay*by) + az*bz
HC_DEV hc_f2 dot2_xnynz(const HcVec2& a, const HcVec2& bn) { return dif2(dif2(mul2(a.x, bn.x), mul2(a.y, bn.y)), mul2(a.z, bn.z)); }

#define HC_CSWAP(ta, ca, tb, cb) { const bool s_ = (tb < ta); const float tt_ = s_ ? tb : ta; const float tu_ = s_ ? ta : tb; \
                                   const unsigned ct_ = s_ ? cb : ca; const unsigned cu_ = s_ ? ca : cb; ta = tt_; tb = tu_; ca = ct_; cb = cu_; }


struct HcF8 { float4 a, b; };
HC_DEV HcF8 ldg256(const void* p)
{
  HcF8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
  return r;
}


struct HcRayTrav
{
  float3 o, d, inv;          // current space (world, or the instance's object space)
  float  t; int primId, geomId, hitInst;
  int    instId;             // instance being traversed (-1 outside)
  int    sp, instTop;
  unsigned node;
  unsigned nearX, nearY, nearZ;   // byte offset of the near row of each axis inside a quad (far row = near ^ 16)
};

// row order in a quad: minx 0, maxx 16, miny 32, maxy 48, minz 64, maxz 80 (bytes).  inv < 0 -> the max plane is the near one.
HC_DEV void SetNearRows(HcRayTrav& r)
{
  r.nearX = (r.inv.x < 0.0f) ? 16u : 0u;
  r.nearY = (r.inv.y < 0.0f) ? 48u : 32u;
  r.nearZ = (r.inv.z < 0.0f) ? 80u : 64u;
}

HC_DEV void TravStart(HcRayTrav& r, float3 o, float3 d, float tFar, int singleLevel = 0)
{
  r.o = o; r.d = d; r.inv = SafeInverse(d); SetNearRows(r);
  r.t = tFar; r.primId = -1; r.hitInst = -1; r.geomId = int(0xC0000000u);      // Make_Lite_Hit(t, -1), cglobals.h:1258-1266
  r.instId = singleLevel ? HC_INST_SINGLE : -1; r.sp = 0; r.instTop = 0; r.node = 1u;
}

// entry key of one child: tmin when (tmin <= tmax) && (tmax >= 0) && (tmin <= tHit), else MAXFLOAT (RayBoxIntersectionLite2 + the visit
// condition of BVH4InstTraverse, ctrace.h:32-53, 880-895).  With tHit >= 0 that condition equals max(tmin, 0) <= min(tmax, tHit).
// n* / f* are the per-axis near / far plane distances.
HC_DEV float ChildKey(float nx, float ny, float nz, float fx, float fy, float fz, float tHit)
{
  const float tmin = max3f(nx, ny, nz), tmax = min3f(fx, fy, fz);
  return (fmaxf(tmin, 0.0f) <= fminf(tmax, tHit)) ? tmin : HC_MAXFLOAT;
}

// slab test of the four children of quad `node` for this lane's ray: entry keys (MAXFLOAT = not to be visited) and child words
HC_DEV void QuadKeys(const HcRayTrav& r, const HcBvh& bvh, const unsigned node, float& t0, float& t1, float& t2, float& t3,
                     unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3)
{
  const char* q = reinterpret_cast<const char*>(bvh.nodes) + size_t(node)*128u;
  const float4 NX = __ldg(reinterpret_cast<const float4*>(q + r.nearX)), FX = __ldg(reinterpret_cast<const float4*>(q + (r.nearX ^ 16u)));
  const float4 NY = __ldg(reinterpret_cast<const float4*>(q + r.nearY)), FY = __ldg(reinterpret_cast<const float4*>(q + (r.nearY ^ 16u)));
  const float4 NZ = __ldg(reinterpret_cast<const float4*>(q + r.nearZ)), FZ = __ldg(reinterpret_cast<const float4*>(q + (r.nearZ ^ 16u)));
  const uint4  ch = __ldg(reinterpret_cast<const uint4*>(q + 96));
  const hc_f2 oX = bc2(r.o.x), oY = bc2(r.o.y), oZ = bc2(r.o.z), iX = bc2(r.inv.x), iY = bc2(r.inv.y), iZ = bc2(r.inv.z);
  // RayBoxIntersectionLite2 (ctrace.h:32-53): t = invDir*(plane - pos)
  float nx0, nx1, nx2, nx3, ny0, ny1, ny2, ny3, nz0, nz1, nz2, nz3, fx0, fx1, fx2, fx3, fy0, fy1, fy2, fy3, fz0, fz1, fz2, fz3;
  upk2(mul2(iX, sub2(lo2(NX), oX)), nx0, nx1); upk2(mul2(iX, sub2(hi2(NX), oX)), nx2, nx3);
  upk2(mul2(iX, sub2(lo2(FX), oX)), fx0, fx1); upk2(mul2(iX, sub2(hi2(FX), oX)), fx2, fx3);
  upk2(mul2(iY, sub2(lo2(NY), oY)), ny0, ny1); upk2(mul2(iY, sub2(hi2(NY), oY)), ny2, ny3);
  upk2(mul2(iY, sub2(lo2(FY), oY)), fy0, fy1); upk2(mul2(iY, sub2(hi2(FY), oY)), fy2, fy3);
  upk2(mul2(iZ, sub2(lo2(NZ), oZ)), nz0, nz1); upk2(mul2(iZ, sub2(hi2(NZ), oZ)), nz2, nz3);
  upk2(mul2(iZ, sub2(lo2(FZ), oZ)), fz0, fz1); upk2(mul2(iZ, sub2(hi2(FZ), oZ)), fz2, fz3);
  t0 = ChildKey(nx0, ny0, nz0, fx0, fy0, fz0, r.t); t1 = ChildKey(nx1, ny1, nz1, fx1, fy1, fz1, r.t);
  t2 = ChildKey(nx2, ny2, nz2, fx2, fy2, fz2, r.t); t3 = ChildKey(nx3, ny3, nz3, fx3, fy3, fz3, r.t);
  c0 = ch.x; c1 = ch.y; c2 = ch.z; c3 = ch.w;
}

// leave the instance: the world-space ray (origin, direction, reciprocal direction) was parked in local memory at fixed slots by HC_ENTER
#define HC_LEAVE(r, saved)                                                                                  \
  {                                                                                                        \
    const uint2 a_ = saved[0], b_ = saved[1], c_ = saved[2], d_ = saved[3], e_ = saved[4];                 \
    r.o = f3(__uint_as_float(a_.x), __uint_as_float(a_.y), __uint_as_float(b_.x));                         \
    r.d = f3(__uint_as_float(b_.y), __uint_as_float(c_.x), __uint_as_float(c_.y));                         \
    r.inv = f3(__uint_as_float(d_.x), __uint_as_float(d_.y), __uint_as_float(e_.x));                       \
    SetNearRows(r); r.instId = -1;                                                                         \
  }

// pop until an entry that can still matter (entry distance <= current hit; t is the same quantity in world and object space because the
// direction is not renormalised); leave the instance when the stack has dropped below its entry level
#define HC_POP(r, stk, saved)                                                                               \
  {                                                                                                        \
    for (;;)                                                                                               \
    {                                                                                                      \
      if (r.sp == 0) { r.node = HC_NODE_SENTINEL; break; }                                                 \
      r.sp--;                                                                                              \
      const uint2 e_ = stk[r.sp];                                                                          \
      if (!(__uint_as_float(e_.y) <= r.t)) continue;                                                       \
      r.node = e_.x; break;                                                                                \
    }                                                                                                      \
    if (r.instId >= 0 && r.sp < r.instTop) HC_LEAVE(r, saved)                                              \
  }

// one interior quad: slab-test four children, sort near to far (the reference's network (0,1)(2,3) (0,2)(1,3) (1,2), ctrace.h:896-957), push
// three, descend into the nearest.  No capacity test on the pushes: hc_set_bvh rejects a tree whose worst-case stack exceeds HC_STACK_CAP
// (the reference instead silently drops children once its 80-entry stack is full, ctrace.h:959-979 - a tree that deep is refused here).
#define HC_QUAD(r, bvh, stk, saved)                                                                         \
  {                                                                                                        \
    float t0, t1, t2, t3; unsigned c0, c1, c2, c3;                                                         \
    QuadKeys(r, bvh, r.node, t0, t1, t2, t3, c0, c1, c2, c3);                                              \
    HC_CSWAP(t0, c0, t1, c1); HC_CSWAP(t2, c2, t3, c3);                                                    \
    HC_CSWAP(t0, c0, t2, c2); HC_CSWAP(t1, c1, t3, c3);                                                    \
    HC_CSWAP(t1, c1, t2, c2);                                                                              \
    if (t3 < HC_MAXFLOAT) { stk[r.sp] = make_uint2(c3, __float_as_uint(t3)); r.sp++; }                     \
    if (t2 < HC_MAXFLOAT) { stk[r.sp] = make_uint2(c2, __float_as_uint(t2)); r.sp++; }                     \
    if (t1 < HC_MAXFLOAT) { stk[r.sp] = make_uint2(c1, __float_as_uint(t1)); r.sp++; }                     \
    if (t0 < HC_MAXFLOAT) r.node = c0;                                                                     \
    else HC_POP(r, stk, saved)                                                                             \
  }

// decompressTexCoord16 (ctrace.h:316-326) and the opacity lookup of IntersectAllPrimitivesInLeafAlpha (ctrace.h:384-398) through
// sample2DLite (cfetch.h:738-760): a hit counts when max(rgb) of the opacity texel exceeds 0.5
HC_DEV float2 DecompressTexCoord16(unsigned packed)
{
  const float fx = (1.0f/65535.0f)*(float)(packed & 0x0000FFFFu), fy = (1.0f/65535.0f)*(float)((packed & 0xFFFF0000u) >> 16);
  return make_float2(2.0f*fx - 1.0f, 2.0f*fy - 1.0f);
}
HC_DEV bool AlphaPass(const HcBvh& bvh, const uint4 a, float u, float v)
{
  if (a.x == 0xFFFFFFFFu || (int)a.x <= 0) return true;                         // no opacity map on this triangle: sample2DLite -> (1,1,1)
  const float2 A = DecompressTexCoord16(a.y), B = DecompressTexCoord16(a.z), C = DecompressTexCoord16(a.w);
  const float w = 1.0f - u - v;
  const float2 tc = make_float2(w*A.x + v*B.x + u*C.x, w*A.y + v*B.y + u*C.y);
  const uint2* sp = bvh.alphaTable + a.x;                                       // SWTexSampler {flags, gamma, texId, dummy, row0, row1} as six uint2
  const uint2 s0 = __ldg(sp), s1 = __ldg(sp + 1), s2 = __ldg(sp + 2), s3 = __ldg(sp + 3), s4 = __ldg(sp + 4), s5 = __ldg(sp + 5);
  const int flags = (int)s0.x, texId = (int)s1.x; const float gamma = __uint_as_float(s0.y);
  if (texId == 0) return true;
  const float2 tct = make_float2(__uint_as_float(s2.x)*tc.x + __uint_as_float(s2.y)*tc.y + __uint_as_float(s3.y),
                                 __uint_as_float(s4.x)*tc.x + __uint_as_float(s4.y)*tc.y + __uint_as_float(s5.y));     // mul2x4, cfetch.h:640-646
  float4 c = ReadImageSw4(bvh.textures + bvh.texturesTable[texId], tct, flags, (gamma != 1.0f));
  if (flags & HC_TEX_ALPHASRC_W) { c.x = c.w; c.y = c.w; c.z = c.w; }
  return fmaxf(c.x, fmaxf(c.y, c.z)) > 0.5f;
}

// triangle leaf: IntersectAllPrimitivesInLeaf (ctrace.h:124-182), ONE pair record (two triangles, three 256-bit loads) per call
template<bool ALPHA>
HC_DEV bool PairTest(HcRayTrav& r, const HcBvh& bvh, const size_t pairIndex)
{
  const char* p = reinterpret_cast<const char*>(bvh.tris) + pairIndex*(HC_PAIR_F4*16);
  HcVec2 O, D;
  O.x = bc2(r.o.x); O.y = bc2(r.o.y); O.z = bc2(r.o.z);
  D.x = bc2(r.d.x); D.y = bc2(r.d.y); D.z = bc2(r.d.z);
  bool found = false;
  const HcF8 R0 = ldg256(p), R1 = ldg256(p + 32), R2 = ldg256(p + 64);
  HcVec2 A, E1, E2;
  A.x  = lo2(R0.a); A.y  = hi2(R0.a); A.z  = lo2(R0.b);
  E1.x = hi2(R0.b); E1.y = lo2(R1.a); E1.z = hi2(R1.a);
  E2.x = lo2(R1.b); E2.y = hi2(R1.b); E2.z = lo2(R2.a);
  const HcVec2 pvecN = cross2_xnynz(D, E2);                                   // (p.x, -p.y, -p.z)
  HcVec2 tvec; tvec.x = sub2(O.x, A.x); tvec.y = sub2(O.y, A.y); tvec.z = sub2(O.z, A.z);
  const HcVec2 qvecN = cross2_xnynz(tvec, E1);                                // (q.x, -q.y, -q.z)
  float det0, det1; upk2(dot2_xnynz(E1, pvecN), det0, det1);
  const hc_f2 invDet = pk2(1.0f/det0, 1.0f/det1);
  float v0, v1, u0, u1, t0, t1;
  upk2(mul2(dot2_xnynz(tvec, pvecN), invDet), v0, v1);
  upk2(mul2(dot2_xnynz(D, qvecN), invDet), u0, u1);
  upk2(mul2(dot2_xnynz(E2, qvecN), invDet), t0, t1);
  if (v0 > -HC_TRI_EPS && u0 > -HC_TRI_EPS && (u0 + v0 < 1.0f + HC_TRI_EPS) && t0 > 0.0f && t0 < r.t && (!ALPHA || AlphaPass(bvh, __ldg(bvh.alphaPairs + 2*pairIndex), u0, v0)))
  {
    r.t = t0; r.primId = __float_as_int(R2.a.z); r.geomId = __float_as_int(R2.b.x); r.hitInst = (r.instId == HC_INST_SINGLE) ? __float_as_int(R2.b.z) : r.instId; found = true;
  }
  if (v1 > -HC_TRI_EPS && u1 > -HC_TRI_EPS && (u1 + v1 < 1.0f + HC_TRI_EPS) && t1 > 0.0f && t1 < r.t     // sequential, like the reference loop
      && (!ALPHA || AlphaPass(bvh, __ldg(bvh.alphaPairs + 2*pairIndex + 1), u1, v1)))
  {
    r.t = t1; r.primId = __float_as_int(R2.a.w); r.geomId = __float_as_int(R2.b.y); r.hitInst = (r.instId == HC_INST_SINGLE) ? __float_as_int(R2.b.w) : r.instId; found = true;
  }
  return found;
}

// instance leaf of the top level: park the world-space ray in local memory, then move the ray into the instance's object space
// (ctrace.h:1020-1046; DON'T normalise the direction)
#define HC_ENTER(r, bvh, saved)                                                                             \
  {                                                                                                        \
    const float4* rec_ = bvh.nodes + size_t(r.node & 0x7fffffffu)*8;                                       \
    HcMat4 m_; m_.c0 = __ldg(rec_ + 0); m_.c1 = __ldg(rec_ + 1); m_.c2 = __ldg(rec_ + 2); m_.c3 = __ldg(rec_ + 3); \
    const float4 w_ = __ldg(rec_ + 4);                                                                     \
    saved[0] = make_uint2(__float_as_uint(r.o.x), __float_as_uint(r.o.y));                                 \
    saved[1] = make_uint2(__float_as_uint(r.o.z), __float_as_uint(r.d.x));                                 \
    saved[2] = make_uint2(__float_as_uint(r.d.y), __float_as_uint(r.d.z));                                 \
    saved[3] = make_uint2(__float_as_uint(r.inv.x), __float_as_uint(r.inv.y));                             \
    saved[4] = make_uint2(__float_as_uint(r.inv.z), 0u);                                                   \
    r.instId = __float_as_int(w_.y); r.instTop = r.sp;                                                     \
    r.o = mul4x3(m_, r.o); r.d = mul3x3(m_, r.d); r.inv = SafeInverse(r.d); SetNearRows(r);                \
    r.node = __float_as_uint(w_.x);                                                                        \
  }

HC_DEV bool RayIsFinite(float3 o, float3 d)
{
  return isfinite(o.x) && isfinite(o.y) && isfinite(o.z) && isfinite(d.x) && isfinite(d.y) && isfinite(d.z);
}
