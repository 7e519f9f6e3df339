// hc_trace2.cuh — K2 / K2s, second generation: the same traversal (BVH4InstTraverse / BVH4InstTraverseShadow, reference
// hydra_drv/ctrace.h:841-1062, 1065-1294; triangle test IntersectAllPrimitivesInLeaf ctrace.h:124-182, bit-exact, shared with
// hc_trace.cuh) re-designed around what ncu showed bounds the first kernel on B200 (profiles/r01_final_ncu_full_summary.md):
// instruction issue at 21 of 32 lanes and the L1 wavefront rate (one wavefront per lane and load instruction for divergent rays).
//
//   * quads are stored CENTRE / HALF-EXTENT: three 32-byte rows {c[4], h[4]} per axis + one row of child words (128 B, one line).
//     Three 256-bit loads (LDG.E.256) at FIXED offsets + one 128-bit load replace seven 128-bit loads through sign-selected row
//     pointers: 4 instead of 7 L1 wavefronts per lane and quad, no row-pointer registers.  near/far = tc -/+ h*|1/d| with
//     tc = (c - o)*(1/d): no near/far selection, 24 packed-FP32 instructions as before.  The box is CONSERVATIVE with respect to
//     the reference's RayBoxIntersectionLite2 (ctrace.h:32-53): h is rounded up and inflated by 2^-21 at upload, and the overlap test
//     carries a relative margin of 2^-19, so every child the reference visits is visited; triangle acceptance is untouched, hence
//     the closest hit is the reference's hit (equal-t ties aside).
//   * triangle pair records (96 B) are fetched by three 256-bit loads instead of five (+1) 128-bit loads.
//   * the traversal stack: the first HC2_SSTK entries of a ray live in SHARED memory laid out [entry][thread], so a push or a pop is
//     conflict-free whatever the lanes' stack depths (local memory costs one L1 wavefront per distinct depth in the warp, and those made up
//     a quarter of the first kernel's wavefronts); deeper entries overflow to local memory.
//   * the world-space ray is parked in local memory while the ray is inside an instance (six registers less).
//   * ray supply: every warp owns a 32-ray chunk and holds the next one, claimed one switch ahead by an atomicAdd whose result is not
//     needed until then, and whose rays were prefetched into L1 meanwhile: a refill costs no memory round trip.
#pragma once
#include "hc_trace.cuh"

#define HC_BOX_MARGIN  1.0000019073486328125f   // 1 + 2^-19

struct HcF8 { float4 a, b; };
HC_DEV HcF8 ldg256(const void* p)
{
  HcF8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
  return r;
}
HC_DEV hc_f2 fma2(hc_f2 a, hc_f2 b, hc_f2 c) { hc_f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

#ifndef HC2_CH
#define HC2_CH 1              // 1: centre / half-extent quads by three 256-bit loads; 0: the first kernel's min / max rows by seven 128-bit loads
#endif
#ifndef HC2_PAIR256
#define HC2_PAIR256 1         // 1: triangle pair records by three 256-bit loads; 0: five (+1) 128-bit loads
#endif
#ifndef HC2_CHUNK
#define HC2_CHUNK 1           // 1: chunked ray supply (claimed and prefetched ahead); 0: one atomicAdd per refill
#endif
#ifndef HC2_SAVED
#define HC2_SAVED 1           // 1: world-space ray parked in local memory inside an instance; 0: kept in registers
#endif
#ifndef HC2_POSTPONE
#define HC2_POSTPONE 0        // 1: one triangle leaf per lane is parked while the lane keeps descending (speculative traversal, Aila & Laine)
#endif
#define HC_PEND_EMPTY 0u              // no parked leaf (a leaf word always has bit 31 set)
#define HC_NODE_WAIT  0xfffffffeu     // leaf-class word: the lane must test its parked leaf before it may leave the instance
#ifndef HC2_SSTK
#define HC2_SSTK 8            // stack entries per ray held in SHARED memory ([entry][thread]: bank = lane whatever the lanes' stack depths, so a
#endif                        // push or pop is one conflict-free wavefront per half-warp; local memory costs one wavefront per distinct depth); deeper entries go to local memory

struct HcRay2
{
  float3 o, d, inv;          // current space (world, or the instance's object space)
  float  t; int primId, geomId, hitInst;
  int    instId;             // instance being traversed (-1 outside)
  int    sp, instTop;
  unsigned node;
#if HC2_POSTPONE
  unsigned pend;             // parked triangle-leaf word (cursor) or HC_PEND_EMPTY
#endif
#if !HC2_CH
  const char* nearX; const char* nearY; const char* nearZ;   // address of the near row of each axis in quad 0 (far row = near ^ 16)
#endif
#if !HC2_SAVED
  float3 wo, wd;
#endif
};

#if !HC2_CH
HC_DEV void SetNearRows2(HcRay2& r, const HcBvh& bvh)
{
  const char* base = reinterpret_cast<const char*>(bvh.nodes);
  r.nearX = base + (r.inv.x < 0.0f ? 16 : 0);
  r.nearY = base + (r.inv.y < 0.0f ? 48 : 32);
  r.nearZ = base + (r.inv.z < 0.0f ? 80 : 64);
}
#define HC_SETROWS2(r, bvh) SetNearRows2(r, bvh);
#else
#define HC_SETROWS2(r, bvh)
#endif

HC_DEV void Trav2Start(HcRay2& r, const HcBvh& bvh, float3 o, float3 d, float tFar)
{
  r.o = o; r.d = d; r.inv = SafeInverse(d);
  HC_SETROWS2(r, bvh)
  r.t = tFar; r.primId = -1; r.hitInst = -1; r.geomId = int(0xC0000000u);      // Make_Lite_Hit(t, -1), cglobals.h:1258-1266
  r.instId = -1; r.sp = 0; r.instTop = 0; r.node = 1u;
#if HC2_POSTPONE
  r.pend = HC_PEND_EMPTY;
#endif
}

// slab test of the four children of quad `node`: entry keys (MAXFLOAT = not to be visited) and child words
HC_DEV void QuadKeys2(const HcRay2& r, const HcBvh& bvh, const unsigned node, float& t0, float& t1, float& t2, float& t3,
                      unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3)
{
#if !HC2_CH
  const size_t qo = size_t(node)*128u;
  const char* ax_ = r.nearX + qo; const char* ay_ = r.nearY + qo; const char* az_ = r.nearZ + qo;
  const float4 NX = __ldg(reinterpret_cast<const float4*>(ax_)), FX = __ldg(reinterpret_cast<const float4*>(size_t(ax_) ^ 16u));
  const float4 NY = __ldg(reinterpret_cast<const float4*>(ay_)), FY = __ldg(reinterpret_cast<const float4*>(size_t(ay_) ^ 16u));
  const float4 NZ = __ldg(reinterpret_cast<const float4*>(az_)), FZ = __ldg(reinterpret_cast<const float4*>(size_t(az_) ^ 16u));
  const uint4  ch = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(bvh.nodes) + qo + 96));
  const hc_f2 oX = bc2(r.o.x), oY = bc2(r.o.y), oZ = bc2(r.o.z), iX = bc2(r.inv.x), iY = bc2(r.inv.y), iZ = bc2(r.inv.z);
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
  return;
#else
  const char* q = reinterpret_cast<const char*>(bvh.nodes) + size_t(node)*128u;
  const HcF8 X = ldg256(q), Y = ldg256(q + 32), Z = ldg256(q + 64);
  const uint4 ch = __ldg(reinterpret_cast<const uint4*>(q + 96));
  const float ax = fabsf(r.inv.x), ay = fabsf(r.inv.y), az = fabsf(r.inv.z);
  const hc_f2 oX = bc2(r.o.x), oY = bc2(r.o.y), oZ = bc2(r.o.z), iX = bc2(r.inv.x), iY = bc2(r.inv.y), iZ = bc2(r.inv.z);
  const hc_f2 pX = bc2(ax), pY = bc2(ay), pZ = bc2(az), mX = bc2(-ax), mY = bc2(-ay), mZ = bc2(-az);
  const hc_f2 tx01 = mul2(sub2(lo2(X.a), oX), iX), tx23 = mul2(sub2(hi2(X.a), oX), iX);
  const hc_f2 ty01 = mul2(sub2(lo2(Y.a), oY), iY), ty23 = mul2(sub2(hi2(Y.a), oY), iY);
  const hc_f2 tz01 = mul2(sub2(lo2(Z.a), oZ), iZ), tz23 = mul2(sub2(hi2(Z.a), oZ), iZ);
  float nx0, nx1, nx2, nx3, ny0, ny1, ny2, ny3, nz0, nz1, nz2, nz3, fx0, fx1, fx2, fx3, fy0, fy1, fy2, fy3, fz0, fz1, fz2, fz3;
  upk2(fma2(lo2(X.b), mX, tx01), nx0, nx1); upk2(fma2(hi2(X.b), mX, tx23), nx2, nx3);
  upk2(fma2(lo2(X.b), pX, tx01), fx0, fx1); upk2(fma2(hi2(X.b), pX, tx23), fx2, fx3);
  upk2(fma2(lo2(Y.b), mY, ty01), ny0, ny1); upk2(fma2(hi2(Y.b), mY, ty23), ny2, ny3);
  upk2(fma2(lo2(Y.b), pY, ty01), fy0, fy1); upk2(fma2(hi2(Y.b), pY, ty23), fy2, fy3);
  upk2(fma2(lo2(Z.b), mZ, tz01), nz0, nz1); upk2(fma2(hi2(Z.b), mZ, tz23), nz2, nz3);
  upk2(fma2(lo2(Z.b), pZ, tz01), fz0, fz1); upk2(fma2(hi2(Z.b), pZ, tz23), fz2, fz3);
  const float tHitK = r.t*HC_BOX_MARGIN;
  float m0, m1, m2, m3;
  upk2(mul2(pk2(min3f(fx0, fy0, fz0), min3f(fx1, fy1, fz1)), bc2(HC_BOX_MARGIN)), m0, m1);
  upk2(mul2(pk2(min3f(fx2, fy2, fz2), min3f(fx3, fy3, fz3)), bc2(HC_BOX_MARGIN)), m2, m3);
  const float n0 = max3f(nx0, ny0, nz0), n1 = max3f(nx1, ny1, nz1), n2 = max3f(nx2, ny2, nz2), n3 = max3f(nx3, ny3, nz3);
  t0 = (fmaxf(n0, 0.0f) <= fminf(m0, tHitK)) ? n0 : HC_MAXFLOAT;
  t1 = (fmaxf(n1, 0.0f) <= fminf(m1, tHitK)) ? n1 : HC_MAXFLOAT;
  t2 = (fmaxf(n2, 0.0f) <= fminf(m2, tHitK)) ? n2 : HC_MAXFLOAT;
  t3 = (fmaxf(n3, 0.0f) <= fminf(m3, tHitK)) ? n3 : HC_MAXFLOAT;
  c0 = ch.x; c1 = ch.y; c2 = ch.z; c3 = ch.w;
#endif
}

// stack access: the first HC2_SSTK entries of a ray live in shared memory, deeper ones in local memory.  `sstk` ([HC2_SSTK][block]) and
// `stk` (local) are names of the kernel's scope.
#if HC2_SSTK > 0
#define HC_STK_ST(i, v) { if ((i) < HC2_SSTK) sstk[(i)][threadIdx.x] = (v); else stk[(i) - HC2_SSTK] = (v); }
#define HC_STK_LD(i)    (((i) < HC2_SSTK) ? sstk[(i)][threadIdx.x] : stk[(i) - HC2_SSTK])
#else
#define HC_STK_ST(i, v) { stk[(i)] = (v); }
#define HC_STK_LD(i)    (stk[(i)])
#endif

// leave the instance: the world-space ray (origin, direction, reciprocal direction) was parked in local memory at fixed slots by HC_ENTER2
#if !HC2_SAVED
#define HC_LEAVE2(r, bvh, saved) { r.o = r.wo; r.d = r.wd; r.inv = SafeInverse(r.d); HC_SETROWS2(r, bvh) r.instId = -1; }
#else
#define HC_LEAVE2(r, bvh, saved)                                                                            \
  {                                                                                                        \
    const uint2 a_ = saved[0], b_ = saved[1], c_ = saved[2], d_ = saved[3], e_ = saved[4];                 \
    r.o = f3(__uint_as_float(a_.x), __uint_as_float(a_.y), __uint_as_float(b_.x));                         \
    r.d = f3(__uint_as_float(b_.y), __uint_as_float(c_.x), __uint_as_float(c_.y));                         \
    r.inv = f3(__uint_as_float(d_.x), __uint_as_float(d_.y), __uint_as_float(e_.x));                       \
    HC_SETROWS2(r, bvh) r.instId = -1;                                                                     \
  }
#endif

// pop until an entry that can still matter: entry distance <= current hit (with the box margin: the stored distance is ours, up to
// 2 ulp above the reference's); leave the instance when the stack has dropped below its entry level
#define HC_POP2(r, bvh, saved)                                                                                  \
  {                                                                                                        \
    const float tK_ = r.t*HC_BOX_MARGIN;                                                                   \
    for (;;)                                                                                               \
    {                                                                                                      \
      if (r.sp == 0) { r.node = HC_NODE_SENTINEL; break; }                                                 \
      r.sp--;                                                                                              \
      const uint2 e_ = HC_STK_LD(r.sp);                                                                    \
      if (!(__uint_as_float(e_.y) <= tK_)) continue;                                                       \
      r.node = e_.x; break;                                                                                \
    }                                                                                                      \
    if (r.instId >= 0 && r.sp < r.instTop) HC_LEAVE2(r, bvh, saved)                                            \
  }

#if HC2_POSTPONE
// Speculative variant.  Pop: a triangle leaf popped while nothing is parked is parked and the pop goes on; the instance is not left while a
// leaf is parked (it has to be tested with the object-space ray): the lane then waits for the next leaf step (HC_NODE_WAIT).
#undef HC_POP2
#define HC_POP2(r, bvh, saved)                                                                              \
  {                                                                                                        \
    const float tK_ = r.t*HC_BOX_MARGIN;                                                                   \
    for (;;)                                                                                               \
    {                                                                                                      \
      if (r.instId >= 0 && r.sp == r.instTop)                                                              \
      {                                                                                                    \
        if (r.pend != HC_PEND_EMPTY) { r.node = HC_NODE_WAIT; break; }                                     \
        HC_LEAVE2(r, bvh, saved)                                                                           \
      }                                                                                                    \
      if (r.sp == 0) { r.node = HC_NODE_SENTINEL; break; }                                                 \
      r.sp--;                                                                                              \
      const uint2 e_ = HC_STK_LD(r.sp);                                                                    \
      if (!(__uint_as_float(e_.y) <= tK_)) continue;                                                       \
      if ((e_.x & HC_LEAF_BIT) && r.instId >= 0 && r.pend == HC_PEND_EMPTY) { r.pend = e_.x; continue; }   \
      r.node = e_.x; break;                                                                                \
    }                                                                                                      \
  }
// Quad: when the nearest child is a triangle leaf and nothing is parked, it is parked and the second nearest child becomes the next node
// (the sorted list is rotated in registers, no stack round trip).
#define HC_QUAD2(r, bvh, saved)                                                                             \
  {                                                                                                        \
    float t0, t1, t2, t3; unsigned c0, c1, c2, c3;                                                         \
    QuadKeys2(r, bvh, r.node, t0, t1, t2, t3, c0, c1, c2, c3);                                             \
    HC_CSWAP(t0, c0, t1, c1); HC_CSWAP(t2, c2, t3, c3);                                                    \
    HC_CSWAP(t0, c0, t2, c2); HC_CSWAP(t1, c1, t3, c3);                                                    \
    HC_CSWAP(t1, c1, t2, c2);                                                                              \
    if (t0 < HC_MAXFLOAT && (c0 & HC_LEAF_BIT) && r.instId >= 0 && r.pend == HC_PEND_EMPTY)                \
    { r.pend = c0; t0 = t1; c0 = c1; t1 = t2; c1 = c2; t2 = t3; c2 = c3; t3 = HC_MAXFLOAT; }               \
    if (t3 < HC_MAXFLOAT) { HC_STK_ST(r.sp, make_uint2(c3, __float_as_uint(t3))) r.sp++; }                 \
    if (t2 < HC_MAXFLOAT) { HC_STK_ST(r.sp, make_uint2(c2, __float_as_uint(t2))) r.sp++; }                 \
    if (t1 < HC_MAXFLOAT) { HC_STK_ST(r.sp, make_uint2(c1, __float_as_uint(t1))) r.sp++; }                 \
    if (t0 < HC_MAXFLOAT) r.node = c0;                                                                     \
    else HC_POP2(r, bvh, saved)                                                                            \
  }
#else
// one interior quad: slab-test four children, sort near to far (the reference's network, ctrace.h:896-957), push three, descend into
// the nearest
#define HC_QUAD2(r, bvh, saved)                                                                             \
  {                                                                                                        \
    float t0, t1, t2, t3; unsigned c0, c1, c2, c3;                                                         \
    QuadKeys2(r, bvh, r.node, t0, t1, t2, t3, c0, c1, c2, c3);                                             \
    HC_CSWAP(t0, c0, t1, c1); HC_CSWAP(t2, c2, t3, c3);                                                    \
    HC_CSWAP(t0, c0, t2, c2); HC_CSWAP(t1, c1, t3, c3);                                                    \
    HC_CSWAP(t1, c1, t2, c2);                                                                              \
    if (t3 < HC_MAXFLOAT) { HC_STK_ST(r.sp, make_uint2(c3, __float_as_uint(t3))) r.sp++; }                 \
    if (t2 < HC_MAXFLOAT) { HC_STK_ST(r.sp, make_uint2(c2, __float_as_uint(t2))) r.sp++; }                 \
    if (t1 < HC_MAXFLOAT) { HC_STK_ST(r.sp, make_uint2(c1, __float_as_uint(t1))) r.sp++; }                 \
    if (t0 < HC_MAXFLOAT) r.node = c0;                                                                     \
    else HC_POP2(r, bvh, saved)                                                                            \
  }
#endif

// triangle pair record by three 256-bit loads; arithmetic identical to PairTest (hc_trace.cuh)
template<bool ALPHA>
HC_DEV bool PairTest2(HcRay2& r, const HcBvh& bvh, const size_t pairIndex)
{
  const char* p = reinterpret_cast<const char*>(bvh.tris) + pairIndex*(HC_PAIR_F4*16);
  HcVec2 O, D;
  O.x = bc2(r.o.x); O.y = bc2(r.o.y); O.z = bc2(r.o.z);
  D.x = bc2(r.d.x); D.y = bc2(r.d.y); D.z = bc2(r.d.z);
  bool found = false;
#if HC2_PAIR256
  const HcF8 R0 = ldg256(p), R1 = ldg256(p + 32), R2 = ldg256(p + 64);
#else
  HcF8 R0, R1, R2;
  R0.a = __ldg(reinterpret_cast<const float4*>(p)); R0.b = __ldg(reinterpret_cast<const float4*>(p) + 1); R1.a = __ldg(reinterpret_cast<const float4*>(p) + 2);
  R1.b = __ldg(reinterpret_cast<const float4*>(p) + 3); R2.a = __ldg(reinterpret_cast<const float4*>(p) + 4); R2.b = __ldg(reinterpret_cast<const float4*>(p) + 5);
#endif
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
    r.t = t0; r.primId = __float_as_int(R2.a.z); r.geomId = __float_as_int(R2.b.x); r.hitInst = r.instId; found = true;
  }
  if (v1 > -HC_TRI_EPS && u1 > -HC_TRI_EPS && (u1 + v1 < 1.0f + HC_TRI_EPS) && t1 > 0.0f && t1 < r.t     // sequential, like the reference loop
      && (!ALPHA || AlphaPass(bvh, __ldg(bvh.alphaPairs + 2*pairIndex + 1), u1, v1)))
  {
    r.t = t1; r.primId = __float_as_int(R2.a.w); r.geomId = __float_as_int(R2.b.y); r.hitInst = r.instId; found = true;
  }
  return found;
}

#if HC2_SAVED
#define HC_SAVE2(r, saved)                                                                                  \
    saved[0] = make_uint2(__float_as_uint(r.o.x), __float_as_uint(r.o.y));                                 \
    saved[1] = make_uint2(__float_as_uint(r.o.z), __float_as_uint(r.d.x));                                 \
    saved[2] = make_uint2(__float_as_uint(r.d.y), __float_as_uint(r.d.z));                                 \
    saved[3] = make_uint2(__float_as_uint(r.inv.x), __float_as_uint(r.inv.y));                             \
    saved[4] = make_uint2(__float_as_uint(r.inv.z), 0u);
#else
#define HC_SAVE2(r, saved) r.wo = r.o; r.wd = r.d;
#endif
// instance leaf of the top level: park the world-space ray in local memory, then move the ray into the instance's object space
// (ctrace.h:1020-1046; the direction is NOT normalised, so t means the same in both spaces)
#define HC_ENTER2(r, bvh, saved)                                                                            \
  {                                                                                                        \
    const float4* rec_ = bvh.nodes + size_t(r.node & 0x7fffffffu)*8;                                       \
    HcMat4 m_; m_.c0 = __ldg(rec_ + 0); m_.c1 = __ldg(rec_ + 1); m_.c2 = __ldg(rec_ + 2); m_.c3 = __ldg(rec_ + 3); \
    const float4 w_ = __ldg(rec_ + 4);                                                                     \
    HC_SAVE2(r, saved)                                                                                     \
    r.instId = __float_as_int(w_.y); r.instTop = r.sp;                                                     \
    r.o = mul4x3(m_, r.o); r.d = mul3x3(m_, r.d); r.inv = SafeInverse(r.d); HC_SETROWS2(r, bvh)            \
    r.node = __float_as_uint(w_.x);                                                                        \
  }
