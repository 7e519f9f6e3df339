// hc_trace.cuh — K2 / K2s: closest-hit and any-hit traversal of the reference's flat two-level BVH4.
//
// Replaces BVH4TraversalInstKernel / BVH4TraversalInstShadowKenrel (reference hydra_drv/shaders/trace.cl:50, 309), i.e. the
// device functions BVH4InstTraverse / BVH4InstTraverseShadow (hydra_drv/ctrace.h:841-1062, 1065-1294) with the triangle test
// IntersectAllPrimitivesInLeaf (ctrace.h:124-182), the slab test RayBoxIntersectionLite2 (ctrace.h:32-53) and SafeInverse
// (cglobals.h:726-735).  Same predicates, same float operations (no FMA contraction), same near-to-far child order, so the
// closest hit is the same hit; what differs is HOW the tree is walked:
//   * stack entries carry the child's entry distance, and an entry whose distance exceeds the current hit is dropped at
//     pop time without fetching its 128-byte quad (the reference re-fetches and re-tests all four children);
//   * one thread = one ray; the 128-byte quad is fetched with eight 128-bit loads on the read-only path;
//   * persistent warps pull rays from a global counter with ballot/popc aggregation (one atomic per refill).
#pragma once
#include "hc_math.cuh"

#define HC_LEAF_BIT   0x80000000u
#define HC_MAXFLOAT   3.402823466e+38f // MAXFLOAT = FLT_MAX from <math.h> / OpenCL (the 1e37f fallback of ctrace.h:665-667 is never taken)
#define HC_TRI_EPS    1e-6f            // barycentric slack, ctrace.h:111
#define HC_STACK_CAP  64               // entries per ray; hc_set_bvh checks the tree against it

struct HcHit { float t; int primId; int instId; int geomId; };   // == Lite_Hit (cglobals.h:1248-1256)

struct HcBvh
{
  const float4* __restrict__ nodes;   // 2 float4 per node, 8 per quad
  const float4* __restrict__ tris;    // leaf header float4 + 3 float4 per triangle
};

HC_DEV float3 SafeInverse(float3 d)
{
  const float ooeps = 1.0e-36f;
  float3 r;
  r.x = 1.0f/(fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
  r.y = 1.0f/(fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
  r.z = 1.0f/(fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
  return r;
}

// one child of a quad: slab test + validity, returns entry distance or HC_MAXFLOAT when the child is not to be visited
HC_DEV float ChildEntry(float4 lo4, float4 hi4, float3 o, float3 inv, float tHit)
{
  const float lo  = inv.x*(lo4.x - o.x), hi  = inv.x*(hi4.x - o.x);
  const float lo1 = inv.y*(lo4.y - o.y), hi1 = inv.y*(hi4.y - o.y);
  const float lo2 = inv.z*(lo4.z - o.z), hi2 = inv.z*(hi4.z - o.z);
  float tmin = fminf(lo, hi), tmax = fmaxf(lo, hi);
  tmin = fmaxf(tmin, fminf(lo1, hi1)); tmax = fminf(tmax, fmaxf(lo1, hi1));
  tmin = fmaxf(tmin, fminf(lo2, hi2)); tmax = fminf(tmax, fmaxf(lo2, hi2));
  const bool valid = !((__float_as_uint(lo4.w) == 0xffffffffu) && (__float_as_uint(hi4.w) == 0xffffffffu));   // IsValidNode, cglobals.h:1321
  const bool hit   = (tmin <= tmax) && (tmax >= 0.0f) && (tmin <= tHit) && valid;                             // t_rayMin == 0
  return hit ? tmin : HC_MAXFLOAT;
}

#define HC_CSWAP(ta, ca, tb, cb) { const bool s_ = (tb < ta); const float tt_ = s_ ? tb : ta; const float tu_ = s_ ? ta : tb; \
                                   const unsigned ct_ = s_ ? cb : ca; const unsigned cu_ = s_ ? ca : cb; ta = tt_; tb = tu_; ca = ct_; cb = cu_; }

template<bool ANYHIT>
HC_DEV HcHit Traverse(const HcBvh bvh, float3 o, float3 d, const float tFar, unsigned* __restrict__ stkNode, float* __restrict__ stkT)
{
  HcHit hit; hit.t = tFar; hit.primId = -1; hit.instId = -1; hit.geomId = 0;     // Make_Lite_Hit(t, -1), cglobals.h:1258-1268
  float3 inv = SafeInverse(d);

  int      sp       = 0;
  unsigned node     = 1u;          // quad 1 = children of the root record (ctrace.h:851)
  bool     inInst   = false;       // instDeep
  int      instTop  = 0;
  int      instId   = -1;
  float3   wo = o, wd = d;         // world-space ray kept while inside an instance
  bool     done     = false;

  while (!done)
  {
    bool needPop = false;

    if (!(node & HC_LEAF_BIT))
    {
      const float4* q = bvh.nodes + size_t(node)*8;
      const float4 a0 = __ldg(q + 0), b0 = __ldg(q + 1), a1 = __ldg(q + 2), b1 = __ldg(q + 3);
      const float4 a2 = __ldg(q + 4), b2 = __ldg(q + 5), a3 = __ldg(q + 6), b3 = __ldg(q + 7);
      float t0 = ChildEntry(a0, b0, o, inv, hit.t), t1 = ChildEntry(a1, b1, o, inv, hit.t);
      float t2 = ChildEntry(a2, b2, o, inv, hit.t), t3 = ChildEntry(a3, b3, o, inv, hit.t);
      unsigned c0 = __float_as_uint(a0.w), c1 = __float_as_uint(a1.w), c2 = __float_as_uint(a2.w), c3 = __float_as_uint(a3.w);
      // 5-comparator network of ctrace.h:900-962 : (x,y)(z,w) (x,z)(y,w) (y,z)
      HC_CSWAP(t0, c0, t1, c1); HC_CSWAP(t2, c2, t3, c3);
      HC_CSWAP(t0, c0, t2, c2); HC_CSWAP(t1, c1, t3, c3);
      HC_CSWAP(t1, c1, t2, c2);
      if (t3 < HC_MAXFLOAT && sp < HC_STACK_CAP) { stkNode[sp] = c3; stkT[sp] = t3; sp++; }
      if (t2 < HC_MAXFLOAT && sp < HC_STACK_CAP) { stkNode[sp] = c2; stkT[sp] = t2; sp++; }
      if (t1 < HC_MAXFLOAT && sp < HC_STACK_CAP) { stkNode[sp] = c1; stkT[sp] = t1; sp++; }
      if (t0 < HC_MAXFLOAT) node = c0; else needPop = true;
    }
    else if (!inInst)
    {
      // instance leaf: read the record, move the ray to object space WITHOUT renormalising (ctrace.h:1019-1046)
      const float4* r = bvh.nodes + size_t(node & 0x7fffffffu)*8;
      const unsigned next = __float_as_uint(__ldg(r + 0).w);
      HcMat4 m; m.c0 = __ldg(r + 2); m.c1 = __ldg(r + 3); m.c2 = __ldg(r + 4); m.c3 = __ldg(r + 5);
      instId = __float_as_int(__ldg(r + 6).x);
      wo = o; wd = d;
      o = mul4x3(m, o); d = mul3x3(m, d); inv = SafeInverse(d);
      inInst = true; instTop = sp;
      node = next;
    }
    else
    {
      // triangle leaf (IntersectAllPrimitivesInLeaf, ctrace.h:124-182)
      const float4* tp  = bvh.tris + size_t(node & 0x7fffffffu);
      const float4 hdr  = __ldg(tp);
      const int first   = __float_as_int(hdr.x);
      const int count   = __float_as_int(hdr.y);
      const float4* tri = bvh.tris + first;
      bool found = false;
      for (int i = 0; i < count; i++, tri += 3)
      {
        const float4 A4 = __ldg(tri + 0), B4 = __ldg(tri + 1), C4 = __ldg(tri + 2);
        const float3 A = f3(A4), edge1 = f3(B4) - A, edge2 = f3(C4) - A;
        const float3 pvec = cross(d, edge2);
        const float3 tvec = o - A;
        const float3 qvec = cross(tvec, edge1);
        const float invDet = 1.0f/dot(edge1, pvec);
        const float v = dot(tvec, pvec)*invDet;
        const float u = dot(qvec, d)*invDet;
        const float t = dot(edge2, qvec)*invDet;
        if (v > -HC_TRI_EPS && u > -HC_TRI_EPS && (u + v < 1.0f + HC_TRI_EPS) && t > 0.0f && t < hit.t)
        {
          hit.t = t; hit.primId = __float_as_int(A4.w); hit.geomId = __float_as_int(B4.w); hit.instId = instId;
          found = true;
        }
      }
      if (ANYHIT && found) { done = true; break; }    // early exit, ctrace.h:1243-1246
      needPop = true;
    }

    if (needPop)
    {
      // drop entries that can no longer contain a closer hit without touching their quads
      for (;;)
      {
        if (sp == 0) { done = true; break; }
        sp--;
        node = stkNode[sp];
        const float te = stkT[sp];
        if (inInst && sp < instTop) { o = wo; d = wd; inv = SafeInverse(d); inInst = false; }   // ctrace.h:999-1005
        if (te <= hit.t) break;
      }
    }
  }
  return hit;
}
