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
