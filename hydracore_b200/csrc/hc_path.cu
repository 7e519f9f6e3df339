// hc_path.cu — wavefront path tracing: K1 (eye paths), K3+K4+K5 (one fused shade kernel per bounce), K6 (warp-aggregated
// compaction inside the shade kernel), K7 (per-pixel HDR accumulation), and the C-ABI entry points that drive them.
//
// What it reproduces: the CPU oracle integrators IntegratorMISPTLoop2::PathTrace (reference hydra_drv/CPUExp_Integrators_PT_Loop.cpp:
// 264-321, the one CPUExpLayer instantiates), IntegratorStupidPT::PathTrace (CPUExp_Integrators_PT.cpp:9-38) and
// IntegratorMISPT_QMC::DoPass (CPUExp_Integrators_PT_QMC.cpp:5-49) under IntegratorCommon::DoPass (CPUExp_Integrators_Common.cpp:278-316),
// with the per-pixel generator rule of InitRandomGen (shaders/trace.cl:6-13): gen[p] = RandomGenInit(seed + p), carried across passes.
//
// How it differs from the OpenCL wavefront (GPUOCLLayerCore.cpp:9-130: 6-8 launches per bounce, ~650 B of state streamed per path per
// bounce, no compaction): per bounce there are three launches — closest-hit traversal, any-hit traversal of the previous bounce's
// shadow rays, and ONE shade kernel that evaluates the surface, emission/MIS, samples the light, evaluates the BSDF for it, samples the
// next direction, updates the path and writes the survivors densely into the other half of a double-buffered SoA state (ballot/popc
// ranks, one atomic per warp).  The unshadowed direct-light term travels with the shadow ray and is added when visibility is known.
// Nothing is read back between bounces: every kernel takes its element count from device memory.
#include "hc_context.h"
#include "hc_shade.cuh"
#include "hc_raygen.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <chrono>
#include <string>

#define HC_SHADE_BLOCK 128
#ifndef HC_SHADE_MINB
#define HC_SHADE_MINB 5      // CTAs per SM: 96 registers. Swept again after the code-size cut (C3 / C4 shade ms): 3 -> 1.81 / 1.19, 4 -> 1.55 / 1.04, 5 -> 1.51 / 1.03, 6 -> 1.49 / 1.06
#endif

// One half of the double-buffered path queue: 128 bytes per path in FOUR arrays of 32-byte elements (one L2 / DRAM sector each):
//   A  rpos.xyz, bsdf pdf      | rdir.xyz, flags            the ray of K2 (one 256-bit load, interleaved {pos, dir} with stride 2)
//   B  throughput.xyz, pixel | specular bit | accum.xyz, qmc position
//   C  spos.xyz, sexp.x        | sdir.xyz, t_far            the shadow ray of K2s (it zeroes t_far when the ray is occluded)
//   D  sexp.y, sexp.z, rng.x, rng.y | hit record            (the hit record is written by K2)
// Unpermuted, the 32 lanes of a warp read 1 KB contiguous per array; through the material-sort permutation every gathered element is exactly
// one sector.  Round 1 kept nine arrays of 16-, 8- and 4-byte elements: the permuted gather pulled 32-byte sectors for 16-byte elements and
// read 2.2 x the bytes it needed (profiles/r01_final_ncu_full_summary.md); one 128-byte record per path (tried first in round 2) reads no
// spare byte either but turns every load instruction into 32 separate L1 wavefronts: C3 shade 1.49 -> 1.61 ms, K2 3.37 -> 3.50 ms.
struct HcPathState { float4* a; float4* b; float4* c; float4* d; };
struct HcF8p { float4 x, y; };
HC_DEV HcF8p LoadPair(const float4* p)
{
  HcF8p r;
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.x.x), "=f"(r.x.y), "=f"(r.x.z), "=f"(r.x.w), "=f"(r.y.x), "=f"(r.y.y), "=f"(r.y.z), "=f"(r.y.w) : "l"(p));
  return r;
}
HC_DEV void StorePair(float4* p, float4 x, float4 y)
{
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w), "f"(y.x), "f"(y.y), "f"(y.z), "f"(y.w) : "memory");
}

struct HcPassParams
{
  int integrator;           // HC_INTEGRATOR_*
  int width, height;
  int maxDepth;             // m_maxDepth of the oracle integrator
  int depth;                // bounce index of this shade launch
  int isLast;               // last bounce of the pass
  unsigned qmcPass;         // passes done so far (QMC sample index = pass*W*H + i)
  int world, rank;
  // sample streams (hc_pt_set_sample_streams): S generators per pixel, pass p draws from stream p mod S, so that up to S consecutive passes of
  // a pixel are independent of each other and can be in flight together as one wavefront (sub-pass = path index / owned pixels)
  int streams, streamBase;  // S; stream of sub-pass 0 of this wavefront (= first pass index mod S)
  int groupPasses;          // sub-passes in this wavefront (1: paths add straight into the frame buffer)
  int nOwned;               // paths per sub-pass
  unsigned pixMask; int subShift;   // path word = pixel | sub-pass << subShift | specular bit 31
};

// ------------------------------------------------------------------------------------------------------------------ K1: eye paths
// PT/MISPT: IntegratorCommon::makeEyeRay (CPUExp_Integrators_Common.cpp:347-359): rndUniform(-1, 1) -> MakeRandEyeRay.
// QMC     : rndLens with the Niederreiter table (crandom.h:369-391) -> MakeEyeRayFromF4Rnd, pixel = (int)fx, (int)fy.
__global__ void __launch_bounds__(256)
k_pt_generate(const HcCamera cam, const HcPassParams pp, const int n, const int first, const int* __restrict__ ownedPixels, uint2* __restrict__ pixelRng,
              const int* __restrict__ rmQMC, const unsigned* __restrict__ qmcTable, HcPathState st, int* __restrict__ pathCount)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i == 0) pathCount[0] = n;
  if (i >= n) return;
  float3 rpos, rdir; int pixel; unsigned qpos = 0xFFFFFFFFu, subBits = 0u;
  HcRng g;
  if (pp.integrator == HC_INTEGRATOR_MISPT_QMC)
  {
    const int sub = (pp.groupPasses > 1) ? i / pp.nOwned : 0;     // sample streams: sub-pass major, as for PT / MISPT
    const int sample = (i - sub*pp.nOwned)*pp.world + pp.rank;    // this GPU's share of the sample indices of the pass
    const uint2 s2 = pixelRng[size_t((pp.streamBase + sub) % pp.streams)*size_t(pp.width*pp.height) + sample]; g.x = s2.x; g.y = s2.y;
    qpos = (pp.qmcPass + (unsigned)sub)*(unsigned)(pp.width*pp.height) + (unsigned)sample;
    subBits = (unsigned)sub << pp.subShift;
    float4 lens;
    lens.y = rndQmcTab(g, rmQMC, qpos, HC_QMC_VAR_SCR_Y, qmcTable);
    lens.z = rndQmcTab(g, rmQMC, qpos, HC_QMC_VAR_DOF_X, qmcTable);
    lens.w = rndQmcTab(g, rmQMC, qpos, HC_QMC_VAR_DOF_Y, qmcTable);
    lens.x = rndQmcTab(g, rmQMC, qpos, HC_QMC_VAR_SCR_X, qmcTable);
    float fx, fy;
    MakeEyeRayFromF4Rnd(lens, cam, rpos, rdir, fx, fy);
    int x = (int)fx, y = (int)fy;
    if (x >= pp.width) x = pp.width - 1;
    if (y >= pp.height) y = pp.height - 1;
    if (x < 0) x = 0;
    if (y < 0) y = 0;
    pixel = y*pp.width + x;
    // the generator of slot `sample` is written back when the path ends; remember the slot in qpos' companion: thr.w keeps the PIXEL,
    // the slot is recoverable from qpos (qpos - pass*W*H)
  }
  else
  {
    const int gi = first + i;                                       // index in the wavefront (this launch may be one half of it): sub-pass major
    const int sub = (pp.groupPasses > 1) ? gi / pp.nOwned : 0;
    pixel = ownedPixels[gi - sub*pp.nOwned];
    const uint2 s2 = pixelRng[size_t((pp.streamBase + sub) % pp.streams)*size_t(pp.width*pp.height) + pixel]; g.x = s2.x; g.y = s2.y;
    subBits = (unsigned)sub << pp.subShift;
    const float4 r = rndFloat4_Pseudo(g);
    const float4 offs = make_float4(-1.0f + 2.0f*r.x, -1.0f + 2.0f*r.y, -1.0f + 2.0f*r.z, -1.0f + 2.0f*r.w);   // rndUniform(gen, -1, 1), crandom.h:617-620
    MakeRandEyeRay(pixel % pp.width, pixel / pp.width, pp.width, pp.height, offs, cam, rpos, rdir);
  }
  StorePair(st.a + 2*size_t(i), make_float4(rpos.x, rpos.y, rpos.z, 1.0f),                                   // makeInitialMisData: matSamplePdf = 1
                                make_float4(rdir.x, rdir.y, rdir.z, __uint_as_float(0u)));                   // flags = 0
  StorePair(st.b + 2*size_t(i), make_float4(1.0f, 1.0f, 1.0f, __uint_as_float((unsigned)pixel | subBits | 0x80000000u)),  // isSpecular = 1
                                make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(qpos)));
  StorePair(st.c + 2*size_t(i), make_float4(0.0f, 0.0f, 0.0f, 0.0f), make_float4(0.0f, 1.0f, 0.0f, 0.0f));     // no pending shadow ray
  st.d[2*size_t(i)] = make_float4(0.0f, 0.0f, __uint_as_float(g.x), __uint_as_float(g.y));
}

// ------------------------------------------------------------------------------------------------------------------ K3+K4+K5+K6+K7: shade
HC_DEV void FinishPath(float4* __restrict__ fb, float4* __restrict__ subSums, uint2* __restrict__ pixelRng, size_t rngSlot, size_t subSlot, int pixel, float3 accum, const HcRng& g)
{
  // K7: per-pixel HDR SUM (the GPU layer keeps sums and divides by spp on read-back, GPUOCLLayer.cpp:1184-1215)
  if (subSums != nullptr)
    subSums[subSlot] = make_float4(accum.x, accum.y, accum.z, 0.0f);     // several passes of the pixel in flight: k_fb_fold adds them in pass order
  else
  {
    float* p = reinterpret_cast<float*>(fb + pixel);
    atomicAdd(p + 0, accum.x); atomicAdd(p + 1, accum.y); atomicAdd(p + 2, accum.z);
  }
  pixelRng[rngSlot] = make_uint2(g.x, g.y);
}

template<bool NMAP>
__global__ void __launch_bounds__(HC_SHADE_BLOCK, HC_SHADE_MINB)
k_pt_shade(const HcScene s, const HcPassParams pp, const int* __restrict__ nIn, int* __restrict__ nOut,
           const HcPathState in, HcPathState out,
           const unsigned* __restrict__ qmcTable, float4* __restrict__ fb, float4* __restrict__ subSums, uint2* __restrict__ pixelRng, const int* __restrict__ perm)
{
  const int tid = blockIdx.x*blockDim.x + threadIdx.x;
  const int n = *nIn;
  const int i = (perm != nullptr && tid < n) ? perm[tid] : tid;       // material-sorted order (k_pt_sort_*), else queue order
  bool alive = false;

  // new state of a surviving path
  float3 nPos = f3(0, 0, 0), nDir = f3(0, 0, 0), thr = f3(0, 0, 0), accum = f3(0, 0, 0);
  float nPdf = 1.0f; unsigned nFlags = 0, pixSpec = 0, qpos = 0xFFFFFFFFu;
  float4 sPos = make_float4(0, 0, 0, 0), sDir = make_float4(0, 1, 0, 0), sExp = make_float4(0, 0, 0, 0);
  HcRng g; g.x = 0; g.y = 0;

  if (tid < n)
  {
    const HcF8p pa = LoadPair(in.a + 2*size_t(i)), pb = LoadPair(in.b + 2*size_t(i)), pc = LoadPair(in.c + 2*size_t(i)), pd = LoadPair(in.d + 2*size_t(i));
    const float4 rp = pa.x, rd = pa.y, th = pb.x, ac = pb.y, r4 = pc.x, r5 = pc.y, r6 = pd.x, r7 = pd.y;
    g.x = __float_as_uint(r6.z); g.y = __float_as_uint(r6.w);
    qpos = __float_as_uint(ac.w);
    const float3 rayPos = f3(rp), rayDir = f3(rd);
    const unsigned flags = __float_as_uint(rd.w);
    pixSpec = __float_as_uint(th.w);
    const int pixel = int(pixSpec & pp.pixMask);
    const int sub = int((pixSpec & 0x7FFFFFFFu) >> pp.subShift);
    const bool prevSpecular = (pixSpec & 0x80000000u) != 0;
    const float prevPdf = rp.w;
    const size_t frame = size_t(pp.width*pp.height);
    const size_t rngSlot = size_t((pp.streamBase + sub) % pp.streams)*frame +                                 // QMC: one generator per SAMPLE index of the stream
                           ((pp.integrator == HC_INTEGRATOR_MISPT_QMC) ? size_t(qpos - (pp.qmcPass + (unsigned)sub)*(unsigned)(pp.width*pp.height)) : size_t(pixel));
    thr = f3(th); accum = f3(ac);

    // pending direct light of the previous bounce: accumColor += accumuThoroughput*explicitColor (PT_Loop.cpp:253), shadow in {0,1}
    if (r5.w > 0.0f) accum += f3(r4.w, r6.x, r6.y);                        // K2s zeroed t_far when the shadow ray was occluded

    HcHit hit; hit.t = r7.x; hit.primId = __float_as_int(r7.y); hit.instId = __float_as_int(r7.z); hit.geomId = __float_as_int(r7.w);
    const bool isPT = (pp.integrator == HC_INTEGRATOR_PT);
    float3 curr = f3(0, 0, 0);
    bool finished = false;

    if (hit.primId == -1 || !isfinite(hit.t))                     // HitNone (cglobals.h:1270) -> environmentColor (black without a sky light)
    {
      // IntegratorStupidPT: the initial MisData at depth 0, a value-initialised MisData() (pdf 0, not specular, material offset 0) deeper
      // (CPUExp_Integrators_PT.cpp:9-38); the MISPT loop carries the real MisData with prevMaterialOffset = -1
      if (isPT) curr = (pp.depth == 0) ? EnvironmentColor(s, rayDir, 1.0f, true, -1, flags) : EnvironmentColor(s, rayDir, 0.0f, false, 0, flags);
      else      curr = EnvironmentColor(s, rayDir, prevPdf, prevSpecular, -1, flags);
      finished = true;
    }
    else
    {
      const HcSurfaceHit sh = SurfaceEval(s, rayPos, rayDir, hit);
      const float3 e = EmissionEval(s, rayPos, rayDir, sh, flags, hit.instId);
      if (dot(e, e) > (isPT ? 1e-6f : 1e-3f))
      {
        if (isPT) curr = e;
        else
        {
          const float* L = (s.lightsNum != 0) ? LightAt(s, s.instLightIds[hit.instId]) : nullptr;
          if (L != nullptr)
          {                                                       // kernel_EvalEmission (PT_Loop.cpp:86-139)
            const float lgtPdf = L[HC_PLIGHT_PICK_PROB_REV]*LightEvalPDF(L, rayPos, rayDir, sh.pos, sh.normal, sh.texCoord, s);
            float w = misWeightHeuristic(prevPdf, lgtPdf);
            if (prevSpecular) w = 1.0f;
            curr = e*w;
          }
          else curr = e;
        }
        finished = true;
      }
      else if (!isPT && pp.depth >= pp.maxDepth - 1)
        finished = true;                                          // curr = 0
      else
      {
        const float* mat = MaterialAt(s, sh.matId);
        float3 explicitColor = f3(0, 0, 0);
        if (!isPT)
        {
          // kernel_LightSelect / LightSample / Shade (PT_Loop.cpp:141-216); .z picks the light AND is the third sample coordinate
          const float4 rl = rndFloat4_Pseudo(g);
          float pick = 1.0f;
          const int lightOffset = SelectRandomLightRev(rl.z, s, pick);
          if (lightOffset >= 0)
          {
            const float* L = LightAt(s, lightOffset);
            HcShadowSample sam;
            LightSampleRev(L, f3(rl.x, rl.y, rl.z), sh.pos, s, sam);
            const float3 sdir = normalize(sam.pos - sh.pos);
            const float3 spos = OffsShadowRayPos(sh.pos, sh.normal, sdir, sh.sRayOff);
            const float tFar = length(spos - sam.pos)*0.995f;
            const HcBxDF ev = MaterialEval<NMAP>(mat, sdir, (-1.0f)*rayDir, sh, s);
            const float c1 = fmaxf(+dot(sdir, sh.normal), 0.0f), c2 = fmaxf(-dot(sdir, sh.normal), 0.0f);
            const float3 bx = (ev.brdf*c1 + ev.btdf*c2);
            const float lgtPdf = sam.pdf*pick;
            float w = misWeightHeuristic(lgtPdf, ev.pdfFwd);
            if (sam.isPoint) w = 1.0f;
            explicitColor = (1.0f/pick)*(sam.color*(1.0f/fmaxf(sam.pdf, HC_DEPSILON2)))*bx*w;
            sPos = make_float4(spos.x, spos.y, spos.z, 0.0f);
            sDir = make_float4(sdir.x, sdir.y, sdir.z, tFar);
          }
        }

        // kernel_NextBounce (PT_Loop.cpp:218-256) / sampleAndEvalBxDF (CPUExp_Integrators_Common.cpp:528-556): RndMatAll then sample
        const unsigned sampFlags = isPT ? 0u : flags;             // IntegratorStupidPT calls sampleAndEvalBxDF with its default flags = 0
        const int bounceNum = int((sampFlags & 0x0000FF00u) >> 8);
        const bool qmc = (pp.integrator == HC_INTEGRATOR_MISPT_QMC) && bounceNum == 0;
        float rands[HC_MMLT_FLOATS_PER_BOUNCE];
        if (qmc)
        {
          rands[0] = rndQmcTab(g, s.globals + HC_EG_rmQMC/4, qpos, HC_QMC_VAR_MAT_0, qmcTable);
          rands[1] = rndQmcTab(g, s.globals + HC_EG_rmQMC/4, qpos, HC_QMC_VAR_MAT_1, qmcTable);
          rands[2] = rndFloat1_Pseudo(g);
          for (int k = 0; k < HC_MMLT_FLOATS_PER_MLAYER; k++) rands[3 + k] = rndQmcTab(g, s.globals + HC_EG_rmQMC/4, qpos, HC_QMC_VAR_MAT_L, qmcTable);
        }
        else
        {
          const float4 rm = rndFloat4_Pseudo(g);
          rands[0] = rm.x; rands[1] = rm.y; rands[2] = rm.z;
          for (int k = 0; k < HC_MMLT_FLOATS_PER_MLAYER; k++) rands[3 + k] = rndFloat1_Pseudo(g);
        }
        HcMatSample ms;
        MaterialSampleAndEval<NMAP>(mat, rands, sh, rayDir, sampFlags, s, ms);
        const float3 bxdfVal = ms.color*(1.0f/fmaxf(ms.pdf, isPT ? HC_DEPSILON2 : 1e-20f));
        const float cosTheta = fabsf(dot(ms.direction, sh.normal));

        sExp = make_float4(thr.x*explicitColor.x, thr.y*explicitColor.y, thr.z*explicitColor.z, 0.0f);
        thr *= cosTheta*bxdfVal;
        nDir = ms.direction;
        nPos = OffsRayPos(sh.pos, sh.normal, ms.direction);
        nPdf = ms.pdf;
        nFlags = FlagsNextBounceLite(flags, ms, s);
        const bool spec = (ms.flags & HC_RAY_EVENT_S) != 0 || (ms.flags & HC_RAY_EVENT_T) != 0;
        pixSpec = (pixSpec & 0x7FFFFFFFu) | (spec ? 0x80000000u : 0u);
        if (pp.isLast) finished = true;                           // PT: the recursion returns 0 one level deeper; draws are already made
        else alive = true;
      }
    }

    if (finished)
    {
      accum += thr*curr;                                          // kernel_AddLastBouceContrib (PT_Loop.cpp:258-262)
      FinishPath(fb, subSums, pixelRng, rngSlot, size_t(sub)*frame + size_t(pixel), pixel, accum, g);
    }
  }

  // K6: warp-aggregated compaction — survivors of a warp take consecutive slots, one atomic per warp
  const unsigned mask = __ballot_sync(0xffffffffu, alive);
  if (mask == 0u) return;
  const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(nOut, __popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (alive)
  {
    const int j = base + __popc(mask & ((1u << lane) - 1u));
    StorePair(out.a + 2*size_t(j), make_float4(nPos.x, nPos.y, nPos.z, nPdf), make_float4(nDir.x, nDir.y, nDir.z, __uint_as_float(nFlags)));
    StorePair(out.b + 2*size_t(j), make_float4(thr.x, thr.y, thr.z, __uint_as_float(pixSpec)), make_float4(accum.x, accum.y, accum.z, __uint_as_float(qpos)));
    StorePair(out.c + 2*size_t(j), make_float4(sPos.x, sPos.y, sPos.z, sExp.x), sDir);
    out.d[2*size_t(j)] = make_float4(sExp.y, sExp.z, __uint_as_float(g.x), __uint_as_float(g.y));
  }
}

// ------------------------------------------------------------------------------------------------------------------ K6b: material sort
// Replaces the role of bitonic_sort_gpu (reference hydra_drv/bitonic_sort_gpu.cpp:90-158, shaders/sort.cl): the live-path queue is put
// in MATERIAL order before shading, so that the lanes of a warp run the same BSDF code and touch the same material record.  Keys are
// small integers (material id, or one extra bucket for "missed"), so this is a counting sort in three short launches — per-CTA
// shared-memory histogram, one-CTA scan, warp-aggregated scatter of path indices — instead of O(N log^2 N) bitonic passes.  The order
// inside a bucket is arbitrary: every path carries its own pixel and generator, so the image does not depend on it.
#define HC_SORT_MAX_KEYS 2048
#define HC_SORT_BLOCK    256

HC_DEV int SortKeyOf(const HcScene& s, const HcHit& h, int numKeys)
{
  if (h.primId == -1 || !isfinite(h.t)) return numKeys - 1;                    // HitNone -> last bucket
  const int meshOff = s.globals[s.geometryTableOffset + h.geomId];
  const float4* mesh = s.geom + meshOff;
  const int mIdxOff = reinterpret_cast<const int4*>(mesh)[2].x;                // PlainMesh::mIndicesOffset (cfetch.h:1038-1059)
  const int matId = RemapMaterialId(s, reinterpret_cast<const int*>(mesh + mIdxOff)[h.primId], h.instId);
  return min(max(matId, 0), numKeys - 2);
}

__global__ void __launch_bounds__(HC_SORT_BLOCK)
k_pt_sort_count(const HcScene s, const int* __restrict__ nIn, const float4* __restrict__ recD, unsigned short* __restrict__ keys,
                int* __restrict__ bucketCount, const int numKeys)
{
  extern __shared__ int sCount[];
  for (int k = threadIdx.x; k < numKeys; k += blockDim.x) sCount[k] = 0;
  __syncthreads();
  const int n = *nIn;
  for (int i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x)
  {
    const float4 h4 = recD[2*size_t(i) + 1];
    HcHit h; h.t = h4.x; h.primId = __float_as_int(h4.y); h.instId = __float_as_int(h4.z); h.geomId = __float_as_int(h4.w);
    const int key = SortKeyOf(s, h, numKeys);
    keys[i] = (unsigned short)key;
    atomicAdd(&sCount[key], 1);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < numKeys; k += blockDim.x) if (sCount[k] != 0) atomicAdd(&bucketCount[k], sCount[k]);
}

__global__ void __launch_bounds__(1024)
k_pt_sort_scan(int* __restrict__ bucketCount, int* __restrict__ bucketCursor, const int numKeys)
{
  // exclusive prefix sum of <= 2048 counters by one CTA (two per thread); the counters are cleared for the next bounce
  __shared__ int sh[HC_SORT_MAX_KEYS];
  const int t = threadIdx.x;
  const int a = (2*t < numKeys) ? bucketCount[2*t] : 0, b = (2*t + 1 < numKeys) ? bucketCount[2*t + 1] : 0;
  sh[t] = a + b;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1)
  {
    const int v = (t >= off) ? sh[t - off] : 0;
    __syncthreads();
    sh[t] += v;
    __syncthreads();
  }
  const int excl = sh[t] - (a + b);
  if (2*t < numKeys)     { bucketCursor[2*t] = excl;         bucketCount[2*t] = 0; }
  if (2*t + 1 < numKeys) { bucketCursor[2*t + 1] = excl + a; bucketCount[2*t + 1] = 0; }
}

__global__ void __launch_bounds__(HC_SORT_BLOCK)
k_pt_sort_scatter(const int* __restrict__ nIn, const unsigned short* __restrict__ keys, int* __restrict__ bucketCursor, int* __restrict__ perm)
{
  const int n = *nIn;
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int key = keys[i];
  // lanes of the warp with the same key take consecutive slots of that bucket: one atomic per (warp, key)
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, key);
  const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(&bucketCursor[key], __popc(peers));
  base = __shfl_sync(peers, base, leader);
  perm[base + __popc(peers & ((1u << lane) - 1u))] = i;
}

// K2 / K2s wrappers reading the ray count from device memory live in hc_api.cu (hc_launch_trace_counted)

// init pixel generators: InitRandomGen (shaders/trace.cl:6-13) with tid = pixel index
__global__ void k_init_rng(uint2* __restrict__ gen, int n, int seed)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  const HcRng g = RandomGenInit(seed + i);
  gen[i] = make_uint2(g.x, g.y);
}

// GetLDRImage: clamp to 1 (ToneMapping4, cglobals.h:698), pow(x, 1/gamma) as the GPU layer does (shaders/screen.cl:457-460), RealColorToUint32 (cglobals.h:711-724)
__global__ void k_hdr_to_ldr(const float4* __restrict__ fb, unsigned* __restrict__ out, int n, float invSpp, float invGamma)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 c = fb[i]*invSpp;
  c.x = powf(fminf(c.x, 1.0f), invGamma); c.y = powf(fminf(c.y, 1.0f), invGamma); c.z = powf(fminf(c.z, 1.0f), invGamma); c.w = fminf(c.w, 1.0f);
  const unsigned r = (unsigned char)(c.x*255.0f), g = (unsigned char)(c.y*255.0f), b = (unsigned char)(c.z*255.0f), a = (unsigned char)(c.w*255.0f);
  out[i] = r | (g << 8) | (b << 16) | (a << 24);
}

// GetHDRImage: the image keeps SUMS, the read-back divides by the sample count (GPUOCLLayer.cpp:1184-1215) - on the device, into a staging
// buffer that is then copied out, instead of a scalar loop over 8.3 M (1080p) / 33 M (4K) floats on the host
__global__ void k_fb_normalize(const float4* __restrict__ fb, float4* __restrict__ out, int n, float invSpp)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) out[i] = fb[i]*invSpp;
}

// ------------------------------------------------------------------------------------------------------------------ host side
// several passes of a pixel in flight (sample streams): their path sums were stored per sub-pass; add them to the frame buffer in PASS order,
// which is the order a pass-after-pass render adds them in (floating-point sums are order-dependent, the image must not depend on the grouping)
__global__ void k_fb_fold(float4* __restrict__ fb, const float4* __restrict__ subSums, const int* __restrict__ owned, int nOwned, int groupPasses, size_t frame)
{
  const int o = blockIdx.x*blockDim.x + threadIdx.x;
  if (o >= nOwned) return;
  const int pixel = owned[o];
  float4 acc = fb[pixel];
  for (int j = 0; j < groupPasses; j++)
  {
    const float4 v = subSums[size_t(j)*frame + size_t(pixel)];
    acc.x += v.x; acc.y += v.y; acc.z += v.z;
  }
  fb[pixel] = acc;
}

struct HcPathHost
{
  HcDevBuf state[2][4];       // the two halves of the path queue: arrays A, B, C, D of 32-byte elements
  HcDevBuf owned, pathCount, ldr, subSums;
  HcDevBuf sortKeys, sortCount, sortCursor, sortPerm;   // material sort: u16 key per path, 2 x HC_SORT_MAX_KEYS counters, int index per path
  std::vector<unsigned char> materialsHost, globalsHost;
  int64_t capacity = 0;
  int nOwned = 0;
  long long ownedKey = -1;                    // (W, H, tile, rank, world) the owned-pixel list was built for
  bool haveNormalMaps = false;                // set by ValidateScene: normal-mapped or anisotropic (Beckmann / TRGGX) materials present - selects the extended k_pt_shade instantiation
  std::vector<cudaEvent_t> evPool;            // per-launch stage timing of the LAST pass of a hc_pt_pass call: {start, stop} pairs
  std::vector<int> evClass;                    // 0 closest, 1 shadow, 2 shade, 3 other
};
static HcPathHost* PH(hc_ctx* c) { return reinterpret_cast<HcPathHost*>(c->pathHost); }

void hc_path_free(hc_ctx* ctx)
{
  HcPathHost* p = PH(ctx);
  if (!p) return;
  for (int b = 0; b < 2; b++) for (int k = 0; k < 4; k++) hc_buf_free(p->state[b][k]);
  hc_buf_free(p->owned); hc_buf_free(p->pathCount); hc_buf_free(p->ldr); hc_buf_free(p->subSums);
  hc_buf_free(p->sortKeys); hc_buf_free(p->sortCount); hc_buf_free(p->sortCursor); hc_buf_free(p->sortPerm);
  for (cudaEvent_t e : p->evPool) cudaEventDestroy(e);
  delete p;
  ctx->pathHost = nullptr;
}

static HcPathHost* EnsureHost(hc_ctx* ctx)
{
  if (!PH(ctx)) ctx->pathHost = new HcPathHost;
  return PH(ctx);
}

static int ReserveState(hc_ctx* ctx, int64_t n, bool qmc)
{
  HcPathHost* p = EnsureHost(ctx);
  int rc;
  (void)qmc;
  for (int b = 0; b < 2; b++) for (int k = 0; k < 4; k++) if ((rc = hc_buf_reserve(ctx, p->state[b][k], uint64_t(n)*32))) return rc;
  if ((rc = hc_buf_reserve(ctx, p->pathCount, 2*256*sizeof(int)))) return rc;      // live counts per bounce, one block per pipeline
  if ((rc = hc_buf_reserve(ctx, p->sortKeys, uint64_t(n)*2))) return rc;
  if ((rc = hc_buf_reserve(ctx, p->sortPerm, uint64_t(n)*4))) return rc;
  if (!p->sortCount.ptr)
  {
    if ((rc = hc_buf_reserve(ctx, p->sortCount, HC_SORT_MAX_KEYS*sizeof(int)))) return rc;
    if ((rc = hc_buf_reserve(ctx, p->sortCursor, HC_SORT_MAX_KEYS*sizeof(int)))) return rc;
    HC_CUDA(cudaMemsetAsync(p->sortCount.ptr, 0, HC_SORT_MAX_KEYS*sizeof(int), ctx->stream));
  }
  p->capacity = n;
  return HC_OK;
}

static HcPathState StateOf(HcPathHost* p, int b)
{
  HcPathState s; s.a = (float4*)p->state[b][0].ptr; s.b = (float4*)p->state[b][1].ptr; s.c = (float4*)p->state[b][2].ptr; s.d = (float4*)p->state[b][3].ptr;
  return s;
}

// pixels of `rank` under the interleaved tile ownership: tile t -> GPU t mod G.  Inside a tile: 8 x 4 pixel blocks, so that the
// 32 paths of a warp start from a compact screen patch (coherent traversal); the image does not depend on this order (per-pixel generators,
// per-pixel accumulation).  Also the order in which hc_fb_reduce (hc_comm.cu) packs a rank's pixels.
void hc_owned_pixels_of(int W, int H, int T, int rank, int world, std::vector<int>& owned)
{
  T = std::max(1, T);
  const int tx = (W + T - 1)/T, ty = (H + T - 1)/T;
  owned.clear(); owned.reserve(size_t(W)*H/std::max(1, world) + 1024);
  for (int t = 0; t < tx*ty; t++)
  {
    if (t % world != rank) continue;
    const int x0 = (t % tx)*T, y0 = (t / tx)*T;
    for (int by = y0; by < std::min(H, y0 + T); by += 4)
      for (int bx = x0; bx < std::min(W, x0 + T); bx += 8)
        for (int y = by; y < std::min(std::min(H, y0 + T), by + 4); y++)
          for (int x = bx; x < std::min(std::min(W, x0 + T), bx + 8); x++) owned.push_back(y*W + x);
  }
}

static int BuildOwnedPixels(hc_ctx* ctx)
{
  HcPathHost* p = EnsureHost(ctx);
  std::vector<int> owned;
  hc_owned_pixels_of(ctx->width, ctx->height, ctx->tileSize, ctx->rank, ctx->worldSize, owned);
  p->nOwned = int(owned.size());
  int rc = hc_buf_reserve(ctx, p->owned, std::max<size_t>(owned.size(), 1)*sizeof(int)); if (rc) return rc;
  if (!owned.empty()) HC_CUDA(cudaMemcpyAsync(p->owned.ptr, owned.data(), owned.size()*sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return HC_OK;
}

// device list of this rank's pixels for callers outside the path tracer (hc_raycast_pass): rebuilt when the screen or the partition changed
int hc_path_owned_pixels(hc_ctx* ctx, const int** outDevicePixels, int* outCount)
{
  HcPathHost* p = EnsureHost(ctx);
  const long long key = (((long long)ctx->width*65536 + ctx->height)*1024 + ctx->tileSize)*4096 + (long long)ctx->rank*64 + ctx->worldSize;
  if (p->ownedKey != key || !p->owned.ptr) { int rc = BuildOwnedPixels(ctx); if (rc) return rc; p->ownedKey = key; }
  *outDevicePixels = (const int*)p->owned.ptr; *outCount = p->nOwned;
  return HC_OK;
}

// reject what the kernels do not implement — loudly, at init time
static int ValidateScene(hc_ctx* ctx, std::string& why)
{
  const unsigned char* g = ctx->globalsHead.data();
  auto gi = [&](int off) { int v; memcpy(&v, g + off, 4); return v; };
  // EngineGlobals::suns (soft directional lights, IHWLayerDataAssembler.cpp:421-450) are read only by sky portals, which are rejected below
  HcPathHost* p = PH(ctx);
  p->haveNormalMaps = false;
  const std::vector<unsigned char>& gl = p->globalsHost;
  const int lightsNum = gi(HC_EG_lightsNum), lightsOffset = gi(HC_EG_lightsOffset);
  std::vector<int> usedTex(ctx->alphaTexIdsHost.begin(), ctx->alphaTexIdsHost.end()), usedAux;      // texture ids the kernels can sample
  for (int l = 0; l < lightsNum; l++)
  {
    const float* L = reinterpret_cast<const float*>(gl.data()) + lightsOffset + l*HC_LIGHT_DATA_SIZE;
    int type, flags, tex, spot; memcpy(&type, L + HC_PLIGHT_TYPE, 4); memcpy(&flags, L + HC_PLIGHT_FLAGS, 4);
    memcpy(&tex, L + HC_PLIGHT_COLOR_TEX, 4); memcpy(&spot, L + HC_AREA_LIGHT_SPOT_DISTR, 4);
    if (type == HC_PLAIN_LIGHT_TYPE_SKY_DOME)
    {
      int tab, aux; memcpy(&tab, L + 30, 4); memcpy(&aux, L + 20, 4);                  // SKY_DOME_PDF_TABLE0, SKY_DOME_COLOR_TEX_AUX (clight.h:138, 153)
      if (l != gi(HC_EG_skyLightId)) { why = "more than one sky-dome light"; return HC_E_ARG; }
      if (!ctx->storage[HC_STORAGE_PDFS].ptr || tab < 0 || tab >= gi(HC_EG_pdfTableTableSize)) { why = "sky-dome light without a pdf table in the pdfs storage"; return HC_E_ARG; }
      int so; memcpy(&so, L + HC_PLIGHT_COLOR_TEX_MATRIX, 4);                           // environment map: sampler at L + HC_SKY_DOME_SAMPLER0 (hc_shade.cuh)
      if (so != HC_INVALID_TEXTURE && so >= 0 && HC_SKY_DOME_SAMPLER0 + so*4 + 12 <= HC_LIGHT_DATA_SIZE)
      { int texId; memcpy(&texId, L + HC_SKY_DOME_SAMPLER0 + so*4 + 2, 4); if (texId > 0) usedTex.push_back(texId); }
      continue;
    }
    if (type == HC_PLAIN_LIGHT_TYPE_POINT_SPOT || type == HC_PLAIN_LIGHT_TYPE_DIRECT) continue;
    if (type == HC_PLAIN_LIGHT_TYPE_CYLINDER)
    {
      int ctex, ctab; memcpy(&ctex, L + 29, 4); memcpy(&ctab, L + 31, 4);                     // CYLINDER_TEX_ID, CYLINDER_PDF_TABLE_ID (clight.h:111-113)
      if (ctex != HC_INVALID_TEXTURE)                                                         // colour texture: sampler inside the light record (CYLINDER_TEXMATRIX_ID, float4 index)
      {
        int so; memcpy(&so, L + 30, 4);
        if (so < 0 || so*4 + 12 > HC_LIGHT_DATA_SIZE) { why = "cylinder light: colour sampler outside the light record"; return HC_E_RANGE; }
        int texId; memcpy(&texId, L + so*4 + 2, 4); if (texId > 0) usedTex.push_back(texId);
      }
      if (ctab < 0 || ctab >= gi(HC_EG_pdfTableTableSize) || !ctx->storage[HC_STORAGE_PDFS].ptr) { why = "cylinder light without the pdf table the driver builds for it (UpdatePdfTablesForLight)"; return HC_E_ARG; }
      continue;
    }
    if (type == HC_PLAIN_LIGHT_TYPE_MESH)
    {
      int meshTab, pdfTab, triNum, mtex; memcpy(&meshTab, L + 14, 4); memcpy(&pdfTab, L + 15, 4); memcpy(&triNum, L + 16, 4); memcpy(&mtex, L + 30, 4);   // MESH_LIGHT_* (clight.h:169-175)
      const int nTab = gi(HC_EG_pdfTableTableSize);
      if (!ctx->storage[HC_STORAGE_PDFS].ptr || meshTab < 0 || meshTab >= nTab || pdfTab < 0 || pdfTab >= nTab || triNum < 1) { why = "mesh light without its mesh copy / triangle table in the pdfs storage"; return HC_E_ARG; }
      if (mtex != HC_INVALID_TEXTURE)                                                         // MESH_LIGHT_TEXMATRIX_ID
      {
        int so; memcpy(&so, L + 31, 4);
        if (so < 0 || so*4 + 12 > HC_LIGHT_DATA_SIZE) { why = "mesh light: colour sampler outside the light record"; return HC_E_RANGE; }
        int texId; memcpy(&texId, L + so*4 + 2, 4); if (texId > 0) usedTex.push_back(texId);
      }
      if (flags & HC_LIGHT_HAS_IES) { why = "IES distributions are not supported yet"; return HC_E_ARG; }
      continue;
    }
    // IES web: a single-channel float image in the pdfs storage, looked up by point and area lights (lightDistributionMask)
    auto iesOk = [&]() -> bool
    {
      int iesTex; memcpy(&iesTex, L + HC_IES_SPHERE_TEX_ID, 4);
      if (!ctx->storage[HC_STORAGE_PDFS].ptr || iesTex < 0 || iesTex >= gi(HC_EG_pdfTableTableSize)) { why = "light with LIGHT_HAS_IES but no IES table in the pdfs storage"; return false; }
      int off; memcpy(&off, gl.data() + (size_t(gi(HC_EG_pdfTableTableOffset)) + size_t(iesTex))*4, 4);
      const size_t pdfBytes = ctx->storage[HC_STORAGE_PDFS].bytes;
      if (off < 0 || size_t(off)*16 + 16 > pdfBytes) { why = "IES table outside the pdfs storage"; return false; }
      int hdr[4] = { 0, 0, 0, 0 };                                          // the storage has no host mirror: read the 16-byte image header back (init time only)
      if (cudaMemcpy(hdr, (const char*)ctx->storage[HC_STORAGE_PDFS].ptr + size_t(off)*16, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { why = "IES table: header read-back failed"; return false; }
      if (hdr[0] <= 0 || hdr[1] <= 0 || hdr[3] != 4 || size_t(off)*16 + 16 + size_t(hdr[0])*size_t(hdr[1])*4 > pdfBytes) { why = "IES table: expected a w x h single-channel float image"; return false; }
      return true;
    };
    if (type == HC_PLAIN_LIGHT_TYPE_SPHERE)
    {
      if (flags & HC_LIGHT_HAS_IES) { why = "IES distributions on sphere lights are not supported yet (point and area lights are)"; return HC_E_ARG; }
      continue;
    }
    if (type == HC_PLAIN_LIGHT_TYPE_POINT_OMNI)
    {
      if ((flags & HC_LIGHT_HAS_IES) && !iesOk()) return HC_E_ARG;
      continue;
    }
    if (type != HC_PLAIN_LIGHT_TYPE_AREA) { why = "unknown light type (area, sphere, cylinder, mesh, point, spot, directional, sky-dome lights are supported)"; return HC_E_ARG; }
    if (flags & HC_AREA_LIGHT_SKY_PORTAL) { why = "sky-portal area lights are not supported yet"; return HC_E_ARG; }
    if ((flags & HC_LIGHT_HAS_IES) && !iesOk()) return HC_E_ARG;
    if (tex != HC_INVALID_TEXTURE) { why = "textured area lights are not supported yet"; return HC_E_ARG; }
  }
  {
    // remap lists: every target material id must exist; opacity samplers of the alpha-tested tree: every texture id must have an image
    const int matTabSize0 = gi(HC_EG_materialsTableSize);
    for (size_t k = 1; k < ctx->remapListsHost.size() && ctx->remapListsSize > 0; k += 2)
      if (ctx->remapListsHost[k] < 0 || ctx->remapListsHost[k] >= matTabSize0) { why = "a material remap list maps to material id " + std::to_string(ctx->remapListsHost[k]) + ", which is not in the materials table"; return HC_E_RANGE; }
    if (ctx->haveAlpha1)
    {
      if (!ctx->storage[HC_STORAGE_TEXTURES].ptr) { why = "opacity maps need the textures storage"; return HC_E_STATE; }
      const int texTabOff = gi(HC_EG_texturesTableOffset), texTabSize = gi(HC_EG_texturesTableSize);
      for (int texId : ctx->alphaTexIdsHost)
      {
        if (texId == 0) continue;
        int off = -1;
        if (texId > 0 && texId < texTabSize) memcpy(&off, gl.data() + 4*size_t(texTabOff + texId), 4);
        if (off < 0) { why = "an opacity sampler of the alpha-tested tree uses texture id " + std::to_string(texId) + ", which has no image in the textures storage"; return HC_E_RANGE; }
      }
    }
  }
  // walk the material nodes REACHABLE from the materials table (the storage may hold stale or unused chunks after in-place updates)
  const size_t nNodes = p->materialsHost.size()/(HC_PLAIN_MATERIAL_DATA_SIZE*4);
  const int matTabOff = gi(HC_EG_materialsTableOffset), matTabSize = gi(HC_EG_materialsTableSize);
  if (size_t(matTabOff) + size_t(std::max(matTabSize, 0)) > gl.size()/4) { why = "materials table lies outside the globals blob"; return HC_E_RANGE; }
  std::vector<long long> todo;
  for (int t = 0; t < matTabSize; t++)
  {
    int off4; memcpy(&off4, gl.data() + (size_t(matTabOff) + size_t(t))*4, 4);      // float4 offset of the head node, or -1
    if (off4 < 0) continue;
    todo.push_back((long long)off4*4);                                               // -> float index
  }
  size_t visited = 0;
  while (!todo.empty())
  {
    const long long f0 = todo.back(); todo.pop_back();
    if (f0 < 0 || size_t(f0) + HC_PLAIN_MATERIAL_DATA_SIZE > p->materialsHost.size()/4) { why = "material node outside the materials storage"; return HC_E_RANGE; }
    if (++visited > 4*nNodes + 64) { why = "material blend tree does not terminate"; return HC_E_ARG; }
    const float* m = reinterpret_cast<const float*>(p->materialsHost.data()) + f0;
    int type, flags, ntex, ptex; memcpy(&type, m + HC_PLAIN_MAT_TYPE_OFFSET, 4); memcpy(&flags, m + HC_PLAIN_MAT_FLAGS_OFFSET, 4);
    memcpy(&ntex, m + HC_NORMAL_TEX_OFFSET, 4); memcpy(&ptex, m + HC_PROC_TEX1_F4_HEAD_OFFSET, 4);
    const bool ok = type == HC_PLAIN_MAT_CLASS_LAMBERT || type == HC_PLAIN_MAT_CLASS_PHONG_SPECULAR || type == HC_PLAIN_MAT_CLASS_BLINN_SPECULAR || type == HC_PLAIN_MAT_CLASS_GGX ||
                    type == HC_PLAIN_MAT_CLASS_PERFECT_MIRROR || type == HC_PLAIN_MAT_CLASS_GLASS || type == HC_PLAIN_MAT_CLASS_BLEND_MASK ||
                    type == HC_PLAIN_MAT_CLASS_EMISSIVE || type == HC_PLAIN_MAT_CLASS_OREN_NAYAR || type == HC_PLAIN_MAT_CLASS_TRANSLUCENT || type == HC_PLAIN_MAT_CLASS_THIN_GLASS ||
                    type == HC_PLAIN_MAT_CLASS_BECKMANN || type == HC_PLAIN_MAT_CLASS_TRGGX;
    if (!ok) { why = "material class " + std::to_string(type) + " is not supported yet (Lambert, Oren-Nayar, translucent, Phong, Blinn, GGX, Beckmann, TRGGX, mirror, glass, thin glass, blend mask are)"; return HC_E_ARG; }
    {
      // texture ids behind the samplers this material class reads (Sample2D call sites of hc_shade.cuh): sampler offset in float4 from the node start
      int slots[5] = { HC_EMISSIVE_TEXMATRIXID_OFFSET, HC_LAMBERT_TEXMATRIXID_OFFSET, -1, -1, -1 };
      if (type == HC_PLAIN_MAT_CLASS_BECKMANN || type == HC_PLAIN_MAT_CLASS_TRGGX)
      {
        slots[2] = HC_BECKMANN_GLOSINESS_TEXMATRIXID_OFFSET; slots[3] = HC_BECKMANN_ANISO_TEXMATRIXID_OFFSET; slots[4] = HC_BECKMANN_ROT_TEXMATRIXID_OFFSET;
        p->haveNormalMaps = true;                      // the anisotropic lobes live in the extended k_pt_shade instantiation only
      }
      if (type == HC_PLAIN_MAT_CLASS_PHONG_SPECULAR || type == HC_PLAIN_MAT_CLASS_BLINN_SPECULAR || type == HC_PLAIN_MAT_CLASS_THIN_GLASS || type == HC_PLAIN_MAT_CLASS_GGX) slots[2] = HC_PHONG_GLOSINESS_TEXMATRIXID_OFFSET;
      if (type == HC_PLAIN_MAT_CLASS_GLASS) slots[2] = HC_GLASS_GLOSINESS_TEXMATRIXID_OFFSET;
      if (type == HC_PLAIN_MAT_CLASS_EMISSIVE) slots[1] = -1;
      for (int sl : slots)
      {
        if (sl < 0) continue;
        int so; memcpy(&so, m + sl, 4);
        if (so == HC_INVALID_TEXTURE || so < 0 || so*4 + 12 > HC_PLAIN_MATERIAL_DATA_SIZE) continue;
        int texId; memcpy(&texId, m + so*4 + 2, 4);
        if (texId > 0) usedTex.push_back(texId);
      }
    }
    if (ntex != HC_INVALID_TEXTURE)            // normal map: image in the "textures_aux" storage, found through the aux texture table
    {
      p->haveNormalMaps = true; usedAux.push_back(ntex);
      if (!ctx->storage[HC_STORAGE_TEXTURES_AUX].ptr || ntex < 0 || ntex >= gi(HC_EG_texturesAuxTableSize)) { why = "normal map without an image in the textures_aux storage"; return HC_E_ARG; }
      int auxOff; memcpy(&auxOff, gl.data() + 4*size_t(gi(HC_EG_texturesAuxTableOffset) + ntex), 4);
      if (auxOff < 0) { why = "normal map texture id has no entry in the aux texture table"; return HC_E_ARG; }
    }
    if (ptex != HC_INVALID_TEXTURE) { why = "procedural textures are not supported yet"; return HC_E_ARG; }
    if (type == HC_PLAIN_MAT_CLASS_BLEND_MASK)
    {
      int bf, o1, o2; memcpy(&bf, m + HC_BLEND_MASK_FLAGS_OFFSET, 4);
      memcpy(&o1, m + HC_BLEND_MASK_MATERIAL1_OFFSET, 4); memcpy(&o2, m + HC_BLEND_MASK_MATERIAL2_OFFSET, 4);
      if (bf & HC_BLEND_MASK_FALOFF) { why = "falloff blend masks are not supported yet"; return HC_E_ARG; }
      if (o1 == 0 || o2 == 0) { why = "blend mask node references itself"; return HC_E_ARG; }
      todo.push_back(f0 + (long long)o1*HC_PLAIN_MATERIAL_DATA_SIZE);               // children by RELATIVE node offset (cmaterial.h:1994-1995)
      todo.push_back(f0 + (long long)o2*HC_PLAIN_MATERIAL_DATA_SIZE);
    }
  }
  // every image the kernels can sample: RGBA8 / float4, or single-channel (depth 1) float / 8-bit - what ReadImageSw4 / ReadImageSw1
  // (hc_texture.cuh) implement - and wholly inside its storage.  (Unreferenced slots of a storage may hold anything.)
  if (ctx->texturesDirty)
  {
    struct Tab { int offOff, sizeOff, slot; const char* name; const std::vector<int>* ids; };
    const Tab tabs[2] = { { HC_EG_texturesTableOffset, HC_EG_texturesTableSize, HC_STORAGE_TEXTURES, "textures", &usedTex },
                          { HC_EG_texturesAuxTableOffset, HC_EG_texturesAuxTableSize, HC_STORAGE_TEXTURES_AUX, "textures_aux", &usedAux } };
    for (const Tab& tb : tabs)
    {
      const int tOff = gi(tb.offOff), tSize = gi(tb.sizeOff);
      const HcDevBuf& st = ctx->storage[tb.slot];
      std::vector<int> ids(*tb.ids); std::sort(ids.begin(), ids.end()); ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
      for (int t : ids)
      {
        if (t <= 0 || t >= tSize || tOff < 0 || size_t(tOff) + size_t(tSize) > gl.size()/4) continue;      // the fetch treats these as "no image"
        int off4; memcpy(&off4, gl.data() + 4*(size_t(tOff) + size_t(t)), 4);
        if (off4 < 0) continue;
        if (!st.ptr || (uint64_t(off4) + 1)*16 > st.bytes) { why = std::string(tb.name) + " table entry " + std::to_string(t) + " points outside the storage"; return HC_E_RANGE; }
        int hdr[4];
        if (cudaMemcpy(hdr, (const char*)st.ptr + size_t(off4)*16, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); why = "reading a texture header back failed"; return HC_E_STATE; }
        const int w = hdr[0], h = hdr[1], d = hdr[2], bpp = hdr[3];
        const bool okFmt = (d == 1) ? (bpp == 1 || bpp == 4) : (bpp == 4 || bpp == 16);
        if (w <= 0 || h <= 0 || !okFmt)
        { why = std::string(tb.name) + " image " + std::to_string(t) + ": " + std::to_string(w) + "x" + std::to_string(h) + ", depth " + std::to_string(d) + ", " + std::to_string(bpp) +
                " bytes per pixel is not a supported format (RGBA8, float4, single-channel float or 8-bit)"; return HC_E_ARG; }
        if ((uint64_t(off4) + 1)*16 + uint64_t(w)*uint64_t(h)*uint64_t(bpp) > st.bytes) { why = std::string(tb.name) + " image " + std::to_string(t) + " does not fit into the storage"; return HC_E_RANGE; }
      }
    }
  }
  return HC_OK;
}

static HcScene MakeScene(hc_ctx* ctx)
{
  const unsigned char* g = ctx->globalsHead.data();
  auto gi = [&](int off) { int v; memcpy(&v, g + off, 4); return v; };
  HcScene s;
  s.globals = (const int*)ctx->globals.ptr;
  s.geom = (const float4*)ctx->storage[HC_STORAGE_GEOM].ptr;
  s.materials = (const float4*)ctx->storage[HC_STORAGE_MATERIALS].ptr;
  s.textures = (const int4*)ctx->storage[HC_STORAGE_TEXTURES].ptr;
  s.texturesAux = (const int4*)ctx->storage[HC_STORAGE_TEXTURES_AUX].ptr;
  s.texturesAuxTableOffset = gi(HC_EG_texturesAuxTableOffset);
  s.pdfs = (const float4*)ctx->storage[HC_STORAGE_PDFS].ptr;
  s.pdfTableTableOffset = gi(HC_EG_pdfTableTableOffset);
  s.instMatrices = (const float4*)ctx->instMatrices.ptr;
  s.instLightIds = (const int*)ctx->instLightIds.ptr;
  s.remapLists = ctx->remapListsSize > 0 ? (const int*)ctx->remapLists.ptr : nullptr;
  s.remapTable = ctx->remapTableSize > 0 ? (const int2*)ctx->remapTable.ptr : nullptr;
  s.remapInst  = ctx->remapInstSize > 0 ? (const int*)ctx->remapInst.ptr : nullptr;
  s.remapListsSize = ctx->remapListsSize; s.remapTableSize = ctx->remapTableSize; s.remapInstSize = ctx->remapInstSize;
  s.materialsTableOffset = gi(HC_EG_materialsTableOffset); s.geometryTableOffset = gi(HC_EG_geometryTableOffset);
  s.texturesTableOffset = gi(HC_EG_texturesTableOffset);
  s.lightSelTableOffsetRev = gi(HC_EG_lightSelectorTableOffsetRev); s.lightSelTableSizeRev = gi(HC_EG_lightSelectorTableSizeRev);
  s.lightsOffset = gi(HC_EG_lightsOffset); s.lightsNum = gi(HC_EG_lightsNum); s.skyLightId = gi(HC_EG_skyLightId);
  s.gflags = gi(HC_EG_g_flags);
  s.traceDepth = gi(HC_EG_varsI + 4*HC_HRT_TRACE_DEPTH); s.diffTraceDepth = gi(HC_EG_varsI + 4*HC_HRT_DIFFUSE_TRACE_DEPTH);
  s.essGgxTableOffsetBytes = HC_EG_m_essGgx2017Table;
  return s;
}

// Niederreiter base-2 table, 11 dimensions x 31 bits (Bratley-Fox-Niederreiter, TOMS 738; what initQuasirandomGenerator builds,
// qmc_sobol_niederreiter.cpp:75-186): per dimension an irreducible polynomial p over GF(2); every deg(p) columns the running power
// b = p^q gains one more factor p and its recurrence generates the next rows.  Only the top 31 of 63 bits are kept.
static void BuildQmcTable(unsigned table[HC_QRNG_DIMENSIONS_K][HC_QRNG_RESOLUTION_K])
{
  static const unsigned long long irred[HC_QRNG_DIMENSIONS_K] = { 2, 3, 7, 11, 13, 19, 25, 31, 37, 41, 47 };
  auto deg = [](unsigned long long p) { int d = -1; while (p) { d++; p >>= 1; } return d; };
  auto mul = [](unsigned long long a, unsigned long long b) { unsigned long long r = 0; while (b) { if (b & 1) r ^= a; a <<= 1; b >>= 1; } return r; };
  for (int dim = 0; dim < HC_QRNG_DIMENSIONS_K; dim++)
  {
    const unsigned long long p = irred[dim]; const int e = deg(p);
    unsigned long long cj[63]; for (auto& c : cj) c = 0;
    unsigned long long b = 1; int m = 0, u = e; int v[96];
    for (int j = 62; j >= 32; --j, ++u)
    {
      if (u == e)
      {
        u = 0; const int m1 = m; b = mul(b, p); m += e;
        for (int i = 0; i < m1; i++) v[i] = 0;
        for (int i = m1; i < m; i++) v[i] = 1;
        for (int i = m; i <= 63 + e - 2; i++) { int a = 0; for (int k = 1; k <= m; k++) a ^= v[i - k] & int((b >> (m - k)) & 1ull); v[i] = a; }
      }
      for (int i = 0; i < 63; i++) cj[i] |= (unsigned long long)v[i + u] << j;
    }
    for (int bit = 0; bit < HC_QRNG_RESOLUTION_K; bit++) table[dim][bit] = (unsigned)((cj[bit] >> 32) & 0x7FFFFFFFu);
  }
}

// host mirrors for validation (materials are small; the globals blob carries the lights), validation, choice of the shade-kernel variant.
// Runs at hc_pt_init and again before the next pass whenever a scene upload entry point was called in between (ctx->sceneDirty): the reference
// driver calls InitPathTracing once but re-uploads lights, materials and globals on later frames.
static int RefreshScene(hc_ctx* ctx, const char* who)
{
  HC_REQUIRE(ctx->globals.ptr && ctx->bvhNodes.ptr && ctx->instMatrices.ptr && ctx->instLightIds.ptr, HC_E_STATE,
             "scene incomplete (globals, BVH, instance matrices and instance light ids are required)");
  HC_REQUIRE(ctx->storage[HC_STORAGE_GEOM].ptr && ctx->storage[HC_STORAGE_MATERIALS].ptr, HC_E_STATE, "geom / materials storage missing");
  HcPathHost* p = EnsureHost(ctx);
  p->materialsHost = ctx->materialsMirror;                    // kept up to date by hc_storage_write / hc_set_globals: no read-back
  p->materialsHost.resize(ctx->storage[HC_STORAGE_MATERIALS].bytes, 0);
  p->globalsHost = ctx->globalsMirror;
  std::string why;
  const int rc = ValidateScene(ctx, why);
  if (rc) { hc_set_error((std::string(who) + ": " + why).c_str()); return rc; }
  ctx->sceneDirty = false; ctx->texturesDirty = false;
  return HC_OK;
}

extern "C"
{
int hc_resize(hc_ctx* ctx, int width, int height)
{
  if (!ctx || width <= 0 || height <= 0) return HC_E_ARG;
  HC_CUDA(cudaSetDevice(ctx->device));
  ctx->width = width; ctx->height = height; ctx->ptReady = false; ctx->combinedValid = false;
  int rc = hc_buf_reserve(ctx, ctx->fbSum, uint64_t(width)*height*16); if (rc) return rc;
  HC_CUDA(cudaMemsetAsync(ctx->fbSum.ptr, 0, uint64_t(width)*height*16, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->spp = 0.0; ctx->passCounter = 0;
  return HC_OK;
}

int hc_pt_set_tiles(hc_ctx* ctx, int tileSize, int rank, int worldSize)
{
  if (!ctx || tileSize <= 0 || worldSize <= 0 || rank < 0 || rank >= worldSize) return HC_E_ARG;
  ctx->tileSize = tileSize; ctx->rank = rank; ctx->worldSize = worldSize; ctx->ptReady = false;
  return HC_OK;
}

int hc_pt_set_sample_streams(hc_ctx* ctx, int streams, int64_t maxPathsInFlight)
{
  if (!ctx || streams < 1 || streams > 64 || maxPathsInFlight < 0) return HC_E_ARG;
  ctx->sampleStreams = streams; ctx->maxPathsInFlight = maxPathsInFlight; ctx->ptReady = false;      // the generator array changes: hc_pt_init again
  return HC_OK;
}

int hc_pt_group_passes(hc_ctx* ctx, int* outPasses)
{
  if (!ctx || !outPasses) return HC_E_ARG;
  HcPathHost* p = PH(ctx);
  HC_REQUIRE(ctx->ptReady && p, HC_E_STATE, "hc_pt_group_passes: call hc_pt_init first");
  if (p->nOwned <= 0) { *outPasses = 1; return HC_OK; }                          // a rank without pixels (more ranks than tiles): its passes are empty
  const int64_t frame = int64_t(ctx->width)*ctx->height;
  // default limit: 8M paths (4 GB of path queue), or one pass of the frame if that is more.  Below 2M paths a launch is latency-bound;
  // beyond that the late bounces still gain from launching over more live paths (scripts/gpu_streams_cap.py, 1080p, 1 / 2 / 4 / 8 passes in
  // flight: C3 6.81 / 6.20 / 5.92 / 5.78 ms per pass, C4 6.33 / 6.03 / 5.86 / 5.78)
  const int64_t cap = ctx->maxPathsInFlight > 0 ? ctx->maxPathsInFlight : std::max<int64_t>(frame, int64_t(1) << 23);
  int m = int(std::min<int64_t>(ctx->sampleStreams, std::max<int64_t>(1, cap/p->nOwned)));
  if (frame > (int64_t(1) << 24) || ctx->sampleStreams > 64) m = 1;              // the sub-pass index shares the path's pixel word (7 bits above 24)
  *outPasses = std::max(1, m);
  return HC_OK;
}

int hc_pt_init(hc_ctx* ctx, int seed)
{
  if (!ctx) return HC_E_ARG;
  HC_REQUIRE(ctx->width > 0 && ctx->height > 0, HC_E_STATE, "hc_pt_init: call hc_resize first");
  HC_CUDA(cudaSetDevice(ctx->device));
  int rc = RefreshScene(ctx, "hc_pt_init"); if (rc) return rc;

  const int n = ctx->width*ctx->height;
  // generator of stream k of pixel p: index k*W*H + p, i.e. stream 0 is the single-stream rule and the further streams continue the numbering
  const int nGen = n*std::max(1, ctx->sampleStreams);
  if ((rc = hc_buf_reserve(ctx, ctx->pixelRng, uint64_t(nGen)*8))) return rc;
  k_init_rng<<<(nGen + 255)/256, 256, 0, ctx->stream>>>((uint2*)ctx->pixelRng.ptr, nGen, seed);
  HC_CUDA(cudaGetLastError());
  unsigned table[HC_QRNG_DIMENSIONS_K][HC_QRNG_RESOLUTION_K];
  BuildQmcTable(table);
  if ((rc = hc_buf_reserve(ctx, ctx->qmcTable, sizeof(table)))) return rc;
  HC_CUDA(cudaMemcpyAsync(ctx->qmcTable.ptr, table, sizeof(table), cudaMemcpyHostToDevice, ctx->stream));
  { const int* px; int cnt; if ((rc = hc_path_owned_pixels(ctx, &px, &cnt))) return rc; }
  HC_CUDA(cudaMemsetAsync(ctx->fbSum.ptr, 0, uint64_t(n)*16, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->seed = seed; ctx->spp = 0.0; ctx->passCounter = 0; ctx->ptReady = true; ctx->combinedValid = false;
  ctx->stats.kernelLaunches++;
  return HC_OK;
}

int hc_pt_pass(hc_ctx* ctx, int integrator, int passes)
{
  if (!ctx || passes < 0) return HC_E_ARG;
  HC_REQUIRE(integrator == HC_INTEGRATOR_PT || integrator == HC_INTEGRATOR_MISPT || integrator == HC_INTEGRATOR_MISPT_QMC, HC_E_ARG, "hc_pt_pass: unknown integrator");
  HC_REQUIRE(ctx->ptReady, HC_E_STATE, "hc_pt_pass: call hc_pt_init first (after the scene, the screen size and the tiles are set)");
  HC_CUDA(cudaSetDevice(ctx->device));
  const auto hostT0 = std::chrono::steady_clock::now();
  if (ctx->sceneDirty) { const int rcv = RefreshScene(ctx, "hc_pt_pass"); if (rcv) return rcv; }      // lights / materials / globals re-uploaded since hc_pt_init
  HcPathHost* p = PH(ctx);
  ctx->combinedValid = false;                                  // new samples: an earlier cross-rank sum (hc_fb_reduce) is stale
  const bool qmc = (integrator == HC_INTEGRATOR_MISPT_QMC);
  const int W = ctx->width, H = ctx->height;
  const int nPerPass = qmc ? ((W*H - ctx->rank + ctx->worldSize - 1)/ctx->worldSize) : p->nOwned;
  if (nPerPass <= 0) return HC_OK;
  // sample streams: up to `groupMax` consecutive passes share one wavefront (hc_pt_group_passes)
  const int S = std::max(1, ctx->sampleStreams);
  int groupMax = 1; { int rcg = hc_pt_group_passes(ctx, &groupMax); if (rcg) return rcg; }
  if (qmc)                                                       // QMC splits SAMPLE indices over the ranks, not tiles: its own pass size
  {
    const int64_t frame = int64_t(W)*H;
    const int64_t cap = ctx->maxPathsInFlight > 0 ? ctx->maxPathsInFlight : std::max<int64_t>(frame, int64_t(1) << 23);
    groupMax = (frame > (int64_t(1) << 24)) ? 1 : int(std::min<int64_t>(S, std::max<int64_t>(1, cap/nPerPass)));
  }
  groupMax = std::max(1, std::min(groupMax, passes));
  int n = nPerPass*groupMax;                                    // paths of the wavefront being enqueued (set per group below)
  int rc = ReserveState(ctx, n, qmc); if (rc) return rc;
  p = PH(ctx);
  // per-sub-pass sums (PT / MISPT: one path per pixel and sub-pass, folded in pass order).  QMC paths land on arbitrary pixels and add with float
  // atomics in any case, so they go straight into the frame
  if (groupMax > 1 && !qmc && (rc = hc_buf_reserve(ctx, p->subSums, uint64_t(groupMax)*uint64_t(W)*uint64_t(H)*16))) return rc;
  HcScene scn = MakeScene(ctx);
  if (integrator == HC_INTEGRATOR_PT) scn.gflags |= HC_HRT_STUPID_PT_MODE;          // IntegratorStupidPT::DoPass sets it in g_flags (CPUExp_Integrators.h:326-330)
  const HcCamera cam = hc_camera_from_globals(ctx->globalsHead.data());
  HcPassParams pp;
  pp.integrator = integrator; pp.width = W; pp.height = H; pp.world = ctx->worldSize; pp.rank = ctx->rank;
  pp.streams = S; pp.streamBase = 0; pp.groupPasses = 1; pp.nOwned = nPerPass; pp.pixMask = 0x7FFFFFFFu; pp.subShift = 31;
  float4* subSums = nullptr;
  pp.maxDepth = (integrator == HC_INTEGRATOR_PT) ? scn.traceDepth + 1 : scn.traceDepth;     // IntegratorStupidPT::SetMaxDepth adds one (CPUExp_Integrators.h:333)
  HC_REQUIRE(pp.maxDepth >= 1 && pp.maxDepth < 250, HC_E_ARG, "hc_pt_pass: HRT_TRACE_DEPTH out of range");
  int* counts = (int*)p->pathCount.ptr;
  const int* rmQMC = (const int*)ctx->globals.ptr + HC_EG_rmQMC/4;
  const unsigned* qtab = (const unsigned*)ctx->qmcTable.ptr;
  const int nBounces = (integrator == HC_INTEGRATOR_PT) ? pp.maxDepth : pp.maxDepth;       // MISPT finishes every path at depth maxDepth-1
  // material sort: one bucket per entry of the materials table + one for rays that missed; off when the table is too large for one CTA scan
  int sortKeys = 0;
  if (ctx->materialSort != 0)
  {
    int matTab; memcpy(&matTab, ctx->globalsHead.data() + HC_EG_materialsTableSize, 4);
    // measured on B200 (scripts/gpu_sort_ab.py): C3 (6 materials) 7.66 -> 7.26 ms per 1080p pass, shade 2.02 -> 1.41 ms, sort 0.29 ms;
    // C4 (one surface material + the light's) 6.54 -> 6.59 ms.  Hence: only when there are at least three materials to tell apart.
    // ... and, unless asked for explicitly, only when the queue is long enough for the three extra launches per bounce to pay
    // (scripts/gpu_sort_small.py, ms per pass with / without: C1 262k paths 0.636 / 0.616, C3 130k 1.320 / 1.269, 518k 2.542 / 2.602, 2.07M 6.66 / 7.26)
    const bool longEnough = (ctx->materialSort == 1) || n >= 384*1024;
    if (matTab >= 3 && matTab + 1 <= HC_SORT_MAX_KEYS && longEnough) sortKeys = matTab + 1;
  }

  // stage timing: one {start, stop} event pair per launch, summed per kernel class after the final synchronise (feeds MRaysStat /
  // hc_stats and the roofline of bench.py).  Events on the launching stream cost ~1 us each, i.e. well below 1 % of a pass.
  size_t evUsed = 0;
  p->evClass.clear();
  bool timed = true;                           // stage events only on the LAST pass of a call: 2 event records per launch are not free for small frames
  auto stageBegin = [&](int cls) -> int
  {
    if (!timed) return 0;
    if (evUsed + 2 > p->evPool.size()) { p->evPool.resize(evUsed + 2, nullptr); for (size_t k = evUsed; k < evUsed + 2; k++) if (cudaEventCreate(&p->evPool[k]) != cudaSuccess) return HC_E_NOMEM; }
    p->evClass.push_back(cls);
    return int(cudaEventRecord(p->evPool[evUsed], ctx->stream));
  };
  auto stageEnd = [&]() -> int { if (!timed) return 0; const int rc = int(cudaEventRecord(p->evPool[evUsed + 1], ctx->stream)); evUsed += 2; return rc; };
#define HC_STAGE(cls, launch) { if ((rc = stageBegin(cls))) return rc; launch; if ((rc = stageEnd())) return rc; }

  // one pass = a fixed sequence of launches (live counts stay on the device), so untimed passes of a multi-pass call are replayed from a
  // CUDA graph captured once per call: at 512x512 (C1) the ~40 launches / memsets of a pass cost more to enqueue one by one than to run.
  // Not for QMC (the pass index is a kernel argument).  The last pass of a call always runs directly, with the stage events.
  // Small frames are latency-bound: a 262k-ray launch is 2048 warps, one 32-ray batch each, and takes as long as its slowest warp
  // (scripts/gpu_small_launch.py: 45-50 us from 65k to 262k rays on the C1 scene).  Passes over at most 1M paths are therefore split into two
  // independent halves of the owned-pixel list - own live counters, state slices and pair of streams - so that one half's shade / sort / launch
  // gaps are filled by the other half's traversal.  Pixels are independent, so the image is unchanged.  Measured (scripts/gpu_graph_ab.py,
  // HC_PT_PIPES=1 for one pipeline): C1 0.619 -> 0.591 ms per pass, C3 at 480x270 1.272 -> 1.216; no more than that, because the warps of both
  // halves were already resident together in the single launch.  The timed last pass of a call runs as one pipeline (stage events).
  struct Pipe { int first, n; int* counts; cudaStream_t s, side; cudaEvent_t evFork, evJoin; };
  auto sliceOf = [&](int b, int first) -> HcPathState
  {
    HcPathState st = StateOf(p, b);
    st.a += 2*size_t(first); st.b += 2*size_t(first); st.c += 2*size_t(first); st.d += 2*size_t(first);
    return st;
  };
  auto genPipe = [&](const Pipe& q) -> int
  {
    int rc = 0;
    HC_CUDA(cudaMemsetAsync(q.counts, 0, 256*sizeof(int), q.s));
    HcPathState st0 = sliceOf(0, q.first);
    HC_STAGE(3, (k_pt_generate<<<(q.n + 255)/256, 256, 0, q.s>>>(cam, pp, q.n, qmc ? 0 : q.first, (const int*)p->owned.ptr, (uint2*)ctx->pixelRng.ptr, rmQMC, qtab, st0, q.counts)));
    HC_CUDA(cudaGetLastError());
    ctx->stats.kernelLaunches++; ctx->stats.paths += (uint64_t)q.n;
    return HC_OK;
  };
  auto bouncePipe = [&](const Pipe& q, int depth, int cur) -> int
  {
    int rc = 0;
    HcPathState in = sliceOf(cur, q.first), out = sliceOf(1 - cur, q.first);
    int* counts = q.counts;
    const int n = q.n;
    // the shadow rays of the previous bounce and the closest-hit rays of this one are independent: two streams, so that the tail of
    // one persistent launch (last warps finishing their rays) is filled by the head of the other; joined before sort / shade
    const bool haveShadow = (depth > 0 && integrator != HC_INTEGRATOR_PT);
    if (haveShadow)
    {
      HC_CUDA(cudaEventRecord(q.evFork, q.s));
      HC_CUDA(cudaStreamWaitEvent(q.side, q.evFork, 0));
    }
    // closest hit: ray = array A (interleaved {pos, dir}), hit record -> second half of the D element
    HC_STAGE(0, if ((rc = hc_launch_trace_counted(ctx, false, in.a, in.a + 1, 2, n, counts + depth, (HcHit*)(in.d + 1), nullptr, 2, q.s))) return rc);
    if (haveShadow)
    {
      // timing: the "shadow" stage is what the any-hit launch ADDS after the closest-hit launch has finished (both measured on the
      // main stream), so that the stage times still add up to the pass
      if ((rc = stageBegin(1))) return rc;
      // any hit: shadow ray = array C (the origin's .w carries sexp.x and is ignored), an occluded ray gets its t_far zeroed
      if ((rc = hc_launch_trace_counted(ctx, true, in.c, in.c + 1, 2, n, counts + depth, nullptr, (unsigned char*)(in.c + 1), 2, q.side))) return rc;
      HC_CUDA(cudaEventRecord(q.evJoin, q.side));
      HC_CUDA(cudaStreamWaitEvent(q.s, q.evJoin, 0));
      if ((rc = stageEnd())) return rc;
    }
    const int* perm = nullptr;
    if (sortKeys > 0 && depth >= ctx->sortFromBounce)        // only single-pipeline passes sort (the buffers below are not sliced)
    {
      // material sort of the live-path queue (skipped while the queue is still in screen order and therefore coherent)
      const int sortGrid = std::min((n + HC_SORT_BLOCK - 1)/HC_SORT_BLOCK, ctx->smCount*8);
      if ((rc = stageBegin(3))) return rc;
      k_pt_sort_count<<<sortGrid, HC_SORT_BLOCK, sortKeys*sizeof(int), q.s>>>(scn, counts + depth, in.d,
                        (unsigned short*)p->sortKeys.ptr, (int*)p->sortCount.ptr, sortKeys);
      k_pt_sort_scan<<<1, 1024, 0, q.s>>>((int*)p->sortCount.ptr, (int*)p->sortCursor.ptr, sortKeys);
      k_pt_sort_scatter<<<(n + HC_SORT_BLOCK - 1)/HC_SORT_BLOCK, HC_SORT_BLOCK, 0, q.s>>>(counts + depth, (const unsigned short*)p->sortKeys.ptr,
                          (int*)p->sortCursor.ptr, (int*)p->sortPerm.ptr);
      if ((rc = stageEnd())) return rc;
      HC_CUDA(cudaGetLastError());
      ctx->stats.kernelLaunches += 3;
      perm = (const int*)p->sortPerm.ptr;
    }
    pp.depth = depth; pp.isLast = (depth == nBounces - 1) ? 1 : 0;
    if (p->haveNormalMaps)
    {
      HC_STAGE(2, (k_pt_shade<true><<<(n + HC_SHADE_BLOCK - 1)/HC_SHADE_BLOCK, HC_SHADE_BLOCK, 0, q.s>>>(scn, pp, counts + depth, counts + depth + 1, in, out,
                   qtab, (float4*)ctx->fbSum.ptr, subSums, (uint2*)ctx->pixelRng.ptr, perm)));
    }
    else
    {
      HC_STAGE(2, (k_pt_shade<false><<<(n + HC_SHADE_BLOCK - 1)/HC_SHADE_BLOCK, HC_SHADE_BLOCK, 0, q.s>>>(scn, pp, counts + depth, counts + depth + 1, in, out,
                   qtab, (float4*)ctx->fbSum.ptr, subSums, (uint2*)ctx->pixelRng.ptr, perm)));
    }
    HC_CUDA(cudaGetLastError());
    ctx->stats.kernelLaunches++;
    return HC_OK;
  };
  static int pipesEnv = -1; if (pipesEnv < 0) { const char* e = getenv("HC_PT_PIPES"); pipesEnv = e ? atoi(e) : 0; }
  auto enqueuePass = [&]() -> int
  {
    int rc = 0;
    const bool twoPipes = !qmc && sortKeys == 0 && n >= 4096 && (pipesEnv == 2 || (pipesEnv == 0 && n <= 1024*1024));
    Pipe pipes[2]; int np = 1;
    pipes[0] = Pipe{ 0, n, counts, ctx->stream, ctx->copyStream, ctx->evFork, ctx->evJoin };
    if (twoPipes && !timed)
    {
      const int nA = ((n/2 + 31)/32)*32;
      pipes[0].n = nA;
      pipes[1] = Pipe{ nA, n - nA, counts + 256, ctx->stream2, ctx->copyStream2, ctx->evFork2, ctx->evJoin2 };
      np = 2;
      HC_CUDA(cudaEventRecord(ctx->evPipeFork, ctx->stream));
      HC_CUDA(cudaStreamWaitEvent(ctx->stream2, ctx->evPipeFork, 0));
    }
    if (timed) HC_CUDA(cudaEventRecord(ctx->evStage[0], ctx->stream));
    for (int k = 0; k < np; k++) if ((rc = genPipe(pipes[k]))) return rc;
    int cur = 0;
    for (int depth = 0; depth < nBounces; depth++)
    {
      for (int k = 0; k < np; k++) if ((rc = bouncePipe(pipes[k], depth, cur))) return rc;
      cur = 1 - cur;
    }
    if (np == 2)
    {
      HC_CUDA(cudaEventRecord(ctx->evPipeJoin, ctx->stream2));
      HC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->evPipeJoin, 0));
    }
    if (pp.groupPasses > 1 && !qmc)
    {
      HC_STAGE(3, (k_fb_fold<<<(nPerPass + 255)/256, 256, 0, ctx->stream>>>((float4*)ctx->fbSum.ptr, subSums, (const int*)p->owned.ptr, nPerPass, pp.groupPasses, size_t(W)*size_t(H))));
      HC_CUDA(cudaGetLastError());
      ctx->stats.kernelLaunches++;
    }
    if (timed) HC_CUDA(cudaEventRecord(ctx->evStage[1], ctx->stream));
    return HC_OK;
  };
  // measured (scripts/gpu_graph_ab.py, ms per pass, graph / direct): C1 512x512 64 passes 0.637 / 0.681, 16 passes 0.659 / 0.690, 4 passes 0.741 / 0.729;
  // C3 1080p 64 passes 6.73 / 6.71, 4 passes 7.07 / 6.78 (capture + instantiate cost about 1 ms): worth it only for long calls
  const bool useGraph = !qmc && S == 1 && passes >= 32 && getenv("HC_PT_NO_GRAPH") == nullptr;        // (with streams the stream base is a kernel argument)
  cudaGraphExec_t gexec = nullptr;
  uint64_t launchesPerPass = 0;
  int lastGroup = 1;
  for (int pass = 0; pass < passes; )
  {
    const int m = std::min(groupMax, passes - pass);
    pp.qmcPass = ctx->passCounter;
    pp.streamBase = int(ctx->passCounter % (unsigned)S);
    pp.groupPasses = m;
    if (m > 1) { pp.pixMask = 0x00FFFFFFu; pp.subShift = 24; subSums = qmc ? nullptr : (float4*)p->subSums.ptr; }
    else       { pp.pixMask = 0x7FFFFFFFu; pp.subShift = 31; subSums = nullptr; }
    n = nPerPass*m;
    lastGroup = m;
    timed = (pass + m >= passes);
    if (useGraph && !timed)
    {
      if (gexec == nullptr)
      {
        const uint64_t l0 = ctx->stats.kernelLaunches, p0 = ctx->stats.paths;
        HC_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        rc = enqueuePass();
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        if (rc != HC_OK || e != cudaSuccess || graph == nullptr)
        {
          if (graph) cudaGraphDestroy(graph);
          cudaGetLastError();
          hc_set_error("hc_pt_pass: capturing the pass into a CUDA graph failed");
          return rc != HC_OK ? rc : HC_E_STATE;
        }
        launchesPerPass = ctx->stats.kernelLaunches - l0;
        ctx->stats.kernelLaunches = l0; ctx->stats.paths = p0;
        const cudaError_t e2 = cudaGraphInstantiate(&gexec, graph, 0);
        cudaGraphDestroy(graph);
        if (e2 != cudaSuccess) { cudaGetLastError(); hc_set_error("hc_pt_pass: cudaGraphInstantiate failed"); return HC_E_STATE; }
      }
      const cudaError_t e3 = cudaGraphLaunch(gexec, ctx->stream);
      if (e3 != cudaSuccess) { cudaGraphExecDestroy(gexec); cudaGetLastError(); hc_set_error("hc_pt_pass: cudaGraphLaunch failed"); return HC_E_STATE; }
      ctx->stats.kernelLaunches += launchesPerPass; ctx->stats.paths += (uint64_t)n;
    }
    else if ((rc = enqueuePass())) { if (gexec) cudaGraphExecDestroy(gexec); return rc; }
    ctx->passCounter += (unsigned)m;
    ctx->spp += qmc ? double(n)*ctx->worldSize/double(W*H) : double(m);
    pass += m;
  }
  if (gexec) { HC_CUDA(cudaStreamSynchronize(ctx->stream)); cudaGraphExecDestroy(gexec); }
#undef HC_STAGE
  const auto hostT1 = std::chrono::steady_clock::now();
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  const auto hostT2 = std::chrono::steady_clock::now();
  float ms = 0.0f; HC_CUDA(cudaEventElapsedTime(&ms, ctx->evStage[0], ctx->evStage[1]));
  ctx->lastTraceMs = ms;              // device time of the LAST wavefront (lastGroup passes)
  ctx->lastGroupPasses = lastGroup;
  float cls[4] = { 0, 0, 0, 0 };
  for (size_t k = 0; k < p->evClass.size(); k++)
  {
    float t = 0.0f; HC_CUDA(cudaEventElapsedTime(&t, p->evPool[2*k], p->evPool[2*k + 1]));
    cls[p->evClass[k]] += t;
  }
  if (getenv("HC_PT_LOG"))
    fprintf(stderr, "[hc_pt_pass] rank %d: %d passes, last wavefront %d passes / %d paths: host enqueue %.3f ms, wait %.3f ms, device span of the last wavefront %.3f ms\n", ctx->rank, passes,
            lastGroup, n, std::chrono::duration<double, std::milli>(hostT1 - hostT0).count(), std::chrono::duration<double, std::milli>(hostT2 - hostT1).count(), ms);
  if (getenv("HC_PT_LOG") && atoi(getenv("HC_PT_LOG")) >= 2)                    // per-launch device times of the last pass (class 0 closest, 1 shadow added, 2 shade, 3 raygen / sort)
  {
    int live[256]; HC_CUDA(cudaMemcpy(live, counts, sizeof(live), cudaMemcpyDeviceToHost));
    for (int d = 0; d < nBounces && d < 255; d++) fprintf(stderr, "[hc_pt_pass] bounce %d live %d\n", d, live[d]);
    for (size_t k = 0; k < p->evClass.size(); k++)
    {
      float t = 0.0f; HC_CUDA(cudaEventElapsedTime(&t, p->evPool[2*k], p->evPool[2*k + 1]));
      fprintf(stderr, "[hc_pt_pass] launch %zu class %d %.1f us\n", k, p->evClass[k], 1e3f*t);
    }
  }
  {
    int live[256]; HC_CUDA(cudaMemcpy(live, counts, sizeof(live), cudaMemcpyDeviceToHost));     // live paths per bounce of the last pass
    uint64_t closest = 0, shadow = 0;
    for (int d = 0; d < nBounces && d < 255; d++) { closest += uint64_t(live[d]); if (d > 0 && integrator != HC_INTEGRATOR_PT) shadow += uint64_t(live[d]); }
    const double scale = double(passes)/double(lastGroup);               // the counts are those of the last wavefront (lastGroup passes)
    ctx->stats.raysClosest += uint64_t(double(closest)*scale); ctx->stats.raysShadow += uint64_t(double(shadow)*scale);
  }
  // the per-class times are those of the last wavefront; scale to all passes of this call
  const float tscale = float(passes)/float(lastGroup);
  ctx->stats.msClosest += cls[0]*tscale; ctx->stats.msShadow += cls[1]*tscale;
  ctx->stats.msShade += cls[2]*tscale; ctx->stats.msOther += cls[3]*tscale;
  return HC_OK;
}

int hc_fb_clear(hc_ctx* ctx)
{
  if (!ctx || !ctx->fbSum.ptr) return HC_E_STATE;
  HC_CUDA(cudaSetDevice(ctx->device));
  HC_CUDA(cudaMemsetAsync(ctx->fbSum.ptr, 0, uint64_t(ctx->width)*ctx->height*16, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->spp = 0.0; ctx->combinedValid = false;
  return HC_OK;
}

int hc_fb_device_ptr(hc_ctx* ctx, float** outSumRGBA, int64_t* outFloats)
{
  if (!ctx || !outSumRGBA || !ctx->fbSum.ptr) return HC_E_STATE;
  *outSumRGBA = (float*)ctx->fbSum.ptr;
  if (outFloats) *outFloats = int64_t(ctx->width)*ctx->height*4;
  return HC_OK;
}

// the image the read-back entry points see: this rank's sums, or - on the destination rank after hc_fb_reduce of full-size buffers - the sum over ranks
static const float4* ReadSource(hc_ctx* ctx) { return (const float4*)((ctx->combinedValid && ctx->fbCombined.ptr) ? ctx->fbCombined.ptr : ctx->fbSum.ptr); }

int hc_fb_read_hdr(hc_ctx* ctx, float* outRGBA, int width, int height)
{
  if (!ctx || !outRGBA) return HC_E_ARG;
  HC_REQUIRE(width == ctx->width && height == ctx->height && ctx->fbSum.ptr, HC_E_ARG, "hc_fb_read_hdr: bad input resolution");
  HC_CUDA(cudaSetDevice(ctx->device));
  const int n = width*height;
  const float4* src = ReadSource(ctx);
  if (ctx->spp > 0.0)
  {
    // normalisation on read-back (GPUOCLLayer.cpp:1184-1215), on the device
    int rc = hc_buf_reserve(ctx, ctx->fbOut, uint64_t(n)*16); if (rc) return rc;
    k_fb_normalize<<<(n + 255)/256, 256, 0, ctx->stream>>>(src, (float4*)ctx->fbOut.ptr, n, float(1.0/ctx->spp));
    HC_CUDA(cudaGetLastError());
    ctx->stats.kernelLaunches++;
    src = (const float4*)ctx->fbOut.ptr;
  }
  HC_CUDA(cudaMemcpyAsync(outRGBA, src, size_t(n)*16, cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return HC_OK;
}

int hc_fb_read_sum(hc_ctx* ctx, float* outRGBA, int width, int height)
{
  if (!ctx || !outRGBA) return HC_E_ARG;
  HC_REQUIRE(width == ctx->width && height == ctx->height && ctx->fbSum.ptr, HC_E_ARG, "hc_fb_read_sum: bad input resolution");
  HC_CUDA(cudaSetDevice(ctx->device));
  HC_CUDA(cudaMemcpyAsync(outRGBA, ReadSource(ctx), size_t(width)*height*16, cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return HC_OK;
}

int hc_fb_read_ldr(hc_ctx* ctx, uint32_t* outRGBA8, int width, int height)
{
  if (!ctx || !outRGBA8) return HC_E_ARG;
  HC_REQUIRE(width == ctx->width && height == ctx->height && ctx->fbSum.ptr, HC_E_ARG, "hc_fb_read_ldr: bad input resolution");
  HC_CUDA(cudaSetDevice(ctx->device));
  HcPathHost* p = EnsureHost(ctx);
  const int n = width*height;
  int rc = hc_buf_reserve(ctx, p->ldr, uint64_t(n)*4); if (rc) return rc;
  const float* varsF = (const float*)(ctx->globalsHead.data() + HC_EG_varsF);
  const float gamma = varsF[HC_HRT_IMAGE_GAMMA] > 0.0f ? varsF[HC_HRT_IMAGE_GAMMA] : 2.2f;
  k_hdr_to_ldr<<<(n + 255)/256, 256, 0, ctx->stream>>>(ReadSource(ctx), (unsigned*)p->ldr.ptr, n, ctx->spp > 0.0 ? float(1.0/ctx->spp) : 1.0f, 1.0f/gamma);
  HC_CUDA(cudaGetLastError());
  ctx->stats.kernelLaunches++;
  HC_CUDA(cudaMemcpyAsync(outRGBA8, p->ldr.ptr, uint64_t(n)*4, cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return HC_OK;
}

int hc_pt_set_material_sort(hc_ctx* ctx, int enable, int fromBounce)
{
  if (!ctx || fromBounce < 0) return HC_E_ARG;
  ctx->materialSort = (enable == 2) ? 2 : (enable ? 1 : 0); ctx->sortFromBounce = fromBounce;
  return HC_OK;
}

int hc_pt_set_shadow_trees(hc_ctx* ctx, int mode)
{
  if (!ctx || (mode != 0 && mode != 1)) return HC_E_ARG;
  ctx->shadowTrees = mode;
  return HC_OK;
}

int hc_get_spp(hc_ctx* ctx, float* outSpp) { if (!ctx || !outSpp) return HC_E_ARG; *outSpp = float(ctx->spp); return HC_OK; }
} // extern "C"
