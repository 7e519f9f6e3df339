// hc_api.cu — C ABI of libhydracore_b200.so: context, MemoryStorageCUDA backing, scene upload and the ray-casting entry points.
// (Path tracing entry points live in hc_path.cu.)  Reference interfaces replaced are cited in include/hydracore_cuda.h.
#include "hc_context.h"
#include "hc_trace.cuh"
#include "hc_raygen.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <limits>

// ------------------------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_lastError;
void hc_set_error(const char* msg) { g_lastError = msg ? msg : ""; }
int hc_cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
  char buf[1024];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", int(e), cudaGetErrorString(e), file, line, what);
  g_lastError = buf;
  return int(e);
}

int hc_buf_reserve(hc_ctx* ctx, HcDevBuf& b, uint64_t bytes)
{
  (void)ctx;
  if (bytes <= b.bytes && b.ptr) return HC_OK;
  if (b.ptr) { HC_CUDA(cudaFree(b.ptr)); b.ptr = nullptr; b.bytes = 0; }
  if (bytes == 0) return HC_OK;
  cudaError_t e = cudaMalloc(&b.ptr, bytes);
  if (e != cudaSuccess) { b.ptr = nullptr; b.bytes = 0; cudaGetLastError(); hc_set_error("cudaMalloc failed"); return HC_E_NOMEM; }
  b.bytes = bytes;
  return HC_OK;
}
void hc_buf_free(HcDevBuf& b) { if (b.ptr) cudaFree(b.ptr); b.ptr = nullptr; b.bytes = 0; }

// ------------------------------------------------------------------------------------------------------------------ kernels
#define HC_TRACE_BLOCK 128
#define HC_TRACE_MINB  7       // 72 registers -> 7 CTAs (28 warps) per SM.  Measured on B200 (profiles/r02_k2_variants.md): 6 / 7 / 8 CTAs per SM give
                               // 4332 / 4278 / 4164 Mrays/s incoherent and 6676 / 6782 / 6638 shadow; throughput follows N/(N + N0) in the resident CTAs (r2g rows)
#define HC_REFILL_MIN  24      // refill a warp from the global ray counter once this many lanes are idle (swept 4..32 on B200 in both rounds: 20-24 is best:
                               // rays started together stay in phase, a warp refilled early runs its quad steps with fewer lanes: 16.8 instead of 19.6)
#define HC_QBIAS       2       // quad steps while  HC_QBIAS x (lanes at a quad) >= 2 x (lanes at a leaf)   (swept 2, 3, 4, 6, 8: 2 is best by 1-2 %)

// K2 / K2s.  Persistent warps: lanes that finished a ray are refilled together (one atomicAdd per refill, ranks by ballot/popc).
// Between refills a lane runs the while-while loop of hc_trace.cuh: descend through interior quads until a leaf is reached, then
// intersect (or enter the instance); the loop is left early once so few lanes are still busy that a refill pays off.
// rpos/rdir are float4 streams with element stride `stride` (2 = interleaved {pos,dir} records, 1 = separate arrays).
// ALPHA (closest hit only): the second, alpha-tested tree of the reference (BVH4InstTraverseAlpha after BVH4InstTraverse in
// IntegratorCommon::rayTrace, CPUExp_Integrators_Common.cpp:122-150): starts from the hit tree 0 left in hitsOut and keeps it unless a closer
// triangle passes the opacity lookup.
// outStride: 1 = results go to dense arrays (hit records of 16 B, visibility bytes); > 1 = the rays live in the path tracer's queue of 32-byte
// elements (hc_path.cu): the hit record is stored at hitsOut + ray*outStride (float4 units) and an OCCLUDED shadow ray gets its t_far zeroed
// (visOut points at the {sdir, t_far} float4 of path 0), which is how the shade kernel reads "no light from this sample".
// RAYGEN (ray-casting pass only): 0 = rays come from memory; 1 = the primary eye ray of pixel idx is generated in the fetch (K1 fused into K2:
// MakeRandEyeRay with zero offsets, the arithmetic of k_make_eye_rays); 2 = the shadow ray towards `light` is generated from the regenerated eye
// ray and the hit record hitsIn[idx] (k_make_shadow_rays fused into K2s).  Saves two launches and 2 x 64 B per pixel of ray traffic.
struct HcRayGen { HcCamera cam; int width, height; long long firstPixel; float3 light; const HcHit* hitsIn; const int* pixels = nullptr; };   // pixels: this rank's pixel list (tile partition), else pixel = index + firstPixel
template<bool ANYHIT, int TREE1 = 0, int RAYGEN = 0>      // TREE1: 0 = first tree, 1 = second tree (hit carried), 2 = second tree with the alpha table
__global__ void __launch_bounds__(HC_TRACE_BLOCK, HC_TRACE_MINB)
k_trace(const HcBvh bvh, const float4* __restrict__ rpos, const float4* __restrict__ rdir, const int stride, const long long nArg,
        const int* __restrict__ nDev, HcHit* __restrict__ hitsOut, unsigned char* __restrict__ visOut, unsigned* __restrict__ counter, const int refillMin, const int qBias,
        const int tileW, const int outStride, const HcRayGen gen = HcRayGen())
{
  const unsigned n = (unsigned)(nDev ? (long long)(*nDev) : nArg);      // the path tracer keeps its live-path count on the device
  uint2 stk[HC_STACK_CAP];                                              // {child word, entry distance}
  uint2 saved[5];                                                       // world-space ray while inside an instance

  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned ltMask = (1u << lane) - 1u;

  bool idle = true, exhausted = false;
  unsigned rayIdx = 0;
  HcRayTrav r;
  TravStart(r, f3(0, 0, 0), f3(0, 0, 1), 0.0f);
  r.node = HC_NODE_SENTINEL;

  for (;;)
  {
    const unsigned idleMask = __ballot_sync(FULL, idle);
    if (idleMask != 0u && !exhausted && (__popc(idleMask) >= refillMin || idleMask == FULL))
    {
      const unsigned nIdle = (unsigned)__popc(idleMask); const int leader = __ffs(idleMask) - 1;
      unsigned base = 0;
      if (lane == leader) base = atomicAdd(counter, nIdle);
      base = __shfl_sync(FULL, base, leader);
      unsigned idx = base + (unsigned)__popc(idleMask & ltMask);
      if (base + nIdle >= n) exhausted = true;
      if (idle && idx < n)
      {
        if (tileW > 0)
        {
          // the ray stream is a W x H image in row-major order: fetch it in 8 x 4 pixel blocks, so that a warp's 32 rays cover a
          // compact screen patch (fewer distinct BVH nodes per warp, more uniform traversal lengths) instead of a 32 x 1 strip
          const unsigned blk = idx >> 5, w = idx & 31u, bpr = (unsigned)tileW >> 3;
          idx = ((blk/bpr)*4u + (w >> 3))*(unsigned)tileW + (blk % bpr)*8u + (w & 7u);
        }
        float4 p, dd;
        if (RAYGEN == 0) { p = __ldg(rpos + size_t(idx)*stride); dd = __ldg(rdir + size_t(idx)*stride); }
        else
        {
          const long long pix = gen.pixels ? (long long)gen.pixels[idx] : (long long)idx + gen.firstPixel;      // this rank's tiles, or a band of the image
          float3 eo, ed;
          MakeRandEyeRay(int(pix % gen.width), int(pix / gen.width), gen.width, gen.height, make_float4(0.0f, 0.0f, 0.0f, 0.0f), gen.cam, eo, ed);
          p = make_float4(eo.x, eo.y, eo.z, 0.0f); dd = make_float4(ed.x, ed.y, ed.z, HC_MAXFLOAT_RAY);
          if (RAYGEN == 2)
          {
            const HcHit h = gen.hitsIn[pix];
            p = make_float4(0, 0, 0, 0); dd = make_float4(0, 1, 0, 0);             // t_far = 0: "no shadow ray" for pixels that hit nothing
            if (h.primId != -1)
            {
              const float3 pos = eo + ed*h.t;
              const float3 sdir = normalize(gen.light - pos);
              const float eps = fmaxf(fmaxf(fabsf(pos.x), fmaxf(fabsf(pos.y), fabsf(pos.z))), 1.0f)*1e-4f;
              const float3 spos = pos + sdir*eps;
              p = make_float4(spos.x, spos.y, spos.z, 0.0f);
              dd = make_float4(sdir.x, sdir.y, sdir.z, length(spos - gen.light)*0.995f);
            }
          }
        }
        if (RAYGEN != 0 && gen.pixels) idx = (unsigned)gen.pixels[idx];             // results are stored per PIXEL (the gather to the destination rank packs them again)
        rayIdx = idx; idle = false;
        TravStart(r, f3(p), f3(dd), ANYHIT ? dd.w : HC_MAXFLOAT, bvh.singleLevel);
        if (TREE1 != 0 && !ANYHIT)
        {
          const float4 h = reinterpret_cast<const float4*>(hitsOut)[size_t(idx)*outStride];      // Lite_Hit carried from tree to tree
          r.t = h.x; r.primId = __float_as_int(h.y); r.hitInst = __float_as_int(h.z); r.geomId = __float_as_int(h.w);
        }
        if (ANYHIT && TREE1 != 0 && outStride == 1 && visOut[idx] == 0) { idle = true; r.node = HC_NODE_SENTINEL; }   // already occluded in the first tree
        else if (ANYHIT && !(dd.w > 0.0f)) { if (outStride == 1) visOut[idx] = 1; idle = true; r.node = HC_NODE_SENTINEL; }     // maxDist <= 0: lit (trace.cl:343-351); in a path record an occluded ray has t_far = 0 already
        else if (!RayIsFinite(r.o, r.d)) r.node = HC_NODE_SENTINEL;              // every comparison of the reference fails on NaN: no hit
      }
    }
    if (__all_sync(FULL, idle)) { if (exhausted) break; else continue; }

    // Vote scheduling: every lane is at an interior quad (Q), at a leaf (L: instance leaf or triangle-pair record) or has nothing
    // to do.  The warp runs the step the majority is waiting for, and keeps doing so until a refill pays off.
    for (;;)
    {
      bool wantQ = !(r.node & HC_LEAF_BIT);                                        // a finished / idle lane carries the sentinel (leaf bit set)
      bool wantL = !wantQ && r.node != HC_NODE_SENTINEL;
      unsigned mQ = __ballot_sync(FULL, wantQ), mL = __ballot_sync(FULL, wantL);
      while (mQ != 0u && qBias*__popc(mQ) >= 2*__popc(mL))
      {
        if (wantQ) HC_QUAD(r, bvh, stk, saved)
        wantQ = !(r.node & HC_LEAF_BIT); wantL = !wantQ && r.node != HC_NODE_SENTINEL;
        mQ = __ballot_sync(FULL, wantQ); mL = __ballot_sync(FULL, wantL);
      }
      while (mL != 0u && __popc(mL) > __popc(mQ))
      {
        if (wantL)
        {
          if (r.instId < 0) HC_ENTER(r, bvh, saved)
          else
          {
            // IntersectAllPrimitivesInLeaf, ONE pair record per step; the leaf word is the cursor (index up, count down)
            const size_t pairIndex = size_t(r.node & HC_LEAF_INDEX_MASK);
            const bool done = ((r.node >> HC_LEAF_PAIRS_SHIFT) & 63u) == 0u;
            r.node = r.node - (1u << HC_LEAF_PAIRS_SHIFT) + 1u;
            const bool found = PairTest<TREE1 == 2>(r, bvh, pairIndex);
            if (ANYHIT && found) r.node = HC_NODE_SENTINEL;
            else if (done) HC_POP(r, stk, saved)
          }
        }
        wantQ = !(r.node & HC_LEAF_BIT); wantL = !wantQ && r.node != HC_NODE_SENTINEL;
        mQ = __ballot_sync(FULL, wantQ); mL = __ballot_sync(FULL, wantL);
      }
      const int busy = __popc(mQ | mL);
      if (busy == 0 || (!exhausted && busy <= 32 - refillMin)) break;         // all done, or enough idle lanes for a refill
    }

    if (!idle && r.node == HC_NODE_SENTINEL)
    {
      if (ANYHIT)
      {
        if (outStride == 1) visOut[rayIdx] = (r.primId != -1) ? 0 : 1;
        else if (r.primId != -1) reinterpret_cast<float*>(visOut)[size_t(rayIdx)*outStride*4 + 3] = 0.0f;
      }
      else reinterpret_cast<float4*>(hitsOut)[size_t(rayIdx)*outStride] = make_float4(r.t, __int_as_float(r.primId), __int_as_float(r.hitInst), __int_as_float(r.geomId));
      idle = true;
    }
  }
}

// shadow rays from closest hits toward one point light (ray-casting benchmark / debug path of the LightSample stage)
static __global__ void k_make_shadow_rays(const float4* __restrict__ rays, const HcHit* __restrict__ hits, const long long n, const float3 L, float4* __restrict__ out)
{
  const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  const HcHit h = hits[i];
  float4 o = make_float4(0, 0, 0, 0), d = make_float4(0, 1, 0, 0);            // d.w = t_far = 0: "no shadow ray"
  if (h.primId != -1)
  {
    const float3 ro = f3(rays[2*i]), rd = f3(rays[2*i + 1]);
    const float3 pos = ro + rd*h.t;
    const float3 sdir = normalize(L - pos);
    const float eps = fmaxf(fmaxf(fabsf(pos.x), fmaxf(fabsf(pos.y), fabsf(pos.z))), 1.0f)*1e-4f;
    const float3 spos = pos + sdir*eps;
    o = make_float4(spos.x, spos.y, spos.z, 0.0f);
    d = make_float4(sdir.x, sdir.y, sdir.z, length(spos - L)*0.995f);
  }
  out[2*i] = o; out[2*i + 1] = d;
}

// ------------------------------------------------------------------------------------------------------------------ helpers
static int TraceGrid(hc_ctx* ctx)
{
  if (ctx->traceGrid > 0) return ctx->traceGrid;
  int perSM = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_trace<false>, HC_TRACE_BLOCK, 0);
  if (perSM < 1) perSM = 1;
  if (const char* e = getenv("HC_TRACE_CTAS")) { const int v = atoi(e); if (v >= 1 && v < perSM) perSM = v; }      // experiments: fewer resident CTAs per SM
  ctx->traceGrid = ctx->smCount*perSM;           // a whole number of waves: persistent CTAs, all resident
  return ctx->traceGrid;
}

// one persistent-thread ray counter per launch in flight: rotate through the counter block so that back-to-back launches never share one
static int NextCounter(hc_ctx* ctx, cudaStream_t stream, unsigned** out)
{
  ctx->traceCounterSlot = (ctx->traceCounterSlot + 1) % 32;
  *out = (unsigned*)((unsigned long long*)ctx->counters.ptr + ctx->traceCounterSlot);
  HC_CUDA(cudaMemsetAsync(*out, 0, sizeof(unsigned long long), stream));
  return HC_OK;
}

// launch K2 (closest) or K2s (any-hit) on device-resident streams; used by hc_trace_* and by the path tracer
static int LaunchTrace(hc_ctx* ctx, bool anyHit, const float4* rpos, const float4* rdir, int stride, long long n, const int* nDev, HcHit* hits, unsigned char* vis, int tileW = 0, cudaStream_t stream = nullptr, int outStride = 1);
int hc_launch_trace(hc_ctx* ctx, bool anyHit, const float4* rpos, const float4* rdir, int stride, long long n, HcHit* hits, unsigned char* vis)
{ return LaunchTrace(ctx, anyHit, rpos, rdir, stride, n, nullptr, hits, vis); }
int hc_launch_trace_counted(hc_ctx* ctx, bool anyHit, const float4* rpos, const float4* rdir, int stride, long long nUpper, const int* nDev, HcHit* hits, unsigned char* vis, int outStride, cudaStream_t stream)
{ return LaunchTrace(ctx, anyHit, rpos, rdir, stride, nUpper, nDev, hits, vis, 0, stream, outStride); }
static int LaunchTrace(hc_ctx* ctx, bool anyHit, const float4* rpos, const float4* rdir, int stride, long long n, const int* nDev, HcHit* hits, unsigned char* vis, int tileW, cudaStream_t stream, int outStride)
{
  if (n <= 0) return HC_OK;
  if (!stream) stream = ctx->stream;
  HC_REQUIRE(ctx->bvhNodes.ptr && ctx->bvhTris.ptr, HC_E_STATE, "hc_trace: no BVH uploaded (hc_set_bvh)");
  HC_REQUIRE(n < 0xffff0000ll, HC_E_ARG, "hc_trace: more than 2^32 rays in one launch");
  HcBvh bvh; bvh.nodes = (const float4*)ctx->bvhNodes.ptr; bvh.tris = (const float4*)ctx->bvhTris.ptr; bvh.singleLevel = ctx->haveInst ? 0 : 1;
  unsigned* counter = nullptr;
  int rc = NextCounter(ctx, stream, &counter); if (rc) return rc;
  const int grid = (int)std::min<long long>(TraceGrid(ctx), (n + HC_TRACE_BLOCK - 1)/HC_TRACE_BLOCK);
  const int rf = ctx->traceRefill > 0 ? ctx->traceRefill : HC_REFILL_MIN, qb = ctx->traceQBias > 0 ? ctx->traceQBias : HC_QBIAS;
  if (anyHit) k_trace<true><<<grid, HC_TRACE_BLOCK, 0, stream>>>(bvh, rpos, rdir, stride, n, nDev, nullptr, vis, counter, rf, qb, tileW, outStride);
  else        k_trace<false><<<grid, HC_TRACE_BLOCK, 0, stream>>>(bvh, rpos, rdir, stride, n, nDev, hits, nullptr, counter, rf, qb, tileW, outStride);
  HC_CUDA(cudaGetLastError());
  ctx->stats.kernelLaunches++;
  if (ctx->haveTree1 && (!anyHit || ctx->shadowTrees == 1))
  {
    // IntegratorCommon::rayTrace walks the trees one after another with the hit carried along (CPUExp_Integrators_Common.cpp:131-147).
    // Shadow rays: the CPU integrators' shadowTrace looks at tree 0 only (:163-171: meshes with opacity maps cast no shadows there), the
    // OpenCL layer walks every tree with the opacity test (GPUOCLKernels.cpp:959-1000, BVH4InstTraverseShadowAlphaS ctrace.h:1748).
    // hc_pt_set_shadow_trees chooses: 1 (default) = every tree, a cut-out occludes where its opacity texel passes the closest-hit test
    // (binary; the smooth-opacity and skip-shadow flags of the OpenCL path are not in the alpha words); 0 = first tree only (oracle parity)
    HcBvh b1; b1.nodes = (const float4*)ctx->bvh1Nodes.ptr; b1.tris = (const float4*)ctx->bvh1Tris.ptr; b1.singleLevel = ctx->haveInst1 ? 0 : 1;
    unsigned* counter1 = nullptr;
    rc = NextCounter(ctx, stream, &counter1); if (rc) return rc;
    if (ctx->haveAlpha1)
    {
      int texTab = 0; memcpy(&texTab, ctx->globalsHead.data() + HC_EG_texturesTableOffset, 4);
      HC_REQUIRE(ctx->globals.ptr && ctx->storage[HC_STORAGE_TEXTURES].ptr, HC_E_STATE, "hc_trace: the alpha-tested tree needs the textures storage and the globals (texture table)");
      b1.alphaPairs = (const uint4*)ctx->bvh1AlphaPairs.ptr; b1.alphaTable = (const uint2*)ctx->bvh1AlphaTable.ptr;
      b1.textures = (const int4*)ctx->storage[HC_STORAGE_TEXTURES].ptr; b1.texturesTable = (const int*)ctx->globals.ptr + texTab;
      if (anyHit) k_trace<true, 2><<<grid, HC_TRACE_BLOCK, 0, stream>>>(b1, rpos, rdir, stride, n, nDev, nullptr, vis, counter1, rf, qb, tileW, outStride);
      else        k_trace<false, 2><<<grid, HC_TRACE_BLOCK, 0, stream>>>(b1, rpos, rdir, stride, n, nDev, hits, nullptr, counter1, rf, qb, tileW, outStride);
    }
    else if (anyHit) k_trace<true, 1><<<grid, HC_TRACE_BLOCK, 0, stream>>>(b1, rpos, rdir, stride, n, nDev, nullptr, vis, counter1, rf, qb, tileW, outStride);
    else             k_trace<false, 1><<<grid, HC_TRACE_BLOCK, 0, stream>>>(b1, rpos, rdir, stride, n, nDev, hits, nullptr, counter1, rf, qb, tileW, outStride);
    HC_CUDA(cudaGetLastError());
    ctx->stats.kernelLaunches++;
  }
  if (!nDev) { if (anyHit) ctx->stats.raysShadow += (uint64_t)n; else ctx->stats.raysClosest += (uint64_t)n; }   // counted launches: hc_pt_pass reads the live counts back
  return HC_OK;
}

// ray-casting pass: K2 / K2s with the eye / shadow rays generated in the fetch (RAYGEN 1 / 2); first tree only (the caller falls back to rays in
// memory when a second tree is present)
static int LaunchTraceGen(hc_ctx* ctx, bool shadow, long long n, HcHit* hitsLocal, unsigned char* vis, int tileW, const HcRayGen& gen)
{
  if (n <= 0) return HC_OK;
  cudaStream_t stream = ctx->stream;
  HcBvh bvh; bvh.nodes = (const float4*)ctx->bvhNodes.ptr; bvh.tris = (const float4*)ctx->bvhTris.ptr; bvh.singleLevel = ctx->haveInst ? 0 : 1;
  unsigned* counter = nullptr;
  int rc = NextCounter(ctx, stream, &counter); if (rc) return rc;
  const int grid = (int)std::min<long long>(TraceGrid(ctx), (n + HC_TRACE_BLOCK - 1)/HC_TRACE_BLOCK);
  const int rf = ctx->traceRefill > 0 ? ctx->traceRefill : HC_REFILL_MIN, qb = ctx->traceQBias > 0 ? ctx->traceQBias : HC_QBIAS;
  if (shadow) k_trace<true, 0, 2><<<grid, HC_TRACE_BLOCK, 0, stream>>>(bvh, nullptr, nullptr, 0, n, nullptr, nullptr, vis, counter, rf, qb, tileW, 1, gen);
  else        k_trace<false, 0, 1><<<grid, HC_TRACE_BLOCK, 0, stream>>>(bvh, nullptr, nullptr, 0, n, nullptr, hitsLocal, nullptr, counter, rf, qb, tileW, 1, gen);
  HC_CUDA(cudaGetLastError());
  ctx->stats.kernelLaunches++;
  if (shadow) ctx->stats.raysShadow += (uint64_t)n; else ctx->stats.raysClosest += (uint64_t)n;
  return HC_OK;
}

// Walk the uploaded reference-layout tree (producer bvh_access_dll2.cpp:199-717) on the host: validate every
// offset, bound the traversal stack, and re-lay it out for the device (formats in hc_trace.cuh).  Quads and instance records keep
// their quad index, so child words of interior nodes are unchanged.
//   interior quad q : float4[8q+0..5] = minx[4] maxx[4] miny[4] maxy[4] minz[4] maxz[4], uint4[8q+6] = child words
//   instance record : float4[8q+0..3] = inverse matrix columns, [8q+4] = {sub-tree word, realInstId, meshId, 0}
//   triangle leaf   : pair records of 6 float4 {Ax0 Ax1 Ay0 Ay1 | Az0 Az1 E1x0 E1x1 | E1y0 E1y1 E1z0 E1z1 | E2x0 E2x1 E2y0 E2y1 |
//                     E2z0 E2z1 prim0 prim1 | geom0 geom1 0 0} with E1 = B - A, E2 = C - A evaluated in float exactly as
//                     IntersectAllPrimitivesInLeaf does (ctrace.h:159-160); an odd leaf is padded with a zero triangle, whose
//                     determinant is 0 -> v = u = t = NaN -> every acceptance test fails.
static int ConvertBvhForDevice(const unsigned char* nodes, int nodesNum, const float* trif4, int trif4Num,
                               std::vector<float>& outNodes, std::vector<float>& outPairs, int* outStackBound,
                               const unsigned* alphaU2 = nullptr, int alphaNum = 0, std::vector<unsigned>* outAlphaPairs = nullptr, bool singleLevel = false)
{
  struct N { float bmin[3]; unsigned lo; float bmax[3]; unsigned esc; };
  const N* nd = (const N*)nodes;
  const int quads = nodesNum/4;
  if (quads < 2) return HC_E_ARG;
  outNodes.assign(size_t(quads)*32, 0.0f);
  outPairs.clear();
  outPairs.reserve(size_t(trif4Num)*4 + 64);
  std::vector<unsigned> leafWord(size_t(trif4Num), 0u);     // float4 offset of a leaf header -> converted child word (0 = not yet)
  const float INF = std::numeric_limits<float>::infinity();

  auto convertLeaf = [&](unsigned off, unsigned* word) -> int
  {
    if (off >= unsigned(trif4Num)) return HC_E_RANGE;
    if (leafWord[off]) { *word = leafWord[off]; return HC_OK; }
    int hdr[4]; memcpy(hdr, trif4 + size_t(off)*4, 16);
    const int first = hdr[0], count = hdr[1];
    if (first < 0 || count < 1 || size_t(first) + size_t(count)*3 > size_t(trif4Num)) return HC_E_RANGE;
    const int pairs = (count + 1)/2;
    if (pairs > HC_LEAF_PAIRS_MAX) return HC_E_RANGE;
    const size_t index = outPairs.size()/(HC_PAIR_F4*4);
    if (index + size_t(pairs) >= HC_LEAF_INDEX_MASK) return HC_E_RANGE;
    outPairs.resize(outPairs.size() + size_t(pairs)*HC_PAIR_F4*4, 0.0f);
    float* P = outPairs.data() + index*HC_PAIR_F4*4;
    if (outAlphaPairs)
    {
      // per triangle {alpha[o].x (sampler offset), alpha[o].y, alpha[o+1].y, alpha[o+2].y (packed uv)}; the empty half of an odd pair never hits
      if (size_t(first) + size_t(count)*3 > size_t(alphaNum)) return HC_E_RANGE;
      outAlphaPairs->resize((index + size_t(pairs))*8, 0xFFFFFFFFu);
      for (int k = 0; k < count; k++)
      {
        const unsigned* a = alphaU2 + (size_t(first) + size_t(k)*3)*2;
        unsigned* o = outAlphaPairs->data() + (index + size_t(k/2))*8 + size_t(k & 1)*4;
        o[0] = a[0]; o[1] = a[1]; o[2] = a[3]; o[3] = a[5];
      }
    }
    for (int k = 0; k < count; k++)
    {
      const float* A = trif4 + (size_t(first) + size_t(k)*3)*4; const float* B = A + 4; const float* C = A + 8;
      float* R = P + size_t(k/2)*HC_PAIR_F4*4; const int s = k & 1;
      const float e1[3] = { B[0] - A[0], B[1] - A[1], B[2] - A[2] }, e2[3] = { C[0] - A[0], C[1] - A[1], C[2] - A[2] };
      R[0 + s] = A[0];  R[2 + s] = A[1];  R[4 + s] = A[2];
      R[6 + s] = e1[0]; R[8 + s] = e1[1]; R[10 + s] = e1[2];
      R[12 + s] = e2[0]; R[14 + s] = e2[1]; R[16 + s] = e2[2];
      memcpy(&R[18 + s], &A[3], 4);            // primId
      memcpy(&R[20 + s], &B[3], 4);            // geomId
      memcpy(&R[22 + s], &C[3], 4);            // instId (single-level trees, IntersectAllPrimitivesInLeaf1 ctrace.h:101; -1 in the two-level layout)
    }
    if (count & 1) { float* R = P + size_t(count/2)*HC_PAIR_F4*4; const int m1 = -1; memcpy(&R[19], &m1, 4); memcpy(&R[21], &m1, 4); memcpy(&R[23], &m1, 4); }
    *word = leafWord[off] = HC_LEAF_BIT | (unsigned(pairs - 1) << HC_LEAF_PAIRS_SHIFT) | unsigned(index);
    return HC_OK;
  };

  struct It { unsigned quad; int depth; bool inst; };
  std::vector<It> st; st.push_back({ 1u, 1, singleLevel });       // single-level trees (no instance records): the tree at quad 1 is a mesh tree
  std::vector<unsigned char> seen(size_t(quads), 0);
  int maxTop = 0, maxMesh = 0;
  while (!st.empty())
  {
    It it = st.back(); st.pop_back();
    if (it.quad >= unsigned(quads) || it.quad == 0) return HC_E_RANGE;
    if (it.inst) maxMesh = std::max(maxMesh, it.depth); else maxTop = std::max(maxTop, it.depth);
    if (seen[it.quad]) continue;     // shared mesh sub-trees: depth of first visit is representative
    seen[it.quad] = 1;
    float* Q = outNodes.data() + size_t(it.quad)*32;
    unsigned words[4];
    for (int i = 0; i < 4; i++)
    {
      const N& c = nd[size_t(it.quad)*4 + i];
      if (c.lo == 0xffffffffu && c.esc == 0xffffffffu)        // IsValidNode (cglobals.h:1321): an x slab at +inf fails for every finite ray
      {
        Q[0 + i] = INF; Q[4 + i] = INF; words[i] = HC_NODE_SENTINEL;
        continue;
      }
      Q[0 + i] = c.bmin[0]; Q[4 + i] = c.bmax[0]; Q[8 + i] = c.bmin[1]; Q[12 + i] = c.bmax[1]; Q[16 + i] = c.bmin[2]; Q[20 + i] = c.bmax[2];
      const unsigned off = c.lo & 0x7fffffffu;
      if (c.lo & 0x80000000u)
      {
        if (!it.inst)
        {
          if (off >= unsigned(quads) || off == 0) return HC_E_RANGE;
          words[i] = HC_LEAF_BIT | off;
          if (seen[off]) continue;
          seen[off] = 1;
          const N* rec = nd + size_t(off)*4;
          float* R = outNodes.data() + size_t(off)*32;
          memcpy(R, (const float*)rec + 8, 64);                // inverse matrix: float4 8Q+2 .. 8Q+5
          unsigned sub = rec[0].lo & 0x7fffffffu, subWord;
          if (rec[0].lo & 0x80000000u) { int rc = convertLeaf(sub, &subWord); if (rc) return rc; maxMesh = std::max(maxMesh, 1); }
          else { subWord = sub; st.push_back({ sub, 1, true }); }
          memcpy(R + 16, &subWord, 4);
          memcpy(R + 17, (const float*)rec + 24, 8);           // {realInstId, meshId}: float4 8Q+6 .xy
        }
        else { int rc = convertLeaf(off, &words[i]); if (rc) return rc; }
      }
      else { words[i] = off; st.push_back({ off, it.depth + 1, it.inst }); }
    }
    memcpy(Q + 24, words, 16);
  }
  *outStackBound = 3*(maxTop + maxMesh) + 2;      // single-level trees: maxMesh only (maxTop stays 0)
  return HC_OK;
}

// ------------------------------------------------------------------------------------------------------------------ C ABI
extern "C"
{
int hc_abi_version(void) { return HC_ABI_VERSION; }
const char* hc_last_error(void) { return g_lastError.c_str(); }

int hc_device_count(int* outCount)
{
  if (!outCount) return HC_E_ARG;
  int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); *outCount = 0; return hc_cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
  *outCount = n; return HC_OK;
}

int hc_ctx_create(int device, hc_ctx** out)
{
  if (!out) return HC_E_ARG;
  *out = nullptr;
  int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { cudaGetLastError(); hc_set_error("hc_ctx_create: no CUDA device (this layer has no CPU fallback)"); return HC_E_NODEVICE; }
  HC_REQUIRE(device >= 0 && device < n, HC_E_ARG, "hc_ctx_create: bad device id");
  HC_CUDA(cudaSetDevice(device));
  hc_ctx* c = new hc_ctx;
  c->device = device;
  HC_CUDA(cudaGetDeviceProperties(&c->prop, device));
  c->smCount = c->prop.multiProcessorCount;
  HC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  HC_CUDA(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
  HC_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking)); HC_CUDA(cudaStreamCreateWithFlags(&c->copyStream2, cudaStreamNonBlocking));
  HC_CUDA(cudaEventCreateWithFlags(&c->evFork2, cudaEventDisableTiming)); HC_CUDA(cudaEventCreateWithFlags(&c->evJoin2, cudaEventDisableTiming));
  HC_CUDA(cudaEventCreateWithFlags(&c->evPipeFork, cudaEventDisableTiming)); HC_CUDA(cudaEventCreateWithFlags(&c->evPipeJoin, cudaEventDisableTiming));
  HC_CUDA(cudaEventCreateWithFlags(&c->evCopy, cudaEventDisableTiming));
  HC_CUDA(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming)); HC_CUDA(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
  HC_CUDA(cudaEventCreate(&c->ev0)); HC_CUDA(cudaEventCreate(&c->ev1));
  for (int i = 0; i < 5; i++) HC_CUDA(cudaEventCreate(&c->evStage[i]));
  int rc = hc_buf_reserve(c, c->counters, 64*sizeof(unsigned long long));
  if (rc != HC_OK) { delete c; return rc; }
  HC_CUDA(cudaMemsetAsync(c->counters.ptr, 0, c->counters.bytes, c->stream));
  c->globalsHead.assign(HC_EG_HEAD_BYTES, 0);
  if (const char* e = getenv("HC_TRACE_REFILL")) { const int v = atoi(e); if (v >= 1 && v <= 32) c->traceRefill = v; }
  if (const char* e = getenv("HC_TRACE_QBIAS")) { const int v = atoi(e); if (v >= 2 && v <= 64) c->traceQBias = v; }      // < 2 would leave states where neither step runs
  *out = c;
  return HC_OK;
}

void hc_ctx_destroy(hc_ctx* c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  hc_path_free(c);
  hc_comm_free(c);
  hc_buf_free(c->fbOut);
  for (int i = 0; i < HC_STORAGE_COUNT; i++) hc_buf_free(c->storage[i]);
  hc_buf_free(c->globals); hc_buf_free(c->bvhNodes); hc_buf_free(c->bvhTris); hc_buf_free(c->bvh1Nodes); hc_buf_free(c->bvh1Tris); hc_buf_free(c->bvh1AlphaPairs); hc_buf_free(c->bvh1AlphaTable); hc_buf_free(c->remapLists); hc_buf_free(c->remapTable); hc_buf_free(c->remapInst); hc_buf_free(c->instMatrices); hc_buf_free(c->instLightIds);
  hc_buf_free(c->fbSum); hc_buf_free(c->scratchRays); hc_buf_free(c->scratchOut); hc_buf_free(c->counters); hc_buf_free(c->pixelRng); hc_buf_free(c->qmcTable);
  hc_buf_free(c->rcRays); hc_buf_free(c->rcHits); hc_buf_free(c->rcSRays); hc_buf_free(c->rcVis);
  for (int i = 0; i < 5; i++) cudaEventDestroy(c->evStage[i]);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->stream2); cudaStreamDestroy(c->copyStream2); cudaEventDestroy(c->evFork2); cudaEventDestroy(c->evJoin2); cudaEventDestroy(c->evPipeFork); cudaEventDestroy(c->evPipeJoin);
  cudaStreamDestroy(c->copyStream); cudaEventDestroy(c->evCopy); cudaEventDestroy(c->evFork); cudaEventDestroy(c->evJoin);
  delete c;
}

int hc_device_name(hc_ctx* ctx, char* buf, int bufSize)
{
  if (!ctx || !buf || bufSize <= 0) return HC_E_ARG;
  snprintf(buf, size_t(bufSize), "%s (sm_%d%d, %d SMs)", ctx->prop.name, ctx->prop.major, ctx->prop.minor, ctx->smCount);
  return HC_OK;
}

int hc_mem_info(hc_ctx* ctx, size_t* freeBytes, size_t* totalBytes)
{
  if (!ctx) return HC_E_ARG;
  HC_CUDA(cudaSetDevice(ctx->device));
  size_t f = 0, t = 0; HC_CUDA(cudaMemGetInfo(&f, &t));
  if (freeBytes) *freeBytes = f; if (totalBytes) *totalBytes = t;
  return HC_OK;
}

int hc_sync(hc_ctx* ctx) { if (!ctx) return HC_E_ARG; HC_CUDA(cudaStreamSynchronize(ctx->stream)); return HC_OK; }
int hc_stream(hc_ctx* ctx, void** outCudaStream) { if (!ctx || !outCudaStream) return HC_E_ARG; *outCudaStream = (void*)ctx->stream; return HC_OK; }

// ---- MemoryStorageCUDA backing
int hc_storage_reserve(hc_ctx* ctx, int slot, uint64_t bytes)
{
  if (!ctx || slot < 0 || slot >= HC_STORAGE_COUNT) return HC_E_ARG;
  HC_CUDA(cudaSetDevice(ctx->device));
  HcDevBuf& b = ctx->storage[slot];
  if (b.bytes == bytes && b.ptr) return HC_OK;              // MemoryStorageOCL::Reserve keeps an equal-size buffer (MemoryStorageOCL.cpp:12-13)
  ctx->sceneDirty = true;
  if (slot == HC_STORAGE_TEXTURES || slot == HC_STORAGE_TEXTURES_AUX) ctx->texturesDirty = true;
  if (slot == HC_STORAGE_MATERIALS) ctx->materialsMirror.clear();       // a re-allocated storage starts empty
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  hc_buf_free(b);
  return hc_buf_reserve(ctx, b, std::max<uint64_t>(bytes, 16));
}

int hc_storage_write(hc_ctx* ctx, int slot, uint64_t offsetBytes, const void* data, uint64_t bytes)
{
  if (!ctx || slot < 0 || slot >= HC_STORAGE_COUNT || (!data && bytes)) return HC_E_ARG;
  HcDevBuf& b = ctx->storage[slot];
  HC_REQUIRE(offsetBytes + bytes <= b.bytes, HC_E_RANGE, "hc_storage_write: beyond reserved capacity");
  if (bytes == 0) return HC_OK;
  ctx->sceneDirty = true;
  if (slot == HC_STORAGE_TEXTURES || slot == HC_STORAGE_TEXTURES_AUX) ctx->texturesDirty = true;
  if (slot == HC_STORAGE_MATERIALS)
  {
    if (ctx->materialsMirror.size() < b.bytes) ctx->materialsMirror.resize(b.bytes, 0);
    memcpy(ctx->materialsMirror.data() + offsetBytes, data, bytes);
  }
  HC_CUDA(cudaSetDevice(ctx->device));
  HC_CUDA(cudaMemcpyAsync((char*)b.ptr + offsetBytes, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));              // blocking write, as clEnqueueWriteBuffer(CL_TRUE) (MemoryStorageOCL.cpp:55)
  return HC_OK;
}

int hc_storage_capacity(hc_ctx* ctx, int slot, uint64_t* outBytes)
{
  if (!ctx || slot < 0 || slot >= HC_STORAGE_COUNT || !outBytes) return HC_E_ARG;
  *outBytes = ctx->storage[slot].bytes; return HC_OK;
}

// ---- scene upload
int hc_set_globals(hc_ctx* ctx, const void* blob, uint64_t bytes)
{
  if (!ctx || !blob) return HC_E_ARG;
  HC_REQUIRE(bytes >= HC_EG_sizeof, HC_E_ARG, "hc_set_globals: blob smaller than EngineGlobals");
  HC_CUDA(cudaSetDevice(ctx->device));
  int rc = hc_buf_reserve(ctx, ctx->globals, bytes); if (rc) return rc;
  HC_CUDA(cudaMemcpyAsync(ctx->globals.ptr, blob, bytes, cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(ctx->globalsHead.data(), blob, HC_EG_HEAD_BYTES);
  ctx->globalsMirror.assign((const unsigned char*)blob, (const unsigned char*)blob + bytes);
  ctx->sceneDirty = true; ctx->texturesDirty = true;        // the texture tables live in the globals blob
  return HC_OK;
}

static int SetBvhTree(hc_ctx* ctx, int treeId, const void* nodes, int nodesNum, const void* trif4, int trif4Num, const void* alphaTable, int alphaNum, int haveInst)
{
  if (!ctx || !nodes || !trif4 || nodesNum < 8 || trif4Num < 4) return HC_E_ARG;
  HC_REQUIRE(treeId == 0 || treeId == 1, HC_E_ARG, "hc_set_bvh: tree 0 (opaque geometry) and tree 1 (meshes with opacity maps) are supported");
  HC_REQUIRE(treeId == 1 || alphaTable == nullptr, HC_E_ARG, "hc_set_bvh: an alpha table on tree 0 is not supported (the driver puts opacity meshes into tree 1)");
  HC_REQUIRE(alphaTable == nullptr || alphaNum >= trif4Num, HC_E_ARG, "hc_set_bvh_alpha: the alpha table must cover every float4 of the triangle list");
  int bound = 0;
  ctx->sceneDirty = true;
  std::vector<float> devNodes, devPairs;
  std::vector<unsigned> devAlpha;
  int rc = ConvertBvhForDevice((const unsigned char*)nodes, nodesNum, (const float*)trif4, trif4Num, devNodes, devPairs, &bound,
                               (const unsigned*)alphaTable, alphaNum, alphaTable ? &devAlpha : nullptr, haveInst == 0);
  HC_REQUIRE(rc == HC_OK, rc, "hc_set_bvh: tree references nodes or triangles out of range (or a leaf holds more than 128 triangles)");
  HC_REQUIRE(bound <= HC_STACK_CAP, HC_E_RANGE, "hc_set_bvh: tree too deep for the traversal stack");
  HC_CUDA(cudaSetDevice(ctx->device));
  HcDevBuf& bn = treeId ? ctx->bvh1Nodes : ctx->bvhNodes;
  HcDevBuf& bt = treeId ? ctx->bvh1Tris : ctx->bvhTris;
  rc = hc_buf_reserve(ctx, bn, devNodes.size()*4); if (rc) return rc;
  rc = hc_buf_reserve(ctx, bt, std::max<size_t>(devPairs.size()*4, 96)); if (rc) return rc;
  HC_CUDA(cudaMemcpyAsync(bn.ptr, devNodes.data(), devNodes.size()*4, cudaMemcpyHostToDevice, ctx->stream));
  if (!devPairs.empty()) HC_CUDA(cudaMemcpyAsync(bt.ptr, devPairs.data(), devPairs.size()*4, cudaMemcpyHostToDevice, ctx->stream));
  if (treeId == 1)
  {
    ctx->haveAlpha1 = false;
    ctx->alphaTexIdsHost.clear();
    if (alphaTable)
    {
      // every triangle's sampler offset must point at a whole SWTexSampler (6 uint2) inside the table; remember the texture ids for hc_pt_init
      const unsigned* au = (const unsigned*)alphaTable;
      for (size_t k = 0; k + 3 < devAlpha.size(); k += 4)
      {
        const unsigned off = devAlpha[k];
        if (off == 0xFFFFFFFFu || (int)off <= 0) continue;
        HC_REQUIRE((long long)off + 6 <= (long long)alphaNum, HC_E_RANGE, "hc_set_bvh_alpha: an opacity sampler offset points outside the alpha table");
        const int texId = (int)au[2*(size_t(off) + 1)];
        if (std::find(ctx->alphaTexIdsHost.begin(), ctx->alphaTexIdsHost.end(), texId) == ctx->alphaTexIdsHost.end()) ctx->alphaTexIdsHost.push_back(texId);
      }
      rc = hc_buf_reserve(ctx, ctx->bvh1AlphaPairs, std::max<size_t>(devAlpha.size()*4, 16)); if (rc) return rc;
      rc = hc_buf_reserve(ctx, ctx->bvh1AlphaTable, size_t(alphaNum)*8); if (rc) return rc;
      if (!devAlpha.empty()) HC_CUDA(cudaMemcpyAsync(ctx->bvh1AlphaPairs.ptr, devAlpha.data(), devAlpha.size()*4, cudaMemcpyHostToDevice, ctx->stream));
      HC_CUDA(cudaMemcpyAsync(ctx->bvh1AlphaTable.ptr, alphaTable, size_t(alphaNum)*8, cudaMemcpyHostToDevice, ctx->stream));
      ctx->haveAlpha1 = true;
    }
  }
  HC_CUDA(cudaStreamSynchronize(ctx->stream));                // ConvertionResult pointers die at ConvertUnmap (RenderDriverRTE.cpp:1436)
  if (treeId == 0) ctx->alphaTexIdsHost.clear();
  if (treeId == 0) { ctx->nodesNum = nodesNum; ctx->trif4Num = trif4Num; ctx->haveInst = haveInst; ctx->bvhDepthBound = bound; ctx->haveTree1 = false; ctx->haveAlpha1 = false; }
  else { ctx->haveTree1 = true; ctx->haveInst1 = haveInst; }
  return HC_OK;
}

// SetAllBVH4 sends every tree again: tree 0 first (which forgets a previous tree 1), then - if the scene has meshes with opacity maps - tree 1
int hc_set_bvh(hc_ctx* ctx, int treeId, const void* nodes, int nodesNum, const void* trif4, int trif4Num, int haveInst)
{ return SetBvhTree(ctx, treeId, nodes, nodesNum, trif4, trif4Num, nullptr, 0, haveInst); }
int hc_set_bvh_alpha(hc_ctx* ctx, int treeId, const void* nodes, int nodesNum, const void* trif4, int trif4Num, const void* alphaTableUint2, int alphaNum, int haveInst)
{ return SetBvhTree(ctx, treeId, nodes, nodesNum, trif4, trif4Num, alphaTableUint2, alphaNum, haveInst); }

// material remap lists: per-instance material overrides looked up in surface evaluation (remapMaterialId, cglobals.h:2931-2983)
int hc_set_remap_lists(hc_ctx* ctx, const int32_t* allLists, const int32_t* tableOffsetAndSize, int allSize, int tableSize)
{
  if (!ctx) return HC_E_ARG;
  ctx->sceneDirty = true;
  if (!allLists || !tableOffsetAndSize || allSize <= 0 || tableSize <= 0) { ctx->remapListsSize = 0; ctx->remapTableSize = 0; ctx->remapListsHost.clear(); return HC_OK; }   // GPUOCLData.cpp:203-210
  for (int i = 0; i < tableSize; i++)
    HC_REQUIRE(tableOffsetAndSize[2*i] >= 0 && tableOffsetAndSize[2*i + 1] >= 0 && (long long)tableOffsetAndSize[2*i] + tableOffsetAndSize[2*i + 1] <= allSize, HC_E_RANGE,
               "hc_set_remap_lists: a remap list lies outside the array of all lists");
  HC_CUDA(cudaSetDevice(ctx->device));
  int rc = hc_buf_reserve(ctx, ctx->remapLists, size_t(allSize)*4); if (rc) return rc;
  rc = hc_buf_reserve(ctx, ctx->remapTable, size_t(tableSize)*8); if (rc) return rc;
  HC_CUDA(cudaMemcpyAsync(ctx->remapLists.ptr, allLists, size_t(allSize)*4, cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaMemcpyAsync(ctx->remapTable.ptr, tableOffsetAndSize, size_t(tableSize)*8, cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->remapListsSize = allSize; ctx->remapTableSize = tableSize;
  ctx->remapListsHost.assign(allLists, allLists + allSize);
  return HC_OK;
}
int hc_set_inst_remap_ids(hc_ctx* ctx, const int32_t* instRemapListId, int n)
{
  if (!ctx) return HC_E_ARG;
  ctx->sceneDirty = true;
  if (!instRemapListId || n <= 0) { ctx->remapInstSize = 0; return HC_OK; }
  HC_CUDA(cudaSetDevice(ctx->device));
  int rc = hc_buf_reserve(ctx, ctx->remapInst, size_t(n)*4); if (rc) return rc;
  HC_CUDA(cudaMemcpyAsync(ctx->remapInst.ptr, instRemapListId, size_t(n)*4, cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->remapInstSize = n;
  return HC_OK;
}

int hc_bvh_device_layout(const void* nodes, int nodesNum, const void* trif4, int trif4Num, float* outNodes, float* outPairs,
                         int64_t outPairsCapacityFloats, int64_t* outPairsFloats, int* outStackBound)
{
  if (!nodes || !trif4 || nodesNum < 8 || trif4Num < 4) return HC_E_ARG;
  std::vector<float> devNodes, devPairs; int bound = 0;
  const int rc = ConvertBvhForDevice((const unsigned char*)nodes, nodesNum, (const float*)trif4, trif4Num, devNodes, devPairs, &bound);
  HC_REQUIRE(rc == HC_OK, rc, "hc_bvh_device_layout: tree references nodes or triangles out of range (or a leaf holds more than 128 triangles)");
  if (outStackBound) *outStackBound = bound;
  if (outPairsFloats) *outPairsFloats = (int64_t)devPairs.size();
  if (outNodes) memcpy(outNodes, devNodes.data(), devNodes.size()*4);
  if (outPairs)
  {
    HC_REQUIRE(outPairsCapacityFloats >= (int64_t)devPairs.size(), HC_E_RANGE, "hc_bvh_device_layout: outPairs too small");
    memcpy(outPairs, devPairs.data(), devPairs.size()*4);
  }
  return HC_OK;
}

int hc_set_inst_matrices(hc_ctx* ctx, const float* invMatrices16, int n)
{
  if (!ctx || !invMatrices16 || n <= 0) return HC_E_ARG;
  HC_CUDA(cudaSetDevice(ctx->device));
  int rc = hc_buf_reserve(ctx, ctx->instMatrices, uint64_t(n)*64); if (rc) return rc;
  HC_CUDA(cudaMemcpyAsync(ctx->instMatrices.ptr, invMatrices16, uint64_t(n)*64, cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->numInst = n;
  return HC_OK;
}

int hc_set_inst_light_ids(hc_ctx* ctx, const int32_t* lightInstId, int n)
{
  if (!ctx || !lightInstId || n <= 0) return HC_E_ARG;
  HC_CUDA(cudaSetDevice(ctx->device));
  int rc = hc_buf_reserve(ctx, ctx->instLightIds, uint64_t(n)*4); if (rc) return rc;
  HC_CUDA(cudaMemcpyAsync(ctx->instLightIds.ptr, lightInstId, uint64_t(n)*4, cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return HC_OK;
}

// ---- ray casting
int hc_make_eye_rays(hc_ctx* ctx, int width, int height, const float* offsets4OrNull, float* rays8Out, int space)
{
  if (!ctx || !rays8Out || width <= 0 || height <= 0) return HC_E_ARG;
  HC_CUDA(cudaSetDevice(ctx->device));
  const long long n = (long long)width*height;
  HcCamera cam = hc_camera_from_globals(ctx->globalsHead.data());
  float4* dRays = (float4*)rays8Out;
  const float4* dOffs = (const float4*)offsets4OrNull;
  if (space == HC_HOST)
  {
    int rc = hc_buf_reserve(ctx, ctx->scratchRays, uint64_t(n)*32 + (offsets4OrNull ? uint64_t(n)*16 : 0)); if (rc) return rc;
    dRays = (float4*)ctx->scratchRays.ptr;
    if (offsets4OrNull)
    {
      dOffs = dRays + n*2;
      HC_CUDA(cudaMemcpyAsync((void*)dOffs, offsets4OrNull, uint64_t(n)*16, cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  const int block = 256; const int grid = int((n + block - 1)/block);
  k_make_eye_rays<<<grid, block, 0, ctx->stream>>>(cam, width, height, dOffs, dRays);
  HC_CUDA(cudaGetLastError());
  ctx->stats.kernelLaunches++;
  if (space == HC_HOST) HC_CUDA(cudaMemcpyAsync(rays8Out, dRays, uint64_t(n)*32, cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return HC_OK;
}

static int TraceEntry(hc_ctx* ctx, bool anyHit, const float* rays8, int64_t n, void* out, int space)
{
  if (!ctx || !rays8 || !out || n < 0) return HC_E_ARG;
  if (n == 0) return HC_OK;
  HC_CUDA(cudaSetDevice(ctx->device));
  const uint64_t outBytes = anyHit ? uint64_t(n) : uint64_t(n)*16;
  const float4* dRays = (const float4*)rays8; void* dOut = out;
  if (space == HC_HOST)
  {
    int rc = hc_buf_reserve(ctx, ctx->scratchRays, uint64_t(n)*32); if (rc) return rc;
    rc = hc_buf_reserve(ctx, ctx->scratchOut, outBytes); if (rc) return rc;
    HC_CUDA(cudaMemcpyAsync(ctx->scratchRays.ptr, rays8, uint64_t(n)*32, cudaMemcpyHostToDevice, ctx->stream));
    dRays = (const float4*)ctx->scratchRays.ptr; dOut = ctx->scratchOut.ptr;
  }
  HC_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  int rc = hc_launch_trace(ctx, anyHit, dRays, dRays + 1, 2, n, anyHit ? nullptr : (HcHit*)dOut, anyHit ? (unsigned char*)dOut : nullptr);
  if (rc) return rc;
  HC_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (space == HC_HOST) HC_CUDA(cudaMemcpyAsync(out, dOut, outBytes, cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  HC_CUDA(cudaEventElapsedTime(&ctx->lastTraceMs, ctx->ev0, ctx->ev1));
  if (anyHit) ctx->stats.msShadow += ctx->lastTraceMs; else ctx->stats.msClosest += ctx->lastTraceMs;
  return HC_OK;
}

int hc_make_shadow_rays(hc_ctx* ctx, const float* rays8, const hc_hit* hits, int64_t n, const float lightPos[3], float* shadowRays8Out, int space)
{
  if (!ctx || !rays8 || !hits || !lightPos || !shadowRays8Out || n < 0) return HC_E_ARG;
  if (n == 0) return HC_OK;
  HC_CUDA(cudaSetDevice(ctx->device));
  const float4* dRays = (const float4*)rays8; const HcHit* dHits = (const HcHit*)hits; float4* dOut = (float4*)shadowRays8Out;
  if (space == HC_HOST)
  {
    int rc = hc_buf_reserve(ctx, ctx->scratchRays, uint64_t(n)*64); if (rc) return rc;
    rc = hc_buf_reserve(ctx, ctx->scratchOut, uint64_t(n)*16); if (rc) return rc;
    HC_CUDA(cudaMemcpyAsync(ctx->scratchRays.ptr, rays8, uint64_t(n)*32, cudaMemcpyHostToDevice, ctx->stream));
    HC_CUDA(cudaMemcpyAsync(ctx->scratchOut.ptr, hits, uint64_t(n)*16, cudaMemcpyHostToDevice, ctx->stream));
    dRays = (const float4*)ctx->scratchRays.ptr; dHits = (const HcHit*)ctx->scratchOut.ptr; dOut = (float4*)ctx->scratchRays.ptr + 2*n;
  }
  const int block = 256; const int grid = int((n + block - 1)/block);
  k_make_shadow_rays<<<grid, block, 0, ctx->stream>>>(dRays, dHits, n, make_float3(lightPos[0], lightPos[1], lightPos[2]), dOut);
  HC_CUDA(cudaGetLastError());
  ctx->stats.kernelLaunches++;
  if (space == HC_HOST) HC_CUDA(cudaMemcpyAsync(shadowRays8Out, dOut, uint64_t(n)*32, cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return HC_OK;
}

int hc_raycast_pass(hc_ctx* ctx, const float lightPos[3], hc_hit* hitsOutOrNull, uint8_t* visibleOutOrNull, int space)
{
  if (!ctx || !lightPos) return HC_E_ARG;
  HC_REQUIRE(ctx->width > 0 && ctx->height > 0, HC_E_STATE, "hc_raycast_pass: call hc_resize first");
  HC_REQUIRE(ctx->globals.ptr != nullptr, HC_E_STATE, "hc_raycast_pass: call hc_set_globals first");
  HC_CUDA(cudaSetDevice(ctx->device));
  const long long n = (long long)ctx->width*ctx->height;
  int rc;
  if ((rc = hc_buf_reserve(ctx, ctx->rcRays, uint64_t(n)*32))) return rc;
  if ((rc = hc_buf_reserve(ctx, ctx->rcHits, uint64_t(n)*16))) return rc;
  if ((rc = hc_buf_reserve(ctx, ctx->rcSRays, uint64_t(n)*32))) return rc;
  if ((rc = hc_buf_reserve(ctx, ctx->rcVis, uint64_t(n)))) return rc;
  float4* rays = (float4*)ctx->rcRays.ptr; HcHit* hits = (HcHit*)ctx->rcHits.ptr; float4* srays = (float4*)ctx->rcSRays.ptr;
  unsigned char* vis = (unsigned char*)ctx->rcVis.ptr;
  const HcCamera cam = hc_camera_from_globals(ctx->globalsHead.data());
  const int block = 256; const int grid = int((n + block - 1)/block);

  // eye rays and shadow rays are generated inside the traversal kernels' ray fetch (no ray buffers, two launches fewer) unless a second BVH
  // tree has to be walked with the same rays
  const bool fused = !ctx->haveTree1 && ctx->bvhNodes.ptr && ctx->bvhTris.ptr && getenv("HC_RAYCAST_UNFUSED") == nullptr;
  HcRayGen gen; gen.cam = cam; gen.width = ctx->width; gen.height = ctx->height; gen.firstPixel = 0;
  gen.light = make_float3(lightPos[0], lightPos[1], lightPos[2]); gen.hitsIn = hits;
  // Multi-GPU: with a tile partition (hc_pt_set_tiles, world size > 1) this rank casts the rays of ITS pixels only; with a communicator
  // (hc_comm_init) the hit records and visibility bytes of all ranks are then gathered on rank 0 - one frame, split over the GPUs.
  const bool split = ctx->worldSize > 1;
  long long nMine = n;
  if (split)
  {
    HC_REQUIRE(fused, HC_E_STATE, "hc_raycast_pass: the tile-partitioned pass needs the fused ray generation (one BVH tree)");
    const int* px = nullptr; int cnt = 0;
    if ((rc = hc_path_owned_pixels(ctx, &px, &cnt))) return rc;
    gen.pixels = px; nMine = cnt;
  }
  HC_CUDA(cudaEventRecord(ctx->evStage[0], ctx->stream));
  if (!fused)
  {
    k_make_eye_rays<<<grid, block, 0, ctx->stream>>>(cam, ctx->width, ctx->height, nullptr, rays);
    HC_CUDA(cudaGetLastError());
    ctx->stats.kernelLaunches++;
  }
  HC_CUDA(cudaEventRecord(ctx->evStage[1], ctx->stream));
  const int tileW = (!split && ctx->width % 8 == 0 && ctx->height % 4 == 0) ? ctx->width : 0;       // rays fetched in 8 x 4 pixel blocks per warp (the owned-pixel list is in that order already)
  // With HOST outputs the closest-hit pass runs in two horizontal bands, so that the read-back of the first band's hit records (16 B per ray,
  // copy stream) overlaps the traversal of the second and the read-back of the second overlaps the shadow pass.  Bands are whole groups of
  // 4 rows, which keeps the 8 x 4 block fetch order valid.
  const int rowGroups = (ctx->height + 3)/4;
  const int bands = (!split && space == HC_HOST && hitsOutOrNull && tileW > 0 && rowGroups >= 16) ? 2 : 1;     // measured on B200 at 1080p: 1 / 2 / 3 / 4 bands -> 1.39 / 1.18 / 1.21 / 1.30 ms
  const cudaMemcpyKind kind = (space == HC_HOST) ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  for (int b = 0; b < bands; b++)
  {
    const long long r0 = (long long)(rowGroups*b/bands)*4, r1 = std::min<long long>((long long)(rowGroups*(b + 1)/bands)*4, ctx->height);
    const long long i0 = split ? 0 : r0*ctx->width, nb = split ? nMine : (r1 - r0)*ctx->width;
    gen.firstPixel = i0;
    if (fused) { if ((rc = LaunchTraceGen(ctx, false, nb, split ? hits : hits + i0, nullptr, tileW, gen))) return rc; }
    else if ((rc = LaunchTrace(ctx, false, rays + 2*i0, rays + 2*i0 + 1, 2, nb, nullptr, hits + i0, nullptr, tileW))) return rc;
    if (hitsOutOrNull && !split)
    {
      HC_CUDA(cudaEventRecord(ctx->evCopy, ctx->stream));
      HC_CUDA(cudaStreamWaitEvent(ctx->copyStream, ctx->evCopy, 0));
      HC_CUDA(cudaMemcpyAsync(hitsOutOrNull + i0, hits + i0, uint64_t(nb)*16, kind, ctx->copyStream));
    }
  }
  HC_CUDA(cudaEventRecord(ctx->evStage[2], ctx->stream));
  if (!fused)
  {
    k_make_shadow_rays<<<grid, block, 0, ctx->stream>>>(rays, hits, n, make_float3(lightPos[0], lightPos[1], lightPos[2]), srays);
    HC_CUDA(cudaGetLastError());
    ctx->stats.kernelLaunches++;
  }
  HC_CUDA(cudaEventRecord(ctx->evStage[3], ctx->stream));
  gen.firstPixel = 0;
  if (fused) { if ((rc = LaunchTraceGen(ctx, true, nMine, nullptr, vis, tileW, gen))) return rc; }
  else if ((rc = LaunchTrace(ctx, true, srays, srays + 1, 2, n, nullptr, nullptr, vis, tileW))) return rc;
  HC_CUDA(cudaEventRecord(ctx->evStage[4], ctx->stream));
  if (split)
  {
    // one frame over all GPUs: the records of every rank's pixels travel to rank 0 (NCCL send / recv of the packed owned pixels)
    if ((rc = hc_comm_gather_raycast(ctx, hits, vis, 0))) return rc;
    if (ctx->rank == 0 && hitsOutOrNull) HC_CUDA(cudaMemcpyAsync(hitsOutOrNull, hits, uint64_t(n)*16, kind, ctx->stream));
    if (ctx->rank == 0 && visibleOutOrNull) HC_CUDA(cudaMemcpyAsync(visibleOutOrNull, vis, uint64_t(n), kind, ctx->stream));
  }
  else if (visibleOutOrNull) HC_CUDA(cudaMemcpyAsync(visibleOutOrNull, vis, uint64_t(n), kind, ctx->stream));      // 1 byte per ray: not worth a band
  HC_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->stats.paths += (uint64_t)nMine;
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  if (hitsOutOrNull && !split) HC_CUDA(cudaStreamSynchronize(ctx->copyStream));
  float ms[5];
  for (int i = 0; i < 4; i++) HC_CUDA(cudaEventElapsedTime(&ms[i], ctx->evStage[i], ctx->evStage[i + 1]));
  HC_CUDA(cudaEventElapsedTime(&ms[4], ctx->evStage[4], ctx->ev1));
  ctx->stats.msOther += ms[0] + ms[2] + (split ? ms[4] : 0.0f); ctx->stats.msClosest += ms[1]; ctx->stats.msShadow += ms[3];
  ctx->lastTraceMs = ms[0] + ms[1] + ms[2] + ms[3] + (split ? ms[4] : 0.0f);      // a split frame is not complete before the gather
  return HC_OK;
}

// ---- memory-system microbenchmarks (denominators of the roofline that MEASURED_PEAKS.json does not hold)
// Streaming 128-bit reads of a buffer, `repeats` sweeps per launch.  With a buffer well below the 126 MB L2 the sweeps after the first are
// served by L2 (-> L2 read bandwidth); with a buffer far above it they come from HBM (cross-check of the driver-measured copy bandwidth).
static __global__ void __launch_bounds__(256) k_stream_read(const uint4* __restrict__ buf, const size_t n16, const int repeats, unsigned* __restrict__ sink)
{
  unsigned acc = 0;
  const size_t stride = size_t(gridDim.x)*blockDim.x;
  for (int r = 0; r < repeats; r++)
    for (size_t i = size_t(blockIdx.x)*blockDim.x + threadIdx.x; i < n16; i += stride)
    {
      uint4 v;
      asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));   // .cg: bypass L1
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x12345678u) sink[0] = acc;       // keeps the loads alive
}

int hc_measure_read_bandwidth(hc_ctx* ctx, uint64_t bytes, int repeats, float* outGBs)
{
  if (!ctx || !outGBs || bytes < 4096 || repeats < 1) return HC_E_ARG;
  HC_CUDA(cudaSetDevice(ctx->device));
  HcDevBuf b, sink;
  int rc = hc_buf_reserve(ctx, b, bytes); if (rc) return rc;
  rc = hc_buf_reserve(ctx, sink, 16); if (rc) { hc_buf_free(b); return rc; }
  HC_CUDA(cudaMemsetAsync(b.ptr, 1, bytes, ctx->stream));
  const size_t n16 = bytes/16;
  const int grid = ctx->smCount*8;
  k_stream_read<<<grid, 256, 0, ctx->stream>>>((const uint4*)b.ptr, n16, 2, (unsigned*)sink.ptr);         // warm-up (fills L2 when it fits)
  float best = 0.0f;
  for (int it = 0; it < 5; it++)
  {
    HC_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    k_stream_read<<<grid, 256, 0, ctx->stream>>>((const uint4*)b.ptr, n16, repeats, (unsigned*)sink.ptr);
    HC_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    HC_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.0f; HC_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    const float gbs = float(double(n16)*16.0*double(repeats)/(double(ms)*1e-3)/1e9);
    if (gbs > best) best = gbs;
  }
  HC_CUDA(cudaGetLastError());
  hc_buf_free(b); hc_buf_free(sink);
  *outGBs = best;
  return HC_OK;
}

int hc_trace_closest(hc_ctx* ctx, const float* rays8, int64_t n, hc_hit* hitsOut, int space) { return TraceEntry(ctx, false, rays8, n, hitsOut, space); }
int hc_trace_shadow(hc_ctx* ctx, const float* rays8, int64_t n, uint8_t* visibleOut, int space) { return TraceEntry(ctx, true, rays8, n, visibleOut, space); }
int hc_trace_last_ms(hc_ctx* ctx, float* outMs) { if (!ctx || !outMs) return HC_E_ARG; *outMs = ctx->lastTraceMs; return HC_OK; }

int hc_get_stats(hc_ctx* ctx, hc_stats* out) { if (!ctx || !out) return HC_E_ARG; *out = ctx->stats; return HC_OK; }
int hc_reset_stats(hc_ctx* ctx) { if (!ctx) return HC_E_ARG; ctx->stats = hc_stats{}; return HC_OK; }

} // extern "C"
