// hc_api.cu — C ABI of libhydracore_b200.so: context, MemoryStorageCUDA backing, scene upload and the ray-casting entry points.
// (Path tracing entry points live in hc_path.cu.)  Reference interfaces replaced are cited in include/hydracore_cuda.h.
#include "hc_context.h"
#include "hc_trace.cuh"
#include "hc_raygen.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>

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
#define HC_REFILL_MIN  12      // refill a warp from the global ray counter once this many lanes are idle

// K2 / K2s.  Persistent warps: lanes that finished a ray are refilled together (one atomicAdd per refill, ranks by ballot/popc).
// rpos/rdir are float4 streams with element stride `stride` (2 = interleaved {pos,dir} records, 1 = separate arrays).
template<bool ANYHIT>
__global__ void __launch_bounds__(HC_TRACE_BLOCK)
k_trace(const HcBvh bvh, const float4* __restrict__ rpos, const float4* __restrict__ rdir, const int stride, const long long nArg,
        const int* __restrict__ nDev, HcHit* __restrict__ hitsOut, unsigned char* __restrict__ visOut, unsigned long long* __restrict__ counter)
{
  const long long n = nDev ? (long long)(*nDev) : nArg;      // the path tracer keeps its live-path count on the device
  unsigned stkNode[HC_STACK_CAP];
  float    stkT[HC_STACK_CAP];

  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned ltMask = (1u << lane) - 1u;

  bool  idle = true, exhausted = false;
  long long rayIdx = -1;

  // per-ray traversal state (see Traverse<> in hc_trace.cuh for the single-ray form used by the path tracer)
  HcHit hit; float3 o, d, inv, wo, wd; int sp = 0, instTop = 0, instId = -1; unsigned node = 1u; bool inInst = false;
  hit.t = 0; hit.primId = -1; hit.instId = -1; hit.geomId = 0;
  o = d = inv = wo = wd = f3(0, 0, 0);

  for (;;)
  {
    const unsigned idleMask = __ballot_sync(FULL, idle);
    if (idleMask != 0u && !exhausted && (__popc(idleMask) >= HC_REFILL_MIN || idleMask == FULL))
    {
      const int nIdle = __popc(idleMask), leader = __ffs(idleMask) - 1;
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(counter, (unsigned long long)nIdle);
      base = __shfl_sync(FULL, base, leader);
      if (idle)
      {
        const long long idx = (long long)base + __popc(idleMask & ltMask);
        if (idx < n)
        {
          const float4 p = __ldg(rpos + idx*stride), dd = __ldg(rdir + idx*stride);
          rayIdx = idx; idle = false;
          o = f3(p); d = f3(dd); inv = SafeInverse(d);
          hit.t = ANYHIT ? dd.w : HC_MAXFLOAT; hit.primId = -1; hit.instId = -1; hit.geomId = int(0xC0000000u);   // Make_Lite_Hit(t, -1)
          sp = 0; node = 1u; inInst = false; instTop = 0; instId = -1;
          if (ANYHIT && !(dd.w > 0.0f)) { visOut[idx] = 1; idle = true; }          // maxDist <= 0: lit (trace.cl:343-351)
        }
      }
      if ((long long)base + nIdle >= n) exhausted = true;
    }
    if (__all_sync(FULL, idle)) { if (exhausted) break; else continue; }
    if (idle) continue;

    bool needPop = false, finished = false;
    if (!(node & HC_LEAF_BIT))
    {
      const float4* q = bvh.nodes + size_t(node)*8;
      const float4 a0 = __ldg(q + 0), b0 = __ldg(q + 1), a1 = __ldg(q + 2), b1 = __ldg(q + 3);
      const float4 a2 = __ldg(q + 4), b2 = __ldg(q + 5), a3 = __ldg(q + 6), b3 = __ldg(q + 7);
      float t0 = ChildEntry(a0, b0, o, inv, hit.t), t1 = ChildEntry(a1, b1, o, inv, hit.t);
      float t2 = ChildEntry(a2, b2, o, inv, hit.t), t3 = ChildEntry(a3, b3, o, inv, hit.t);
      unsigned c0 = __float_as_uint(a0.w), c1 = __float_as_uint(a1.w), c2 = __float_as_uint(a2.w), c3 = __float_as_uint(a3.w);
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
      if (ANYHIT && found) finished = true; else needPop = true;
    }

    if (needPop)
    {
      for (;;)
      {
        if (sp == 0) { finished = true; break; }
        sp--;
        node = stkNode[sp];
        const float te = stkT[sp];
        if (inInst && sp < instTop) { o = wo; d = wd; inv = SafeInverse(d); inInst = false; }
        if (te <= hit.t) break;
      }
    }

    if (finished)
    {
      if (ANYHIT) visOut[rayIdx] = (hit.primId != -1) ? 0 : 1;
      else        hitsOut[rayIdx] = hit;
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
  ctx->traceGrid = ctx->smCount*perSM;           // a whole number of waves: persistent CTAs, all resident
  return ctx->traceGrid;
}

// launch K2 (closest) or K2s (any-hit) on device-resident streams; used by hc_trace_* and by the path tracer
static int LaunchTrace(hc_ctx* ctx, bool anyHit, const float4* rpos, const float4* rdir, int stride, long long n, const int* nDev, HcHit* hits, unsigned char* vis);
int hc_launch_trace(hc_ctx* ctx, bool anyHit, const float4* rpos, const float4* rdir, int stride, long long n, HcHit* hits, unsigned char* vis)
{ return LaunchTrace(ctx, anyHit, rpos, rdir, stride, n, nullptr, hits, vis); }
int hc_launch_trace_counted(hc_ctx* ctx, bool anyHit, const float4* rpos, const float4* rdir, long long nUpper, const int* nDev, HcHit* hits, unsigned char* vis)
{ return LaunchTrace(ctx, anyHit, rpos, rdir, 1, nUpper, nDev, hits, vis); }
static int LaunchTrace(hc_ctx* ctx, bool anyHit, const float4* rpos, const float4* rdir, int stride, long long n, const int* nDev, HcHit* hits, unsigned char* vis)
{
  if (n <= 0) return HC_OK;
  HC_REQUIRE(ctx->bvhNodes.ptr && ctx->bvhTris.ptr, HC_E_STATE, "hc_trace: no BVH uploaded (hc_set_bvh)");
  HC_REQUIRE(ctx->haveInst != 0, HC_E_STATE, "hc_trace: only the two-level (instanced) layout is supported");
  HcBvh bvh; bvh.nodes = (const float4*)ctx->bvhNodes.ptr; bvh.tris = (const float4*)ctx->bvhTris.ptr;
  // one persistent-thread ray counter per launch in flight: rotate through the counter block so that back-to-back launches never share one
  ctx->traceCounterSlot = (ctx->traceCounterSlot + 1) % 32;
  unsigned long long* counter = (unsigned long long*)ctx->counters.ptr + ctx->traceCounterSlot;
  HC_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx->stream));
  const int grid = (int)std::min<long long>(TraceGrid(ctx), (n + HC_TRACE_BLOCK - 1)/HC_TRACE_BLOCK);
  if (anyHit) k_trace<true><<<grid, HC_TRACE_BLOCK, 0, ctx->stream>>>(bvh, rpos, rdir, stride, n, nDev, nullptr, vis, counter);
  else        k_trace<false><<<grid, HC_TRACE_BLOCK, 0, ctx->stream>>>(bvh, rpos, rdir, stride, n, nDev, hits, nullptr, counter);
  HC_CUDA(cudaGetLastError());
  ctx->stats.kernelLaunches++;
  if (anyHit) ctx->stats.raysShadow += (uint64_t)n; else ctx->stats.raysClosest += (uint64_t)n;
  return HC_OK;
}

// walk the uploaded tree on the host: validates offsets and bounds the traversal stack
static int ValidateBvh(const unsigned char* nodes, int nodesNum, int trif4Num, int* outStackBound)
{
  struct N { float bmin[3]; unsigned lo; float bmax[3]; unsigned esc; };
  const N* nd = (const N*)nodes;
  const int quads = nodesNum/4;
  if (quads < 2) return HC_E_ARG;
  struct It { unsigned quad; int depth; bool inst; };
  std::vector<It> st; st.push_back({ 1u, 1, false });
  std::vector<unsigned char> seen(size_t(quads), 0);
  int maxTop = 0, maxMesh = 0;
  while (!st.empty())
  {
    It it = st.back(); st.pop_back();
    if (it.quad >= unsigned(quads)) return HC_E_RANGE;
    if (it.inst) maxMesh = std::max(maxMesh, it.depth); else maxTop = std::max(maxTop, it.depth);
    if (seen[it.quad]) continue;     // shared mesh sub-trees: depth of first visit is representative
    seen[it.quad] = 1;
    for (int i = 0; i < 4; i++)
    {
      const N& c = nd[size_t(it.quad)*4 + i];
      if (c.lo == 0xffffffffu && c.esc == 0xffffffffu) continue;
      const unsigned off = c.lo & 0x7fffffffu;
      if (c.lo & 0x80000000u)
      {
        if (!it.inst)
        {
          if (off >= unsigned(quads)) return HC_E_RANGE;
          const N& rec = nd[size_t(off)*4];
          const unsigned sub = rec.lo & 0x7fffffffu;
          if (rec.lo & 0x80000000u) { if (sub >= unsigned(trif4Num)) return HC_E_RANGE; maxMesh = std::max(maxMesh, 1); }
          else st.push_back({ sub, 1, true });
        }
        else if (off >= unsigned(trif4Num)) return HC_E_RANGE;
      }
      else st.push_back({ off, it.depth + 1, it.inst });
    }
  }
  *outStackBound = 3*(maxTop + maxMesh) + 2;
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
  HC_CUDA(cudaEventCreate(&c->ev0)); HC_CUDA(cudaEventCreate(&c->ev1));
  for (int i = 0; i < 5; i++) HC_CUDA(cudaEventCreate(&c->evStage[i]));
  int rc = hc_buf_reserve(c, c->counters, 64*sizeof(unsigned long long));
  if (rc != HC_OK) { delete c; return rc; }
  HC_CUDA(cudaMemsetAsync(c->counters.ptr, 0, c->counters.bytes, c->stream));
  c->globalsHead.assign(HC_EG_HEAD_BYTES, 0);
  *out = c;
  return HC_OK;
}

void hc_ctx_destroy(hc_ctx* c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  hc_path_free(c);
  for (int i = 0; i < HC_STORAGE_COUNT; i++) hc_buf_free(c->storage[i]);
  hc_buf_free(c->globals); hc_buf_free(c->bvhNodes); hc_buf_free(c->bvhTris); hc_buf_free(c->instMatrices); hc_buf_free(c->instLightIds);
  hc_buf_free(c->fbSum); hc_buf_free(c->scratchRays); hc_buf_free(c->scratchOut); hc_buf_free(c->counters); hc_buf_free(c->pixelRng); hc_buf_free(c->qmcTable);
  hc_buf_free(c->rcRays); hc_buf_free(c->rcHits); hc_buf_free(c->rcSRays); hc_buf_free(c->rcVis);
  for (int i = 0; i < 5; i++) cudaEventDestroy(c->evStage[i]);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
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
  return HC_OK;
}

int hc_set_bvh(hc_ctx* ctx, int treeId, const void* nodes, int nodesNum, const void* trif4, int trif4Num, int haveInst)
{
  if (!ctx || !nodes || !trif4 || nodesNum < 8 || trif4Num < 4) return HC_E_ARG;
  HC_REQUIRE(treeId == 0, HC_E_ARG, "hc_set_bvh: only tree 0 (opaque geometry) is supported; tree 1 holds alpha-tested meshes");
  HC_REQUIRE(haveInst != 0, HC_E_ARG, "hc_set_bvh: single-level trees (bvhType \"triangle4v\") are not supported, pass the two-level layout");
  int bound = 0;
  int rc = ValidateBvh((const unsigned char*)nodes, nodesNum, trif4Num, &bound);
  HC_REQUIRE(rc == HC_OK, rc, "hc_set_bvh: tree references nodes or triangles out of range");
  HC_REQUIRE(bound <= HC_STACK_CAP, HC_E_RANGE, "hc_set_bvh: tree too deep for the traversal stack");
  HC_CUDA(cudaSetDevice(ctx->device));
  rc = hc_buf_reserve(ctx, ctx->bvhNodes, uint64_t(nodesNum)*32); if (rc) return rc;
  rc = hc_buf_reserve(ctx, ctx->bvhTris, uint64_t(trif4Num)*16); if (rc) return rc;
  HC_CUDA(cudaMemcpyAsync(ctx->bvhNodes.ptr, nodes, uint64_t(nodesNum)*32, cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaMemcpyAsync(ctx->bvhTris.ptr, trif4, uint64_t(trif4Num)*16, cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));                // ConvertionResult pointers die at ConvertUnmap (RenderDriverRTE.cpp:1436)
  ctx->nodesNum = nodesNum; ctx->trif4Num = trif4Num; ctx->haveInst = haveInst; ctx->bvhDepthBound = bound;
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

  HC_CUDA(cudaEventRecord(ctx->evStage[0], ctx->stream));
  k_make_eye_rays<<<grid, block, 0, ctx->stream>>>(cam, ctx->width, ctx->height, nullptr, rays);
  HC_CUDA(cudaGetLastError());
  HC_CUDA(cudaEventRecord(ctx->evStage[1], ctx->stream));
  if ((rc = hc_launch_trace(ctx, false, rays, rays + 1, 2, n, hits, nullptr))) return rc;
  HC_CUDA(cudaEventRecord(ctx->evStage[2], ctx->stream));
  k_make_shadow_rays<<<grid, block, 0, ctx->stream>>>(rays, hits, n, make_float3(lightPos[0], lightPos[1], lightPos[2]), srays);
  HC_CUDA(cudaGetLastError());
  HC_CUDA(cudaEventRecord(ctx->evStage[3], ctx->stream));
  if ((rc = hc_launch_trace(ctx, true, srays, srays + 1, 2, n, nullptr, vis))) return rc;
  HC_CUDA(cudaEventRecord(ctx->evStage[4], ctx->stream));
  ctx->stats.kernelLaunches += 2;
  ctx->stats.paths += (uint64_t)n;
  if (hitsOutOrNull) HC_CUDA(cudaMemcpyAsync(hitsOutOrNull, hits, uint64_t(n)*16, space == HC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
  if (visibleOutOrNull) HC_CUDA(cudaMemcpyAsync(visibleOutOrNull, vis, uint64_t(n), space == HC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms[4];
  for (int i = 0; i < 4; i++) HC_CUDA(cudaEventElapsedTime(&ms[i], ctx->evStage[i], ctx->evStage[i + 1]));
  ctx->stats.msOther += ms[0] + ms[2]; ctx->stats.msClosest += ms[1]; ctx->stats.msShadow += ms[3];
  ctx->lastTraceMs = ms[0] + ms[1] + ms[2] + ms[3];
  return HC_OK;
}

int hc_trace_closest(hc_ctx* ctx, const float* rays8, int64_t n, hc_hit* hitsOut, int space) { return TraceEntry(ctx, false, rays8, n, hitsOut, space); }
int hc_trace_shadow(hc_ctx* ctx, const float* rays8, int64_t n, uint8_t* visibleOut, int space) { return TraceEntry(ctx, true, rays8, n, visibleOut, space); }
int hc_trace_last_ms(hc_ctx* ctx, float* outMs) { if (!ctx || !outMs) return HC_E_ARG; *outMs = ctx->lastTraceMs; return HC_OK; }

int hc_get_stats(hc_ctx* ctx, hc_stats* out) { if (!ctx || !out) return HC_E_ARG; *out = ctx->stats; return HC_OK; }
int hc_reset_stats(hc_ctx* ctx) { if (!ctx) return HC_E_ARG; ctx->stats = hc_stats{}; return HC_OK; }

} // extern "C"
