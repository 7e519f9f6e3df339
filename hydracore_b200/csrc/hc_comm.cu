// hc_comm.cu — multi-GPU exchange of the HDR framebuffer inside the product library (C ABI: hc_comm_*, hc_fb_reduce).
//
// Replaces the reference's way of combining the partial images of its one-process-per-GPU mode: every process adds its buffer into an OS
// shared-memory image under a mutex (GPUOCLLayer::ContribToExternalImageAccumulator, hydra_drv/GPUOCLLayerOther.cpp:365-430).  Here the
// processes (still one per GPU, scene replicated) exchange over NVLink with NCCL:
//   * PT / MISPT: the image plane is partitioned in interleaved tiles (hc_pt_set_tiles), so the per-rank buffers are DISJOINT.  Only the
//     pixels a rank owns travel: they are packed densely (owned-pixel order), sent point to point to the destination rank and scattered
//     into its SUM buffer - 1/G of the volume of a full-buffer reduce per rank, and assignment instead of addition, so the call can be
//     repeated during progressive rendering without counting anything twice.
//   * MISPT-QMC: samples land on arbitrary pixels, every rank holds a full-size buffer: ncclReduce(sum) into a SEPARATE buffer of the
//     destination rank (the local sums stay untouched), which the read-back entry points then use.
// NCCL is bound at run time (dlopen of libnccl.so.2): a host process that already carries an NCCL (e.g. a PyTorch process) shares it,
// a plain C++ host gets the system library; without NCCL the single-GPU paths are unaffected and hc_comm_* fail loudly.
#include "hc_context.h"
#include <dlfcn.h>
#include <cstring>
#include <algorithm>

void hc_owned_pixels_of(int W, int H, int T, int rank, int world, std::vector<int>& owned);      // hc_path.cu

namespace
{
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7 };           // ncclDataType_t (nccl.h): ncclInt8 0, ncclUint8 1, ncclInt32 2, ncclUint32 3, ncclInt64 4, ncclUint64 5, ncclFloat16 6, ncclFloat32 7
enum { ncclSum = 0 };

struct NcclApi
{
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;

int LoadNccl()
{
  if (g_nccl.lib) return HC_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { hc_set_error("hc_comm: libnccl.so.2 not found (multi-GPU exchange needs NCCL)"); return HC_E_STATE; }
#define HC_SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { hc_set_error("hc_comm: NCCL symbol missing: " name); dlclose(h); return HC_E_STATE; }
  HC_SYM(GetUniqueId, "ncclGetUniqueId") HC_SYM(CommInitRank, "ncclCommInitRank") HC_SYM(CommDestroy, "ncclCommDestroy")
  HC_SYM(GroupStart, "ncclGroupStart") HC_SYM(GroupEnd, "ncclGroupEnd") HC_SYM(Send, "ncclSend") HC_SYM(Recv, "ncclRecv")
  HC_SYM(Reduce, "ncclReduce") HC_SYM(GetErrorString, "ncclGetErrorString") HC_SYM(GetVersion, "ncclGetVersion")
#undef HC_SYM
  g_nccl.lib = h;
  return HC_OK;
}

int NcclFail(int e, const char* what)
{
  std::string m = std::string("NCCL error in ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
  hc_set_error(m.c_str());
  return HC_E_STATE;
}
#define HC_NCCL(call, what) do { const int e_ = (call); if (e_ != ncclSuccess) return NcclFail(e_, what); } while (0)

// dense <-> per-pixel buffer by a pixel list (the owned pixels of one rank, in the order hc_owned_pixels_of produces); T = float4 for the
// framebuffer and the hit records, unsigned char for the visibility bytes
template<class T> __global__ void k_px_pack(const T* __restrict__ buf, const int* __restrict__ pixels, const int n, T* __restrict__ out)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) out[i] = buf[pixels[i]];
}
template<class T> __global__ void k_px_unpack(T* __restrict__ buf, const int* __restrict__ pixels, const int n, const T* __restrict__ in)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) buf[pixels[i]] = in[i];
}

// pixel lists of the exchange: mine on a source rank, every source rank's (back to back) on the destination
int EnsurePixelLists(hc_ctx* ctx, int dstRank)
{
  const int W = ctx->width, H = ctx->height, T = std::max(1, ctx->tileSize), G = ctx->commSize;
  const long long key = (((long long)W*65536 + H)*1024 + T)*64 + G + 1000000000000000ll*(dstRank + 1);
  if (ctx->commPixelsKey == key) return HC_OK;
  std::vector<int> all;
  if (ctx->commRank != dstRank) { hc_owned_pixels_of(W, H, T, ctx->commRank, G, all); ctx->commCount.assign(1, (int)all.size()); }
  else
  {
    ctx->commCount.assign(G, 0);
    for (int g = 0; g < G; g++)
    {
      std::vector<int> px; if (g != dstRank) hc_owned_pixels_of(W, H, T, g, G, px);
      ctx->commCount[g] = (int)px.size();
      all.insert(all.end(), px.begin(), px.end());
    }
  }
  int rc = hc_buf_reserve(ctx, ctx->commPixels, std::max<size_t>(all.size(), 1)*4); if (rc) return rc;
  if (!all.empty()) HC_CUDA(cudaMemcpyAsync(ctx->commPixels.ptr, all.data(), all.size()*4, cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->commPixelsKey = key;
  return HC_OK;
}

// gather the owned pixels of every rank's per-pixel buffer `buf` (elemBytes 16 or 1) into the destination rank's buffer; enqueued on the context's stream
template<class T> int GatherOwned(hc_ctx* ctx, T* buf, int dstRank, HcDevBuf& stage)
{
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  cudaStream_t s = ctx->stream;
  const int G = ctx->commSize;
  int rc = EnsurePixelLists(ctx, dstRank); if (rc) return rc;
  if (ctx->commRank != dstRank)
  {
    const int n = ctx->commCount[0];
    rc = hc_buf_reserve(ctx, stage, std::max<size_t>(n, 1)*sizeof(T)); if (rc) return rc;
    if (n > 0)
    {
      k_px_pack<T><<<(n + 255)/256, 256, 0, s>>>(buf, (const int*)ctx->commPixels.ptr, n, (T*)stage.ptr);
      HC_CUDA(cudaGetLastError());
      ctx->stats.kernelLaunches++;
      HC_NCCL(g_nccl.Send(stage.ptr, size_t(n)*sizeof(T), 1 /* ncclUint8 */, dstRank, comm, s), "ncclSend");
    }
    return HC_OK;
  }
  size_t total = 0; for (int g = 0; g < G; g++) total += size_t(ctx->commCount[g]);
  rc = hc_buf_reserve(ctx, stage, std::max<size_t>(total, 1)*sizeof(T)); if (rc) return rc;
  HC_NCCL(g_nccl.GroupStart(), "ncclGroupStart");
  size_t off = 0;
  for (int g = 0; g < G; g++)
  {
    const int n = ctx->commCount[g];
    if (g == dstRank || n == 0) continue;
    const int e = g_nccl.Recv((T*)stage.ptr + off, size_t(n)*sizeof(T), 1 /* ncclUint8 */, g, comm, s);
    if (e != ncclSuccess) { g_nccl.GroupEnd(); return NcclFail(e, "ncclRecv"); }
    off += size_t(n);
  }
  HC_NCCL(g_nccl.GroupEnd(), "ncclGroupEnd");
  if (total > 0)
  {
    k_px_unpack<T><<<(int)((total + 255)/256), 256, 0, s>>>(buf, (const int*)ctx->commPixels.ptr, (int)total, (const T*)stage.ptr);
    HC_CUDA(cudaGetLastError());
    ctx->stats.kernelLaunches++;
  }
  return HC_OK;
}
}

// hit records (16 B) and visibility bytes of a rank's pixels packed into ONE message: [n x 16 B][n x 1 B]
__global__ void k_rc_pack(const float4* __restrict__ hits, const unsigned char* __restrict__ vis, const int* __restrict__ pixels, const int n,
                          float4* __restrict__ outHits, unsigned char* __restrict__ outVis)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i < n) { const int p = pixels[i]; outHits[i] = hits[p]; outVis[i] = vis[p]; }
}
// destination: segment g of the staging buffer starts at byte segOff[g] and holds cnt[g] hit records followed by cnt[g] bytes
__global__ void k_rc_unpack(float4* __restrict__ hits, unsigned char* __restrict__ vis, const int* __restrict__ pixels, const int total,
                            const unsigned char* __restrict__ stage, const long long* __restrict__ segOff, const int* __restrict__ segFirst, const int nSeg)
{
  const int i = blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= total) return;
  int g = 0;
  while (g + 1 < nSeg && i >= segFirst[g + 1]) g++;
  const int k = i - segFirst[g], cnt = ((g + 1 < nSeg) ? segFirst[g + 1] : total) - segFirst[g];
  const unsigned char* seg = stage + segOff[g];
  const int p = pixels[i];
  hits[p] = reinterpret_cast<const float4*>(seg)[k];
  vis[p] = seg[size_t(cnt)*16 + k];
}

// ray-casting results of a tile-partitioned pass (hc_raycast_pass): hit records and visibility bytes of the owned pixels -> destination rank,
// one message per source rank (17 bytes per pixel), received in one NCCL group
int hc_comm_gather_raycast(hc_ctx* ctx, void* hits16, unsigned char* vis, int dstRank)
{
  if (!ctx->comm || ctx->commSize == 1) return HC_OK;
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  cudaStream_t s = ctx->stream;
  const int G = ctx->commSize;
  int rc = EnsurePixelLists(ctx, dstRank); if (rc) return rc;
  if (ctx->commRank != dstRank)
  {
    const int n = ctx->commCount[0];
    const size_t bytes = ((size_t(n)*17 + 15)/16)*16;
    rc = hc_buf_reserve(ctx, ctx->commStage, std::max<size_t>(bytes, 16)); if (rc) return rc;
    if (n > 0)
    {
      k_rc_pack<<<(n + 255)/256, 256, 0, s>>>((const float4*)hits16, vis, (const int*)ctx->commPixels.ptr, n, (float4*)ctx->commStage.ptr, (unsigned char*)ctx->commStage.ptr + size_t(n)*16);
      HC_CUDA(cudaGetLastError());
      ctx->stats.kernelLaunches++;
      HC_NCCL(g_nccl.Send(ctx->commStage.ptr, bytes, 1 /* ncclUint8 */, dstRank, comm, s), "ncclSend");
    }
    return HC_OK;
  }
  // destination: segment table (host side is tiny: G entries), one receive per source rank
  std::vector<long long> segOff; std::vector<int> segFirst; std::vector<int> src;
  size_t off = 0; int first = 0;
  for (int g = 0; g < G; g++)
  {
    const int n = ctx->commCount[g];
    if (g == dstRank || n == 0) continue;
    segOff.push_back((long long)off); segFirst.push_back(first); src.push_back(g);
    off += ((size_t(n)*17 + 15)/16)*16; first += n;
  }
  if (src.empty()) return HC_OK;
  rc = hc_buf_reserve(ctx, ctx->commStage, std::max<size_t>(off, 16)); if (rc) return rc;
  rc = hc_buf_reserve(ctx, ctx->commStage2, 64*12 + 64); if (rc) return rc;
  HC_REQUIRE(src.size() <= 64, HC_E_RANGE, "hc_raycast_pass: more than 64 source ranks");
  long long* dOff = (long long*)ctx->commStage2.ptr; int* dFirst = (int*)((char*)ctx->commStage2.ptr + 64*8);
  if (ctx->commSegKey != ctx->commPixelsKey)
  {
    HC_CUDA(cudaMemcpyAsync(dOff, segOff.data(), segOff.size()*8, cudaMemcpyHostToDevice, s));
    HC_CUDA(cudaMemcpyAsync(dFirst, segFirst.data(), segFirst.size()*4, cudaMemcpyHostToDevice, s));
    HC_CUDA(cudaStreamSynchronize(s));               // the vectors are locals
    ctx->commSegKey = ctx->commPixelsKey;
  }
  HC_NCCL(g_nccl.GroupStart(), "ncclGroupStart");
  for (size_t k = 0; k < src.size(); k++)
  {
    const size_t bytes = ((size_t(ctx->commCount[src[k]])*17 + 15)/16)*16;
    const int e = g_nccl.Recv((unsigned char*)ctx->commStage.ptr + segOff[k], bytes, 1 /* ncclUint8 */, src[k], comm, s);
    if (e != ncclSuccess) { g_nccl.GroupEnd(); return NcclFail(e, "ncclRecv"); }
  }
  HC_NCCL(g_nccl.GroupEnd(), "ncclGroupEnd");
  k_rc_unpack<<<(first + 255)/256, 256, 0, s>>>((float4*)hits16, vis, (const int*)ctx->commPixels.ptr, first, (const unsigned char*)ctx->commStage.ptr, dOff, dFirst, (int)src.size());
  HC_CUDA(cudaGetLastError());
  ctx->stats.kernelLaunches++;
  return HC_OK;
}

void hc_comm_free(hc_ctx* ctx)
{
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->comm);
  ctx->comm = nullptr;
  hc_buf_free(ctx->commStage); hc_buf_free(ctx->commStage2); hc_buf_free(ctx->commPixels); hc_buf_free(ctx->fbCombined);
  ctx->combinedValid = false;
}

extern "C"
{
int hc_comm_unique_id(void* out128)
{
  if (!out128) return HC_E_ARG;
  int rc = LoadNccl(); if (rc) return rc;
  ncclUniqueId id;
  HC_NCCL(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out128, &id, 128);
  return HC_OK;
}

int hc_comm_init(hc_ctx* ctx, const void* uniqueId128, int rank, int nranks)
{
  if (!ctx || !uniqueId128 || nranks < 1 || rank < 0 || rank >= nranks) return HC_E_ARG;
  int rc = LoadNccl(); if (rc) return rc;
  HC_CUDA(cudaSetDevice(ctx->device));
  if (ctx->comm) { g_nccl.CommDestroy((ncclComm_t)ctx->comm); ctx->comm = nullptr; }
  ncclUniqueId id; memcpy(&id, uniqueId128, 128);
  ncclComm_t c = nullptr;
  HC_NCCL(g_nccl.CommInitRank(&c, nranks, id, rank), "ncclCommInitRank");
  ctx->comm = c; ctx->commRank = rank; ctx->commSize = nranks;
  return HC_OK;
}

int hc_comm_version(int* outVersion)
{
  if (!outVersion) return HC_E_ARG;
  int rc = LoadNccl(); if (rc) return rc;
  HC_NCCL(g_nccl.GetVersion(outVersion), "ncclGetVersion");
  return HC_OK;
}

// Combine the per-rank SUM buffers on `dstRank` (see the header comment).  mode 0 = tile partition (gather of owned tiles), 1 = full-size sum.
// Asynchronous on the context's stream up to the final synchronise; *outMs (optional) = device time of the exchange on this rank.
int hc_fb_reduce(hc_ctx* ctx, int dstRank, int mode, float* outMs)
{
  if (!ctx || (mode != 0 && mode != 1)) return HC_E_ARG;
  HC_REQUIRE(ctx->fbSum.ptr && ctx->width > 0 && ctx->height > 0, HC_E_STATE, "hc_fb_reduce: no framebuffer (hc_resize)");
  if (outMs) *outMs = 0.0f;
  const int G = ctx->comm ? ctx->commSize : 1;
  if (G == 1) return HC_OK;
  HC_REQUIRE(dstRank >= 0 && dstRank < G, HC_E_ARG, "hc_fb_reduce: bad destination rank");
  HC_REQUIRE(mode == 1 || (ctx->worldSize == G && ctx->rank == ctx->commRank), HC_E_STATE,
             "hc_fb_reduce: the tile partition (hc_pt_set_tiles) must use the communicator's rank and size");
  HC_CUDA(cudaSetDevice(ctx->device));
  ncclComm_t comm = (ncclComm_t)ctx->comm;
  cudaStream_t s = ctx->stream;
  const int W = ctx->width, H = ctx->height;
  const size_t nPix = size_t(W)*H;
  HC_CUDA(cudaEventRecord(ctx->ev0, s));
  if (mode == 1)
  {
    // out of place on the destination: the local sums stay untouched, so the call can be repeated during progressive rendering
    // (4K, 133 MB, two B200s: 0.26 ms, the same as torch.distributed.reduce in place - scripts/gpu_reduce_probe.py)
    if (ctx->commRank == dstRank) { int rc = hc_buf_reserve(ctx, ctx->fbCombined, nPix*16); if (rc) return rc; }
    HC_NCCL(g_nccl.Reduce(ctx->fbSum.ptr, ctx->commRank == dstRank ? ctx->fbCombined.ptr : nullptr, nPix*4, ncclFloat32, ncclSum, dstRank, comm, s), "ncclReduce");
    if (ctx->commRank == dstRank) ctx->combinedValid = true;
  }
  else
  {
    int rc = GatherOwned<float4>(ctx, (float4*)ctx->fbSum.ptr, dstRank, ctx->commStage); if (rc) return rc;
  }
  HC_CUDA(cudaEventRecord(ctx->ev1, s));
  HC_CUDA(cudaStreamSynchronize(s));
  if (outMs) HC_CUDA(cudaEventElapsedTime(outMs, ctx->ev0, ctx->ev1));
  return HC_OK;
}
} // extern "C"
