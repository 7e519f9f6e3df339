// hc_context.h — device-side state of one CUDA layer instance (what GPUOCLLayer keeps in m_scene / m_rays / m_screen,
// reference hydra_drv/GPUOCLLayer.h:300-470), shared by the translation units of libhydracore_b200.so.
#pragma once
#include "../../include/hydracore_cuda.h"
#include "hc_layout.h"
#include <cuda_runtime.h>
#include <string>
#include <vector>

void hc_set_error(const char* msg);
int  hc_cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define HC_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return hc_cuda_fail(e_, #call, __FILE__, __LINE__); } while (0)
#define HC_REQUIRE(cond, code, msg) do { if (!(cond)) { hc_set_error(msg); return (code); } } while (0)

struct HcHit;

struct HcDevBuf
{
  void*    ptr = nullptr;
  uint64_t bytes = 0;
};

struct hc_ctx
{
  int          device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copyStream = nullptr;       // read-backs that overlap the next kernel (hc_raycast_pass)
  cudaEvent_t  evCopy = nullptr;
  cudaStream_t stream2 = nullptr, copyStream2 = nullptr;   // second path pipeline of small frames (hc_pt_pass)
  cudaEvent_t  evFork2 = nullptr, evJoin2 = nullptr, evPipeFork = nullptr, evPipeJoin = nullptr;
  cudaEvent_t  evFork = nullptr, evJoin = nullptr;   // closest-hit and shadow traversal of one bounce run on two streams (hc_pt_pass)
  cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
  cudaDeviceProp prop{};
  int          smCount = 0;

  HcDevBuf storage[HC_STORAGE_COUNT];
  HcDevBuf globals;                        // EngineGlobals + tables blob
  std::vector<unsigned char> globalsHead;  // host copy of the first HC_EG_HEAD_BYTES bytes
  std::vector<unsigned char> globalsMirror, materialsMirror;   // host copies of the globals blob and of the materials storage, kept by the upload entry points (scene validation reads them)
  bool     texturesDirty = true;           // a texture storage was written since the image headers were last validated
  HcDevBuf bvhNodes, bvhTris;              // tree 0 (opaque geometry)
  HcDevBuf bvh1Nodes, bvh1Tris, bvh1AlphaPairs, bvh1AlphaTable;   // tree 1 (meshes with opacity maps), its per-pair alpha words and the reference's alpha table
  bool     haveTree1 = false, haveAlpha1 = false;
  int      shadowTrees = 1;                // 1: shadow rays walk every tree (GPUOCLLayer), 0: the first tree only (CPU integrators); hc_pt_set_shadow_trees
  HcDevBuf remapLists, remapTable, remapInst;   // material remap lists (SetAllRemapLists / SetAllInstIdToRemapId)
  int      remapListsSize = 0, remapTableSize = 0, remapInstSize = 0;
  std::vector<int> remapListsHost;              // host copy for validation at hc_pt_init (every 'to' id must exist in the materials table)
  std::vector<int> alphaTexIdsHost;             // texture ids used by the opacity samplers of tree 1, validated against the texture table
  int      nodesNum = 0, trif4Num = 0, haveInst = 1, haveInst1 = 1, bvhDepthBound = 0;     // haveInst = 0: single-level tree (bvhType "triangle4v")
  HcDevBuf instMatrices, instLightIds;
  int      numInst = 0;

  int width = 0, height = 0;
  HcDevBuf fbSum;                          // float4 per pixel: SUM of samples
  double   spp = 0.0;

  // ray-casting scratch (hc_trace_* with HC_HOST buffers)
  HcDevBuf scratchRays, scratchOut;
  HcDevBuf rcRays, rcHits, rcSRays, rcVis;  // hc_raycast_pass working set (device resident)
  cudaEvent_t evStage[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
  HcDevBuf counters;                       // device: [0] persistent-thread ray counter, [1..] compaction counters
  float    lastTraceMs = 0.0f;
  int      lastGroupPasses = 1;            // passes the last wavefront of hc_pt_pass carried (sample streams)
  int      sampleStreams = 1;              // generators per pixel (hc_pt_set_sample_streams)
  int64_t  maxPathsInFlight = 0;           // 0: max(W*H, 8M)

  // path tracing
  void*    pathHost = nullptr;             // HcPathHost (hc_path.cu): double-buffered SoA path state, hit / visibility buffers, tile ownership
  int      seed = 0;
  bool     ptReady = false;
  bool     sceneDirty = true;              // set by every scene upload entry point: hc_pt_pass re-validates the scene (and re-selects the shade kernel variant) before the next pass
  int      tileSize = 32, rank = 0, worldSize = 1;
  int      materialSort = 2 /* 0 off, 1 on, 2 auto: on for >= 3 materials and >= 384k paths per pass */, sortFromBounce = 1;   // K6b: sort the live-path queue by material before shading, from this bounce on
  HcDevBuf pixelRng;                       // uint2 per pixel: generator state carried across passes (trace.cl:6-13)
  HcDevBuf qmcTable;                       // Niederreiter table, 11 x 31 uint (qmc_sobol_niederreiter.cpp:179-186)
  unsigned passCounter = 0;

  // multi-GPU exchange (hc_comm.cu)
  void*    comm = nullptr;                 // ncclComm_t
  int      commRank = 0, commSize = 1;
  HcDevBuf commStage, commStage2, commPixels;          // dense staging of owned pixels; pixel lists (mine, or every source rank's on the destination)
  std::vector<int> commCount;              // pixels per source rank
  long long commPixelsKey = -1;            // (W, H, tile, G) the lists were built for
  long long commSegKey = -2;               // ... and the destination's segment table of the ray-casting gather
  HcDevBuf fbCombined;                     // destination rank, full-size sums (sample partition): sum over ranks, separate from fbSum
  bool     combinedValid = false;
  HcDevBuf fbOut;                          // read-back staging: normalised float4 image

  hc_stats stats{};
  int traceGrid = 0;
  int traceCounterSlot = 0;
  int traceRefill = 0, traceQBias = 0;     // 0 = built-in defaults; HC_TRACE_REFILL / HC_TRACE_QBIAS override
};

int hc_buf_reserve(hc_ctx* ctx, HcDevBuf& b, uint64_t bytes);   // grow-only
void hc_buf_free(HcDevBuf& b);

void hc_path_free(hc_ctx* ctx);   // hc_path.cu
void hc_comm_free(hc_ctx* ctx);   // hc_comm.cu
int  hc_comm_gather_raycast(hc_ctx* ctx, void* hits16, unsigned char* vis, int dstRank);
int  hc_path_owned_pixels(hc_ctx* ctx, const int** outDevicePixels, int* outCount);   // hc_path.cu: device list of this rank's pixels (hc_pt_set_tiles)
int  hc_launch_trace_counted(hc_ctx* ctx, bool anyHit, const float4* rpos, const float4* rdir, int stride, long long nUpper, const int* nDev, HcHit* hits, unsigned char* vis, int outStride, cudaStream_t stream = nullptr);
int  hc_launch_trace(hc_ctx* ctx, bool anyHit, const float4* rpos, const float4* rdir, int stride, long long n, HcHit* hits, unsigned char* vis);   // hc_api.cu
