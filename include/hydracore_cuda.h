/* hydracore_cuda.h — the thin C ABI between the C++ CUDA layer (GPUCUDALayer : IHWLayer, MemoryStorageCUDA : IMemoryStorage)
 * and the sm_100a kernels of libhydracore_b200.so.  POD only: plain pointers, sizes, ints.  Every entry point returns
 * 0 on success, a positive cudaError_t value, or a negative HC_E_* code; hc_last_error() gives the text.  The C++ layer turns
 * a non-zero status into RUN_TIME_ERROR (reference hydra_drv/globals_sys.h:56-62) or size_t(-1) where the reference does
 * (MemoryStorageOCL.cpp:21-38).  All file:line citations are relative to the reference tree (Ray-Tracing-Systems/HydraCore).
 *
 * Blob formats are the reference's own: offsets in float4 (16 B) units, BVHNode 32 B in quads of 4,
 * triangles 3 x float4 behind a one-float4 leaf header, PlainMaterial 192 floats, PlainLight 128 floats, EngineGlobals blob.
 */
#ifndef HYDRACORE_CUDA_H
#define HYDRACORE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HC_ABI_VERSION 2

enum { HC_OK = 0, HC_E_ARG = -1, HC_E_STATE = -2, HC_E_NOMEM = -3, HC_E_NODEVICE = -4, HC_E_RANGE = -5 };

/* storage slots: the five IMemoryStorage objects RenderDriverRTE creates (RenderDriverRTE.cpp:705-709) */
enum { HC_STORAGE_TEXTURES = 0, HC_STORAGE_TEXTURES_AUX = 1, HC_STORAGE_GEOM = 2, HC_STORAGE_MATERIALS = 3, HC_STORAGE_PDFS = 4,
       HC_STORAGE_COUNT = 5 };

/* integrators of the CPU oracle this layer reproduces (CPUExp_Integrators_PT.cpp:9, CPUExp_Integrators_PT_Loop.cpp:264,
 * CPUExp_Integrators_PT_QMC.cpp:5) */
enum { HC_INTEGRATOR_PT = 0, HC_INTEGRATOR_MISPT = 2, HC_INTEGRATOR_MISPT_QMC = 3 };

/* memory space of a caller-supplied buffer */
enum { HC_HOST = 0, HC_DEVICE = 1 };

typedef struct hc_ctx hc_ctx;            /* one per CUDA device: what GPUOCLLayer holds in m_globals/m_scene/m_rays (GPUOCLLayer.h) */
typedef struct hc_bvh hc_bvh;            /* host-side BVH4 builder: stands where IBVHBuilder2 stands (IBVHBuilderAPI.h:35-68)       */

/* 16-byte hit record == Lite_Hit (cglobals.h:1248-1256) */
typedef struct hc_hit { float t; int32_t primId; int32_t instId; int32_t geomId; } hc_hit;

/* counters of one traversal launch / one pass (MRaysStat, cglobals.h:1764-1787, is filled from these) */
typedef struct hc_stats
{
  uint64_t raysClosest;      /* closest-hit rays traced                         */
  uint64_t raysShadow;       /* any-hit rays traced                             */
  uint64_t paths;            /* eye paths started                               */
  uint64_t kernelLaunches;   /* launches of OUR kernels since the last reset    */
  float    msClosest;        /* device time (CUDA events) in closest-hit kernel */
  float    msShadow;         /* ... any-hit kernel                              */
  float    msShade;          /* ... surface/light/BSDF kernel                   */
  float    msOther;          /* ray generation, compaction, accumulation        */
} hc_stats;

/* ---------------------------------------------------------------- process / device ---------------------------------------- */
int         hc_abi_version(void);
const char* hc_last_error(void);
int         hc_device_count(int* outCount);                                  /* IHWLayer::ListDevices, IHWLayer.h:160          */
int         hc_ctx_create(int device, hc_ctx** out);                          /* CreateCudaImpl beside IHWLayer.h:256-257       */
void        hc_ctx_destroy(hc_ctx* ctx);
int         hc_device_name(hc_ctx* ctx, char* buf, int bufSize);             /* IHWLayer::GetDeviceName, IHWLayer.h:158        */
int         hc_mem_info(hc_ctx* ctx, size_t* freeBytes, size_t* totalBytes); /* GetAvaliableMemoryAmount, IHWLayer.h:152       */
int         hc_sync(hc_ctx* ctx);                                            /* IHWLayer::FinishAll, IHWLayer.h:137            */
int         hc_stream(hc_ctx* ctx, void** outCudaStream);                    /* the stream all kernels of ctx are launched on  */

/* ---------------------------------------------------------------- MemoryStorageCUDA --------------------------------------- */
int hc_storage_reserve(hc_ctx* ctx, int slot, uint64_t bytes);               /* IMemoryStorage::Reserve, MemoryStorageOCL.cpp:10 */
int hc_storage_write(hc_ctx* ctx, int slot, uint64_t offsetBytes, const void* data, uint64_t bytes); /* MemCopyAt, :53-56     */
int hc_storage_capacity(hc_ctx* ctx, int slot, uint64_t* outBytes);

/* ---------------------------------------------------------------- scene upload -------------------------------------------- */
int hc_set_globals(hc_ctx* ctx, const void* blob, uint64_t bytes);           /* PrepareEngineGlobals/Tables upload, GPUOCLData.cpp:289-327 */
int hc_set_bvh(hc_ctx* ctx, int treeId, const void* nodes, int nodesNum, const void* trif4, int trif4Num, int haveInst);
                                                                              /* SetAllBVH4(ConvertionResult), GPUOCLData.cpp:88-160       */
int hc_set_bvh_alpha(hc_ctx* ctx, int treeId, const void* nodes, int nodesNum, const void* trif4, int trif4Num,
                     const void* alphaTableUint2, int alphaNum, int haveInst);
                                                                              /* tree 1 of the ConvertionResult: meshes with opacity maps + pTriangleAlpha / triAfNum
                                                                                 (RenderDriverRTE_AlphaTestTable.cpp:66-221; consumer BVH4InstTraverseAlpha, ctrace.h:1297).
                                                                                 Send tree 0 first (hc_set_bvh forgets an earlier tree 1).  Closest-hit launches then walk
                                                                                 tree 0 and tree 1 with the hit carried along (IntegratorCommon::rayTrace,
                                                                                 CPUExp_Integrators_Common.cpp:122-150); shadow rays: see hc_pt_set_shadow_trees             */
int hc_bvh_device_layout(const void* nodes, int nodesNum, const void* trif4, int trif4Num, float* outNodesOrNull, float* outPairsOrNull,
                         int64_t outPairsCapacityFloats, int64_t* outPairsFloats, int* outStackBound);
                                                                              /* host only, no device needed: the re-layout hc_set_bvh applies before the upload (SoA quads, triangle
                                                                                 pair records) and the worst-case traversal stack; outNodes = nodesNum*8 floats */
int hc_set_inst_matrices(hc_ctx* ctx, const float* invMatrices16, int n);    /* SetAllInstMatrices, IHWLayer.h:117                        */
int hc_set_inst_light_ids(hc_ctx* ctx, const int32_t* lightInstId, int n);   /* SetAllInstLightInstId, IHWLayer.h:118                     */
int hc_set_remap_lists(hc_ctx* ctx, const int32_t* allLists, const int32_t* tableOffsetAndSize, int allSize, int tableSize);
                                                                              /* SetAllRemapLists, IHWLayer.h:122 (GPUOCLData.cpp:200-250): {from, to} material id pairs of
                                                                                 all lists back to back + int2 {offset, size} per list; nullptr / 0 clears                  */
int hc_set_inst_remap_ids(hc_ctx* ctx, const int32_t* instRemapListId, int n); /* SetAllInstIdToRemapId, IHWLayer.h:123: remap list id per instance or -1                 */
int hc_resize(hc_ctx* ctx, int width, int height);                           /* ResizeScreen, IHWLayer.h:147                              */

/* ---------------------------------------------------------------- ray casting (kernels K1, K2, K2s) ----------------------- */
/* rays: n x 8 floats {pos.xyz, tNear(unused, 0) | dir.xyz, tFar}.  `space` says where rays/out live (HC_HOST: copied inside). */
int hc_make_eye_rays(hc_ctx* ctx, int width, int height, const float* offsets4OrNull, float* rays8Out, int space);
                                                                              /* MakeEyeRaysUnifiedSampling, screen.cl:280 ; MakeRandEyeRay, cfetch.h:877 */
int hc_trace_closest(hc_ctx* ctx, const float* rays8, int64_t n, hc_hit* hitsOut, int space);  /* BVH4TraversalInstKernel, trace.cl:50  */
int hc_trace_shadow(hc_ctx* ctx, const float* rays8, int64_t n, uint8_t* visibleOut, int space);/* BVH4TraversalInstShadowKenrel, trace.cl:309 */
int hc_make_shadow_rays(hc_ctx* ctx, const float* rays8, const hc_hit* hits, int64_t n, const float lightPos[3], float* shadowRays8Out, int space);
                                                                              /* shadow-ray construction of LightSample for one point light (light.cl:140, clight.h:1561;
                                                                                 t_far = 0.995*distance, CPUExp_Integrators_PT_Loop.cpp:176); missed rays get t_far = 0 */
int hc_raycast_pass(hc_ctx* ctx, const float lightPos[3], hc_hit* hitsOutOrNull, uint8_t* visibleOutOrNull, int space);
                                                                              /* one ray-casting pass over the whole screen, all on the device: K1 eye rays -> K2 closest hit ->
                                                                                 shadow rays to lightPos -> K2s any hit.  The RT-mode pass of the OpenCL layer
                                                                                 (GPUOCLLayer::trace1DPrimaryOnly, GPUOCLLayerCore.cpp:294) plus the shadow step of ShadePass (:1007).
                                                                                 Results stay on the device unless out pointers are given (space says where they live). */
int hc_trace_last_ms(hc_ctx* ctx, float* outMs);
int hc_measure_read_bandwidth(hc_ctx* ctx, uint64_t bytes, int repeats, float* outGBs);
                                                                              /* streaming-read microbenchmark (L1 bypassed): bytes << 126 MB and repeats > 1 gives the L2 read
                                                                                 bandwidth, bytes >> 126 MB the HBM read bandwidth; roofline denominators measured in place */                             /* device time of the last hc_trace_* launch (CUDA events)   */

/* ---------------------------------------------------------------- path tracing -------------------------------------------- */
int hc_pt_init(hc_ctx* ctx, int seed);                                       /* InitPathTracing, IHWLayer.h:139 ; InitRandomGen, trace.cl:6 */
int hc_pt_set_tiles(hc_ctx* ctx, int tileSize, int rank, int worldSize);     /* interleaved tile ownership for multi-GPU        */
int hc_pt_set_material_sort(hc_ctx* ctx, int enable, int fromBounce);       /* material sort of the live-path queue before shading: replaces
                                                                                 bitonic_sort_gpu (bitonic_sort_gpu.cpp:90-158); enable 0 = off, 1 = on, 2 = auto (the default:
                                                                                 on with >= 3 materials and >= 384k paths per pass), from bounce 1 */
int hc_pt_set_shadow_trees(hc_ctx* ctx, int mode);                            /* shadow rays and the second (alpha-tested) BVH tree: 1 (default) = every tree is walked and a
                                                                                 cut-out occludes where its opacity texel passes, as GPUOCLLayer does (GPUOCLKernels.cpp:959-1000,
                                                                                 BVH4InstTraverseShadowAlphaS ctrace.h:1748; binary opacity only); 0 = first tree only, as the CPU
                                                                                 integrators do (IntegratorCommon::shadowTrace, CPUExp_Integrators_Common.cpp:163-171)        */
int hc_pt_set_sample_streams(hc_ctx* ctx, int streams, int64_t maxPathsInFlight);
                                                                              /* S generators per pixel (1..64, default 1): pass p of a pixel draws from stream p mod S, generator
                                                                                 index k*W*H + pixel.  Up to S consecutive passes are then independent and hc_pt_pass keeps them in
                                                                                 flight as ONE wavefront (small frames, or a GPU that owns 1/G of the tiles, no longer run
                                                                                 latency-bound launches); their sums are added to the frame in pass order, so the image depends on
                                                                                 (seed, S) only, not on how many passes shared a wavefront, the tile split or the GPU count.
                                                                                 MISPT-QMC: one generator per SAMPLE index and stream, the Sobol index keeps running over
                                                                                 pass*W*H + sample; its samples add with float atomics, so groupings agree to rounding only.
                                                                                 The OpenCL layer has the same degree of freedom: one RandomGen per slot of its ray block,
                                                                                 randGenState[MEGABLOCKSIZE] (GPUOCLLayer.cpp:131), whatever the frame size.
                                                                                 maxPathsInFlight: 0 = max(W*H, 8M).  Call before hc_pt_init.                                 */
int hc_pt_group_passes(hc_ctx* ctx, int* outPasses);                         /* passes one wavefront carries with the current streams / tiles / limit (after hc_pt_init)    */
int hc_pt_pass(hc_ctx* ctx, int integrator, int passes);                     /* BeginTracingPass+EndTracingPass, IHWLayer.h:133-134         */
int hc_fb_clear(hc_ctx* ctx);                                                /* ClearAccumulatedColor, IHWLayer.h:140                       */
int hc_fb_device_ptr(hc_ctx* ctx, float** outSumRGBA, int64_t* outFloats);   /* per-pixel SUM buffer (for the NCCL reduce over NVLink)      */
int hc_fb_read_hdr(hc_ctx* ctx, float* outRGBA, int width, int height);      /* GetHDRImage: sum / spp, GPUOCLLayer.cpp:1184-1215           */
int hc_fb_read_sum(hc_ctx* ctx, float* outRGBA, int width, int height);      /* the raw per-pixel SUMS (what the OpenCL layer adds into the shared image,
                                                                                 GPUOCLLayerOther.cpp:365-430)                               */
int hc_fb_read_ldr(hc_ctx* ctx, uint32_t* outRGBA8, int width, int height);  /* GetLDRImage, IHWLayer.h:149                                 */
/* multi-GPU: one process (or host thread) per GPU, scene replicated, framebuffers combined over NVLink by NCCL inside the library.
 * Replaces the shared-memory image + mutex of the reference's process-per-GPU mode (GPUOCLLayerOther.cpp:365-430). */
int hc_comm_unique_id(void* out128);                                         /* ncclGetUniqueId: rank 0 creates it, the host distributes the 128 bytes */
int hc_comm_init(hc_ctx* ctx, const void* uniqueId128, int rank, int nranks); /* ncclCommInitRank on the context's device                              */
int hc_comm_version(int* outVersion);                                        /* version of the NCCL bound at run time (dlopen libnccl.so.2)            */
int hc_fb_reduce(hc_ctx* ctx, int dstRank, int mode, float* outMsOrNull);    /* combine the per-rank SUM buffers on dstRank.  mode 0: interleaved tile partition
                                                                                 (hc_pt_set_tiles with the communicator's rank / size): every rank sends only the pixels
                                                                                 it owns (1/G of the image), the destination scatters them into its own SUM buffer -
                                                                                 repeatable, nothing is counted twice.  mode 1: full-size buffers (sample partition of
                                                                                 MISPT-QMC): ncclReduce(sum) into a separate buffer that hc_fb_read_* then read.  Without a
                                                                                 communicator (one GPU) it is a no-op.  outMs = device time of the exchange on this rank */
int hc_get_spp(hc_ctx* ctx, float* outSpp);                                  /* GetSPP, IHWLayer.h:207                                      */
int hc_get_stats(hc_ctx* ctx, hc_stats* out);                                /* GetRaysStat, IHWLayer.h:155                                 */
int hc_reset_stats(hc_ctx* ctx);                                             /* ResetPerfCounters, IHWLayer.h:145                           */

/* ---------------------------------------------------------------- BVH4 builder (host) ------------------------------------- */
/* Stands in for libhydrabvhbuilder (bvh_builder/bvh_access_dll2.cpp) whose Embree 2.17 back end is not available: binned-SAH
 * BVH4 per mesh + BVH4 over instances, flattened to the SAME two-level layout ConvertMap() emits (bvh_access_dll2.cpp:388-717). */
int  hc_bvh_create(hc_bvh** out);                                            /* CreateBuilder2, bvh_access_dll2.cpp:801        */
void hc_bvh_destroy(hc_bvh* b);
int  hc_bvh_add_mesh(hc_bvh* b, const float* vert4f, int numVert, const int32_t* indices, int numIndices, int* outMeshId);
int  hc_bvh_add_instance(hc_bvh* b, int meshId, const float* matrixRowMajor16, int* outInstId);
                                                                              /* InstanceTriangleMeshes, bvh_access_dll2.cpp:143 */
int  hc_bvh_add_instance_id(hc_bvh* b, int meshId, const float* matrixRowMajor16, int realInstId);
                                                                              /* the same with the scene-wide instance id given by the caller
                                                                               * (a_realInstIdBase of InstanceTriangleMeshes): one builder per tree,
                                                                               * opaque meshes in tree 0, meshes with opacity maps in tree 1
                                                                               * (RenderDriverRTE.cpp:1989-1991)                               */
int  hc_bvh_commit(hc_bvh* b);                                               /* CommitScene + ConvertMap                        */
int  hc_bvh_result(hc_bvh* b, const void** nodes, int* nodesNum, const void** trif4, int* trif4Num,
                   const float** invMatrices16, int* numInst, int* maxStackDepth);
int  hc_bvh_bounds(hc_bvh* b, float bmin[3], float bmax[3]);                 /* IBVHBuilder2::GetBounds                         */

#ifdef __cplusplus
}
#endif
#endif /* HYDRACORE_CUDA_H */
