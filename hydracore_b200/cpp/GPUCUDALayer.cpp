#include "GPUCUDALayer.h"
#include <cstring>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cwchar>
#include <sstream>
#include <stdexcept>

static const char* const kStorageNames[HC_STORAGE_COUNT] = { "textures", "textures_aux", "geom", "materials", "pdfs" };   // RenderDriverRTE.cpp:705-709

void GPUCUDALayer::Check(int rc, const char* what) const
{
  if (rc == HC_OK) return;
  std::ostringstream s;
  s << "GPUCUDALayer::" << what << ": status " << rc << ": " << hc_last_error();
  throw std::runtime_error(s.str());                            // RUN_TIME_ERROR semantics (globals_sys.h:56-62), caught in main.cpp:331-338
}

GPUCUDALayer::GPUCUDALayer(int w, int h, int a_flags, int a_deviceId)
  : m_ctx(nullptr), m_initFlags(a_flags), m_integratorOverride(-1), m_globalsDirty(true), m_ptInitialised(false), m_seed(0), m_spp(0.0f),
    m_sppContributed(0.0f)
{
  Check(hc_ctx_create(a_deviceId, &m_ctx), "GPUCUDALayer (hc_ctx_create)");
  memset(&m_stat, 0, sizeof(m_stat));
  if (w > 0 && h > 0) ResizeScreen(w, h, a_flags);
}

GPUCUDALayer::~GPUCUDALayer()
{
  if (m_ctx) hc_ctx_destroy(m_ctx);                             // the storages belong to the driver (IHWLayerDataAssembler.cpp:66-78)
  m_ctx = nullptr;
}

void GPUCUDALayer::Clear(CLEAR_FLAGS a_flags)
{
  // what GPUOCLLayer does: the storages are cleared by the driver through IMemoryStorage::Clear; here only derived state goes
  if (a_flags & CLEAR_GEOMETRY) m_ptInitialised = false;
  m_globalsDirty = true;
}

IMemoryStorage* GPUCUDALayer::CreateMemStorage(uint64_t a_maxSizeInBytes, const char* a_name)
{
  int slot = -1;
  for (int i = 0; i < HC_STORAGE_COUNT; i++) if (std::string(a_name) == kStorageNames[i]) slot = i;
  if (slot < 0) Check(HC_E_ARG, "CreateMemStorage (unknown storage name)");
  IMemoryStorage* pStorage = nullptr;
  if (slot == HC_STORAGE_GEOM || slot == HC_STORAGE_TEXTURES)   // host mirror needed by the driver (GPUOCLData.cpp:56-63)
    pStorage = new MemoryStorageBothCPUAndCUDA(new LinearStorageCPU(), new MemoryStorageCUDA(m_ctx, slot));
  else
    pStorage = new MemoryStorageCUDA(m_ctx, slot);
  if (pStorage->Reserve(a_maxSizeInBytes) == size_t(-1)) { delete pStorage; Check(HC_E_NOMEM, "CreateMemStorage (device allocation failed)"); }
  m_allMemStorages[a_name] = pStorage;
  return pStorage;
}

void GPUCUDALayer::PrepareEngineGlobals()
{
  Base::PrepareEngineGlobals();                                 // SetQMCVarRemapTable + header copy into m_cdataPrepared (IHWLayerDataAssembler.cpp:326-339)
  m_globalsDirty = true;
}

void GPUCUDALayer::PrepareEngineTables()
{
  Base::PrepareEngineTables();                                  // id -> offset tables of the five storages + light selector tables (:346-388)
  m_globalsDirty = true;
  UploadGlobalsIfDirty();
}

void GPUCUDALayer::UploadGlobalsIfDirty()
{
  if (!m_globalsDirty || m_cdataPrepared.empty()) return;
  memcpy(m_cdataPrepared.data(), &m_globsBuffHeader, sizeof(EngineGlobals));   // the header may have changed since (SetAllFlagsAndVars, SetCamMatrices)
  Check(hc_set_globals(m_ctx, m_cdataPrepared.data(), uint64_t(m_cdataPrepared.size())*sizeof(int)), "PrepareEngineGlobals (hc_set_globals)");
  m_globalsDirty = false;
}

void GPUCUDALayer::SetAllBVH4(const ConvertionResult& a_convertedBVH, IBVHBuilder2* a_inBuilderAPI, int a_flags)
{
  (void)a_inBuilderAPI; (void)a_flags;
  if (a_convertedBVH.treesNum < 1) Check(HC_E_ARG, "SetAllBVH4 (no trees)");
  if (a_convertedBVH.treesNum > 2) Check(HC_E_ARG, "SetAllBVH4 (more than two BVH trees: the driver uses tree 0 for opaque meshes and tree 1 for meshes with opacity maps)");
  if (a_convertedBVH.pTriangleAlpha[0] != nullptr) Check(HC_E_ARG, "SetAllBVH4 (alpha-test table on tree 0 is not supported)");
  // the ConvertionResult pointers die at ConvertUnmap (RenderDriverRTE.cpp:1436): hc_set_bvh converts and uploads before it returns
  for (int i = 0; i < a_convertedBVH.treesNum; i++)
  {
    if (i > 0 && a_convertedBVH.nodesNum[i] <= 4) continue;                                                                  // empty second tree
    const bool haveInst = (std::string(a_convertedBVH.bvhType[i] ? a_convertedBVH.bvhType[i] : "") != "triangle4v");       // GPUOCLData.cpp:118
    if (i == 0)
      Check(hc_set_bvh(m_ctx, 0, a_convertedBVH.pBVH[0], a_convertedBVH.nodesNum[0], a_convertedBVH.pTriangleData[0], a_convertedBVH.trif4Num[0], haveInst ? 1 : 0),
            "SetAllBVH4 (hc_set_bvh)");
    else                                                                                                                     // tree 1 + pTriangleAlpha (may be null), GPUOCLData.cpp:103-116
      Check(hc_set_bvh_alpha(m_ctx, i, a_convertedBVH.pBVH[i], a_convertedBVH.nodesNum[i], a_convertedBVH.pTriangleData[i], a_convertedBVH.trif4Num[i],
                             a_convertedBVH.pTriangleAlpha[i], a_convertedBVH.triAfNum[i], haveInst ? 1 : 0), "SetAllBVH4 (hc_set_bvh_alpha)");
  }
}

void GPUCUDALayer::SetAllInstMatrices(const float4x4* a_matrices, int32_t a_matrixNum)
{
  if (a_matrices == nullptr || a_matrixNum == 0) return;
  Check(hc_set_inst_matrices(m_ctx, reinterpret_cast<const float*>(a_matrices), a_matrixNum), "SetAllInstMatrices");   // float4x4 = 4 column float4 (cglobals.h:206-209)
}

void GPUCUDALayer::SetAllInstLightInstId(const int32_t* a_lightInstIds, int32_t a_instNum)
{
  if (a_lightInstIds == nullptr || a_instNum == 0) return;
  Check(hc_set_inst_light_ids(m_ctx, a_lightInstIds, a_instNum), "SetAllInstLightInstId");
}

void GPUCUDALayer::SetAllRemapLists(const int* a_allLists, const int2* a_table, int a_allSize, int a_tableSize)
{
  // per-instance material overrides (HydraAPI remap lists); empty input clears them, as GPUOCLData.cpp:203-210 does
  Check(hc_set_remap_lists(m_ctx, a_allLists, reinterpret_cast<const int32_t*>(a_table), a_allSize, a_tableSize), "SetAllRemapLists");
}

void GPUCUDALayer::SetAllInstIdToRemapId(const int* a_allInstId, int a_instNum)
{
  Check(hc_set_inst_remap_ids(m_ctx, a_allInstId, a_instNum), "SetAllInstIdToRemapId");
}

void GPUCUDALayer::SetAllPODLights(PlainLight* a_lights2, size_t a_number)
{
  Base::SetAllPODLights(a_lights2, a_number);                   // lights + sky / sun bookkeeping into m_cdataPrepared (IHWLayerDataAssembler.cpp:390-452)
  m_globalsDirty = true;
}

void GPUCUDALayer::SetAllFlagsAndVars(const AllRenderVarialbes& a_vars)
{
  Base::SetAllFlagsAndVars(a_vars);
  m_globalsDirty = true;                                        // re-uploaded before the next pass (UpdateVarsOnGPU of the OpenCL layer, GPUOCLData.cpp:277-287)
}

void GPUCUDALayer::ResizeScreen(int w, int h, int a_flags)
{
  Base::ResizeScreen(w, h, a_flags);
  Check(hc_resize(m_ctx, w, h), "ResizeScreen");
  m_ptInitialised = false;
  m_spp = 0.0f;
}

int GPUCUDALayer::IntegratorFromState() const
{
  if (m_integratorOverride >= 0) return m_integratorOverride;
  if (m_vars.m_flags & HRT_STUPID_PT_MODE) return HC_INTEGRATOR_PT;                 // IntegratorStupidPT
  if (m_vars.m_varsI[HRT_KMLT_OR_QMC_MAT_BOUNCES] != 0) return HC_INTEGRATOR_MISPT_QMC;   // the OpenCL layer's QMC switch (GPUOCLLayer.cpp:1436-1446)
  return HC_INTEGRATOR_MISPT;                                                       // what CPUExpLayer instantiates (IHWLayerDataAssembler.cpp:579)
}

void GPUCUDALayer::InitPathTracing(int seed, std::vector<int32_t>* pInstRemapTable)
{
  (void)pInstRemapTable;
  UploadGlobalsIfDirty();
  m_seed = seed;
  Check(hc_pt_init(m_ctx, seed), "InitPathTracing");
  m_ptInitialised = true;
  m_spp = 0.0f; m_sppContributed = 0.0f;
}

void GPUCUDALayer::BeginTracingPass()
{
  UploadGlobalsIfDirty();
  if (!(m_vars.m_flags & HRT_UNIFIED_IMAGE_SAMPLING))
  {
    // Without unified image sampling the OpenCL layer casts the primary rays and draws debug normals (DrawNormals, GPUOCLLayer.cpp:1455-1458).
    // Here the same branch is the ray-casting pass: primary rays + one shadow ray per hit towards the point set by
    // CallNamedFunc("raycast_light"); hit records and visibility go to the host buffers registered by CallNamedFunc("raycast_results").
    if (m_rcLightSet)
      Check(hc_raycast_pass(m_ctx, m_rcLight, static_cast<hc_hit*>(m_rcHits), static_cast<uint8_t*>(m_rcVis), HC_HOST), "BeginTracingPass (ray casting)");
    return;
  }
  if (!m_ptInitialised) InitPathTracing(m_seed);
  // one call = one wavefront: with sample streams (CallNamedFunc("sample_streams")) it carries several passes of the owned pixels, as one
  // ray block of the OpenCL layer carries more than one sample per pixel of a small frame (MEGABLOCKSIZE, GPUOCLLayer.cpp:103); GetSPP() tells
  int passes = 1;
  Check(hc_pt_group_passes(m_ctx, &passes), "BeginTracingPass (hc_pt_group_passes)");
  Check(hc_pt_pass(m_ctx, IntegratorFromState(), passes), "BeginTracingPass");
}

void GPUCUDALayer::EndTracingPass()
{
  Check(hc_sync(m_ctx), "EndTracingPass");                      // clFinish of the OpenCL layer (GPUOCLLayer.cpp:1486-1491)
  Check(hc_get_spp(m_ctx, &m_spp), "EndTracingPass (hc_get_spp)");
  if (m_pExternalImage != nullptr) ContribToExternalImageAccumulator(m_pExternalImage);
}

void GPUCUDALayer::FinishAll() { Check(hc_sync(m_ctx), "FinishAll"); }

void GPUCUDALayer::ClearAccumulatedColor()
{
  Check(hc_fb_clear(m_ctx), "ClearAccumulatedColor");
  m_spp = 0.0f;
}

void GPUCUDALayer::ResetPerfCounters()
{
  memset(&m_stat, 0, sizeof(MRaysStat));
  Check(hc_reset_stats(m_ctx), "ResetPerfCounters");
}

void GPUCUDALayer::GetLDRImage(uint32_t* data, int width, int height) const { Check(hc_fb_read_ldr(m_ctx, data, width, height), "GetLDRImage"); }
void GPUCUDALayer::GetHDRImage(float4* data, int width, int height)   const { Check(hc_fb_read_hdr(m_ctx, reinterpret_cast<float*>(data), width, height), "GetHDRImage"); }

size_t GPUCUDALayer::GetAvaliableMemoryAmount(bool allMem)
{
  size_t f = 0, t = 0;
  Check(hc_mem_info(m_ctx, &f, &t), "GetAvaliableMemoryAmount");
  return allMem ? t : f;
}

MRaysStat GPUCUDALayer::GetRaysStat()
{
  hc_stats s; Check(hc_get_stats(m_ctx, &s), "GetRaysStat");
  const float total = s.msClosest + s.msShadow + s.msShade + s.msOther;
  m_stat.traversalTimeMs  = s.msClosest;
  m_stat.shadowTimeMs     = s.msShadow;
  m_stat.shadeTimeMs      = s.msShade;
  m_stat.traceTimePerCent = total > 0.0f ? int(100.0f*(s.msClosest + s.msShadow)/total) : 0;
  m_stat.raysPerSec       = total > 0.0f ? float(double(s.raysClosest + s.raysShadow)/(1e-3*double(total))) : 0.0f;
  m_stat.samplesPerSec    = total > 0.0f ? float(double(s.paths)/(1e-3*double(total))) : 0.0f;
  return m_stat;
}

const char* GPUCUDALayer::GetDeviceName(int* pOCLVer) const
{
  if (pOCLVer) *pOCLVer = 0;
  char buf[256]; Check(hc_device_name(m_ctx, buf, 256), "GetDeviceName");
  m_deviceName = buf;
  return m_deviceName.c_str();
}

const HRRenderDeviceInfoListElem* GPUCUDALayer::ListDevices() const
{
  int n = 0; hc_device_count(&n);
  m_deviceList.assign(size_t(n), HRRenderDeviceInfoListElem());
  for (int i = 0; i < n; i++)
  {
    HRRenderDeviceInfoListElem& e = m_deviceList[size_t(i)];
    memset(&e, 0, sizeof(e));
    e.id = i;
    swprintf(e.name, 256, L"CUDA device %d", i);
    wcsncpy(e.driver, L"CUDA (sm_100a)", 255);
    e.isCPU = false; e.isEnabled = false;
    e.next = (i + 1 < n) ? &m_deviceList[size_t(i) + 1] : nullptr;
  }
  return m_deviceList.empty() ? nullptr : m_deviceList.data();
}

void GPUCUDALayer::CallNamedFunc(const char* a_name, const char* a_args)
{
  const std::string name = a_name ? a_name : "", args = a_args ? a_args : "";
  if (name == "integrator")
  {
    if      (args == "pt")    m_integratorOverride = HC_INTEGRATOR_PT;
    else if (args == "mispt") m_integratorOverride = HC_INTEGRATOR_MISPT;
    else if (args == "qmc")   m_integratorOverride = HC_INTEGRATOR_MISPT_QMC;
    else if (args == "auto")  m_integratorOverride = -1;
    else Check(HC_E_ARG, "CallNamedFunc(integrator): expected pt | mispt | qmc | auto");
  }
  else if (name == "tiles")
  {
    int tile = 32, rank = 0, world = 1;
    if (sscanf(args.c_str(), "%d %d %d", &tile, &rank, &world) != 3) Check(HC_E_ARG, "CallNamedFunc(tiles): expected \"<tileSize> <rank> <worldSize>\"");
    Check(hc_pt_set_tiles(m_ctx, tile, rank, world), "CallNamedFunc(tiles)");
    m_ptInitialised = false;
  }
  else if (name == "raycast_light")
  {
    if (sscanf(args.c_str(), "%f %f %f", &m_rcLight[0], &m_rcLight[1], &m_rcLight[2]) != 3) Check(HC_E_ARG, "CallNamedFunc(raycast_light): expected \"<x> <y> <z>\"");
    m_rcLightSet = true;
  }
  else if (name == "raycast_results")
  {
    // "<hits> <visibility>": host addresses (decimal) of W*H hit records (16 B each, Lite_Hit) and W*H bytes; 0 = keep that result on the device
    unsigned long long a = 0, b = 0;
    if (sscanf(args.c_str(), "%llu %llu", &a, &b) != 2) Check(HC_E_ARG, "CallNamedFunc(raycast_results): expected \"<hitsAddress> <visibilityAddress>\"");
    m_rcHits = reinterpret_cast<void*>(static_cast<uintptr_t>(a)); m_rcVis = reinterpret_cast<void*>(static_cast<uintptr_t>(b));
  }
  else if (name == "shadow_trees")
  {
    int mode = 1;
    if (sscanf(args.c_str(), "%d", &mode) != 1) Check(HC_E_ARG, "CallNamedFunc(shadow_trees): expected 0 | 1");
    Check(hc_pt_set_shadow_trees(m_ctx, mode), "CallNamedFunc(shadow_trees)");
  }
  else if (name == "sample_streams")
  {
    int streams = 1; long long limit = 0;
    if (sscanf(args.c_str(), "%d %lld", &streams, &limit) < 1) Check(HC_E_ARG, "CallNamedFunc(sample_streams): expected \"<streams> [maxPathsInFlight]\"");
    Check(hc_pt_set_sample_streams(m_ctx, streams, limit), "CallNamedFunc(sample_streams)");
    m_ptInitialised = false;                                  // the generator array changes: the next pass initialises it again
  }
  else if (name == "comm_id")
  {
    // rank 0 of a multi-process render creates the NCCL unique id; the host hands its 256 hex digits to every process ("comm" below)
    unsigned char id[128];
    Check(hc_comm_unique_id(id), "CallNamedFunc(comm_id)");
    static const char* hex = "0123456789abcdef";
    m_commIdHex.assign(256, '0');
    for (int i = 0; i < 128; i++) { m_commIdHex[2*i] = hex[id[i] >> 4]; m_commIdHex[2*i + 1] = hex[id[i] & 15]; }
  }
  else if (name == "comm")
  {
    // "<rank> <nranks> <256 hex digits>": join the communicator of the one-process-per-GPU render (what the shared-memory image is in the
    // reference, GPUOCLLayerOther.cpp:365-430)
    int rank = 0, world = 1; char hexId[260] = { 0 };
    if (sscanf(args.c_str(), "%d %d %256s", &rank, &world, hexId) != 3 || strlen(hexId) != 256) Check(HC_E_ARG, "CallNamedFunc(comm): expected \"<rank> <nranks> <256 hex digits>\"");
    unsigned char id[128];
    auto nib = [](char c) -> int { return (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : 0; };
    for (int i = 0; i < 128; i++) id[i] = (unsigned char)((nib(hexId[2*i]) << 4) | nib(hexId[2*i + 1]));
    Check(hc_comm_init(m_ctx, id, rank, world), "CallNamedFunc(comm)");
  }
  else if (name == "reduce")
  {
    // "<dstRank> <mode>": combine the framebuffers of all processes on dstRank (mode 0 tile partition, 1 full-size sum); GetHDRImage /
    // GetLDRImage on dstRank then return the whole frame
    int dst = 0, mode = 0;
    if (sscanf(args.c_str(), "%d %d", &dst, &mode) != 2) Check(HC_E_ARG, "CallNamedFunc(reduce): expected \"<dstRank> <mode>\"");
    Check(hc_fb_reduce(m_ctx, dst, mode, nullptr), "CallNamedFunc(reduce)");
  }
}

// One process per GPU adds its SUM buffer into the shared image under the image's lock and hands over its sample count: the
// multi-process mode of the reference (GPUOCLLayerOther.cpp:365-430; README.md:99-103).  Every process renders the FULL frame with its
// own seed in this mode (no tile ownership), so that Header()->spp, the sum of the contributed sample counts, normalises every pixel.
void GPUCUDALayer::ContribToExternalImageAccumulator(IHRSharedAccumImage* a_pImage)
{
  if (a_pImage == nullptr || m_spp <= 0.0f) return;
  std::vector<float> sums(size_t(m_width)*size_t(m_height)*4);
  Check(hc_fb_read_sum(m_ctx, sums.data(), m_width, m_height), "ContribToExternalImageAccumulator");
  if (!a_pImage->Lock(100)) return;                             // try again after the next pass, like the reference
  HRSharedBufferHeader* h = a_pImage->Header();
  float* dst = a_pImage->ImageData(0);
  bool done = false;
  if (h != nullptr && dst != nullptr && h->width == m_width && h->height == m_height)
  {
    const size_t n = size_t(m_width)*size_t(m_height)*4;
    for (size_t i = 0; i < n; i++) dst[i] += sums[i];
    h->counterRcv++;
    h->spp += m_spp;
    done = true;
  }
  a_pImage->Unlock();
  if (done)
  {
    m_sppContributed += m_spp;
    ClearAccumulatedColor();                                   // the device buffer starts over: only not-yet-contributed samples live there
  }
}

IHWLayer* CreateCudaImpl(int w, int h, int a_flags, int a_deviceId) { return new GPUCUDALayer(w, h, a_flags, a_deviceId); }
