// layer_harness.cpp — TEST HARNESS ONLY.  A flat C interface over the IHWLayer VIRTUAL interface (reference hydra_drv/IHWLayer.h:97-246)
// so that the Python GPU tests can drive GPUCUDALayer exactly the way RenderDriverRTE drives a layer: every call below goes through an
// IHWLayer* / IMemoryStorage* obtained from CreateCudaImpl, never through GPUCUDALayer's own type.  The call order in the tests follows
// RenderDriverRTE (CreateMemStorage x5 RenderDriverRTE.cpp:705-709, Update* per object, SetAllBVH4 :1436-1460, SetAllInstMatrices /
// SetAllInstLightInstId :1997-2008, SetAllPODLights / SetAllLightsSelectTable :1499-1521, PrepareEngineGlobals/Tables :1723-1725,
// InitPathTracing, BeginTracingPass / EndTracingPass :1846-1848, GetHDRImage).
#include "GPUCUDALayer.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <string>

namespace
{
  thread_local std::string g_err;
  struct Session
  {
    IHWLayer* layer = nullptr;
    std::map<std::string, IMemoryStorage*> storages;      // owned here, as the render driver owns them (RenderDriverRTE.cpp:406-410)
    ~Session() { delete layer; for (auto& kv : storages) delete kv.second; }
  };
  template<class F> int Guard(F&& f)
  {
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
    catch (...) { g_err = "unknown exception"; return -1; }
  }
}

namespace
{
  // in-process stand-in for HydraAPI's shared-memory accumulation image (one layer of float4 sums, a header, a lock)
  struct FakeSharedImage : public IHRSharedAccumImage
  {
    HRSharedBufferHeader hdr{}; std::vector<float> data; bool locked = false; int lockCalls = 0;
    bool  Create(int w, int h, int d, const char*, char[256]) override { hdr = HRSharedBufferHeader{}; hdr.width = w; hdr.height = h; hdr.depth = d; hdr.channels = 4; data.assign(size_t(w)*h*4, 0.0f); return true; }
    bool  Attach(const char*, char[256]) override { return true; }
    void  Clear() override { std::fill(data.begin(), data.end(), 0.0f); hdr.spp = 0.0f; hdr.counterRcv = 0; }
    bool  Lock(int) override { lockCalls++; if (locked) return false; locked = true; return true; }
    void  Unlock() override { locked = false; }
    float* ImageData(int) override { return data.data(); }
    char*  MessageSendData() override { return nullptr; }
    char*  MessageRcvData() override { return nullptr; }
    HRSharedBufferHeader* Header() override { return &hdr; }
  };
}

extern "C"
{
const char* hl_last_error() { return g_err.c_str(); }

void* hl_shared_image_create(int w, int h) { FakeSharedImage* im = new FakeSharedImage; char err[256]; im->Create(w, h, 1, "test", err); return im; }
void  hl_shared_image_destroy(void* im) { delete static_cast<FakeSharedImage*>(im); }
int   hl_shared_image_read(void* im, float* out, float* spp, int* counterRcv)
{
  FakeSharedImage* s = static_cast<FakeSharedImage*>(im);
  memcpy(out, s->data.data(), s->data.size()*4); *spp = s->hdr.spp; *counterRcv = s->hdr.counterRcv; return s->locked ? 1 : 0;
}
int hl_contribute(void* p, void* im)
{ Session* s = static_cast<Session*>(p); return Guard([&] { s->layer->ContribToExternalImageAccumulator(static_cast<FakeSharedImage*>(im)); }); }
float hl_get_spp_contrib(void* p) { return static_cast<Session*>(p)->layer->GetSPPContrib(); }

void* hl_create(int w, int h, int flags, int deviceId)
{
  Session* s = new Session;
  if (Guard([&] { s->layer = CreateCudaImpl(w, h, flags, deviceId); }) != 0) { delete s; return nullptr; }
  return s;
}
void hl_destroy(void* p) { delete static_cast<Session*>(p); }

int hl_create_storage(void* p, const char* name, uint64_t bytes)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&] { s->storages[name] = s->layer->CreateMemStorage(bytes, name); });
}
int hl_storage_update(void* p, const char* name, int id, const void* data, uint64_t bytes, int* outOffset)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&] { const int32_t off = s->storages.at(name)->Update(id, data, bytes); if (outOffset) *outOffset = off; });
}
int hl_storage_info(void* p, const char* name, uint64_t* size, uint64_t* capacity, int* hasHostMirror)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&] { IMemoryStorage* m = s->storages.at(name); *size = m->GetSize(); *capacity = m->GetCapacity(); *hasHostMirror = (m->GetBegin() != nullptr) ? 1 : 0; });
}
int hl_resize_tables(void* p, int geomNum, int imgNum, int matNum, int lightNum)
{ Session* s = static_cast<Session*>(p); return Guard([&] { s->layer->ResizeTablesForEngineGlobals(geomNum, imgNum, matNum, lightNum); }); }

int hl_set_bvh(void* p, const void* nodes, int nodesNum, const float* trif4, int trif4Num, const char* bvhType)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&]
  {
    ConvertionResult r;
    r.treesNum = 1; r.bvhType[0] = bvhType; r.pBVH[0] = static_cast<const BVHNode*>(nodes); r.nodesNum[0] = nodesNum;
    r.pTriangleData[0] = trif4; r.trif4Num[0] = trif4Num;
    s->layer->SetAllBVH4(r, nullptr, 0);
  });
}
int hl_set_bvh2(void* p, const void* nodes, int nodesNum, const float* trif4, int trif4Num, const void* nodes1, int nodesNum1, const float* trif41, int trif4Num1,
                const void* alpha1, int alphaNum1, const char* bvhType)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&]
  {
    ConvertionResult r;
    r.treesNum = 2;
    r.bvhType[0] = bvhType; r.pBVH[0] = static_cast<const BVHNode*>(nodes); r.nodesNum[0] = nodesNum; r.pTriangleData[0] = trif4; r.trif4Num[0] = trif4Num;
    r.bvhType[1] = bvhType; r.pBVH[1] = static_cast<const BVHNode*>(nodes1); r.nodesNum[1] = nodesNum1; r.pTriangleData[1] = trif41; r.trif4Num[1] = trif4Num1;
    r.pTriangleAlpha[1] = static_cast<const uint2*>(alpha1); r.triAfNum[1] = alphaNum1;
    s->layer->SetAllBVH4(r, nullptr, 0);
  });
}
int hl_set_remap(void* p, const int* allLists, int allSize, const int* tableOffsetAndSize, int tableSize, const int* instRemapId, int nInst)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&] { s->layer->SetAllRemapLists(allLists, reinterpret_cast<const int2*>(tableOffsetAndSize), allSize, tableSize);
                     s->layer->SetAllInstIdToRemapId(instRemapId, nInst); });
}
int hl_set_instances(void* p, const float* invMatrices16, const int32_t* lightInstIds, int n)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&] { s->layer->SetAllInstMatrices(reinterpret_cast<const float4x4*>(invMatrices16), n); s->layer->SetAllInstLightInstId(lightInstIds, n); });
}
int hl_set_lights(void* p, const float* lights128, int n, const float* selectTable, int tableSize)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&]
  {
    std::vector<PlainLight> tmp(size_t(n > 0 ? n : 1));
    if (n > 0) memcpy(tmp.data(), lights128, size_t(n)*sizeof(PlainLight));
    s->layer->SetAllLightsSelectTable(selectTable, tableSize, false);       // driver order: tables first, then the lights (RenderDriverRTE.cpp:1517-1519)
    s->layer->SetAllLightsSelectTable(selectTable, tableSize, true);
    s->layer->SetAllPODLights(n > 0 ? tmp.data() : nullptr, size_t(n));
  });
}
int hl_set_camera(void* p, const float projInv[16], const float worldViewInv[16], const float proj[16], const float worldView[16], float aspect, float fovX,
                  const float lookAt[3])
{
  Session* s = static_cast<Session*>(p);
  return Guard([&]
  {
    float a[16], b[16], c[16], d[16];
    memcpy(a, projInv, 64); memcpy(b, worldViewInv, 64); memcpy(c, proj, 64); memcpy(d, worldView, 64);
    s->layer->SetCamMatrices(a, b, c, d, aspect, fovX, float3(lookAt[0], lookAt[1], lookAt[2]));
  });
}
int hl_get_vars(void* p, int* varsI64, float* varsF64, unsigned* flags)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&] { const AllRenderVarialbes v = s->layer->GetAllFlagsAndVars(); memcpy(varsI64, v.m_varsI, 64*4); memcpy(varsF64, v.m_varsF, 64*4); *flags = v.m_flags; });
}
int hl_set_vars(void* p, const int* varsI64, const float* varsF64, unsigned flags)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&] { AllRenderVarialbes v; memcpy(v.m_varsI, varsI64, 64*4); memcpy(v.m_varsF, varsF64, 64*4); v.m_flags = flags; s->layer->SetAllFlagsAndVars(v); });
}
int hl_prepare(void* p) { Session* s = static_cast<Session*>(p); return Guard([&] { s->layer->PrepareEngineGlobals(); s->layer->PrepareEngineTables(); }); }
int hl_globals_blob(void* p, int* out, int maxInts, int* outInts)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&]
  {
    const EngineGlobals* g = s->layer->GetEngineGlobals();
    const int n = g->lightsOffset + g->lightsSize;           // the last region of the blob (CalcConstGlobDataOffsets, IHWLayerDataAssembler.cpp:149-170)
    *outInts = n;
    if (out && maxInts >= n) memcpy(out, g, size_t(n)*4);
  });
}
int hl_comm_id_hex(void* p, char* out257)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&] { s->layer->CallNamedFunc("comm_id", ""); const std::string& h = static_cast<GPUCUDALayer*>(s->layer)->CommIdHex(); memcpy(out257, h.c_str(), 257); });
}
int hl_call(void* p, const char* name, const char* args) { Session* s = static_cast<Session*>(p); return Guard([&] { s->layer->CallNamedFunc(name, args); }); }
int hl_init_path_tracing(void* p, int seed) { Session* s = static_cast<Session*>(p); return Guard([&] { s->layer->InitPathTracing(seed, nullptr); }); }
int hl_passes(void* p, int n) { Session* s = static_cast<Session*>(p); return Guard([&] { for (int i = 0; i < n; i++) { s->layer->BeginTracingPass(); s->layer->EndTracingPass(); } }); }
int hl_clear_accumulated(void* p) { Session* s = static_cast<Session*>(p); return Guard([&] { s->layer->ClearAccumulatedColor(); }); }
int hl_get_hdr(void* p, float* out, int w, int h) { Session* s = static_cast<Session*>(p); return Guard([&] { s->layer->GetHDRImage(reinterpret_cast<float4*>(out), w, h); }); }
int hl_get_ldr(void* p, uint32_t* out, int w, int h) { Session* s = static_cast<Session*>(p); return Guard([&] { s->layer->GetLDRImage(out, w, h); }); }
float hl_get_spp(void* p) { return static_cast<Session*>(p)->layer->GetSPP(); }
int hl_device_name(void* p, char* buf, int n) { Session* s = static_cast<Session*>(p); return Guard([&] { strncpy(buf, s->layer->GetDeviceName(nullptr), size_t(n) - 1); buf[n - 1] = 0; }); }
int hl_device_count(void* p) { int n = 0; for (const HRRenderDeviceInfoListElem* e = static_cast<Session*>(p)->layer->ListDevices(); e; e = e->next) n++; return n; }
int hl_rays_stat(void* p, float* raysPerSec, float* samplesPerSec, int* tracePerCent)
{
  Session* s = static_cast<Session*>(p);
  return Guard([&] { const MRaysStat st = s->layer->GetRaysStat(); *raysPerSec = st.raysPerSec; *samplesPerSec = st.samplesPerSec; *tracePerCent = st.traceTimePerCent; });
}
uint64_t hl_available_memory(void* p, int all) { return static_cast<Session*>(p)->layer->GetAvaliableMemoryAmount(all != 0); }
int hl_store_cpu_data(void* p) { return static_cast<Session*>(p)->layer->StoreCPUData() ? 1 : 0; }
}
