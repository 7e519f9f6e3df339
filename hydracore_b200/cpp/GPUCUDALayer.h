// GPUCUDALayer — the B200 (sm_100a) implementation of the reference's compute-layer interface IHWLayer
// (reference hydra_drv/IHWLayer.h:97-246).  It stands beside GPUOCLLayer (hydra_drv/GPUOCLLayer.h:28) and CPUExpLayer
// (hydra_drv/CPUExpLayer.cpp:9): RenderDriverRTE creates it through CreateCudaImpl() where it creates the other two
// (RenderDriverRTE.cpp:57-62, 85-88) and drives it with the same calls.  All device work goes through the C ABI of
// libhydracore_b200.so (include/hydracore_cuda.h); this class only keeps the host-side bookkeeping the base class expects
// (m_allMemStorages, m_cdataPrepared, m_vars) and turns a non-zero status into RUN_TIME_ERROR semantics (std::runtime_error,
// reference globals_sys.h:56-62).  There is no CPU fallback: without a CUDA device the constructor throws.
#pragma once
#include "IHWLayer.h"
#include "MemoryStorageCUDA.h"
#include <string>
#include <vector>

class GPUCUDALayer : public IHWLayer
{
public:
  typedef IHWLayer Base;

  GPUCUDALayer(int w, int h, int a_flags, int a_deviceId);
  ~GPUCUDALayer() override;

  void Clear(CLEAR_FLAGS a_flags) override;
  IMemoryStorage* CreateMemStorage(uint64_t a_maxSizeInBytes, const char* a_name) override;

  void PrepareEngineGlobals() override;
  void PrepareEngineTables()  override;

  void SetAllBVH4(const ConvertionResult& a_convertedBVH, IBVHBuilder2* a_inBuilderAPI, int a_flags) override;
  void SetAllInstMatrices(const float4x4* a_matrices, int32_t a_matrixNum) override;
  void SetAllInstLightInstId(const int32_t* a_lightInstIds, int32_t a_instNum) override;
  void SetAllRemapLists(const int* a_allLists, const int2* a_table, int a_allSize, int a_tableSize) override;
  void SetAllInstIdToRemapId(const int* a_allInstId, int a_instNum) override;
  void SetAllPODLights(PlainLight* a_lights2, size_t a_number) override;
  void SetAllFlagsAndVars(const AllRenderVarialbes& a_vars) override;

  void BeginTracingPass() override;
  void EndTracingPass()   override;
  void FinishAll()        override;

  void InitPathTracing(int seed, std::vector<int32_t>* pInstRemapTable = nullptr) override;
  void ClearAccumulatedColor() override;
  void ResetPerfCounters()     override;
  void ResizeScreen(int w, int h, int a_flags) override;

  void GetLDRImage(uint32_t* data, int width, int height) const override;
  void GetHDRImage(float4* data, int width, int height)   const override;

  size_t    GetAvaliableMemoryAmount(bool allMem = false) override;
  MRaysStat GetRaysStat() override;
  int32_t   GetRayBuffSize() const override { return m_width*m_height; }

  const char* GetDeviceName(int* pOCLVer = nullptr) const override;
  const HRRenderDeviceInfoListElem* ListDevices() const override;

  // custom pipeline hooks of the interface, used for what the reference API has no slot for:
  //   CallNamedFunc("integrator", "pt" | "mispt" | "qmc")        which CPUExpLayer integrator to reproduce (default: by flags)
  //   CallNamedFunc("tiles", "<tileSize> <rank> <worldSize>")    interleaved tile ownership of this process (one process per GPU)
  //   CallNamedFunc("raycast_light", "<x> <y> <z>")               ray-casting mode (passes without HRT_UNIFIED_IMAGE_SAMPLING): shadow rays go to this point
  //   CallNamedFunc("raycast_results", "<hitsAddr> <visAddr>")    host buffers the ray-casting pass fills (W*H Lite_Hit records, W*H bytes)
  //   CallNamedFunc("shadow_trees", "0" | "1")                    shadow rays through the alpha-tested tree (1, default: as GPUOCLLayer) or not (0: as CPUExpLayer)
  //   CallNamedFunc("sample_streams", "<S> [maxPathsInFlight]")  S generators per pixel: one BeginTracingPass then carries up to S passes (hc_pt_set_sample_streams)
  //   CallNamedFunc("comm_id", "") -> CommIdHex()                 rank 0: create the NCCL unique id of a multi-process render
  //   CallNamedFunc("comm", "<rank> <nranks> <256 hex digits>")   join the communicator
  //   CallNamedFunc("reduce", "<dstRank> <mode>")                 combine the framebuffers on dstRank (0 = tile partition, 1 = full-size sum)
  void CallNamedFunc(const char* a_name, const char* a_args) override;

  bool StoreCPUData() const override { return false; }          // no CPU integrator behind this layer (IHWLayerDataAssembler.cpp:574)

  void ContribToExternalImageAccumulator(IHRSharedAccumImage* a_pImage) override;

  float GetSPP() const override { return m_spp; }
  float GetSPPContrib() const override { return m_sppContributed; }

  hc_ctx* Context() const { return m_ctx; }
  const std::string& CommIdHex() const { return m_commIdHex; }

protected:
  void Check(int rc, const char* what) const;
  void UploadGlobalsIfDirty();
  int  IntegratorFromState() const;

  hc_ctx* m_ctx;
  int     m_initFlags;
  int     m_integratorOverride;      // -1: derive from m_vars
  bool    m_globalsDirty;
  bool    m_ptInitialised;
  int     m_seed;
  float   m_spp;
  float   m_sppContributed;
  MRaysStat m_stat;
  std::string m_commIdHex;
  float   m_rcLight[3] = { 0.0f, 0.0f, 0.0f };
  bool    m_rcLightSet = false;
  void*   m_rcHits = nullptr;
  void*   m_rcVis = nullptr;
  mutable std::string m_deviceName;
  mutable std::vector<HRRenderDeviceInfoListElem> m_deviceList;
};

IHWLayer* CreateCudaImpl(int w, int h, int a_flags, int a_deviceId);   // beside CreateOclImpl / CreateCPUExpImpl (IHWLayer.h:256-257)
