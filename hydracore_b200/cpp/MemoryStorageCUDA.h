// MemoryStorageCUDA — device-resident IMemoryStorage for the CUDA layer; sits where MemoryStorageOCL sits
// (reference hydra_drv/MemoryStorageOCL.h:7-37) and keeps its contract:
//   Reserve(total) -> capacity, or size_t(-1) when the device allocation fails (MemoryStorageOCL.cpp:10-29)
//   Resize(n)      -> n when it fits the reservation, else size_t(-1)                  (MemoryStorageOCL.cpp:31-40)
//   GetBegin()     -> nullptr: device-only                                              (MemoryStorageOCL.cpp:42-45)
//   MemCopyAt(off, data, bytes): blocking host->device write at a byte offset            (MemoryStorageOCL.cpp:57-60)
// Update / UpdatePartial / GetTable / AppendToTheEnd are inherited from IMemoryStorage (MemoryStorageCPU.cpp:53-127) unchanged.
// MemoryStorageBothCPUAndCUDA mirrors MemoryStorageBothCPUAndGPU (MemoryStorageOCL.h:40-71): "geom" and "textures" keep a host copy
// because RenderDriverRTE reads them back (InstanceMeshes / CreateAlphaTestTable, RenderDriverRTE.cpp:1976-1980).
#pragma once
#include "IMemoryStorage.h"
#include "MemoryStorageCPU.h"
#include "../../include/hydracore_cuda.h"

struct MemoryStorageCUDA : public IMemoryStorage
{
  MemoryStorageCUDA(hc_ctx* a_ctx, int a_slot) : m_ctx(a_ctx), m_slot(a_slot), m_currSize(0), m_totalSize(0) {}
  ~MemoryStorageCUDA() { Clear(); }

  void   Clear()                       override;
  size_t Reserve(uint64_t a_totalSize) override;
  size_t Resize(uint64_t a_totalSize)  override;

  const void*   GetBegin()    const override { return nullptr; }
  const size_t  GetSize()     const override { return size_t(m_currSize); }
  const size_t  GetCapacity() const override { return size_t(m_totalSize); }

  void MemCopyAt(uint64_t a_offsetInBytes, const void* a_data, uint64_t a_sizeInBytes) override;
  void DebugSaveToFile(const char* a_fileName) override;

  int  Slot() const { return m_slot; }

protected:
  hc_ctx*  m_ctx;
  int      m_slot;        // HC_STORAGE_*: which of the layer's five device blobs this object fronts
  uint64_t m_currSize;
  uint64_t m_totalSize;
};

struct MemoryStorageBothCPUAndCUDA : public IMemoryStorage
{
  MemoryStorageBothCPUAndCUDA(LinearStorageCPU* a_pStorageCPU, MemoryStorageCUDA* a_pStorageGPU) : m_pStorageCPU(a_pStorageCPU), m_pStorageGPU(a_pStorageGPU) {}
  ~MemoryStorageBothCPUAndCUDA() { delete m_pStorageCPU; m_pStorageCPU = nullptr; delete m_pStorageGPU; m_pStorageGPU = nullptr; }

  void   Clear()                       override { if (m_pStorageCPU) m_pStorageCPU->Clear(); if (m_pStorageGPU) m_pStorageGPU->Clear(); objects.clear(); maxId = 0; }
  size_t Reserve(uint64_t a_totalSize) override { if (m_pStorageCPU) m_pStorageCPU->Reserve(a_totalSize); return m_pStorageGPU ? m_pStorageGPU->Reserve(a_totalSize) : 0; }
  size_t Resize(uint64_t a_totalSize)  override { if (m_pStorageCPU) m_pStorageCPU->Resize(a_totalSize);  return m_pStorageGPU ? m_pStorageGPU->Resize(a_totalSize) : 0; }

  const void*   GetBegin()    const override { return m_pStorageCPU ? m_pStorageCPU->GetBegin() : nullptr; }
  const size_t  GetSize()     const override { return m_pStorageGPU ? m_pStorageGPU->GetSize() : 0; }
  const size_t  GetCapacity() const override { return m_pStorageGPU ? m_pStorageGPU->GetCapacity() : 0; }

  void MemCopyAt(uint64_t a_offsetInBytes, const void* a_data, uint64_t a_sizeInBytes) override
  {
    if (m_pStorageCPU) m_pStorageCPU->MemCopyAt(a_offsetInBytes, a_data, a_sizeInBytes);
    if (m_pStorageGPU) m_pStorageGPU->MemCopyAt(a_offsetInBytes, a_data, a_sizeInBytes);
  }
  void DebugSaveToFile(const char* a_fileName) override { if (m_pStorageGPU) m_pStorageGPU->DebugSaveToFile(a_fileName); }
  void FreeHostMem() override { delete m_pStorageCPU; m_pStorageCPU = nullptr; }     // RenderDriverRTE::FreeCPUMem, RenderDriverRTE.cpp:1552-1558

protected:
  LinearStorageCPU*  m_pStorageCPU;
  MemoryStorageCUDA* m_pStorageGPU;
};
