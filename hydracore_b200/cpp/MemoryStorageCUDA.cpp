#include "MemoryStorageCUDA.h"
#include <fstream>
#include <stdexcept>
#include <string>

void MemoryStorageCUDA::Clear()
{
  m_currSize = 0;                 // like MemoryStorageOCL::Clear the device buffer is kept; the id -> chunk table starts over
  objects.clear();
  maxId = 0;
}

size_t MemoryStorageCUDA::Reserve(uint64_t a_totalSize)
{
  if (m_totalSize == a_totalSize && a_totalSize != 0)
    return size_t(m_totalSize);
  if (hc_storage_reserve(m_ctx, m_slot, a_totalSize) != HC_OK)
    return size_t(-1);
  m_totalSize = a_totalSize;
  m_currSize  = 0;
  return size_t(m_totalSize);
}

size_t MemoryStorageCUDA::Resize(uint64_t a_size)
{
  if (a_size <= m_totalSize) { m_currSize = a_size; return size_t(m_currSize); }
  return size_t(-1);
}

void MemoryStorageCUDA::MemCopyAt(uint64_t a_offsetInBytes, const void* a_data, uint64_t a_sizeInBytes)
{
  const int rc = hc_storage_write(m_ctx, m_slot, a_offsetInBytes, a_data, a_sizeInBytes);
  if (rc != HC_OK)
    throw std::runtime_error(std::string("MemoryStorageCUDA::MemCopyAt: ") + hc_last_error());
}

void MemoryStorageCUDA::DebugSaveToFile(const char* a_fileName)
{
  std::ofstream fout(a_fileName);
  fout << "MemoryStorageCUDA slot " << m_slot << ": " << m_currSize << " of " << m_totalSize << " bytes used (device resident)" << std::endl;
}
