// Stand-in for HydraAPI's public header (HydraAPI is a sibling checkout of HydraCore and is not vendored in the reference tree:
// reference CMakeLists.txt:11-15).  ONLY the types that cross the IHWLayer boundary are declared, with the members the layer uses
// (SURVEY.md 8b).  A real build drops this directory from the include path and picks up the real HydraAPI headers.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include "LiteMath.h"


struct HRRenderDeviceInfoListElem      // HydraAPI.h: list node returned by IHRRenderDriver::DeviceList()
{
  int32_t  id;
  wchar_t  name[256];
  wchar_t  driver[256];
  bool     isCPU;
  bool     isEnabled;
  const HRRenderDeviceInfoListElem* next;
};

namespace pugi
{
  struct xml_node                      // opaque handle: the CUDA layer stores it (SetCamNode / SetSettingsNode) and never reads it
  {
    void* impl = nullptr;
    explicit operator bool() const { return impl != nullptr; }
  };
}
