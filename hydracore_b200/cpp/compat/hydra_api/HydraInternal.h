// Stand-in for HydraAPI's HydraInternal.h: the shared accumulation image through which several render processes add their
// partial framebuffers (reference GPUOCLLayerOther.cpp:232-430, main.cpp:226-238).  Members used by IHWLayer implementations only.
#pragma once
#include <cstdint>

struct HRSharedBufferHeader
{
  int32_t width, height, depth, channels;
  float   spp;
  int32_t counterRcv, counterSnd;
  float   avgImageB;
  int32_t totalByteSize, messageSendOffset, messageRcvOffset, imageDataOffset;
};

struct IHRSharedAccumImage
{
  virtual ~IHRSharedAccumImage() {}
  virtual bool  Create(int w, int h, int d, const char* name, char errMsg[256]) = 0;
  virtual bool  Attach(const char* name, char errMsg[256]) = 0;
  virtual void  Clear() = 0;
  virtual bool  Lock(int a_miliseconds) = 0;
  virtual void  Unlock() = 0;
  virtual float* ImageData(int layerNum) = 0;
  virtual char*  MessageSendData() = 0;
  virtual char*  MessageRcvData() = 0;
  virtual HRSharedBufferHeader* Header() = 0;
};
