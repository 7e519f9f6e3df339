// hc_raygen.cuh — K1: eye-ray generation.
// Replaces MakeEyeRaysUnifiedSampling / MakeEyeRaysSamplesOnly (reference hydra_drv/shaders/screen.cl:280, 236) and restates
// MakeRandEyeRay / MakeEyeRayFromF4Rnd (hydra_drv/cfetch.h:877-968) with EyeRayDirNormalized, matrix4x4f_mult_ray3
// (cglobals.h:1069-1089), tiltCorrection (cfetch.h:832-863) and MapSamplesToDisc (cglobals.h:1609-1655).
#pragma once
#include "hc_math.cuh"
#include "hc_layout.h"
#include <cstring>
#include <cmath>

#define HC_MAXFLOAT_RAY 3.402823466e+38f   // MAXFLOAT (FLT_MAX)

struct HcCamera
{
  HcMat4 projInv, worldViewInv;       // EngineGlobals::mProjInverse / mWorldViewInverse, four columns each (make_float4x4, cglobals.h:790-798)
  float  sinHalfFov;                  // sin(0.5f*varsF[HRT_CAM_FOV]) evaluated on the host exactly like the oracle does (double sin, rounded)
  float  tiltX, tiltY;
  float  lensRadius, focalDist;
  float  fwidth, fheight;
  int    enableDof;
};

static inline HcCamera hc_camera_from_globals(const unsigned char* head)
{
  HcCamera c;
  std::memcpy(&c.projInv, head + HC_EG_mProjInverse, 64);
  std::memcpy(&c.worldViewInv, head + HC_EG_mWorldViewInverse, 64);
  const float* varsF = (const float*)(head + HC_EG_varsF);
  const int*   varsI = (const int*)(head + HC_EG_varsI);
  c.sinHalfFov = (float)std::sin((double)(0.5f*varsF[HC_HRT_CAM_FOV]));
  c.tiltX = varsF[HC_HRT_TILT_ROT_X]; c.tiltY = varsF[HC_HRT_TILT_ROT_Y];
  c.lensRadius = varsF[HC_HRT_DOF_LENS_RADIUS]; c.focalDist = varsF[HC_HRT_DOF_FOCAL_PLANE_DIST];
  c.fwidth = varsF[HC_HRT_WIDTH_F]; c.fheight = varsF[HC_HRT_HEIGHT_F];
  c.enableDof = (varsI[HC_HRT_ENABLE_DOF] == 1) ? 1 : 0;
  return c;
}

HC_DEV float3 EyeRayDirNormalized(float x, float y, const HcMat4& projInv)     // cglobals.h:1069-1078
{
  float4 pos = make_float4(2.0f*x - 1.0f, 2.0f*y - 1.0f, 0.0f, 1.0f);
  pos = mul4x4(projInv, pos);
  return normalize(f3(pos.x/pos.w, pos.y/pos.w, pos.z/pos.w));
}

HC_DEV float2 MapSamplesToDisc(float2 xy)                                      // cglobals.h:1609-1655
{
  const float x = xy.x, y = xy.y;
  float r = 0.0f, phi = 0.0f;
  if (x > y && x > -y)  { r = x;  phi = 0.25f*3.141592654f*(y/x); }
  if (x < y && x > -y)  { r = y;  phi = 0.25f*3.141592654f*(2.0f - x/y); }
  if (x < y && x < -y)  { r = -x; phi = 0.25f*3.141592654f*(4.0f + y/x); }
  if (x > y && x < -y)  { r = -y; phi = 0.25f*3.141592654f*(6 - x/y); }
  return f2(r*hc_sin(phi), r*hc_cos(phi));
}

HC_DEV float3 TiltCorrection(float3 pos, float3 dir, const HcCamera& cam)      // cfetch.h:832-863
{
  const float tiltX = cam.tiltX, tiltY = cam.tiltY;
  if ((fabsf(tiltX) > 0.0f || fabsf(tiltY) > 0.0f) && fabsf(dir.z) > 0.0f)
  {
    const float t = (-1.0f - pos.z)/dir.z;
    float3 p = pos + t*dir;
    p.z += 1.0f;
    if (fabsf(tiltY) > 0.0f)        // make_matrix_rotationY(-tiltY), cglobals.h:813-824
    {
      const float s = hc_sin(-tiltY), c = hc_cos(-tiltY);
      p = f3(p.x*c + p.y*0.0f + p.z*s + 0.0f, p.x*0.0f + p.y*1.0f + p.z*0.0f + 0.0f, p.x*(-s) + p.y*0.0f + p.z*c + 0.0f);
    }
    if (fabsf(tiltX) > 0.0f)        // make_matrix_rotationX(-tiltX), cglobals.h:800-811
    {
      const float s = hc_sin(-tiltX), c = hc_cos(-tiltX);
      p = f3(p.x*1.0f + p.y*0.0f + p.z*0.0f + 0.0f, p.x*0.0f + p.y*c + p.z*(-s) + 0.0f, p.x*0.0f + p.y*s + p.z*c + 0.0f);
    }
    p.z -= 1.0f;
    dir = normalize(p - pos);
  }
  return dir;
}

HC_DEV void MultRay3(const HcMat4& m, float3& pos, float3& dir)                // matrix4x4f_mult_ray3, cglobals.h:1080-1087
{
  const float3 p  = mul4x3(m, pos);
  const float3 p2 = mul4x3(m, pos + 100.0f*dir);
  pos = p; dir = normalize(p2 - p);
}

// MakeRandEyeRay (cfetch.h:877-931): pixel (x, y), offsets in [-1, 1]^4
HC_DEV void MakeRandEyeRay(int x, int y, int w, int h, float4 offsets, const HcCamera& cam, float3& rpos, float3& rdir)
{
  float3 pos = f3(0.0f, 0.0f, 0.0f);
  float3 dir = EyeRayDirNormalized(((float)x + 0.5f)/(float)w, ((float)y + 0.5f)/(float)h, cam.projInv);
  const float pxSizeX = cam.sinHalfFov*(1.0f/(float)w);
  const float pxSizeY = cam.sinHalfFov*(1.0f/(float)h);
  dir.x += pxSizeX*offsets.x;
  dir.y += pxSizeY*offsets.y;
  dir.z = -sqrtf(1.0f - (dir.x*dir.x + dir.y*dir.y));
  dir = TiltCorrection(pos, dir, cam);
  if (cam.enableDof)
  {
    const float  tFocus = cam.focalDist/(-dir.z);
    const float3 focus  = pos + dir*tFocus;
    const float2 xy     = cam.lensRadius*MapSamplesToDisc(1.0f*f2(offsets.z, offsets.w));
    pos.x += xy.x; pos.y += xy.y;
    dir = normalize(focus - pos);
  }
  MultRay3(cam.worldViewInv, pos, dir);
  rpos = pos; rdir = dir;
}

// MakeEyeRayFromF4Rnd (cfetch.h:933-968): lens sample in [0, 1]^4 -> ray + continuous pixel coordinates
HC_DEV void MakeEyeRayFromF4Rnd(float4 lens, const HcCamera& cam, float3& rpos, float3& rdir, float& fx, float& fy)
{
  const float x = cam.fwidth*lens.x, y = cam.fheight*lens.y;
  float3 pos = f3(0.0f, 0.0f, 0.0f);
  float3 dir = EyeRayDirNormalized(x/cam.fwidth, y/cam.fheight, cam.projInv);
  dir = TiltCorrection(pos, dir, cam);
  if (cam.enableDof)
  {
    const float  tFocus = cam.focalDist/(-dir.z);
    const float3 focus  = pos + dir*tFocus;
    const float2 xy     = cam.lensRadius*2.0f*MapSamplesToDisc(f2(lens.z - 0.5f, lens.w - 0.5f));
    pos.x += xy.x; pos.y += xy.y;
    dir = normalize(focus - pos);
  }
  MultRay3(cam.worldViewInv, pos, dir);
  fx = lens.x*cam.fwidth; fy = lens.y*cam.fheight;
  rpos = pos; rdir = dir;
}

// stand-alone K1 for the ray-casting entry point: one thread per pixel, row-major, 32-byte {pos, dir} records
static __global__ void k_make_eye_rays(const HcCamera cam, const int w, const int h, const float4* __restrict__ offsets, float4* __restrict__ raysOut)
{
  const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= (long long)w*h) return;
  const int x = int(i % w), y = int(i / w);
  const float4 offs = offsets ? offsets[i] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  float3 p, d;
  MakeRandEyeRay(x, y, w, h, offs, cam, p, d);
  raysOut[2*i + 0] = make_float4(p.x, p.y, p.z, 0.0f);
  raysOut[2*i + 1] = make_float4(d.x, d.y, d.z, HC_MAXFLOAT_RAY);
}
