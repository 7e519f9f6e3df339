// hydracore-b200 — software texture fetch of the reference (cfetch.h:301-584), shared by the shading code (hc_shade.cuh) and by the
// alpha-tested closest-hit traversal (hc_trace.cuh).  Same arithmetic and operation order as the reference's host build.
#pragma once
#include "hc_math.cuh"
#include "hc_layout.h"

HC_DEV float sRGBToLinear(float x)                                      // cglobals.h:3024-3030
{
  if (x <= 0.0404482362771082f) return x*0.077399381f;
  return hc_pow((x + 0.055f)*0.947867299f, 2.4f);
}
HC_DEV float4 ReadUchar4(const uchar4* data, int offset)               // read_array_uchar4, cfetch.h:301-306
{
  const float mult = 0.003921568f;
  const uchar4 c = data[offset];
  return mult*make_float4((float)c.x, (float)c.y, (float)c.z, (float)c.w);
}
HC_DEV int4 BilinearOffsets(float ffx, float ffy, int flags, int w, int h)     // cfetch.h:315-368
{
  const int sx = (ffx > 0.0f) ? 1 : -1, sy = (ffy > 0.0f) ? 1 : -1;
  const int px = (int)ffx, py = (int)ffy;
  int px0, px1, py0, py1;
  if (flags & HC_TEX_CLAMP_U)
  {
    px0 = (px >= w) ? w - 1 : px;         px1 = (px + 1 >= w) ? w - 1 : px + 1;
    px0 = (px0 < 0) ? 0 : px0;            px1 = (px1 < 0) ? 0 : px1;
  }
  else
  {
    px0 = px % w;                         px1 = (px + sx) % w;
    px0 = (px0 < 0) ? px0 + w : px0;      px1 = (px1 < 0) ? px1 + w : px1;
  }
  if (flags & HC_TEX_CLAMP_V)
  {
    py0 = (py >= h) ? h - 1 : py;         py1 = (py + 1 >= h) ? h - 1 : py + 1;
    py0 = (py0 < 0) ? 0 : py0;            py1 = (py1 < 0) ? 0 : py1;
  }
  else
  {
    py0 = py % h;                         py1 = (py + sy) % h;
    py0 = (py0 < 0) ? py0 + h : py0;      py1 = (py1 < 0) ? py1 + h : py1;
  }
  return make_int4(py0*w + px0, py0*w + px1, py1*w + px0, py1*w + px1);
}

// read_imagef_sw1 (cfetch.h:364-459): single-channel images, float (bpp 4) or 8-bit sRGB (bpp 1).  Other formats are rejected when
// the scene is validated (ValidateScene, hc_path.cu), so `res` is always written.
HC_DEV float ReadImageSw1(const int4* tex, float2 tc, int flags)
{
  const int4 header = *tex;
  const int w = header.x, h = header.y, bpp = header.w;
  float ffx = tc.x*(float)w - 0.5f, ffy = tc.y*(float)h - 0.5f;
  if ((flags & HC_TEX_CLAMP_U) != 0 && ffx < 0) ffx = 0.0f;
  if ((flags & HC_TEX_CLAMP_V) != 0 && ffy < 0) ffy = 0.0f;
  const float* fdata = reinterpret_cast<const float*>(tex + 1);
  const unsigned char* bdata = reinterpret_cast<const unsigned char*>(tex + 1);
  const float mult = 0.003921568f;                                       // read_array_uchar, cfetch.h:305-310
  float res = 0.0f;
  if (flags & HC_TEX_POINT_SAM)
  {
    int px = (int)(ffx + 0.5f), py = (int)(ffy + 0.5f);
    if (flags & HC_TEX_CLAMP_U) { px = (px >= w) ? w - 1 : px; px = (px < 0) ? 0 : px; } else { px = px % w; px = (px < 0) ? px + w : px; }
    if (flags & HC_TEX_CLAMP_V) { py = (py >= h) ? h - 1 : py; py = (py < 0) ? 0 : py; } else { py = py % h; py = (py < 0) ? py + h : py; }
    const int offset = py*w + px;
    if (bpp == 4) res = fdata[offset];
    else if (bpp == 1) res = sRGBToLinear(mult*(float)bdata[offset]);
  }
  else
  {
    const int px = (int)ffx, py = (int)ffy;
    const float fx = fabsf(ffx - (float)px), fy = fabsf(ffy - (float)py);
    const float fx1 = 1.0f - fx, fy1 = 1.0f - fy;
    const float w1 = fx1*fy1, w2 = fx*fy1, w3 = fx1*fy, w4 = fx*fy;
    const int4 o = BilinearOffsets(ffx, ffy, flags, w, h);
    if (bpp == 4) res = fdata[o.x]*w1 + fdata[o.y]*w2 + fdata[o.z]*w3 + fdata[o.w]*w4;
    else if (bpp == 1)
      res = sRGBToLinear(mult*(float)bdata[o.x])*w1 + sRGBToLinear(mult*(float)bdata[o.y])*w2 + sRGBToLinear(mult*(float)bdata[o.z])*w3 + sRGBToLinear(mult*(float)bdata[o.w])*w4;
  }
  return res;
}

// read_imagef_sw4 (cfetch.h:461-584): RGBA8 (bpp 4) and float4 (bpp 16) images; single-channel images (header.z == 1, what
// RenderDriverRTE::UpdateImage writes for grey-scale textures) go through read_imagef_sw1 and come back as (v, v, v, 1)
HC_DEV float4 ReadImageSw4(const int4* tex, float2 tc, int flags, bool srgb)
{
  const int4 header = *tex;
  const int w = header.x, h = header.y, bpp = header.w;
  if (w <= 0 || h <= 0) return make_float4(1.0f, 1.0f, 1.0f, 1.0f);      // empty slot: nothing to read (the reference would divide by zero)
  if (header.z == 1) { const float v = ReadImageSw1(tex, tc, flags); return make_float4(v, v, v, 1.0f); }
  float ffx = tc.x*(float)w - 0.5f, ffy = tc.y*(float)h - 0.5f;
  if ((flags & HC_TEX_CLAMP_U) != 0 && ffx < 0) ffx = 0.0f;
  if ((flags & HC_TEX_CLAMP_V) != 0 && ffy < 0) ffy = 0.0f;
  float4 res = make_float4(0, 0, 0, 0);
  if (flags & HC_TEX_POINT_SAM)
  {
    int px = (int)(ffx + 0.5f), py = (int)(ffy + 0.5f);
    if (flags & HC_TEX_CLAMP_U) { px = (px >= w) ? w - 1 : px; px = (px < 0) ? 0 : px; } else { px = px % w; px = (px < 0) ? px + w : px; }
    if (flags & HC_TEX_CLAMP_V) { py = (py >= h) ? h - 1 : py; py = (py < 0) ? 0 : py; } else { py = py % h; py = (py < 0) ? py + h : py; }
    const int offset = py*w + px;
    if (bpp == 4)
    {
      res = ReadUchar4(reinterpret_cast<const uchar4*>(tex + 1), offset);
      if (srgb) res = make_float4(sRGBToLinear(res.x), sRGBToLinear(res.y), sRGBToLinear(res.z), sRGBToLinear(res.w));
    }
    else if (bpp == 16) res = reinterpret_cast<const float4*>(tex + 1)[offset];
  }
  else
  {
    const int px = (int)ffx, py = (int)ffy;
    const float fx = fabsf(ffx - (float)px), fy = fabsf(ffy - (float)py);
    const float fx1 = 1.0f - fx, fy1 = 1.0f - fy;
    const float w1 = fx1*fy1, w2 = fx*fy1, w3 = fx1*fy, w4 = fx*fy;
    const int4 o = BilinearOffsets(ffx, ffy, flags, w, h);
    float4 f1, f2, f3v, f4;
    if (bpp == 4)
    {
      const uchar4* d = reinterpret_cast<const uchar4*>(tex + 1);
      f1 = ReadUchar4(d, o.x); f2 = ReadUchar4(d, o.y); f3v = ReadUchar4(d, o.z); f4 = ReadUchar4(d, o.w);
      if (srgb)
      {
        f1 = make_float4(sRGBToLinear(f1.x), sRGBToLinear(f1.y), sRGBToLinear(f1.z), sRGBToLinear(f1.w));
        f2 = make_float4(sRGBToLinear(f2.x), sRGBToLinear(f2.y), sRGBToLinear(f2.z), sRGBToLinear(f2.w));
        f3v = make_float4(sRGBToLinear(f3v.x), sRGBToLinear(f3v.y), sRGBToLinear(f3v.z), sRGBToLinear(f3v.w));
        f4 = make_float4(sRGBToLinear(f4.x), sRGBToLinear(f4.y), sRGBToLinear(f4.z), sRGBToLinear(f4.w));
      }
    }
    else
    {
      const float4* d = reinterpret_cast<const float4*>(tex + 1);
      f1 = d[o.x]; f2 = d[o.y]; f3v = d[o.z]; f4 = d[o.w];
    }
    res = make_float4(f1.x*w1 + f2.x*w2 + f3v.x*w3 + f4.x*w4, f1.y*w1 + f2.y*w2 + f3v.y*w3 + f4.y*w4,
                      f1.z*w1 + f2.z*w2 + f3v.z*w3 + f4.z*w4, f1.w*w1 + f2.w*w2 + f3v.w*w3 + f4.w*w4);
  }
  return res;
}

