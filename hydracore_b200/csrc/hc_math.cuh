// hc_math.cuh — small vector algebra for the sm_100a kernels.
// The whole library is compiled with --fmad=false and without fast-math: the CPU oracle is built with plain SSE4.2
// (reference hydra_drv/CMakeLists.txt:69: no FMA contraction), so products and sums must round separately for hit ids
// and t to come out bit-identical.  Operation ORDER follows the reference's formulas (cited per function).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HC_DEV __device__ __forceinline__

HC_DEV float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
HC_DEV float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
HC_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
HC_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
HC_DEV float3 operator*(float3 a, float3 b) { return f3(a.x*b.x, a.y*b.y, a.z*b.z); }
HC_DEV float3 operator*(float3 a, float s)  { return f3(a.x*s, a.y*s, a.z*s); }
HC_DEV float3 operator*(float s, float3 a)  { return f3(a.x*s, a.y*s, a.z*s); }
HC_DEV float3 operator/(float3 a, float s)  { return f3(a.x/s, a.y/s, a.z/s); }
HC_DEV float3 operator-(float3 a)           { return f3(-a.x, -a.y, -a.z); }
HC_DEV float3& operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
HC_DEV float3& operator*=(float3& a, float3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
HC_DEV float3& operator*=(float3& a, float s)  { a.x *= s; a.y *= s; a.z *= s; return a; }
HC_DEV float2 f2(float x, float y) { return make_float2(x, y); }
HC_DEV float2 operator+(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
HC_DEV float2 operator-(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }
HC_DEV float2 operator*(float2 a, float s)  { return f2(a.x*s, a.y*s); }
HC_DEV float2 operator*(float s, float2 a)  { return f2(a.x*s, a.y*s); }
HC_DEV float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
HC_DEV float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
HC_DEV float4 operator*(float4 a, float s)  { return make_float4(a.x*s, a.y*s, a.z*s, a.w*s); }
HC_DEV float4 operator*(float s, float4 a)  { return make_float4(a.x*s, a.y*s, a.z*s, a.w*s); }

HC_DEV float  dot(float3 a, float3 b)   { return a.x*b.x + a.y*b.y + a.z*b.z; }
HC_DEV float3 cross(float3 a, float3 b) { return f3(a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x); }
HC_DEV float  length(float3 a)          { return sqrtf(a.x*a.x + a.y*a.y + a.z*a.z); }
HC_DEV float3 normalize(float3 a)       { return a/length(a); }     // LiteMath stand-in: v / |v| (oracle/ref_shim LiteMath.h)
HC_DEV float  maxcomp(float3 v)         { return fmaxf(v.x, fmaxf(v.y, v.z)); }
HC_DEV float  clampf(float u, float a, float b) { return fminf(fmaxf(a, u), b); }
HC_DEV float3 lerp3(float3 u, float3 v, float t) { return u + t*(v - u); }
HC_DEV float  lerpf(float u, float v, float t)   { return u + t*(v - u); }
HC_DEV float  sqrf(float x) { return x*x; }

// column-storage 4x4 (reference cglobals.h:206-209); mul4x3 / mul3x3 follow cglobals.h:306-322
struct HcMat4 { float4 c0, c1, c2, c3; };
HC_DEV float3 mul4x3(const HcMat4& m, float3 v)
{
  return f3(v.x*m.c0.x + v.y*m.c1.x + v.z*m.c2.x + m.c3.x,
            v.x*m.c0.y + v.y*m.c1.y + v.z*m.c2.y + m.c3.y,
            v.x*m.c0.z + v.y*m.c1.z + v.z*m.c2.z + m.c3.z);
}
HC_DEV float3 mul3x3(const HcMat4& m, float3 v)
{
  return f3(v.x*m.c0.x + v.y*m.c1.x + v.z*m.c2.x,
            v.x*m.c0.y + v.y*m.c1.y + v.z*m.c2.y,
            v.x*m.c0.z + v.y*m.c1.z + v.z*m.c2.z);
}
HC_DEV float4 mul4x4(const HcMat4& m, float4 v)      // mul4x4x4, cglobals.h:826-834
{
  return make_float4(v.x*m.c0.x + v.y*m.c1.x + v.z*m.c2.x + v.w*m.c3.x,
                     v.x*m.c0.y + v.y*m.c1.y + v.z*m.c2.y + v.w*m.c3.y,
                     v.x*m.c0.z + v.y*m.c1.z + v.z*m.c2.z + v.w*m.c3.z,
                     v.x*m.c0.w + v.y*m.c1.w + v.z*m.c2.w + v.w*m.c3.w);
}
HC_DEV HcMat4 transpose(const HcMat4& a)            // cglobals.h:1040-1048
{
  HcMat4 r;
  r.c0 = make_float4(a.c0.x, a.c1.x, a.c2.x, a.c3.x);
  r.c1 = make_float4(a.c0.y, a.c1.y, a.c2.y, a.c3.y);
  r.c2 = make_float4(a.c0.z, a.c1.z, a.c2.z, a.c3.z);
  r.c3 = make_float4(a.c0.w, a.c1.w, a.c2.w, a.c3.w);
  return r;
}
HC_DEV HcMat4 loadMat4(const float4* p) { HcMat4 m; m.c0 = p[0]; m.c1 = p[1]; m.c2 = p[2]; m.c3 = p[3]; return m; }
HC_DEV HcMat4 loadMat4(const float* p)  { return loadMat4(reinterpret_cast<const float4*>(p)); }

// cofactor inverse in the operation order of inverse4x4 (cglobals.h:917-1001)
HC_DEV HcMat4 inverse4x4(const HcMat4& a)
{
  const float4 c[4] = { a.c0, a.c1, a.c2, a.c3 };
  float t[12]; float4 m[4];
  t[0]=c[2].z*c[3].w; t[1]=c[3].z*c[2].w; t[2]=c[1].z*c[3].w; t[3]=c[3].z*c[1].w; t[4]=c[1].z*c[2].w;  t[5]=c[2].z*c[1].w;
  t[6]=c[0].z*c[3].w; t[7]=c[3].z*c[0].w; t[8]=c[0].z*c[2].w; t[9]=c[2].z*c[0].w; t[10]=c[0].z*c[1].w; t[11]=c[1].z*c[0].w;
  m[0].x  = t[0]*c[1].y + t[3]*c[2].y + t[4]*c[3].y;   m[0].x -= t[1]*c[1].y + t[2]*c[2].y + t[5]*c[3].y;
  m[0].y  = t[1]*c[0].y + t[6]*c[2].y + t[9]*c[3].y;   m[0].y -= t[0]*c[0].y + t[7]*c[2].y + t[8]*c[3].y;
  m[0].z  = t[2]*c[0].y + t[7]*c[1].y + t[10]*c[3].y;  m[0].z -= t[3]*c[0].y + t[6]*c[1].y + t[11]*c[3].y;
  m[0].w  = t[5]*c[0].y + t[8]*c[1].y + t[11]*c[2].y;  m[0].w -= t[4]*c[0].y + t[9]*c[1].y + t[10]*c[2].y;
  m[1].x  = t[1]*c[1].x + t[2]*c[2].x + t[5]*c[3].x;   m[1].x -= t[0]*c[1].x + t[3]*c[2].x + t[4]*c[3].x;
  m[1].y  = t[0]*c[0].x + t[7]*c[2].x + t[8]*c[3].x;   m[1].y -= t[1]*c[0].x + t[6]*c[2].x + t[9]*c[3].x;
  m[1].z  = t[3]*c[0].x + t[6]*c[1].x + t[11]*c[3].x;  m[1].z -= t[2]*c[0].x + t[7]*c[1].x + t[10]*c[3].x;
  m[1].w  = t[4]*c[0].x + t[9]*c[1].x + t[10]*c[2].x;  m[1].w -= t[5]*c[0].x + t[8]*c[1].x + t[11]*c[2].x;
  t[0]=c[2].x*c[3].y; t[1]=c[3].x*c[2].y; t[2]=c[1].x*c[3].y; t[3]=c[3].x*c[1].y; t[4]=c[1].x*c[2].y;  t[5]=c[2].x*c[1].y;
  t[6]=c[0].x*c[3].y; t[7]=c[3].x*c[0].y; t[8]=c[0].x*c[2].y; t[9]=c[2].x*c[0].y; t[10]=c[0].x*c[1].y; t[11]=c[1].x*c[0].y;
  m[2].x  = t[0]*c[1].w + t[3]*c[2].w + t[4]*c[3].w;   m[2].x -= t[1]*c[1].w + t[2]*c[2].w + t[5]*c[3].w;
  m[2].y  = t[1]*c[0].w + t[6]*c[2].w + t[9]*c[3].w;   m[2].y -= t[0]*c[0].w + t[7]*c[2].w + t[8]*c[3].w;
  m[2].z  = t[2]*c[0].w + t[7]*c[1].w + t[10]*c[3].w;  m[2].z -= t[3]*c[0].w + t[6]*c[1].w + t[11]*c[3].w;
  m[2].w  = t[5]*c[0].w + t[8]*c[1].w + t[11]*c[2].w;  m[2].w -= t[4]*c[0].w + t[9]*c[1].w + t[10]*c[2].w;
  m[3].x  = t[2]*c[2].z + t[5]*c[3].z + t[1]*c[1].z;   m[3].x -= t[4]*c[3].z + t[0]*c[1].z + t[3]*c[2].z;
  m[3].y  = t[8]*c[3].z + t[0]*c[0].z + t[7]*c[2].z;   m[3].y -= t[6]*c[2].z + t[9]*c[3].z + t[1]*c[0].z;
  m[3].z  = t[6]*c[1].z + t[11]*c[3].z + t[3]*c[0].z;  m[3].z -= t[10]*c[3].z + t[2]*c[0].z + t[7]*c[1].z;
  m[3].w  = t[10]*c[2].z + t[4]*c[0].z + t[9]*c[1].z;  m[3].w -= t[8]*c[1].z + t[11]*c[2].z + t[5]*c[0].z;
  const float k = 1.0f/(c[0].x*m[0].x + c[1].x*m[0].y + c[2].x*m[0].z + c[3].x*m[0].w);
  HcMat4 r; r.c0 = m[0]*k; r.c1 = m[1]*k; r.c2 = m[2]*k; r.c3 = m[3]*k;
  return r;
}

// Transcendentals.  The oracle build (oracle/_ref: reference headers compiled by g++ with only <cmath> in scope) resolves the
// reference's unqualified sin/cos/pow/exp/acos/atan2/tan calls on float arguments to the C double functions and rounds the
// result to float, i.e. it gets (almost always) the correctly rounded float.  B200 keeps a full-rate-ish FP64 pipe, and these
// calls are a handful per path vertex, so we evaluate them the same way: double in, double out, round once.  sqrt and
// division are IEEE-exact in float already (default -prec-sqrt/-prec-div), min/max/abs/floor are exact in either type.
// double-precision libm calls of the reference's host build, by name (kept inline: called out of line the shade kernel shrinks by another
// 20 % but C1 / C4 get 2-4 % slower; the texture fetch, hc_shade.cuh Sample2DFetch, is what pays to keep out of line)
HC_DEV double hc_d_sin(double x)             { return sin(x); }
HC_DEV double hc_d_cos(double x)             { return cos(x); }
HC_DEV double hc_d_tan(double x)             { return tan(x); }
HC_DEV double hc_d_exp(double x)             { return exp(x); }
HC_DEV double hc_d_log(double x)             { return log(x); }
HC_DEV double hc_d_acos(double x)            { return acos(x); }
HC_DEV double hc_d_atan2(double y, double x) { return atan2(y, x); }
HC_DEV double hc_d_pow(double x, double y)   { return pow(x, y); }
HC_DEV float hc_sin(float x)  { return (float)hc_d_sin((double)x); }
HC_DEV float hc_cos(float x)  { return (float)hc_d_cos((double)x); }
HC_DEV float hc_tan(float x)  { return (float)hc_d_tan((double)x); }
HC_DEV float hc_exp(float x)  { return (float)hc_d_exp((double)x); }
HC_DEV float hc_log(float x)  { return (float)hc_d_log((double)x); }
HC_DEV float hc_acos(float x) { return (float)hc_d_acos((double)x); }
HC_DEV float hc_asin(float x) { return (float)asin((double)x); }
HC_DEV float hc_atan(float x) { return (float)atan((double)x); }
HC_DEV float hc_atan2(float y, float x) { return (float)hc_d_atan2((double)y, (double)x); }
HC_DEV float hc_pow(float x, float y)   { return (float)hc_d_pow((double)x, (double)y); }
HC_DEV float hc_sqrt(float x) { return sqrtf(x); }
