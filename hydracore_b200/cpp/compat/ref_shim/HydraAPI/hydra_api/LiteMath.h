// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Minimal stand-in for HydraAPI's LiteMath.h, which
// is not vendored in the reference tree (reference CMakeLists.txt:11-15 expects ../HydraAPI).
// Written from scratch; conventions follow what the reference's own device headers require:
//   * float4x4 stores four COLUMN float4 in m_col[] (reference hydra_drv/cglobals.h:206-209, ctrace.h:1030-1033)
//   * float4x4(const float[16]) takes ROW-major input (scene XML matrices are row-major)
//   * mul4x3 / mul3x3 / mul / inverse4x4 / transpose / lookAt follow the OpenCL-branch formulas in
//     reference hydra_drv/cglobals.h:306-322, 828-847, 917-1049 (same operation order).
// "parity unpinned": the real LiteMath is unavailable, so rounding of normalize()/length() is ours.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

#ifndef SQR
#define SQR(x) ((x)*(x))
#endif
#ifndef MAXFLOAT
#define MAXFLOAT 1e37f
#endif

using std::isfinite;
using std::isnan;

namespace LiteMath
{
  typedef unsigned int   uint;
  typedef unsigned short ushort;
  typedef unsigned char  uchar;

  const float EPSILON    = 1e-6f;

  struct int2  { int2():x(0),y(0){} int2(int a,int b):x(a),y(b){} int x,y; };
  struct int3  { int3():x(0),y(0),z(0){} int3(int a,int b,int c):x(a),y(b),z(c){} int x,y,z; };
  struct int4  { int4():x(0),y(0),z(0),w(0){} int4(int a,int b,int c,int d):x(a),y(b),z(c),w(d){} int x,y,z,w; };
  struct uint2 { uint2():x(0),y(0){} uint2(uint a,uint b):x(a),y(b){} uint x,y; };
  struct uint3 { uint3():x(0),y(0),z(0){} uint3(uint a,uint b,uint c):x(a),y(b),z(c){} uint x,y,z; };
  struct uint4 { uint4():x(0),y(0),z(0),w(0){} uint4(uint a,uint b,uint c,uint d):x(a),y(b),z(c),w(d){} uint x,y,z,w; };
  struct uchar4  { uchar4():x(0),y(0),z(0),w(0){} uchar4(uchar a,uchar b,uchar c,uchar d):x(a),y(b),z(c),w(d){} uchar x,y,z,w; };
  struct ushort2 { ushort2():x(0),y(0){} ushort2(ushort a,ushort b):x(a),y(b){} ushort x,y; };
  struct ushort4 { ushort4():x(0),y(0),z(0),w(0){} ushort4(ushort a,ushort b,ushort c,ushort d):x(a),y(b),z(c),w(d){} ushort x,y,z,w; };

  struct float2 { float2():x(0),y(0){} float2(float a,float b):x(a),y(b){} float x,y; };
  struct float3 { float3():x(0),y(0),z(0){} float3(float a,float b,float c):x(a),y(b),z(c){} explicit float3(const float* p):x(p[0]),y(p[1]),z(p[2]){} float x,y,z; };
  struct float4 { float4():x(0),y(0),z(0),w(0){} float4(float a,float b,float c,float d):x(a),y(b),z(c),w(d){} explicit float4(const float* p):x(p[0]),y(p[1]),z(p[2]),w(p[3]){} float x,y,z,w; };
  struct double3 { double3():x(0),y(0),z(0){} double3(double a,double b,double c):x(a),y(b),z(c){} double x,y,z; };

  static inline float2 make_float2(float a,float b)                 { return float2(a,b); }
  static inline float3 make_float3(float a,float b,float c)         { return float3(a,b,c); }
  static inline float4 make_float4(float a,float b,float c,float d) { return float4(a,b,c,d); }
  static inline int3   make_int3(int a,int b,int c)                 { return int3(a,b,c); }
  static inline int4   make_int4(int a,int b,int c,int d)           { return int4(a,b,c,d); }
  static inline uint2  make_uint2(uint a,uint b)                    { return uint2(a,b); }
  static inline uint4  make_uint4(uint a,uint b,uint c,uint d)      { return uint4(a,b,c,d); }
  static inline uchar4 make_uchar4(uchar a,uchar b,uchar c,uchar d) { return uchar4(a,b,c,d); }
  static inline ushort2 make_ushort2(ushort a,ushort b)             { return ushort2(a,b); }
  static inline ushort4 make_ushort4(ushort a,ushort b,ushort c,ushort d) { return ushort4(a,b,c,d); }
  static inline double3 make_double3(double a,double b,double c)    { return double3(a,b,c); }

  static inline float2 to_float2(float4 v)          { return float2(v.x,v.y); }
  static inline float2 to_float2(float3 v)          { return float2(v.x,v.y); }
  static inline float3 to_float3(float4 v)          { return float3(v.x,v.y,v.z); }
  static inline float4 to_float4(float3 v,float w)  { return float4(v.x,v.y,v.z,w); }
  static inline double3 to_double3(float3 v)        { return double3((double)v.x,(double)v.y,(double)v.z); }
  static inline float3 to_float3(double3 v)         { return float3((float)v.x,(float)v.y,(float)v.z); }

  // ---- float2
  static inline float2 operator+(float2 a,float2 b){ return float2(a.x+b.x,a.y+b.y); }
  static inline float2 operator-(float2 a,float2 b){ return float2(a.x-b.x,a.y-b.y); }
  static inline float2 operator*(float2 a,float2 b){ return float2(a.x*b.x,a.y*b.y); }
  static inline float2 operator/(float2 a,float2 b){ return float2(a.x/b.x,a.y/b.y); }
  static inline float2 operator*(float2 a,float s) { return float2(a.x*s,a.y*s); }
  static inline float2 operator*(float s,float2 a) { return float2(a.x*s,a.y*s); }
  static inline float2 operator/(float2 a,float s) { return float2(a.x/s,a.y/s); }
  static inline float2 operator-(float2 a)         { return float2(-a.x,-a.y); }
  static inline float2& operator+=(float2& a,float2 b){ a.x+=b.x; a.y+=b.y; return a; }
  static inline float2& operator-=(float2& a,float2 b){ a.x-=b.x; a.y-=b.y; return a; }
  static inline float2& operator*=(float2& a,float s){ a.x*=s; a.y*=s; return a; }
  static inline float2& operator*=(float2& a,float2 b){ a.x*=b.x; a.y*=b.y; return a; }
  static inline float2& operator/=(float2& a,float s){ a.x/=s; a.y/=s; return a; }
  static inline float  dot(float2 a,float2 b)      { return a.x*b.x + a.y*b.y; }
  static inline float  length(float2 a)            { return sqrtf(a.x*a.x + a.y*a.y); }
  static inline float2 normalize(float2 a)         { return a/length(a); }

  // ---- float3
  static inline float3 operator+(float3 a,float3 b){ return float3(a.x+b.x,a.y+b.y,a.z+b.z); }
  static inline float3 operator-(float3 a,float3 b){ return float3(a.x-b.x,a.y-b.y,a.z-b.z); }
  static inline float3 operator*(float3 a,float3 b){ return float3(a.x*b.x,a.y*b.y,a.z*b.z); }
  static inline float3 operator/(float3 a,float3 b){ return float3(a.x/b.x,a.y/b.y,a.z/b.z); }
  static inline float3 operator*(float3 a,float s) { return float3(a.x*s,a.y*s,a.z*s); }
  static inline float3 operator*(float s,float3 a) { return float3(a.x*s,a.y*s,a.z*s); }
  static inline float3 operator/(float3 a,float s) { return float3(a.x/s,a.y/s,a.z/s); }
  static inline float3 operator/(float s,float3 a) { return float3(s/a.x,s/a.y,s/a.z); }
  static inline float3 operator+(float s,float3 a) { return float3(a.x+s,a.y+s,a.z+s); }
  static inline float3 operator-(float s,float3 a) { return float3(s-a.x,s-a.y,s-a.z); }
  static inline float3 operator+(float3 a,float s) { return float3(a.x+s,a.y+s,a.z+s); }
  static inline float3 operator-(float3 a,float s) { return float3(a.x-s,a.y-s,a.z-s); }
  static inline float3 operator-(float3 a)         { return float3(-a.x,-a.y,-a.z); }
  static inline float3& operator+=(float3& a,float3 b){ a.x+=b.x; a.y+=b.y; a.z+=b.z; return a; }
  static inline float3& operator-=(float3& a,float3 b){ a.x-=b.x; a.y-=b.y; a.z-=b.z; return a; }
  static inline float3& operator*=(float3& a,float3 b){ a.x*=b.x; a.y*=b.y; a.z*=b.z; return a; }
  static inline float3& operator*=(float3& a,float s){ a.x*=s; a.y*=s; a.z*=s; return a; }
  static inline float3& operator/=(float3& a,float s){ a.x/=s; a.y/=s; a.z/=s; return a; }
  static inline float  dot(float3 a,float3 b)      { return a.x*b.x + a.y*b.y + a.z*b.z; }
  static inline float3 cross(float3 a,float3 b)    { return float3(a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x); }
  static inline float  length(float3 a)            { return sqrtf(a.x*a.x + a.y*a.y + a.z*a.z); }
  static inline float3 normalize(float3 a)         { return a/length(a); }
  static inline float  maxcomp(float3 v)           { return fmaxf(v.x, fmaxf(v.y, v.z)); }
  static inline float  mincomp(float3 v)           { return fminf(v.x, fminf(v.y, v.z)); }

  // ---- float4
  static inline float4 operator+(float4 a,float4 b){ return float4(a.x+b.x,a.y+b.y,a.z+b.z,a.w+b.w); }
  static inline float4 operator-(float4 a,float4 b){ return float4(a.x-b.x,a.y-b.y,a.z-b.z,a.w-b.w); }
  static inline float4 operator*(float4 a,float4 b){ return float4(a.x*b.x,a.y*b.y,a.z*b.z,a.w*b.w); }
  static inline float4 operator/(float4 a,float4 b){ return float4(a.x/b.x,a.y/b.y,a.z/b.z,a.w/b.w); }
  static inline float4 operator*(float4 a,float s) { return float4(a.x*s,a.y*s,a.z*s,a.w*s); }
  static inline float4 operator*(float s,float4 a) { return float4(a.x*s,a.y*s,a.z*s,a.w*s); }
  static inline float4 operator/(float4 a,float s) { return float4(a.x/s,a.y/s,a.z/s,a.w/s); }
  static inline float4 operator-(float4 a)         { return float4(-a.x,-a.y,-a.z,-a.w); }
  static inline float4& operator+=(float4& a,float4 b){ a.x+=b.x; a.y+=b.y; a.z+=b.z; a.w+=b.w; return a; }
  static inline float4& operator-=(float4& a,float4 b){ a.x-=b.x; a.y-=b.y; a.z-=b.z; a.w-=b.w; return a; }
  static inline float4& operator*=(float4& a,float4 b){ a.x*=b.x; a.y*=b.y; a.z*=b.z; a.w*=b.w; return a; }
  static inline float4& operator*=(float4& a,float s){ a.x*=s; a.y*=s; a.z*=s; a.w*=s; return a; }
  static inline float4& operator/=(float4& a,float s){ a.x/=s; a.y/=s; a.z/=s; a.w/=s; return a; }
  static inline float  dot(float4 a,float4 b)      { return a.x*b.x + a.y*b.y + a.z*b.z + a.w*b.w; }
  static inline float  dot3(float4 a,float4 b)     { return a.x*b.x + a.y*b.y + a.z*b.z; }
  static inline float  length(float4 a)            { return sqrtf(a.x*a.x + a.y*a.y + a.z*a.z + a.w*a.w); }
  static inline float4 normalize(float4 a)         { return a/length(a); }

  static inline void store_u(float* p, float4 v) { p[0]=v.x; p[1]=v.y; p[2]=v.z; p[3]=v.w; }
  static inline void store(float* p, float4 v)   { p[0]=v.x; p[1]=v.y; p[2]=v.z; p[3]=v.w; }

  // ---- double3 (used by the DOUBLE_RAY_TRIANGLE variant, reference ctrace.h:186-313)
  static inline double3 operator+(double3 a,double3 b){ return double3(a.x+b.x,a.y+b.y,a.z+b.z); }
  static inline double3 operator-(double3 a,double3 b){ return double3(a.x-b.x,a.y-b.y,a.z-b.z); }
  static inline double3 operator*(double3 a,double s) { return double3(a.x*s,a.y*s,a.z*s); }
  static inline double  dot(double3 a,double3 b)      { return a.x*b.x + a.y*b.y + a.z*b.z; }
  static inline double3 cross(double3 a,double3 b)    { return double3(a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x); }

  // ---- scalar helpers
  static inline float clamp(float u,float a,float b) { return fminf(fmaxf(a,u),b); }
  static inline int   clamp(int u,int a,int b)       { return std::min(std::max(a,u),b); }
  static inline float lerp(float u,float v,float t)  { return u + t*(v-u); }
  static inline float3 lerp(float3 u,float3 v,float t){ return u + t*(v-u); }
  static inline float4 lerp(float4 u,float4 v,float t){ return u + t*(v-u); }
  static inline float3 clamp(float3 u,float a,float b){ return float3(clamp(u.x,a,b),clamp(u.y,a,b),clamp(u.z,a,b)); }
  static inline float4 clamp(float4 u,float a,float b){ return float4(clamp(u.x,a,b),clamp(u.y,a,b),clamp(u.z,a,b),clamp(u.w,a,b)); }
  static inline float  rnd(float s,float e)          { return s + (e-s)*(float(rand())/float(RAND_MAX)); }
  static inline float3 min(float3 a,float3 b)        { return float3(fminf(a.x,b.x),fminf(a.y,b.y),fminf(a.z,b.z)); }
  static inline float3 max(float3 a,float3 b)        { return float3(fmaxf(a.x,b.x),fmaxf(a.y,b.y),fmaxf(a.z,b.z)); }

  // ---- float4x4 : columns in m_col[]
  struct float4x4
  {
    float4x4() { identity(); }
    explicit float4x4(const float A[16])   // ROW-major input -> columns
    {
      m_col[0] = float4(A[0], A[4], A[8],  A[12]);
      m_col[1] = float4(A[1], A[5], A[9],  A[13]);
      m_col[2] = float4(A[2], A[6], A[10], A[14]);
      m_col[3] = float4(A[3], A[7], A[11], A[15]);
    }
    void identity()
    {
      m_col[0] = float4(1,0,0,0); m_col[1] = float4(0,1,0,0);
      m_col[2] = float4(0,0,1,0); m_col[3] = float4(0,0,0,1);
    }
    float4 get_col(int i) const { return m_col[i]; }
    void   set_col(int i, float4 c) { m_col[i] = c; }
    float4 get_row(int i) const
    {
      const float* c0=&m_col[0].x; const float* c1=&m_col[1].x; const float* c2=&m_col[2].x; const float* c3=&m_col[3].x;
      return float4(c0[i],c1[i],c2[i],c3[i]);
    }
    float4 m_col[4];
  };

  static inline float3 mul4x3(float4x4 m, float3 v)
  {
    float3 res;
    res.x = v.x * m.m_col[0].x + v.y * m.m_col[1].x + v.z * m.m_col[2].x + m.m_col[3].x;
    res.y = v.x * m.m_col[0].y + v.y * m.m_col[1].y + v.z * m.m_col[2].y + m.m_col[3].y;
    res.z = v.x * m.m_col[0].z + v.y * m.m_col[1].z + v.z * m.m_col[2].z + m.m_col[3].z;
    return res;
  }
  static inline float3 mul3x3(float4x4 m, float3 v)
  {
    float3 res;
    res.x = v.x * m.m_col[0].x + v.y * m.m_col[1].x + v.z * m.m_col[2].x;
    res.y = v.x * m.m_col[0].y + v.y * m.m_col[1].y + v.z * m.m_col[2].y;
    res.z = v.x * m.m_col[0].z + v.y * m.m_col[1].z + v.z * m.m_col[2].z;
    return res;
  }
  static inline float3 mul(float4x4 m, float3 v) { return mul4x3(m, v); }
  static inline float4 mul(float4x4 m, float4 v)
  {
    float4 res;
    res.x = v.x * m.m_col[0].x + v.y * m.m_col[1].x + v.z * m.m_col[2].x + v.w * m.m_col[3].x;
    res.y = v.x * m.m_col[0].y + v.y * m.m_col[1].y + v.z * m.m_col[2].y + v.w * m.m_col[3].y;
    res.z = v.x * m.m_col[0].z + v.y * m.m_col[1].z + v.z * m.m_col[2].z + v.w * m.m_col[3].z;
    res.w = v.x * m.m_col[0].w + v.y * m.m_col[1].w + v.z * m.m_col[2].w + v.w * m.m_col[3].w;
    return res;
  }
  static inline float4x4 mul(float4x4 a, float4x4 b)   // a*b (apply b first)
  {
    float4x4 r;
    for (int i = 0; i < 4; i++) r.m_col[i] = mul(a, b.m_col[i]);
    return r;
  }
  static inline float4 operator*(const float4x4& m, const float4& v) { return mul(m, v); }
  static inline float3 operator*(const float4x4& m, const float3& v) { return mul4x3(m, v); }
  static inline float4x4 operator*(const float4x4& a, const float4x4& b) { return mul(a, b); }

  static inline float4x4 transpose(const float4x4 a)
  {
    float4x4 r;
    r.m_col[0] = float4(a.m_col[0].x, a.m_col[1].x, a.m_col[2].x, a.m_col[3].x);
    r.m_col[1] = float4(a.m_col[0].y, a.m_col[1].y, a.m_col[2].y, a.m_col[3].y);
    r.m_col[2] = float4(a.m_col[0].z, a.m_col[1].z, a.m_col[2].z, a.m_col[3].z);
    r.m_col[3] = float4(a.m_col[0].w, a.m_col[1].w, a.m_col[2].w, a.m_col[3].w);
    return r;
  }

  // cofactor expansion in the same operation order as the reference's OpenCL branch (cglobals.h:917-1001)
  static inline float4x4 inverse4x4(float4x4 m1)
  {
    float t[12]; float4x4 m;
    const float4 *c = m1.m_col;
    t[0]=c[2].z*c[3].w; t[1]=c[3].z*c[2].w; t[2]=c[1].z*c[3].w; t[3]=c[3].z*c[1].w; t[4]=c[1].z*c[2].w;  t[5]=c[2].z*c[1].w;
    t[6]=c[0].z*c[3].w; t[7]=c[3].z*c[0].w; t[8]=c[0].z*c[2].w; t[9]=c[2].z*c[0].w; t[10]=c[0].z*c[1].w; t[11]=c[1].z*c[0].w;
    m.m_col[0].x  = t[0]*c[1].y + t[3]*c[2].y + t[4]*c[3].y;   m.m_col[0].x -= t[1]*c[1].y + t[2]*c[2].y + t[5]*c[3].y;
    m.m_col[0].y  = t[1]*c[0].y + t[6]*c[2].y + t[9]*c[3].y;   m.m_col[0].y -= t[0]*c[0].y + t[7]*c[2].y + t[8]*c[3].y;
    m.m_col[0].z  = t[2]*c[0].y + t[7]*c[1].y + t[10]*c[3].y;  m.m_col[0].z -= t[3]*c[0].y + t[6]*c[1].y + t[11]*c[3].y;
    m.m_col[0].w  = t[5]*c[0].y + t[8]*c[1].y + t[11]*c[2].y;  m.m_col[0].w -= t[4]*c[0].y + t[9]*c[1].y + t[10]*c[2].y;
    m.m_col[1].x  = t[1]*c[1].x + t[2]*c[2].x + t[5]*c[3].x;   m.m_col[1].x -= t[0]*c[1].x + t[3]*c[2].x + t[4]*c[3].x;
    m.m_col[1].y  = t[0]*c[0].x + t[7]*c[2].x + t[8]*c[3].x;   m.m_col[1].y -= t[1]*c[0].x + t[6]*c[2].x + t[9]*c[3].x;
    m.m_col[1].z  = t[3]*c[0].x + t[6]*c[1].x + t[11]*c[3].x;  m.m_col[1].z -= t[2]*c[0].x + t[7]*c[1].x + t[10]*c[3].x;
    m.m_col[1].w  = t[4]*c[0].x + t[9]*c[1].x + t[10]*c[2].x;  m.m_col[1].w -= t[5]*c[0].x + t[8]*c[1].x + t[11]*c[2].x;
    t[0]=c[2].x*c[3].y; t[1]=c[3].x*c[2].y; t[2]=c[1].x*c[3].y; t[3]=c[3].x*c[1].y; t[4]=c[1].x*c[2].y;  t[5]=c[2].x*c[1].y;
    t[6]=c[0].x*c[3].y; t[7]=c[3].x*c[0].y; t[8]=c[0].x*c[2].y; t[9]=c[2].x*c[0].y; t[10]=c[0].x*c[1].y; t[11]=c[1].x*c[0].y;
    m.m_col[2].x  = t[0]*c[1].w + t[3]*c[2].w + t[4]*c[3].w;   m.m_col[2].x -= t[1]*c[1].w + t[2]*c[2].w + t[5]*c[3].w;
    m.m_col[2].y  = t[1]*c[0].w + t[6]*c[2].w + t[9]*c[3].w;   m.m_col[2].y -= t[0]*c[0].w + t[7]*c[2].w + t[8]*c[3].w;
    m.m_col[2].z  = t[2]*c[0].w + t[7]*c[1].w + t[10]*c[3].w;  m.m_col[2].z -= t[3]*c[0].w + t[6]*c[1].w + t[11]*c[3].w;
    m.m_col[2].w  = t[5]*c[0].w + t[8]*c[1].w + t[11]*c[2].w;  m.m_col[2].w -= t[4]*c[0].w + t[9]*c[1].w + t[10]*c[2].w;
    m.m_col[3].x  = t[2]*c[2].z + t[5]*c[3].z + t[1]*c[1].z;   m.m_col[3].x -= t[4]*c[3].z + t[0]*c[1].z + t[3]*c[2].z;
    m.m_col[3].y  = t[8]*c[3].z + t[0]*c[0].z + t[7]*c[2].z;   m.m_col[3].y -= t[6]*c[2].z + t[9]*c[3].z + t[1]*c[0].z;
    m.m_col[3].z  = t[6]*c[1].z + t[11]*c[3].z + t[3]*c[0].z;  m.m_col[3].z -= t[10]*c[3].z + t[2]*c[0].z + t[7]*c[1].z;
    m.m_col[3].w  = t[10]*c[2].z + t[4]*c[0].z + t[9]*c[1].z;  m.m_col[3].w -= t[8]*c[1].z + t[11]*c[2].z + t[5]*c[0].z;
    const float k = 1.0f / (c[0].x*m.m_col[0].x + c[1].x*m.m_col[0].y + c[2].x*m.m_col[0].z + c[3].x*m.m_col[0].w);
    const float4 vK(k,k,k,k);
    for (int i = 0; i < 4; i++) m.m_col[i] = m.m_col[i]*vK;
    return m;
  }

  static inline float4x4 lookAt(float3 eye, float3 center, float3 up)
  {
    float3 z = normalize(eye - center);
    float3 y = up;
    float3 x = cross(y, z);
    y = cross(z, x);
    x = normalize(x); y = normalize(y);
    float4x4 M;
    M.m_col[0] = float4(x.x, y.x, z.x, 0.0f);
    M.m_col[1] = float4(x.y, y.y, z.y, 0.0f);
    M.m_col[2] = float4(x.z, y.z, z.z, 0.0f);
    M.m_col[3] = float4(-x.x*eye.x - x.y*eye.y - x.z*eye.z, -y.x*eye.x - y.y*eye.y - y.z*eye.z, -z.x*eye.x - z.y*eye.y - z.z*eye.z, 1.0f);
    return M;
  }

  // OpenGL-style frustum, column storage
  static inline float4x4 projectionMatrix(float fovyDeg, float aspect, float zNear, float zFar)
  {
    const float ymax = zNear * tanf(fovyDeg * 3.14159265358979323846f / 360.0f);
    const float xmax = ymax * aspect;
    const float l = -xmax, r = xmax, b = -ymax, t = ymax;
    float4x4 M;
    M.m_col[0] = float4(2.0f*zNear/(r-l), 0, 0, 0);
    M.m_col[1] = float4(0, 2.0f*zNear/(t-b), 0, 0);
    M.m_col[2] = float4((r+l)/(r-l), (t+b)/(t-b), -(zFar+zNear)/(zFar-zNear), -1.0f);
    M.m_col[3] = float4(0, 0, -2.0f*zFar*zNear/(zFar-zNear), 0);
    return M;
  }
  static inline float4x4 translate4x4(float3 t) { float4x4 m; m.m_col[3] = float4(t.x,t.y,t.z,1.0f); return m; }
  static inline float4x4 scale4x4(float3 s)     { float4x4 m; m.m_col[0].x=s.x; m.m_col[1].y=s.y; m.m_col[2].z=s.z; return m; }
}
