// hc_microfacet.cuh — anisotropic Beckmann and Trowbridge-Reitz microfacet lobes in the local (PBRT) frame: D, Lambda, G, BRDF, pdf and the
// visible-normal slope sampling.  Device side of the reference's PLAIN_MAT_CLASS_BECKMANN / PLAIN_MAT_CLASS_TRGGX materials; the
// formulas are those of hydra_drv/cmatpbrt.h:103-524 (PBRT-v3's microfacet.cpp restated there for OpenCL / C++).
//
// Parity with the reference's HOST build is the point of every odd-looking cast below: there the unqualified calls fmax / fmin / fabs /
// sqrt / exp / log / sin / cos / acos / pow on float arguments resolve to the DOUBLE functions and M_PI is <cmath>'s double (checked by
// compiling a probe with the oracle's flags: sizeof(fmax(a, b)) == 8, sizeof(M_PI) == 8, sizeof(M_TWOPI) == 4), so an expression is
// carried in double from the first such call to the next assignment to a float.  A single + - * / or sqrt done in double and rounded
// to float equals the float operation (innocuous double rounding), two or more do not - those are the places written in double here.
//
// Pure arithmetic, no memory access: the header also compiles as plain C++ (tests/host_microfacet.cpp checks it on the CPU against the
// reference's own functions, oracle/ref_driver.cpp ref_pbrt_*), where the only difference from the device build is libm vs libdevice
// in the double transcendentals.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HC_MF __host__ __device__ __forceinline__
#else
#define HC_MF inline
#endif

#define HC_MF_PI_D    3.14159265358979323846
#define HC_MF_TWOPI_F 6.28318530717958647692f

struct HcMf3 { float x, y, z; };
HC_MF HcMf3 mf3(float x, float y, float z) { HcMf3 r; r.x = x; r.y = y; r.z = z; return r; }
HC_MF float mfDot(HcMf3 a, HcMf3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
HC_MF HcMf3 mfNormalize(HcMf3 a) { const float len = sqrtf(a.x*a.x + a.y*a.y + a.z*a.z); return mf3(a.x/len, a.y/len, a.z/len); }   // v / |v| (LiteMath stand-in)
HC_MF float mfClamp(float u, float lo, float hi) { return fminf(fmaxf(lo, u), hi); }
HC_MF bool  mfFinite(float x) { return fabsf(x) <= 3.402823466e+38f; }     // false for inf and NaN

// ---- spherical helpers of the local frame (cmatpbrt.h:103-137); every helper returns a float, i.e. rounds its own result
HC_MF float mfCos2Theta(HcMf3 w) { return w.z*w.z; }
HC_MF float mfSin2Theta(HcMf3 w) { return fmaxf(0.0f, 1.0f - mfCos2Theta(w)); }
HC_MF float mfSinTheta(HcMf3 w)  { return sqrtf(mfSin2Theta(w)); }
HC_MF float mfTanTheta(HcMf3 w)  { return (fabsf(w.z) < 1e-6f) ? 0.0f : mfSinTheta(w)/w.z; }
HC_MF float mfTan2Theta(HcMf3 w) { return mfSin2Theta(w)/fmaxf(mfCos2Theta(w), 1e-6f); }
HC_MF float mfCosPhi(HcMf3 w) { const float s = mfSinTheta(w); return (s == 0.0f) ? 1.0f : mfClamp(w.x/s, -1.0f, 1.0f); }
HC_MF float mfSinPhi(HcMf3 w) { const float s = mfSinTheta(w); return (s == 0.0f) ? 0.0f : mfClamp(w.y/s, -1.0f, 1.0f); }
HC_MF float mfCos2Phi(HcMf3 w) { const float c = mfCosPhi(w); return c*c; }
HC_MF float mfSin2Phi(HcMf3 w) { const float s = mfSinPhi(w); return s*s; }

// ---- erf (Abramowitz-Stegun 7.1.26) and its inverse (cmatpbrt.h:139-193)
HC_MF float mfErf(float x)
{
  const float sign = (x < 0.0f) ? -1.0f : 1.0f;
  x = fabsf(x);
  const float t = 1.0f/(1.0f + 0.3275911f*x);
  const float poly = ((((1.061405429f*t + -1.453152027f)*t) + 1.421413741f)*t + -0.284496736f)*t + 0.254829592f;
  const float y = (float)(1.0 - (double)(poly*t)*exp((double)(-x*x)));                  // the product with exp() and the difference stay in double
  return sign*y;
}
HC_MF float mfErfInv(float x)
{
  x = mfClamp(x, -0.99999f, 0.99999f);
  float w = (float)(-log((double)((1.0f - x)*(1.0f + x))));
  float p;
  if (w < 5.0f)
  {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = 3.43273939e-07f + p*w;
    p = -3.5233877e-06f + p*w;
    p = -4.39150654e-06f + p*w;
    p = 0.00021858087f + p*w;
    p = -0.00125372503f + p*w;
    p = -0.00417768164f + p*w;
    p = 0.246640727f + p*w;
    p = 1.50140941f + p*w;
  }
  else
  {
    w = (float)(sqrt((double)w) - 3.0);                                                   // root and difference in double
    p = -0.000200214257f;
    p = 0.000100950558f + p*w;
    p = 0.00134934322f + p*w;
    p = -0.00367342844f + p*w;
    p = 0.00573950773f + p*w;
    p = -0.0076224613f + p*w;
    p = 0.00943887047f + p*w;
    p = 1.00167406f + p*w;
    p = 2.83297682f + p*w;
  }
  return p*x;
}

// roughness -> alpha fit shared by both lobes (cmatpbrt.h:334-338, 494-498)
HC_MF float mfRoughnessToAlpha(float roughness)
{
  const float x = (float)log((double)fmaxf(roughness, 1.0e-4f));
  return 1.62142f + 0.819955f*x + 0.1734f*x*x + 0.0171201f*x*x*x + 0.000640711f*x*x*x*x;
}

// KIND 0: Beckmann, 1: Trowbridge-Reitz (GGX)
template<int KIND>
HC_MF float mfD(HcMf3 wh, float ax, float ay)
{
  const float tan2 = mfTan2Theta(wh);
  const float cos4 = mfCos2Theta(wh)*mfCos2Theta(wh);
  if (KIND == 0)                                                                          // cmatpbrt.h:195-200
  {
    const double q = (double)mfCos2Phi(wh)/(double)fmaxf(ax*ax, 1e-6f) + (double)mfSin2Phi(wh)/(double)fmaxf(ay*ay, 1e-6f);
    const double num = exp((double)((-1.0f)*tan2)*q);
    const double den = fmax(HC_MF_PI_D*(double)ax*(double)ay*(double)cos4, (double)1e-6f);
    return (float)(num/den);
  }
  if (!mfFinite(tan2)) return 0.0f;                                                       // cmatpbrt.h:365-375
  const float e = (mfCos2Phi(wh)/(ax*ax) + mfSin2Phi(wh)/(ay*ay))*tan2;
  const float ope = 1.0f + e;
  return (float)((double)1.0f/(HC_MF_PI_D*(double)ax*(double)ay*(double)cos4*(double)ope*(double)ope));
}

template<int KIND>
HC_MF float mfLambda(HcMf3 w, float ax, float ay)
{
  const float absTan = fabsf(mfTanTheta(w));
  if (KIND == 0)                                                                          // cmatpbrt.h:202-218
  {
    if (!mfFinite(absTan) || absTan == 0.0f) return 0.0f;
    const float alpha = sqrtf(fmaxf(mfCos2Phi(w)*ax*ax + mfSin2Phi(w)*ay*ay, 1e-6f));
    const float a = 1.0f/fmaxf(alpha*absTan, 1e-6f);
    if (a >= 1.6f) return 0.0f;
    return (1.0f - 1.259f*a + 0.396f*a*a)/(3.535f*a + 2.181f*a*a);
  }
  if (!mfFinite(absTan)) return 0.0f;                                                     // cmatpbrt.h:377-388
  const float alpha = sqrtf(mfCos2Phi(w)*ax*ax + mfSin2Phi(w)*ay*ay);
  const float a2t2 = (alpha*absTan)*(alpha*absTan);
  return (float)((-1.0 + sqrt((double)(1.0f + a2t2)))/2.0);
}

template<int KIND> HC_MF float mfG1(HcMf3 w, float ax, float ay) { return 1.0f/(1.0f + mfLambda<KIND>(w, ax, ay)); }
template<int KIND> HC_MF float mfG(HcMf3 wo, HcMf3 wi, float ax, float ay) { return 1.0f/(1.0f + mfLambda<KIND>(wo, ax, ay) + mfLambda<KIND>(wi, ax, ay)); }

// pdf of the visible-normal sampling with respect to wh (cmatpbrt.h:329-332, 489-492)
template<int KIND>
HC_MF float mfPdf(HcMf3 wo, HcMf3 wh, float ax, float ay)
{
  return mfD<KIND>(wh, ax, ay)*mfG1<KIND>(wo, ax, ay)/fmaxf(4.0f*fabsf(wo.z), 1e-6f);
}

// reflection BRDF without Fresnel (the blend above the leaf carries it), cmatpbrt.h:345-363, 505-524
template<int KIND>
HC_MF float mfBrdf(HcMf3 wo, HcMf3 wi, float ax, float ay)
{
  const float cosO = fabsf(wo.z), cosI = fabsf(wi.z);
  HcMf3 wh = mf3(wi.x + wo.x, wi.y + wo.y, wi.z + wo.z);
  if (cosI <= 1e-6f || cosO <= 1e-6f) return 0.0f;
  if (fabsf(wh.x) <= 1e-6f && fabsf(wh.y) <= 1e-6f && fabsf(wh.z) <= 1e-6f) return 0.0f;
  wh = mfNormalize(wh);
  if (KIND == 0) return mfD<0>(wh, ax, ay)*mfG<0>(wo, wi, ax, ay)*1.0f/fmaxf(4.0f*cosI*cosO, 1e-6f);
  return mfD<1>(wh, ax, ay)*mfG<1>(wo, wi, ax, ay)/fmaxf(4.0f*cosI*cosO, 1e-6f);
}

// slopes of the visible normal for unit roughness and incidence cosine cosT (cmatpbrt.h:220-298, 391-448)
template<int KIND>
HC_MF void mfSample11(float cosT, float u1, float u2, float& slopeX, float& slopeY)
{
  if (KIND == 0)
  {
    if (cosT > 0.9999f)                                                                   // normal incidence
    {
      const float r = (float)sqrt(log((double)(1.0f - u1))*(double)(-1.0f));
      const float ang = HC_MF_TWOPI_F*u2;
      const float sinPhi = (float)sin((double)ang), cosPhi = (float)cos((double)ang);
      slopeX = r*cosPhi; slopeY = r*sinPhi;
      return;
    }
    // numerical inversion of the slope CDF in the erf domain: Newton steps kept inside a shrinking bracket [lo, hi]
    const float sinT = sqrtf(fmaxf(0.0f, 1.0f - cosT*cosT));
    const float tanT = sinT/fmaxf(cosT, 1e-6f);
    const float cotT = 1.0f/fmaxf(tanT, 1e-6f);
    float lo = -1.0f, hi = mfErf(cotT);
    const float sx = fmaxf(u1, 1e-6f);
    const float theta = (float)acos((double)cosT);
    const float fit = 1.0f + theta*(-0.876f + theta*(0.4265f - 0.0594f*theta));            // initial guess
    float b = (float)((double)hi - (double)(1.0f + hi)*pow((double)(1.0f - sx), (double)fit));
    const float invSqrtPi = (float)((double)1.0f/sqrt(HC_MF_PI_D));
    const float norm = (float)((double)1.0f/fmax((double)(1.0f + hi) + (double)(invSqrtPi*tanT)*exp((double)((-1.0f)*cotT*cotT)), (double)1e-6f));
    for (int it = 1; it < 10; it++)
    {
      if (!(b >= lo && b <= hi)) b = 0.5f*(lo + hi);
      const float ie = mfErfInv(b);
      const float value = (float)((double)norm*((double)(1.0f + b) + (double)(invSqrtPi*tanT)*exp((double)((-1.0f)*ie*ie))) - (double)sx);
      const float deriv = norm*(1.0f - ie*tanT);
      if (fabsf(value) < 1e-5f) break;
      if (value > 0.0f) hi = b; else lo = b;
      b = (float)((double)b - (double)value/(double)fmaxf(deriv, 1e-6f));
    }
    slopeX = mfErfInv(b);
    slopeY = mfErfInv((float)(2.0*(double)fmaxf(u2, 1e-6f) - 1.0));
    return;
  }
  if (cosT > 0.9999f)
  {
    const float r = (float)sqrt((double)u1/(double)fmaxf(1.0f - u1, 1e-6f));
    const float phi = HC_MF_TWOPI_F*u2;
    slopeX = (float)((double)r*cos((double)phi));
    slopeY = (float)((double)r*sin((double)phi));
    return;
  }
  const float sinT = sqrtf(fmaxf(0.0f, 1.0f - cosT*cosT));
  const float tanT = sinT/cosT;
  const float a = 1.0f/tanT;
  const float g1 = (float)((double)2.0f/((double)1.0f + sqrt((double)(1.0f + 1.0f/(a*a)))));
  const float A = 2.0f*u1/g1 - 1.0f;
  float tmp = 1.0f/(A*A - 1.0f);
  if (tmp > 1e10f) tmp = 1e10f;
  const float B = tanT;
  const float Dq = sqrtf(fmaxf(B*B*tmp*tmp - (A*A - B*B)*tmp, 0.0f));
  const float s1 = B*tmp - Dq, s2 = B*tmp + Dq;
  slopeX = (A < 0.0f || s2 > 1.0f/tanT) ? s1 : s2;
  float S;
  if (u2 > 0.5f) { S = 1.0f;  u2 = 2.0f*(u2 - 0.5f); }
  else           { S = -1.0f; u2 = 2.0f*(0.5f - u2); }
  const float z = (u2*(u2*(u2*0.27385f - 0.73369f) + 0.46341f))/(u2*(u2*(u2*0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
  slopeY = (float)((double)(S*z)*sqrt((double)(1.0f + slopeX*slopeX)));
}

// half vector seen from wo: stretch, sample unit-roughness slopes, rotate, unstretch (cmatpbrt.h:300-327, 450-482)
template<int KIND>
HC_MF HcMf3 mfSampleWh(HcMf3 wo, float u1, float u2, float ax, float ay)
{
  const bool flip = (wo.z < 0.0f);
  const HcMf3 wi = flip ? mf3(-1.0f*wo.x, -1.0f*wo.y, -1.0f*wo.z) : wo;
  const HcMf3 ws = mfNormalize(mf3(ax*wi.x, ay*wi.y, wi.z));
  float sx, sy;
  mfSample11<KIND>(ws.z, u1, u2, sx, sy);
  const float cp = mfCosPhi(ws), sp = mfSinPhi(ws);
  const float rx = cp*sx - sp*sy;
  sy = sp*sx + cp*sy;
  sx = rx;
  sx = ax*sx;
  sy = ay*sy;
  HcMf3 wh = mfNormalize(mf3(sx*(-1.0f), sy*(-1.0f), 1.0f));
  if (flip) wh = mf3(wh.x*(-1.0f), wh.y*(-1.0f), wh.z*(-1.0f));
  return wh;
}
