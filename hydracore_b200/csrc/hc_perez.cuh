// hc_perez.cuh — Perez all-weather sky luminance (Preetham's fit) with a sun disk, as the reference's sky-dome light evaluates it when
// SKY_LIGHT_USE_PEREZ_ENVIRONMENT is set: skyLightPerezColor, hydra_drv/clight.h:178-283.  Pure arithmetic, host + device like
// hc_microfacet.cuh (same rules: the reference's host build evaluates tan / acos / exp / pow / fmax on floats in DOUBLE, and an expression stays
// in double until it is assigned to a float); tests/host_microfacet.cpp + tests/test_microfacet.py compare it with the reference on the CPU.
#pragma once
#include "hc_microfacet.cuh"

struct HcPerezYxy { float Y, x, y; };

// zenith luminance and chromaticity for turbidity t and sun zenith angle thetaSun (clight.h:178-198)
HC_MF HcPerezYxy mfPerezZenith(float t, float thetaSun)
{
  const float pi = 3.1415926f;
  const float t2 = t*t;
  const float chi = (4.0f/9.0f - t/120.0f)*(pi - 2.0f*thetaSun);
  const float th1 = thetaSun, th2 = thetaSun*thetaSun, th3 = thetaSun*thetaSun*thetaSun;
  // dot(float4 coefficients, (1, th, th^2, th^3)) in the order x + y + z + w
  const float dx1 = 0.0f*1.0f + 0.00209f*th1 + -0.00375f*th2 + 0.00165f*th3;
  const float dx2 = 0.00394f*1.0f + -0.03202f*th1 + 0.06377f*th2 + -0.02903f*th3;
  const float dx3 = 0.25886f*1.0f + 0.06052f*th1 + -0.21196f*th2 + 0.11693f*th3;
  const float dy1 = 0.0f*1.0f + 0.00317f*th1 + -0.00610f*th2 + 0.00275f*th3;
  const float dy2 = 0.00516f*1.0f + -0.04153f*th1 + 0.08970f*th2 + -0.04214f*th3;
  const float dy3 = 0.26688f*1.0f + 0.06670f*th1 + -0.26756f*th2 + 0.15346f*th3;
  HcPerezYxy r;
  r.Y = (float)((double)(4.0453f*t - 4.9710f)*tan((double)chi) - (double)(0.2155f*t) + (double)2.4192f);
  r.x = t2*dx1 + t*dx2 + dx3;
  r.y = t2*dy1 + t*dy2 + dy3;
  return r;
}

// the five-coefficient distribution for (Y, x, y) (clight.h:204-229)
HC_MF HcPerezYxy mfPerezFunc(float t, float cosTheta, float cosGamma)
{
  const float gamma = (float)acos((double)cosGamma);
  const float cg2 = cosGamma*cosGamma;
  const float aY = 0.17872f*t - 1.46303f, bY = -0.35540f*t + 0.42749f, cY = -0.02266f*t + 5.32505f, dY = 0.12064f*t - 2.57705f, eY = -0.06696f*t + 0.37027f;
  const float ax = -0.01925f*t - 0.25922f, bx = -0.06651f*t + 0.00081f, cx = -0.00041f*t + 0.21247f, dx = -0.06409f*t - 0.89887f, ex = -0.00325f*t + 0.04517f;
  const float ay = -0.01669f*t - 0.26078f, by = -0.09495f*t + 0.00921f, cy = -0.00792f*t + 0.21023f, dy = -0.04405f*t - 1.65369f, ey = -0.01092f*t + 0.05291f;
  HcPerezYxy r;
  r.Y = (float)(((double)1.0f + (double)aY*exp((double)(bY/cosTheta)))*((double)1.0f + (double)cY*exp((double)(dY*gamma)) + (double)(eY*cg2)));
  r.x = (float)(((double)1.0f + (double)ax*exp((double)(bx/cosTheta)))*((double)1.0f + (double)cx*exp((double)(dx*gamma)) + (double)(ex*cg2)));
  r.y = (float)(((double)1.0f + (double)ay*exp((double)(by/cosTheta)))*((double)1.0f + (double)cy*exp((double)(dy*gamma)) + (double)(ey*cg2)));
  return r;
}

// sky colour seen along rayDir (pointing away from the viewer) for a sun shining along sunDir (pointing away from the sun): Yxy -> XYZ -> RGB,
// display gamma undone, sun disk blended in over the last 0.0005-0.0015 of the cosine (clight.h:231-283)
HC_MF HcMf3 mfPerezSkyColor(HcMf3 sunDir, float turbidity, HcMf3 sunColorIn, HcMf3 rayDir)
{
  const float cosTheta = fmaxf(rayDir.y, 0.0f) + 0.05f;
  const HcMf3 minusRay = mf3((-1.0f)*rayDir.x, (-1.0f)*rayDir.y, (-1.0f)*rayDir.z);
  const float cosGamma = fmaxf(mfDot(sunDir, minusRay), 0.0f);
  const float cosThetaSun = fmaxf(-sunDir.y, 0.0f);
  const HcPerezYxy z = mfPerezZenith(turbidity, (float)acos((double)cosThetaSun));
  const HcPerezYxy f = mfPerezFunc(turbidity, cosTheta, cosGamma), f0 = mfPerezFunc(turbidity, 1.0f, cosThetaSun);
  float Y = (z.Y*f.Y)/f0.Y;
  const float cx = (z.x*f.x)/f0.x, cy = (z.y*f.y)/f0.y;
  Y = (float)((double)1.0f - exp((double)(-Y/20.0f)));
  const float ratio = Y/fmaxf(cy, 1e-10f);
  const float X = cx*ratio, Yv = Y, Z = ratio - X - Yv;
  HcMf3 rgb = mf3(mfClamp(3.240479f*X + -1.53715f*Yv + -0.49853f*Z, 0.0f, 1.0f), mfClamp(-0.969256f*X + 1.875991f*Yv + 0.041556f*Z, 0.0f, 1.0f),
                  mfClamp(0.055684f*X + -0.204043f*Yv + 1.057311f*Z, 0.0f, 1.0f));
  rgb.x = (float)pow((double)rgb.x, (double)2.2f);
  rgb.y = (float)pow((double)rgb.y, (double)2.2f);
  rgb.z = (float)pow((double)rgb.z, (double)2.2f);
  const float tSunAngle = fmaxf(-sunDir.y, 0.0f);
  const float threshold = 0.9985f + tSunAngle*(0.9995f - 0.9985f);
  const float tSun = mfDot(sunDir, minusRay);
  if (tSun >= threshold)
  {
    const float k = 2.0f + 2.0f*tSunAngle;
    const HcMf3 sunColor = mf3(k*sunColorIn.x, k*sunColorIn.y, k*sunColorIn.z);
    float tSun2 = (tSun - threshold)/(1.0f - threshold);
    tSun2 = tSun2*tSun2;
    const float om = 1.0f - tSun2;
    rgb = mf3(sunColor.x*tSun2 + om*rgb.x, sunColor.y*tSun2 + om*rgb.y, sunColor.z*tSun2 + om*rgb.z);
  }
  return rgb;
}
