// TEST INFRASTRUCTURE: the product header hydracore_b200/csrc/hc_microfacet.cuh compiled as plain C++ (it is pure arithmetic), behind the
// same signatures as oracle/ref_driver.cpp's ref_pbrt_* so that tests/test_microfacet.py can compare the two on the CPU, bit for bit.
#include "../hydracore_b200/csrc/hc_microfacet.cuh"
#include "../hydracore_b200/csrc/hc_perez.cuh"

template<int KIND>
static void Run(const float* wo3, const float* wi3, const float* u2, const float* alpha2, int n, float* out8)
{
  for (int i = 0; i < n; i++)
  {
    const HcMf3 wo = mf3(wo3[3*i], wo3[3*i + 1], wo3[3*i + 2]), wi = mf3(wi3[3*i], wi3[3*i + 1], wi3[3*i + 2]);
    const float ax = alpha2[2*i], ay = alpha2[2*i + 1];
    const HcMf3 whe = mfNormalize(mf3(wo.x + wi.x, wo.y + wi.y, wo.z + wi.z));
    const HcMf3 wh = mfSampleWh<KIND>(wo, u2[2*i], u2[2*i + 1], ax, ay);
    float* o = out8 + 8*i;
    o[0] = mfBrdf<KIND>(wo, wi, ax, ay); o[1] = mfPdf<KIND>(wo, whe, ax, ay);
    o[2] = wh.x; o[3] = wh.y; o[4] = wh.z; o[5] = mfD<KIND>(wh, ax, ay);
    o[6] = mfLambda<KIND>(wo, ax, ay); o[7] = mfRoughnessToAlpha(u2[2*i]);
  }
}

extern "C" void host_pbrt_microfacet(int kind, const float* wo3, const float* wi3, const float* u2, const float* alpha2, int n, float* out8)
{
  if (kind == 0) Run<0>(wo3, wi3, u2, alpha2, n, out8); else Run<1>(wo3, wi3, u2, alpha2, n, out8);
}
extern "C" void host_pbrt_erf(const float* x, int n, float* erfOut, float* erfInvOut)
{
  for (int i = 0; i < n; i++) { erfOut[i] = mfErf(x[i]); erfInvOut[i] = mfErfInv(x[i]); }
}
extern "C" void host_perez_sky(const float* sunDir3, float turbidity, const float* sunColor3, const float* dirs3, int n, float* out3)
{
  for (int i = 0; i < n; i++)
  {
    const HcMf3 c = mfPerezSkyColor(mf3(sunDir3[0], sunDir3[1], sunDir3[2]), turbidity, mf3(sunColor3[0], sunColor3[1], sunColor3[2]), mf3(dirs3[3*i], dirs3[3*i + 1], dirs3[3*i + 2]));
    out3[3*i] = c.x; out3[3*i + 1] = c.y; out3[3*i + 2] = c.z;
  }
}
