// TEST INFRASTRUCTURE ONLY (oracle/_ref build): stand-in for HydraAPI's HR_HDRImage.h (not vendored).
// Only what reference hydra_drv/CPUExp_Integrators*.cpp name (AQMC helper paths that the oracle never runs).
#pragma once
#include <vector>
namespace HydraRender
{
  class HDRImage4f
  {
  public:
    HDRImage4f() : m_w(0), m_h(0) {}
    HDRImage4f(int w, int h) { resize(w, h); }
    void resize(int w, int h) { m_w = w; m_h = h; m_data.assign(size_t(w)*size_t(h)*4, 0.0f); }
    int width()  const { return m_w; }
    int height() const { return m_h; }
    float*       data()       { return m_data.data(); }
    const float* data() const { return m_data.data(); }
    void resampleTo(HDRImage4f& dst) const { (void)dst; }
    void medianFilterInPlace(float a = 0.0f, float b = 0.0f) { (void)a; (void)b; }
    void gaussBlur(int, float) {}
  private:
    int m_w, m_h;
    std::vector<float> m_data;
  };
}
