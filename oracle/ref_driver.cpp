// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path (hydracore_b200/).
//
// oracle/_ref/libhydra_ref.so = the reference's OWN device headers and CPU integrators
//   hydra_drv/{cglobals,cfetch,crandom,ctrace,cmaterial,cmatpbrt,clight,cbidir}.h
//   hydra_drv/CPUExp_Integrators_{Common,PT,PT_Loop,PT_QMC}.cpp, qmc_sobol_niederreiter.cpp
//   bakeBrdfEnergy/MSTables{GGX2017,Transp}.cpp
// compiled IN PLACE from /root/reference (see oracle/Makefile) against the small LiteMath stand-in
// under hydracore_b200/cpp/compat/ref_shim/, plus this driver which only (a) feeds them scene blobs through a C ABI and
// (b) replaces the non-deterministic GetTickCount() seeding by the per-pixel rule of SURVEY.md 8c:
//      gen[p] = RandomGenInit(seed + p), state carried across passes      (cf. reference shaders/trace.cl:6-13)
// and drops the debug red pixel of reference CPUExp_Integrators_Common.cpp:295-301.
// Ray casting runs through the reference's BVH4InstTraverse (ctrace.h:841) because Embree is not
// available (pExternalImpl == nullptr branch of IntegratorCommon::rayTrace, CPUExp_Integrators_Common.cpp:128-151).
#include "CPUExp_Integrators.h"

#include <omp.h>
#include <cstdio>
#include <cstring>
#include <vector>
#include <memory>

unsigned long GetTickCount() { return 12345ul; }              // reference globals_sys.h:107 (non-WIN32 replacement)
std::vector<float> PrefixSumm(const std::vector<float>& a_vec) // only named by the AQMC integrator, never run here
{
  std::vector<float> r(a_vec.size() + 1, 0.0f);
  for (size_t i = 0; i < a_vec.size(); i++) r[i + 1] = r[i] + a_vec[i];
  return r;
}

extern "C" void initQuasirandomGenerator(unsigned int table[QRNG_DIMENSIONS_K][QRNG_RESOLUTION_K]);
const ushort* getGgxTable();
const ushort* getTranspTable();

namespace
{
  struct RefScene
  {
    std::vector<int>      globals;   // EngineGlobals + tables blob (copied: integrators write g_flags)
    const float4*         geom      = nullptr;
    const float4*         materials = nullptr;
    const float4*         textures  = nullptr;
    const float4*         texturesAux = nullptr;
    const float4*         pdfs      = nullptr;
    const BVHNode*        nodes     = nullptr;
    const float4*         tris      = nullptr;
    int                   haveInst  = 1;
    const BVHNode*        nodes1    = nullptr;   // tree 1: meshes with opacity maps (RenderDriverRTE.cpp:1989-1991)
    const float4*         tris1     = nullptr;
    const uint2*          alpha1    = nullptr;   // RenderDriverRTE::CreateAlphaTestTable
    std::vector<float4x4> matrices;
    std::vector<int32_t>  lightInstId;
    std::vector<int>      remapLists, remapInst;   // material remap lists (SetMaterialRemapListPtrs, CPUExp_Integrators_Common.cpp:103-112)
    std::vector<int2>     remapTable;
    int w = 0, h = 0;

    EngineGlobals* G() { return (EngineGlobals*)globals.data(); }

    SceneGeomPointers Pointers()
    {
      SceneGeomPointers p;
      p.nodesPtr[0] = nodes;
      p.primsPtr[0] = tris;
      p.alphaTbl[0] = nullptr;
      p.haveInst[0] = (haveInst != 0);
      p.meshes          = geom;
      p.matrices        = matrices.data();
      p.instLightInstId = lightInstId.data();
      p.pExternalImpl   = nullptr;
      p.bvhTreesNumber  = 1;
      if (nodes1 != nullptr)
      {
        p.nodesPtr[1] = nodes1; p.primsPtr[1] = tris1; p.alphaTbl[1] = alpha1; p.haveInst[1] = true;
        p.bvhTreesNumber = 2;
      }
      p.matrixNum       = int(matrices.size());
      return p;
    }
  };

  // deterministic per-pixel driver over any of the reference integrators
  template<class Base>
  struct Det : public Base
  {
    template<class... A> Det(A... a) : Base(a...) {}
    std::vector<RandomGen> pix;      // generators of the stream the current pass draws from
    std::vector<std::vector<RandomGen>> parked;   // the other streams (sample streams: pass p uses stream p mod S, generator index k*W*H + pixel)
    std::vector<float4>    sum;      // per-pixel SUM of samples (the GPU layer keeps sums too, screen.cl:409-463)
    int passes = 0, streams = 1, seed0 = 0;

    void Seed(int seed) { seed0 = seed; SetStreams(1); }
    void SetStreams(int S)
    {
      const size_t n = size_t(this->m_width)*this->m_height;
      streams = S < 1 ? 1 : S;
      parked.assign(size_t(streams), std::vector<RandomGen>());
      for (int k = 0; k < streams; k++)
      {
        parked[k].resize(n);
        for (size_t i = 0; i < n; i++) parked[k][i] = RandomGenInit(seed0 + int(size_t(k)*n + i));
      }
      pix.swap(parked[0]);
      sum.assign(n, float4(0, 0, 0, 0));
      passes = 0;
    }
    void NextPass()                  // park the stream of the pass just done, take the one of the next pass
    {
      pix.swap(parked[size_t(passes % streams)]);
      passes++;
      pix.swap(parked[size_t(passes % streams)]);
    }

    // mirrors IntegratorCommon::DoPass (CPUExp_Integrators_Common.cpp:278-316) with the per-pixel generator swapped in
    void PassPixels(int x0, int y0, int x1, int y1, unsigned long long* rayCount)
    {
      const int W = this->m_width;
      #pragma omp parallel for collapse(2) schedule(dynamic, 64)
      for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++)
        {
          auto& pt  = this->PerThread();
          pt.gen    = pix[size_t(y)*W + x];
          pt.qmcPos = -1;
          float3 ray_pos, ray_dir;
          std::tie(ray_pos, ray_dir) = this->makeEyeRay(x, y);
          const float3 c = this->PathTrace(ray_pos, ray_dir, makeInitialMisData(), 0, 0);
          pix[size_t(y)*W + x] = pt.gen;
          sum[size_t(y)*W + x] += to_float4(c, 0.0f);
        }
      NextPass();
      (void)rayCount;
    }

    // mirrors IntegratorMISPT_QMC::DoPass (CPUExp_Integrators_PT_QMC.cpp:5-49): sample i of pass s -> qmcPos = s*W*H + i,
    // pseudo-random dims come from the generator of SLOT i (the reference uses the per-thread one).
    void PassQMC(const unsigned int* table, int first, int count)
    {
      const int W = this->m_width, H = this->m_height;
      const int qmcOffset = int(size_t(W)*H)*passes;
      #pragma omp parallel for schedule(dynamic, 64)
      for (int i = first; i < first + count; ++i)
      {
        auto& pt  = this->PerThread();
        pt.gen    = pix[i];
        pt.qmcPos = qmcOffset + i;
        float4 lensOffs = rndLens(&pt.gen, nullptr, float2(1, 1), this->m_pGlobals->rmQMC, pt.qmcPos, table);
        float fx, fy; float3 ray_pos, ray_dir;
        MakeEyeRayFromF4Rnd(lensOffs, this->m_pGlobals, &ray_pos, &ray_dir, &fx, &fy);
        int x = (int)(fx), y = (int)(fy);
        if (x >= W) x = W - 1;
        if (y >= H) y = H - 1;
        if (x < 0) x = 0;
        if (y < 0) y = 0;
        const float3 c = this->PathTrace(ray_pos, ray_dir, makeInitialMisData(), 0, 0);
        pix[i] = pt.gen;
        float4& d = sum[size_t(y)*W + x];
        #pragma omp atomic
        d.x += c.x;
        #pragma omp atomic
        d.y += c.y;
        #pragma omp atomic
        d.z += c.z;
      }
      NextPass();
    }
    const unsigned int* QmcTable() const { return (const unsigned int*)this->m_tableQMC; }
  };

  struct QMCOn : public IntegratorMISPTLoop2   // MISPT body + QMC table enabled, as IntegratorMISPT_QMC does
  {
    QMCOn(int w, int h, EngineGlobals* g, int f) : IntegratorMISPTLoop2(w, h, g, f) {}
    const unsigned int* GetQMCTableIfEnabled() const override { return (const unsigned int*)m_tableQMC; }
  };

  struct RefRender
  {
    RefScene* scn = nullptr;
    int kind = 0;
    std::unique_ptr<Det<IntegratorStupidPT>>   pt;
    std::unique_ptr<Det<IntegratorMISPT>>      mis;
    std::unique_ptr<Det<IntegratorMISPTLoop2>> loop;
    std::unique_ptr<Det<QMCOn>>                qmc;
  };

  template<class I> void Setup(I* p, RefScene* s, int maxDepth)
  {
    p->SetSceneGeomPtrs(s->Pointers());
    p->SetMaterialStoragePtr(s->materials);
    p->SetTexturesStoragePtr(s->textures);
    p->SetTexturesStorageAuxPtr(s->texturesAux);
    p->SetPdfStoragePtr(s->pdfs);
    p->SetMaxDepth(maxDepth);                       // CPUExpLayer.cpp:130: m_maxDepth = HRT_TRACE_DEPTH (PT adds 1 itself)
    if (!s->remapLists.empty() && !s->remapTable.empty() && !s->remapInst.empty())
      p->SetMaterialRemapListPtrs(s->remapLists.data(), s->remapTable.data(), s->remapInst.data(),
                                  int(s->remapLists.size()), int(s->remapTable.size()), int(s->remapInst.size()));
  }
}

extern "C"
{

// ---------------------------------------------------------------------------------------------- layout facts
// name/value pairs of every struct size, field offset and enum the packers on the product side mirror.
struct RefConst { const char* name; long long value; };
#define RC(x) { #x, (long long)(x) }
#define RCO(s, f) { "offsetof_" #s "_" #f, (long long)offsetof(s, f) }
static const RefConst g_consts[] = {
  {"sizeof_EngineGlobals", (long long)sizeof(EngineGlobals)}, {"sizeof_PlainLight", (long long)sizeof(PlainLight)},
  {"sizeof_PlainMaterial", (long long)sizeof(PlainMaterial)}, {"sizeof_PlainMesh", (long long)sizeof(PlainMesh)},
  {"sizeof_BVHNode", (long long)sizeof(BVHNode)}, {"sizeof_Lite_Hit", (long long)sizeof(Lite_Hit)},
  {"sizeof_SWTexSampler", (long long)sizeof(SWTexSampler)}, {"sizeof_RandomGen", (long long)sizeof(RandomGen)},
  RCO(EngineGlobals, mProj), RCO(EngineGlobals, mWorldView), RCO(EngineGlobals, mProjInverse), RCO(EngineGlobals, mWorldViewInverse),
  RCO(EngineGlobals, varsI), RCO(EngineGlobals, varsF), RCO(EngineGlobals, rmQMC), RCO(EngineGlobals, camForward),
  RCO(EngineGlobals, imagePlaneDist), RCO(EngineGlobals, texturesTableOffset), RCO(EngineGlobals, materialsTableOffset),
  RCO(EngineGlobals, pdfTableTableOffset), RCO(EngineGlobals, geometryTableOffset), RCO(EngineGlobals, texturesAuxTableOffset),
  RCO(EngineGlobals, texturesTableSize), RCO(EngineGlobals, materialsTableSize), RCO(EngineGlobals, pdfTableTableSize),
  RCO(EngineGlobals, geometryTableSize), RCO(EngineGlobals, texturesAuxTableSize), RCO(EngineGlobals, floatArraysOffset),
  RCO(EngineGlobals, floatsArraysSize), RCO(EngineGlobals, lightSelectorTableOffsetRev), RCO(EngineGlobals, lightSelectorTableSizeRev),
  RCO(EngineGlobals, lightSelectorTableOffsetFwd), RCO(EngineGlobals, lightSelectorTableSizeFwd), RCO(EngineGlobals, g_flags),
  RCO(EngineGlobals, skyLightId), RCO(EngineGlobals, lightsOffset), RCO(EngineGlobals, lightsSize), RCO(EngineGlobals, lightsNum),
  RCO(EngineGlobals, sunNumber), RCO(EngineGlobals, suns), RCO(EngineGlobals, m_allTablesAreReady),
  RCO(EngineGlobals, m_essGgx2017Table), RCO(EngineGlobals, m_essTranspTable),
#include "ref_consts.inc"
};
int ref_num_consts() { return int(sizeof(g_consts)/sizeof(g_consts[0])); }
const char* ref_const_name(int i) { return g_consts[i].name; }
long long   ref_const_value(int i) { return g_consts[i].value; }

void ref_init_engine_globals(void* out) { InitEngineGlobals((EngineGlobals*)out, getGgxTable(), getTranspTable()); }
void ref_ms_tables(unsigned short* ggx, unsigned short* transp)
{
  memcpy(ggx, getGgxTable(), sizeof(ushort)*64*64);
  memcpy(transp, getTranspTable(), sizeof(ushort)*64*64*64);
}

// ---------------------------------------------------------------------------------------------- a1/a2: samplers
void ref_rng_init(int seed, unsigned int* state2) { RandomGen g = RandomGenInit(seed); state2[0] = g.state.x; state2[1] = g.state.y; }
void ref_rng_float4(unsigned int* state2, int n, float* out4n)
{
  RandomGen g; g.state.x = state2[0]; g.state.y = state2[1]; g.rptr = 0;
  for (int i = 0; i < n; i++) { float4 r = rndFloat4_Pseudo(&g); out4n[4*i+0] = r.x; out4n[4*i+1] = r.y; out4n[4*i+2] = r.z; out4n[4*i+3] = r.w; }
  state2[0] = g.state.x; state2[1] = g.state.y;
}
void ref_rng_float1(unsigned int* state2, int n, float* out)
{
  RandomGen g; g.state.x = state2[0]; g.state.y = state2[1]; g.rptr = 0;
  for (int i = 0; i < n; i++) out[i] = rndFloat1_Pseudo(&g);
  state2[0] = g.state.x; state2[1] = g.state.y;
}
void ref_qmc_table(unsigned int* out) { initQuasirandomGenerator((unsigned int (*)[QRNG_RESOLUTION_K])out); }
float ref_qmc_sobol(unsigned int pos, int dim, const unsigned int* table) { return rndQmcSobolN(pos, dim, table); }

// ---------------------------------------------------------------------------------------------- a4: eye rays
void ref_make_rand_eye_rays(const void* globals, int w, int h, const int* xy, const float* offsets4, int n, float* outPosDir6)
{
  const EngineGlobals* G = (const EngineGlobals*)globals;
  for (int i = 0; i < n; i++)
  {
    float3 p, d;
    MakeRandEyeRay(xy[2*i], xy[2*i+1], w, h, float4(offsets4[4*i], offsets4[4*i+1], offsets4[4*i+2], offsets4[4*i+3]), G, &p, &d);
    float* o = outPosDir6 + 6*i; o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = d.x; o[4] = d.y; o[5] = d.z;
  }
}
void ref_make_eye_rays_f4(const void* globals, const float* lens4, int n, float* outPosDir6, float* outXY)
{
  const EngineGlobals* G = (const EngineGlobals*)globals;
  for (int i = 0; i < n; i++)
  {
    float3 p, d; float fx, fy;
    MakeEyeRayFromF4Rnd(float4(lens4[4*i], lens4[4*i+1], lens4[4*i+2], lens4[4*i+3]), G, &p, &d, &fx, &fy);
    float* o = outPosDir6 + 6*i; o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = d.x; o[4] = d.y; o[5] = d.z;
    outXY[2*i] = fx; outXY[2*i+1] = fy;
  }
}

// ---------------------------------------------------------------------------------------------- a7/a9: traversal
// rays: n x 8 floats (pos.xyz, tNear | dir.xyz, tFar) ; hits: n x Lite_Hit
void ref_trace_closest(const void* nodes, const void* tris, int haveInst, const float* rays8, long long n, void* hitsOut)
{
  Lite_Hit* out = (Lite_Hit*)hitsOut;
  #pragma omp parallel for schedule(dynamic, 256)
  for (long long i = 0; i < n; i++)
  {
    const float* r = rays8 + 8*i;
    Lite_Hit h = Make_Lite_Hit(MAXFLOAT, -1);    // IntegratorCommon::rayTrace, CPUExp_Integrators_Common.cpp:131
    if (haveInst) h = BVH4InstTraverse(float3(r[0], r[1], r[2]), float3(r[4], r[5], r[6]), 0.0f, h, (const float4*)nodes, (const float4*)tris);
    else          h = BVH4Traverse    (float3(r[0], r[1], r[2]), float3(r[4], r[5], r[6]), 0.0f, h, (const float4*)nodes, (const float4*)tris);
    out[i] = h;
  }
}
// CPU shadow semantics = full closest hit then 0 < t < t_far (IntegratorCommon::shadowTrace, CPUExp_Integrators_Common.cpp:156-180)
void ref_trace_shadow(const void* nodes, const void* tris, int haveInst, const float* rays8, long long n, unsigned char* visibleOut)
{
  #pragma omp parallel for schedule(dynamic, 256)
  for (long long i = 0; i < n; i++)
  {
    const float* r = rays8 + 8*i;
    Lite_Hit h = Make_Lite_Hit(MAXFLOAT, -1);
    if (haveInst) h = BVH4InstTraverse(float3(r[0], r[1], r[2]), float3(r[4], r[5], r[6]), 0.0f, h, (const float4*)nodes, (const float4*)tris);
    else          h = BVH4Traverse    (float3(r[0], r[1], r[2]), float3(r[4], r[5], r[6]), 0.0f, h, (const float4*)nodes, (const float4*)tris);
    visibleOut[i] = (HitSome(h) && h.t > 0.0f && h.t < r[7]) ? 0 : 1;
  }
}
// the OpenCL layer's any-hit kernel body (BVH4InstTraverseShadow, ctrace.h:1065-1294): returns hit with early exit
void ref_trace_shadow_anyhit(const void* nodes, const void* tris, const float* rays8, long long n, unsigned char* visibleOut)
{
  #pragma omp parallel for schedule(dynamic, 256)
  for (long long i = 0; i < n; i++)
  {
    const float* r = rays8 + 8*i;
    Lite_Hit h = Make_Lite_Hit(r[7], -1);
    const float3 sh = BVH4InstTraverseShadow(float3(r[0], r[1], r[2]), float3(r[4], r[5], r[6]), 0.0f, h, (const float4*)nodes, (const float4*)tris, -1);
    visibleOut[i] = (sh.x > 0.5f) ? 1 : 0;
  }
}

// OpenMP control for the timed baselines: a launcher (torchrun) exports OMP_NUM_THREADS=1, which libgomp reads when it is loaded
// microfacet lobes of cmatpbrt.h in the local (PBRT) frame, kind 0 = Beckmann (:195-363), 1 = Trowbridge-Reitz (:365-524).  Per item:
// out[0] BRDF(wo, wi), out[1] pdf(wo, wh = normalize(wo + wi)), out[2..4] the half vector sampled from (wo, u), out[5] D(that wh),
// out[6] Lambda(wo), out[7] RoughnessToAlpha(u.x)
void ref_pbrt_microfacet(int kind, const float* wo3, const float* wi3, const float* u2, const float* alpha2, int n, float* out8)
{
  for (int i = 0; i < n; i++)
  {
    const float3 wo = make_float3(wo3[3*i], wo3[3*i + 1], wo3[3*i + 2]), wi = make_float3(wi3[3*i], wi3[3*i + 1], wi3[3*i + 2]);
    const float2 u = make_float2(u2[2*i], u2[2*i + 1]);
    const float ax = alpha2[2*i], ay = alpha2[2*i + 1];
    const float3 whe = normalize(wo + wi);
    float* o = out8 + 8*i;
    if (kind == 0)
    {
      const float3 wh = BeckmannDistributionSampleWH(wo, u, ax, ay);
      o[0] = BeckmannBRDF_PBRT(wo, wi, ax, ay); o[1] = BeckmannDistributionPdf(wo, whe, ax, ay);
      o[2] = wh.x; o[3] = wh.y; o[4] = wh.z; o[5] = BeckmannDistributionD(wh, ax, ay);
      o[6] = BeckmannDistributionLambda(wo, ax, ay); o[7] = BeckmannRoughnessToAlpha(u.x);
    }
    else
    {
      const float3 wh = TrowbridgeReitzDistributionSampleWH(wo, u, ax, ay);
      o[0] = TrowbridgeReitzBRDF_PBRT(wo, wi, ax, ay); o[1] = TrowbridgeReitzDistributionPdf(wo, whe, ax, ay);
      o[2] = wh.x; o[3] = wh.y; o[4] = wh.z; o[5] = TrowbridgeReitzDistributionD(wh, ax, ay);
      o[6] = TrowbridgeReitzDistributionLambda(wo, ax, ay); o[7] = TrowbridgeReitzRoughnessToAlpha(u.x);
    }
  }
}
// skyLightPerezColor (clight.h:255-283) of a sky-dome light with the given sun direction, turbidity and sun colour, for n ray directions
void ref_perez_sky(const float* sunDir3, float turbidity, const float* sunColor3, const float* dirs3, int n, float* out3)
{
  PlainLight L; memset(&L, 0, sizeof(L));
  L.data[SKY_DOME_SUN_DIR_X] = sunDir3[0]; L.data[SKY_DOME_SUN_DIR_Y] = sunDir3[1]; L.data[SKY_DOME_SUN_DIR_Z] = sunDir3[2];
  L.data[SKY_DOME_TURBIDITY] = turbidity;
  L.data[SKY_SUN_COLOR_X] = sunColor3[0]; L.data[SKY_SUN_COLOR_Y] = sunColor3[1]; L.data[SKY_SUN_COLOR_Z] = sunColor3[2];
  for (int i = 0; i < n; i++)
  {
    const float3 c = skyLightPerezColor(&L, make_float3(dirs3[3*i], dirs3[3*i + 1], dirs3[3*i + 2]));
    out3[3*i] = c.x; out3[3*i + 1] = c.y; out3[3*i + 2] = c.z;
  }
}
void ref_pbrt_erf(const float* x, int n, float* erfOut, float* erfInvOut)
{
  for (int i = 0; i < n; i++) { erfOut[i] = ErfPBRT(x[i]); erfInvOut[i] = ErfInvPBRT(x[i]); }
}

int ref_omp_max_threads() { return omp_get_max_threads(); }
void ref_omp_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// One ray-casting step of the C2 workload entirely inside the reference code: closest hit of every ray (BVH4InstTraverse), then one shadow
// ray per hit towards `light` (t_far = 0.995 x distance, CPUExp_Integrators_PT_Loop.cpp:176; origin pushed off the surface by 1e-4 x the
// largest coordinate, as hc_make_shadow_rays does), answered by closest-hit-then-compare (IntegratorCommon::shadowTrace).  Returns the
// number of rays traced (primary + shadow).
long long ref_raycast_step(const void* nodes, const void* tris, const float* rays8, long long n, const float* light3, void* hitsOut, unsigned char* visibleOut)
{
  Lite_Hit* out = (Lite_Hit*)hitsOut;
  long long traced = 0;
  #pragma omp parallel for schedule(dynamic, 256) reduction(+:traced)
  for (long long i = 0; i < n; i++)
  {
    const float* r = rays8 + 8*i;
    const float3 ro(r[0], r[1], r[2]), rd(r[4], r[5], r[6]);
    Lite_Hit h = BVH4InstTraverse(ro, rd, 0.0f, Make_Lite_Hit(MAXFLOAT, -1), (const float4*)nodes, (const float4*)tris);
    out[i] = h; traced++;
    unsigned char vis = 1;
    if (HitSome(h))
    {
      const float3 pos = ro + rd*h.t;
      const float3 L(light3[0], light3[1], light3[2]);
      const float3 sdir = normalize(L - pos);
      const float eps = fmaxf(fmaxf(fabsf(pos.x), fmaxf(fabsf(pos.y), fabsf(pos.z))), 1.0f)*1e-4f;
      const float3 spos = pos + sdir*eps;
      const float tFar = length(spos - L)*0.995f;
      const Lite_Hit sh = BVH4InstTraverse(spos, sdir, 0.0f, Make_Lite_Hit(MAXFLOAT, -1), (const float4*)nodes, (const float4*)tris);
      vis = (HitSome(sh) && sh.t > 0.0f && sh.t < tFar) ? 0 : 1;
      traced++;
    }
    visibleOut[i] = vis;
  }
  return traced;
}

// ---------------------------------------------------------------------------------------------- scene + integrators
void* ref_scene_create(const int* globalsBlob, long long nInts,
                       const void* geom, const void* materials, const void* textures, const void* texturesAux, const void* pdfs,
                       const void* nodes, const void* tris, int haveInst,
                       const float* invMatrices16, int nInst, const int* lightInstId, int w, int h)
{
  RefScene* s = new RefScene;
  s->globals.assign(globalsBlob, globalsBlob + nInts);
  s->geom = (const float4*)geom; s->materials = (const float4*)materials; s->textures = (const float4*)textures;
  s->texturesAux = (const float4*)texturesAux; s->pdfs = (const float4*)pdfs;
  s->nodes = (const BVHNode*)nodes; s->tris = (const float4*)tris; s->haveInst = haveInst;
  s->matrices.resize(nInst);
  for (int i = 0; i < nInst; i++) memcpy(&s->matrices[i], invMatrices16 + 16*i, 64);   // already column storage (4 x float4)
  s->lightInstId.assign(lightInstId, lightInstId + nInst);
  s->w = w; s->h = h;
  return s;
}
// second tree of the ConvertionResult with its alpha table (nullptr = plain second tree); call before ref_render_create
void ref_scene_set_tree1(void* p, const void* nodes, const void* tris, const void* alphaUint2)
{
  RefScene* s = (RefScene*)p;
  s->nodes1 = (const BVHNode*)nodes; s->tris1 = (const float4*)tris; s->alpha1 = (const uint2*)alphaUint2;
}
// material remap lists: what the driver hands to SetAllRemapLists / SetAllInstIdToRemapId (RenderDriverRTE.cpp:1340-1376, 1478)
void ref_scene_set_remap(void* p, const int* allLists, int allSize, const int* tableOffsetAndSize, int tableSize, const int* instRemapId, int nInst)
{
  RefScene* s = (RefScene*)p;
  s->remapLists.assign(allLists, allLists + allSize);
  s->remapTable.resize(tableSize);
  for (int i = 0; i < tableSize; i++) s->remapTable[i] = int2(tableOffsetAndSize[2*i], tableOffsetAndSize[2*i + 1]);
  s->remapInst.assign(instRemapId, instRemapId + nInst);
}
void ref_scene_destroy(void* p) { delete (RefScene*)p; }

// kind: 0 = IntegratorStupidPT (PT), 1 = IntegratorMISPT (recursive), 2 = IntegratorMISPTLoop2 (the one CPUExpLayer instantiates,
// IHWLayerDataAssembler.cpp:579), 3 = MISPT body with the QMC table on (IntegratorMISPT_QMC semantics)
void* ref_render_create(void* scene, int kind, int seed)
{
  RefScene* s = (RefScene*)scene;
  RefRender* r = new RefRender; r->scn = s; r->kind = kind;
  const int depth = s->G()->varsI[HRT_TRACE_DEPTH];
  if (kind != 0) s->G()->g_flags &= ~HRT_STUPID_PT_MODE;     // a PT render on this scene object leaves the flag behind: renders must not depend on their order
  if (kind == 0)      { r->pt.reset(new Det<IntegratorStupidPT>(s->w, s->h, s->G()));        Setup(r->pt.get(), s, depth);   r->pt->Seed(seed);
                        s->G()->g_flags |= HRT_STUPID_PT_MODE; }                              // IntegratorStupidPT::DoPass, CPUExp_Integrators.h:326-330
  else if (kind == 1) { r->mis.reset(new Det<IntegratorMISPT>(s->w, s->h, s->G(), 0));       Setup(r->mis.get(), s, depth);  r->mis->Seed(seed); }
  else if (kind == 2) { r->loop.reset(new Det<IntegratorMISPTLoop2>(s->w, s->h, s->G(), 0)); Setup(r->loop.get(), s, depth); r->loop->Seed(seed); }
  else                { r->qmc.reset(new Det<QMCOn>(s->w, s->h, s->G(), 0));                 Setup(r->qmc.get(), s, depth);  r->qmc->Seed(seed); }
  return r;
}
void ref_render_destroy(void* p) { delete (RefRender*)p; }
// S generators per pixel, pass p draws from stream p mod S (the rule of hc_pt_set_sample_streams); resets the render
void ref_render_set_streams(void* p, int S)
{
  RefRender* r = (RefRender*)p;
  if (r->kind == 0) r->pt->SetStreams(S);
  else if (r->kind == 1) r->mis->SetStreams(S);
  else if (r->kind == 2) r->loop->SetStreams(S);
  else r->qmc->SetStreams(S);
}

void ref_render_pass(void* p, int x0, int y0, int x1, int y1)
{
  RefRender* r = (RefRender*)p;
  if (r->kind == 0) r->pt->PassPixels(x0, y0, x1, y1, nullptr);
  else if (r->kind == 1) r->mis->PassPixels(x0, y0, x1, y1, nullptr);
  else if (r->kind == 2) r->loop->PassPixels(x0, y0, x1, y1, nullptr);
  else r->qmc->PassQMC(r->qmc->QmcTable(), 0, r->scn->w*r->scn->h);
}
void ref_render_pass_qmc_range(void* p, int first, int count)
{
  RefRender* r = (RefRender*)p;
  r->qmc->PassQMC(r->qmc->QmcTable(), first, count);
}
// per-pixel SUM of samples and number of passes
int ref_render_get_sum(void* p, float* out4)
{
  RefRender* r = (RefRender*)p;
  const std::vector<float4>* s; int n;
  if (r->kind == 0) { s = &r->pt->sum; n = r->pt->passes; } else if (r->kind == 1) { s = &r->mis->sum; n = r->mis->passes; }
  else if (r->kind == 2) { s = &r->loop->sum; n = r->loop->passes; } else { s = &r->qmc->sum; n = r->qmc->passes; }
  memcpy(out4, s->data(), s->size()*sizeof(float4));
  return n;
}

// ---------------------------------------------------------------------------------------------- stage probes (a10..a14)
// SurfaceHit flattened to 24 floats: pos3 normal3 flatNormal3 tangent3 biTangent3 texCoord2 matId t sRayOff hfi texCoordCamProj2
static void PackSurf(const SurfaceHit& s, float* o)
{
  o[0]=s.pos.x; o[1]=s.pos.y; o[2]=s.pos.z; o[3]=s.normal.x; o[4]=s.normal.y; o[5]=s.normal.z;
  o[6]=s.flatNormal.x; o[7]=s.flatNormal.y; o[8]=s.flatNormal.z; o[9]=s.tangent.x; o[10]=s.tangent.y; o[11]=s.tangent.z;
  o[12]=s.biTangent.x; o[13]=s.biTangent.y; o[14]=s.biTangent.z; o[15]=s.texCoord.x; o[16]=s.texCoord.y;
  o[17]=as_float(s.matId); o[18]=s.t; o[19]=s.sRayOff; o[20]=s.hfi ? 1.0f : 0.0f; o[21]=s.texCoordCamProj.x; o[22]=s.texCoordCamProj.y; o[23]=0;
}
static SurfaceHit UnpackSurf(const float* o)
{
  SurfaceHit s;
  s.pos=float3(o[0],o[1],o[2]); s.normal=float3(o[3],o[4],o[5]); s.flatNormal=float3(o[6],o[7],o[8]); s.tangent=float3(o[9],o[10],o[11]);
  s.biTangent=float3(o[12],o[13],o[14]); s.texCoord=float2(o[15],o[16]); s.matId=as_int(o[17]); s.t=o[18]; s.sRayOff=o[19]; s.hfi=(o[20]!=0.0f);
  s.texCoordCamProj=float2(o[21],o[22]);
  return s;
}

struct Probe : public IntegratorMISPTLoop2
{
  Probe(RefScene* s) : IntegratorMISPTLoop2(s->w, s->h, s->G(), 0) { Setup(this, s, s->G()->varsI[HRT_TRACE_DEPTH]); }
  using IntegratorCommon::surfaceEval;
  float3 Emission(float3 p, float3 d, const SurfaceHit& s, uint flags, MisData mis, int instId) { return IntegratorCommon::emissionEval(p, d, s, flags, mis, instId); }
  const PlainMaterial* Mat(int id) { return materialAt(m_pGlobals, m_matStorage, id); }
  const EngineGlobals* Glob() { return m_pGlobals; }
  const int4* Tex() { return m_texStorage; }
  const int4* TexAux() { return m_texStorageAux; }
  const float4* Pdf() { return m_pdfStorage; }
  ProcTextureList* Ptl() { return &m_ptlDummy; }
};

void ref_surface_eval(void* scene, const float* rays8, const void* hits, int n, float* surf24)
{
  RefScene* s = (RefScene*)scene; Probe pr(s);
  const Lite_Hit* H = (const Lite_Hit*)hits;
  for (int i = 0; i < n; i++)
  {
    const float* r = rays8 + 8*i;
    if (HitNone(H[i])) { memset(surf24 + 24*i, 0, 96); continue; }
    { Lite_Hit hh = H[i]; PackSurf(pr.surfaceEval(float3(r[0], r[1], r[2]), float3(r[4], r[5], r[6]), hh), surf24 + 24*i); }
  }
}

// MaterialSampleAndEvalBxDF (cmaterial.h:2345): out 8 floats = color3 dir3 pdf flags(as int bits) ; + matOffset
void ref_material_sample(void* scene, const float* surf24, const float* rayDir3, const float* rands10, const unsigned int* flags, int n, float* out8, int* matOffsetOut)
{
  RefScene* s = (RefScene*)scene; Probe pr(s);
  for (int i = 0; i < n; i++)
  {
    SurfaceHit sh = UnpackSurf(surf24 + 24*i);
    float rands[MMLT_FLOATS_PER_BOUNCE]; memcpy(rands, rands10 + 10*i, sizeof(rands));
    MatSample ms; int matOffset = 0;
    MaterialSampleAndEvalBxDF(pr.Mat(sh.matId), rands, &sh, float3(rayDir3[3*i], rayDir3[3*i+1], rayDir3[3*i+2]), make_float3(0,0,0), flags[i], false,
                              pr.Glob(), pr.Tex(), pr.TexAux(), pr.Ptl(), &ms, &matOffset);
    float* o = out8 + 8*i;
    o[0]=ms.color.x; o[1]=ms.color.y; o[2]=ms.color.z; o[3]=ms.direction.x; o[4]=ms.direction.y; o[5]=ms.direction.z; o[6]=ms.pdf; o[7]=as_float(ms.flags);
    matOffsetOut[i] = matOffset;
  }
}

// materialEval (cmaterial.h:2554): out 8 floats = brdf3 btdf3 pdfFwd pdfRev
void ref_material_eval(void* scene, const float* surf24, const float* l3, const float* v3, int n, float* out8)
{
  RefScene* s = (RefScene*)scene; Probe pr(s);
  for (int i = 0; i < n; i++)
  {
    SurfaceHit sh = UnpackSurf(surf24 + 24*i);
    ShadeContext sc;
    sc.wp = sh.pos; sc.l = float3(l3[3*i], l3[3*i+1], l3[3*i+2]); sc.v = float3(v3[3*i], v3[3*i+1], v3[3*i+2]);
    sc.n = sh.normal; sc.fn = sh.flatNormal; sc.tg = sh.tangent; sc.bn = sh.biTangent; sc.tc = sh.texCoord;
    const PlainMaterial* m = pr.Mat(sh.matId);
    ProcTextureList ptl = *pr.Ptl();
    GetProcTexturesIdListFromMaterialHead(m, &ptl);
    BxDFResult ev = materialEval(m, &sc, (EVAL_FLAG_DEFAULT), pr.Glob(), pr.Tex(), pr.TexAux(), &ptl);
    float* o = out8 + 8*i;
    o[0]=ev.brdf.x; o[1]=ev.brdf.y; o[2]=ev.brdf.z; o[3]=ev.btdf.x; o[4]=ev.btdf.y; o[5]=ev.btdf.z; o[6]=ev.pdfFwd; o[7]=ev.pdfRev;
  }
}

// SelectRandomLightRev + LightSampleRev (clight.h:1774, 1561): out 12 floats = pos3 color3 pdf maxDist cosAtLight isPoint lightOffset pickProb
void ref_light_sample(void* scene, const float* hitPos3, const float* rnd4, int n, float* out12)
{
  RefScene* s = (RefScene*)scene; Probe pr(s);
  for (int i = 0; i < n; i++)
  {
    float3 pos(hitPos3[3*i], hitPos3[3*i+1], hitPos3[3*i+2]);
    float pick = 1.0f;
    int lightOffset = SelectRandomLightRev(rnd4[4*i+2], pos, pr.Glob(), &pick);     // .z selects: CPUExp_Integrators_PT_Loop.cpp:149
    float* o = out12 + 12*i; memset(o, 0, 48);
    o[10] = as_float(lightOffset); o[11] = pick;
    if (lightOffset < 0) continue;
    ShadowSample sam;
    LightSampleRev(lightAt(pr.Glob(), lightOffset), float3(rnd4[4*i], rnd4[4*i+1], rnd4[4*i+2]), pos, pr.Glob(), pr.Pdf(), pr.Tex(), &sam);
    o[0]=sam.pos.x; o[1]=sam.pos.y; o[2]=sam.pos.z; o[3]=sam.color.x; o[4]=sam.color.y; o[5]=sam.color.z; o[6]=sam.pdf; o[7]=sam.maxDist; o[8]=sam.cosAtLight;
    o[9]=sam.isPoint ? 1.0f : 0.0f;
  }
}

// emissionEval through the integrator wrapper (CPUExp_Integrators_Common.cpp:509-520) ; prevSpecular -> MisData.isSpecular
void ref_emission_eval(void* scene, const float* rays8, const float* surf24, const int* instId, const unsigned int* flags, const int* prevSpecular, int n, float* out3)
{
  RefScene* s = (RefScene*)scene; Probe pr(s);
  for (int i = 0; i < n; i++)
  {
    SurfaceHit sh = UnpackSurf(surf24 + 24*i);
    MisData mis = makeInitialMisData(); mis.isSpecular = prevSpecular[i];
    const float* r = rays8 + 8*i;
    float3 e = pr.Emission(float3(r[0], r[1], r[2]), float3(r[4], r[5], r[6]), sh, flags[i], mis, instId[i]);
    out3[3*i]=e.x; out3[3*i+1]=e.y; out3[3*i+2]=e.z;
  }
}

} // extern "C"
