// BVH4 builder for the CUDA layer (host side, C++17 + OpenMP tasks).
//
// Stands where libhydrabvhbuilder stands in the reference (bvh_builder/bvh_access_dll2.cpp, IBVHBuilder2 in
// hydra_drv/IBVHBuilderAPI.h:35-68).  The reference builds with Embree 2.17 internals, which are not available, so
// this is our own binned-SAH builder; only the OUTPUT layout is the reference's (what ConvertMap() emits,
// bvh_access_dll2.cpp:388-717, consumed by BVH4InstTraverse, hydra_drv/ctrace.h:841-1062):
//   * 32-byte nodes {boxMin.xyz, leftOffsetAndLeaf | boxMax.xyz, escapeIndex}, allocated in quads of 4 (128 B)
//   * quad 0 = tree root record (scene box, leftOffset = 1, identity matrix in nodes 1..2), traversal starts at quad 1
//   * top level: BVH4 over instance world boxes; an instance leaf points at a 4-node instance record
//       node0 = {mesh root box, -> mesh sub-tree}, nodes1..2 = inverse instance matrix (4 column float4),
//       node3 = int4{instId, meshId, 0, 0}; mesh sub-trees are shared by all instances of a mesh
//   * triangle leaves point (in float4 units) at a header int4{first, count, -1, -1} followed by count x 3 float4
//       (A.xyz, primId) (B.xyz, meshId) (C.xyz, -1); zero-area triangles are dropped
//   * unused child slots: both uints 0xFFFFFFFF, box (+inf, -inf)   (BVHNodeT ctor, cglobals.h:1284-1290)
#include "../../include/hydracore_cuda.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace hcb
{
struct V3 { float x, y, z; };
static inline V3 vmin(V3 a, V3 b) { return { std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z) }; }
static inline V3 vmax(V3 a, V3 b) { return { std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z) }; }

struct Box
{
  V3 lo{ INFINITY, INFINITY, INFINITY }, hi{ -INFINITY, -INFINITY, -INFINITY };
  void grow(V3 p) { lo = vmin(lo, p); hi = vmax(hi, p); }
  void grow(const Box& b) { lo = vmin(lo, b.lo); hi = vmax(hi, b.hi); }
  float area() const
  {
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    if (!(dx >= 0.0f && dy >= 0.0f && dz >= 0.0f)) return 0.0f;
    return 2.0f*(dx*dy + dy*dz + dz*dx);
  }
};

struct Node32 { float bmin[3]; uint32_t leftOffsetAndLeaf; float bmax[3]; uint32_t escapeIndex; };
static_assert(sizeof(Node32) == 32, "BVHNode is 32 bytes (cglobals.h:1280-1317)");
struct F4 { float x, y, z, w; };

static inline Node32 InvalidNode()
{
  Node32 n;
  n.bmin[0] = n.bmin[1] = n.bmin[2] = INFINITY;
  n.bmax[0] = n.bmax[1] = n.bmax[2] = -INFINITY;
  n.leftOffsetAndLeaf = 0xFFFFFFFFu; n.escapeIndex = 0xFFFFFFFFu;
  return n;
}
static inline float AsFloat(int32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline void SetBox(Node32& n, const Box& b) { n.bmin[0] = b.lo.x; n.bmin[1] = b.lo.y; n.bmin[2] = b.lo.z; n.bmax[0] = b.hi.x; n.bmax[1] = b.hi.y; n.bmax[2] = b.hi.z; }

// ---------------------------------------------------------------------------------------------------------------- build tree
struct Prim { Box box; V3 c; int32_t id; };

struct BNode                      // temporary wide node
{
  Box     box;
  int32_t child[4] = { -1, -1, -1, -1 };   // indices into tree nodes
  int32_t first = 0, count = 0;            // leaf: range in prim order
  bool    leaf = false;
};

struct Tree
{
  std::vector<BNode> nodes;       // nodes[0] = root
  std::vector<Prim>  prims;       // reordered
  int depth = 0;
};

struct Range { int32_t first, count; Box box; };

// Builder parameters (defaults from the sweep in DESIGN.md section 3: 32 bins, leaves of at most 2 triangles = one device pair record, SAH leaf
// cost counted in pair records); the HC_BVH_* environment variables exist for that sweep only.
static constexpr int kMaxBins = 64;
static int EnvInt(const char* name, int dflt) { const char* e = std::getenv(name); return e ? std::atoi(e) : dflt; }
static int kBins  = std::min(kMaxBins, std::max(4, EnvInt("HC_BVH_BINS", 32)));
static int kBlock = std::max(1, EnvInt("HC_BVH_BLOCK", 2));
static int kPick  = EnvInt("HC_BVH_PICK", 0);
static int kMaxLeafEnv = EnvInt("HC_BVH_MAXLEAF", 2);
static int kSweep = EnvInt("HC_BVH_SWEEP", 0);

// one binned-SAH split of prims[first, first+count) ; returns the size of the left part (0 < L < count)
static int32_t SplitSAH(std::vector<Prim>& P, int32_t first, int32_t count)
{
  if (count <= kSweep)                                 // small range: exact sweep over the centroid order on each axis
  {
    float bestCost = INFINITY; int bestAxis = -1; int32_t bestL = -1;
    std::vector<float> rA(count);
    for (int axis = 0; axis < 3; axis++)
    {
      std::sort(P.begin() + first, P.begin() + first + count, [axis](const Prim& a, const Prim& b) { return (&a.c.x)[axis] < (&b.c.x)[axis]; });
      Box acc;
      for (int32_t i = count - 1; i > 0; i--) { acc.grow(P[first + i].box); rA[i] = acc.area(); }
      acc = Box();
      for (int32_t i = 0; i < count - 1; i++)
      {
        acc.grow(P[first + i].box);
        const int32_t nl = i + 1, nr = count - nl;
        const float cost = acc.area()*float((nl + kBlock - 1)/kBlock) + rA[i + 1]*float((nr + kBlock - 1)/kBlock);
        if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestL = nl; }
      }
    }
    if (bestAxis != 2)
      std::sort(P.begin() + first, P.begin() + first + count, [bestAxis](const Prim& a, const Prim& b) { return (&a.c.x)[bestAxis] < (&b.c.x)[bestAxis]; });
    return bestL;
  }
  Box cb;
  for (int32_t i = first; i < first + count; i++) cb.grow(P[i].c);
  const float ext[3] = { cb.hi.x - cb.lo.x, cb.hi.y - cb.lo.y, cb.hi.z - cb.lo.z };

  float bestCost = INFINITY; int bestAxis = -1, bestBin = -1;
  for (int axis = 0; axis < 3; axis++)
  {
    if (!(ext[axis] > 0.0f)) continue;
    const float lo = (&cb.lo.x)[axis];
    const float k  = float(kBins)*(1.0f - 1e-6f)/ext[axis];
    Box bb[kMaxBins]; int32_t bc[kMaxBins] = { 0 };
    for (int32_t i = first; i < first + count; i++)
    {
      int b = int(k*((&P[i].c.x)[axis] - lo)); b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
      bb[b].grow(P[i].box); bc[b]++;
    }
    float rArea[kMaxBins]; int32_t rCnt[kMaxBins];
    Box acc; int32_t n = 0;
    for (int b = kBins - 1; b > 0; b--) { acc.grow(bb[b]); n += bc[b]; rArea[b] = acc.area(); rCnt[b] = n; }
    acc = Box(); n = 0;
    for (int b = 0; b < kBins - 1; b++)
    {
      acc.grow(bb[b]); n += bc[b];
      if (n == 0 || rCnt[b + 1] == 0) continue;
      // leaf cost in units of started kBlock-triangle records (the device tests triangles in pairs)
      const float cost = acc.area()*float((n + kBlock - 1)/kBlock) + rArea[b + 1]*float((rCnt[b + 1] + kBlock - 1)/kBlock);
      if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestBin = b; }
    }
  }

  if (bestAxis < 0)                                    // all centroids coincide: split in the middle
    return count/2;

  const float lo = (&cb.lo.x)[bestAxis];
  const float k  = float(kBins)*(1.0f - 1e-6f)/ext[bestAxis];
  auto mid = std::partition(P.begin() + first, P.begin() + first + count, [&](const Prim& p)
  {
    int b = int(k*((&p.c.x)[bestAxis] - lo)); b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
    return b <= bestBin;
  });
  int32_t L = int32_t(mid - (P.begin() + first));
  if (L <= 0 || L >= count) L = count/2;
  return L;
}

static Box RangeBox(const std::vector<Prim>& P, int32_t first, int32_t count)
{
  Box b; for (int32_t i = first; i < first + count; i++) b.grow(P[i].box); return b;
}

// quad levels a balanced (median-split) tree over `count` primitives still needs
static int LevelsNeeded(int32_t count, int maxLeaf)
{
  int levels = 0; long long cap = maxLeaf;
  while (cap < count) { cap *= 4; levels++; }
  return levels;
}

// recursive wide build; node slots are reserved up front so that tasks can write without locking.
// maxDepth bounds the tree: the traversal stack of the kernels (HC_STACK_CAP = 80 = 3 x (top levels + mesh levels) + 2, hc_api.cu) must hold
// the worst case, so once the levels left are only just enough for a balanced tree the splits become median splits (an unbalanced SAH tree
// over long thin geometry could otherwise grow deeper than the stack and the scene would be refused at upload).
static void BuildRec(Tree& T, int32_t nodeIdx, int32_t first, int32_t count, int maxLeaf, int depth, std::vector<int>& depthOut, int maxDepth)
{
  BNode& N = T.nodes[nodeIdx];   // NOTE: T.nodes is pre-sized; never reallocates during the build
  N.box = RangeBox(T.prims, first, count);
  if (count <= maxLeaf) { N.leaf = true; N.first = first; N.count = count; depthOut[nodeIdx] = depth; return; }

  Range parts[4]; int np = 1;
  parts[0] = { first, count, N.box };
  while (np < 4)
  {
    int best = -1; float bestA = -1.0f;
    for (int i = 0; i < np; i++)
      if (parts[i].count > maxLeaf) { const float a = kPick ? parts[i].box.area() : parts[i].box.area()*float(parts[i].count); if (a > bestA) { bestA = a; best = i; } }
    if (best < 0) break;
    const Range r = parts[best];
    const bool balanced = LevelsNeeded(count, maxLeaf) + 1 >= maxDepth - depth;       // no slack left for uneven splits below this node
    const int32_t L = balanced ? r.count/2 : SplitSAH(T.prims, r.first, r.count);
    parts[best] = { r.first, L, RangeBox(T.prims, r.first, L) };
    parts[np++] = { r.first + L, r.count - L, RangeBox(T.prims, r.first + L, r.count - L) };
  }
  std::sort(parts, parts + np, [](const Range& a, const Range& b) { return a.first < b.first; });

  // child node indices: every interior node has >= 2 children, so a sub-tree over c prims needs at most 2c-1 nodes;
  // carve disjoint slot ranges (1 + sum(2c_i - 1) <= 2n - 1 slots for this sub-tree)
  int32_t slot = nodeIdx + 1;
  for (int i = 0; i < np; i++) { N.child[i] = slot; slot += 2*parts[i].count - 1; }
  for (int i = 0; i < np; i++)
  {
    const int32_t ci = N.child[i]; const Range r = parts[i];
    if (r.count > 8192)
    {
      #pragma omp task shared(T, depthOut) firstprivate(ci, r, maxLeaf, depth, maxDepth)
      BuildRec(T, ci, r.first, r.count, maxLeaf, depth + 1, depthOut, maxDepth);
    }
    else
      BuildRec(T, ci, r.first, r.count, maxLeaf, depth + 1, depthOut, maxDepth);
  }
  #pragma omp taskwait
}

static void BuildTree(Tree& T, int maxLeaf, int maxDepth)
{
  const int32_t n = int32_t(T.prims.size());
  T.nodes.assign(size_t(std::max(2*n, 2)), BNode());
  std::vector<int> depthOut(T.nodes.size(), 0);
  #pragma omp parallel
  #pragma omp single
  BuildRec(T, 0, 0, n, maxLeaf, 0, depthOut, maxDepth);
  T.depth = *std::max_element(depthOut.begin(), depthOut.end());
}

// ---------------------------------------------------------------------------------------------------------------- builder object
struct Mesh
{
  std::vector<F4>      vert;
  std::vector<int32_t> idx;
  Tree  tree;
  Box   box;
  bool  built = false;
  int64_t subtreeRef = -1;        // leftOffsetAndLeaf of the flattened mesh root (shared by its instances)
};

struct Instance { int meshId; int realId = -1; /* id written into the instance record; -1 = index in this builder */ float m[16]; /* row-major */ float inv[16]; /* 4 columns */ Box worldBox; };

static void Inverse4x4Columns(const float rowMajor[16], float outCols[16])
{
  // same cofactor expansion and operation order as inverse4x4 in the reference (cglobals.h:917-1001), on column storage
  float c[4][4];                                   // c[col][row]
  for (int col = 0; col < 4; col++) for (int row = 0; row < 4; row++) c[col][row] = rowMajor[row*4 + col];
  #define X 0
  #define Y 1
  #define Z 2
  #define W 3
  float t[12], m[4][4];
  t[0]=c[2][Z]*c[3][W]; t[1]=c[3][Z]*c[2][W]; t[2]=c[1][Z]*c[3][W]; t[3]=c[3][Z]*c[1][W]; t[4]=c[1][Z]*c[2][W];  t[5]=c[2][Z]*c[1][W];
  t[6]=c[0][Z]*c[3][W]; t[7]=c[3][Z]*c[0][W]; t[8]=c[0][Z]*c[2][W]; t[9]=c[2][Z]*c[0][W]; t[10]=c[0][Z]*c[1][W]; t[11]=c[1][Z]*c[0][W];
  m[0][X]  = t[0]*c[1][Y] + t[3]*c[2][Y] + t[4]*c[3][Y];   m[0][X] -= t[1]*c[1][Y] + t[2]*c[2][Y] + t[5]*c[3][Y];
  m[0][Y]  = t[1]*c[0][Y] + t[6]*c[2][Y] + t[9]*c[3][Y];   m[0][Y] -= t[0]*c[0][Y] + t[7]*c[2][Y] + t[8]*c[3][Y];
  m[0][Z]  = t[2]*c[0][Y] + t[7]*c[1][Y] + t[10]*c[3][Y];  m[0][Z] -= t[3]*c[0][Y] + t[6]*c[1][Y] + t[11]*c[3][Y];
  m[0][W]  = t[5]*c[0][Y] + t[8]*c[1][Y] + t[11]*c[2][Y];  m[0][W] -= t[4]*c[0][Y] + t[9]*c[1][Y] + t[10]*c[2][Y];
  m[1][X]  = t[1]*c[1][X] + t[2]*c[2][X] + t[5]*c[3][X];   m[1][X] -= t[0]*c[1][X] + t[3]*c[2][X] + t[4]*c[3][X];
  m[1][Y]  = t[0]*c[0][X] + t[7]*c[2][X] + t[8]*c[3][X];   m[1][Y] -= t[1]*c[0][X] + t[6]*c[2][X] + t[9]*c[3][X];
  m[1][Z]  = t[3]*c[0][X] + t[6]*c[1][X] + t[11]*c[3][X];  m[1][Z] -= t[2]*c[0][X] + t[7]*c[1][X] + t[10]*c[3][X];
  m[1][W]  = t[4]*c[0][X] + t[9]*c[1][X] + t[10]*c[2][X];  m[1][W] -= t[5]*c[0][X] + t[8]*c[1][X] + t[11]*c[2][X];
  t[0]=c[2][X]*c[3][Y]; t[1]=c[3][X]*c[2][Y]; t[2]=c[1][X]*c[3][Y]; t[3]=c[3][X]*c[1][Y]; t[4]=c[1][X]*c[2][Y];  t[5]=c[2][X]*c[1][Y];
  t[6]=c[0][X]*c[3][Y]; t[7]=c[3][X]*c[0][Y]; t[8]=c[0][X]*c[2][Y]; t[9]=c[2][X]*c[0][Y]; t[10]=c[0][X]*c[1][Y]; t[11]=c[1][X]*c[0][Y];
  m[2][X]  = t[0]*c[1][W] + t[3]*c[2][W] + t[4]*c[3][W];   m[2][X] -= t[1]*c[1][W] + t[2]*c[2][W] + t[5]*c[3][W];
  m[2][Y]  = t[1]*c[0][W] + t[6]*c[2][W] + t[9]*c[3][W];   m[2][Y] -= t[0]*c[0][W] + t[7]*c[2][W] + t[8]*c[3][W];
  m[2][Z]  = t[2]*c[0][W] + t[7]*c[1][W] + t[10]*c[3][W];  m[2][Z] -= t[3]*c[0][W] + t[6]*c[1][W] + t[11]*c[3][W];
  m[2][W]  = t[5]*c[0][W] + t[8]*c[1][W] + t[11]*c[2][W];  m[2][W] -= t[4]*c[0][W] + t[9]*c[1][W] + t[10]*c[2][W];
  m[3][X]  = t[2]*c[2][Z] + t[5]*c[3][Z] + t[1]*c[1][Z];   m[3][X] -= t[4]*c[3][Z] + t[0]*c[1][Z] + t[3]*c[2][Z];
  m[3][Y]  = t[8]*c[3][Z] + t[0]*c[0][Z] + t[7]*c[2][Z];   m[3][Y] -= t[6]*c[2][Z] + t[9]*c[3][Z] + t[1]*c[0][Z];
  m[3][Z]  = t[6]*c[1][Z] + t[11]*c[3][Z] + t[3]*c[0][Z];  m[3][Z] -= t[10]*c[3][Z] + t[2]*c[0][Z] + t[7]*c[1][Z];
  m[3][W]  = t[10]*c[2][Z] + t[4]*c[0][Z] + t[9]*c[1][Z];  m[3][W] -= t[8]*c[1][Z] + t[11]*c[2][Z] + t[5]*c[0][Z];
  const float k = 1.0f/(c[0][X]*m[0][X] + c[1][X]*m[0][Y] + c[2][X]*m[0][Z] + c[3][X]*m[0][W]);
  for (int col = 0; col < 4; col++) for (int row = 0; row < 4; row++) outCols[col*4 + row] = m[col][row]*k;
  #undef X
  #undef Y
  #undef Z
  #undef W
}

struct Builder
{
  std::vector<Mesh>     meshes;
  std::vector<Instance> insts;
  std::vector<Node32>   nodes;
  std::vector<F4>       tris;
  std::vector<float>    invMatrices;
  Box  sceneBox;
  int  maxStack = 0;
  bool committed = false;

  size_t Alloc4()
  {
    const size_t at = nodes.size();
    for (int i = 0; i < 4; i++) nodes.push_back(InvalidNode());
    return at;
  }

  // triangle leaf -> float4 offset of the header
  uint32_t EmitLeaf(const Mesh& M, int meshId, int32_t first, int32_t count)
  {
    const size_t hdr = tris.size();
    tris.push_back(F4{ 0, 0, 0, 0 });
    for (int32_t i = first; i < first + count; i++)
    {
      const int32_t t = M.tree.prims[i].id;
      const F4 A = M.vert[M.idx[3*t + 0]], B = M.vert[M.idx[3*t + 1]], C = M.vert[M.idx[3*t + 2]];
      tris.push_back(F4{ A.x, A.y, A.z, AsFloat(t) });
      tris.push_back(F4{ B.x, B.y, B.z, AsFloat(meshId) });
      tris.push_back(F4{ C.x, C.y, C.z, AsFloat(-1) });
    }
    F4 h; h.x = AsFloat(int32_t(hdr + 1)); h.y = AsFloat(count); h.z = AsFloat(-1); h.w = AsFloat(-1);
    tris[hdr] = h;
    return uint32_t(hdr);
  }

  // flatten the children of wide node `n` of mesh M into a fresh quad; returns the quad index
  uint32_t EmitMeshQuad(const Mesh& M, int meshId, int32_t n)
  {
    const size_t q = Alloc4();
    const BNode& N = M.tree.nodes[n];
    for (int i = 0; i < 4; i++)
    {
      if (N.child[i] < 0) continue;
      const BNode& C = M.tree.nodes[N.child[i]];
      Node32 out = InvalidNode();
      SetBox(out, C.box);
      out.escapeIndex = 0;
      if (C.leaf) out.leftOffsetAndLeaf = 0x80000000u | (EmitLeaf(M, meshId, C.first, C.count) & 0x7fffffffu);
      else        out.leftOffsetAndLeaf = EmitMeshQuad(M, meshId, N.child[i]) & 0x7fffffffu;
      nodes[q + i] = out;
    }
    return uint32_t(q/4);
  }

  uint32_t MeshSubtreeRef(int meshId)
  {
    Mesh& M = meshes[meshId];
    if (M.subtreeRef >= 0) return uint32_t(M.subtreeRef);
    const BNode& R = M.tree.nodes[0];
    uint32_t ref;
    if (R.leaf) ref = 0x80000000u | (EmitLeaf(M, meshId, R.first, R.count) & 0x7fffffffu);
    else        ref = EmitMeshQuad(M, meshId, 0) & 0x7fffffffu;
    M.subtreeRef = ref;
    return ref;
  }

  uint32_t EmitInstanceRecord(int instId)
  {
    const Instance& I = insts[instId];
    const uint32_t sub = MeshSubtreeRef(I.meshId);   // may grow `nodes`: take it before allocating the record
    const size_t q = Alloc4();
    Node32 n0 = InvalidNode(); SetBox(n0, meshes[I.meshId].box); n0.leftOffsetAndLeaf = sub; n0.escapeIndex = 0;
    nodes[q] = n0;
    std::memcpy(&nodes[q + 1], I.inv, 64);                      // float4 idx 8Q+2 .. 8Q+5 = inverse matrix columns
    int32_t rec[8] = { I.realId >= 0 ? I.realId : instId, I.meshId, 0, 0, 0, 0, 0, 0 };
    std::memcpy(&nodes[q + 3], rec, 32);                        // float4 idx 8Q+6 = int4{instId, meshId, 0, 0}
    return uint32_t(q/4);
  }

  uint32_t EmitTopQuad(const Tree& T, int32_t n)
  {
    const size_t q = Alloc4();
    const BNode& N = T.nodes[n];
    for (int i = 0; i < 4; i++)
    {
      if (N.child[i] < 0) continue;
      const BNode& C = T.nodes[N.child[i]];
      Node32 out = InvalidNode();
      SetBox(out, C.box);
      if (C.leaf) { out.leftOffsetAndLeaf = 0x80000000u | (EmitInstanceRecord(T.prims[C.first].id) & 0x7fffffffu); out.escapeIndex = 1; }
      else        { out.leftOffsetAndLeaf = EmitTopQuad(T, N.child[i]) & 0x7fffffffu; out.escapeIndex = 0; }
      nodes[q + i] = out;
    }
    return uint32_t(q/4);
  }

  int Commit()
  {
    nodes.clear(); tris.clear(); invMatrices.clear(); sceneBox = Box();
    if (insts.empty()) return HC_E_STATE;

    int meshDepth = 0;
    for (size_t mi = 0; mi < meshes.size(); mi++)
    {
      Mesh& M = meshes[mi];
      M.subtreeRef = -1;
      bool used = false;
      for (const Instance& I : insts) if (I.meshId == int(mi)) { used = true; break; }
      if (!used) continue;                       // meshes of the other tree (one builder per tree, mesh ids shared: they are the geomId of the triangles)
      if (M.built) { meshDepth = std::max(meshDepth, M.tree.depth); continue; }
      const int32_t nt = int32_t(M.idx.size()/3);
      M.tree.prims.clear(); M.tree.prims.reserve(nt);
      M.box = Box();
      for (int32_t t = 0; t < nt; t++)
      {
        const F4 A = M.vert[M.idx[3*t + 0]], B = M.vert[M.idx[3*t + 1]], C = M.vert[M.idx[3*t + 2]];
        // drop zero-area triangles like the reference converter (bvh_access_dll2.cpp:354-355)
        const float e1[3] = { B.x - A.x, B.y - A.y, B.z - A.z }, e2[3] = { C.x - A.x, C.y - A.y, C.z - A.z };
        const float cx = e1[1]*e2[2] - e1[2]*e2[1], cy = e1[2]*e2[0] - e1[0]*e2[2], cz = e1[0]*e2[1] - e1[1]*e2[0];
        if (!(0.5f*std::sqrt(cx*cx + cy*cy + cz*cz) > 0.0f)) continue;
        Prim p; p.id = t;
        p.box.grow(V3{ A.x, A.y, A.z }); p.box.grow(V3{ B.x, B.y, B.z }); p.box.grow(V3{ C.x, C.y, C.z });
        p.c = V3{ 0.5f*(p.box.lo.x + p.box.hi.x), 0.5f*(p.box.lo.y + p.box.hi.y), 0.5f*(p.box.lo.z + p.box.hi.z) };
        M.tree.prims.push_back(p); M.box.grow(p.box);
      }
      if (M.tree.prims.empty()) return HC_E_ARG;
      BuildTree(M.tree, kMaxLeafEnv, 17);        // mesh levels 17 + top levels 9 = 26 -> 3 x 26 + 2 = the 80 entries of the traversal stack
      M.built = true;
      meshDepth = std::max(meshDepth, M.tree.depth);
    }

    Tree top; top.prims.reserve(insts.size());
    for (size_t i = 0; i < insts.size(); i++)
    {
      Instance& I = insts[i];
      Inverse4x4Columns(I.m, I.inv);
      const Box& mb = meshes[I.meshId].box;
      Box wb;
      for (int k = 0; k < 8; k++)
      {
        const float px = (k & 1) ? mb.hi.x : mb.lo.x, py = (k & 2) ? mb.hi.y : mb.lo.y, pz = (k & 4) ? mb.hi.z : mb.lo.z;
        wb.grow(V3{ I.m[0]*px + I.m[1]*py + I.m[2]*pz + I.m[3], I.m[4]*px + I.m[5]*py + I.m[6]*pz + I.m[7], I.m[8]*px + I.m[9]*py + I.m[10]*pz + I.m[11] });
      }
      // the world box is rounded independently of the object-space test the ray will get inside the instance: pad by a few ulps
      const float ex = std::fmax(std::fmax(std::fabs(wb.lo.x), std::fabs(wb.hi.x)), std::fmax(std::fmax(std::fabs(wb.lo.y), std::fabs(wb.hi.y)), std::fmax(std::fabs(wb.lo.z), std::fabs(wb.hi.z))));
      const float pad = 4.0f*std::numeric_limits<float>::epsilon()*std::fmax(ex, 1e-30f);
      wb.lo = V3{ wb.lo.x - pad, wb.lo.y - pad, wb.lo.z - pad }; wb.hi = V3{ wb.hi.x + pad, wb.hi.y + pad, wb.hi.z + pad };
      I.worldBox = wb;
      Prim p; p.id = int32_t(i); p.box = wb;
      p.c = V3{ 0.5f*(wb.lo.x + wb.hi.x), 0.5f*(wb.lo.y + wb.hi.y), 0.5f*(wb.lo.z + wb.hi.z) };
      top.prims.push_back(p); sceneBox.grow(wb);
      invMatrices.insert(invMatrices.end(), I.inv, I.inv + 16);
    }
    BuildTree(top, 1, 9);

    // quad 0 : root record (bvh_access_dll2.cpp:637-644)
    const size_t q0 = Alloc4();
    Node32 root = InvalidNode(); SetBox(root, sceneBox); root.leftOffsetAndLeaf = 1; root.escapeIndex = 0;
    nodes[q0] = root;
    const float ident[16] = { 1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1 };
    std::memcpy(&nodes[q0 + 1], ident, 64);

    if (top.nodes[0].leaf)            // single instance: quad 1 holds just that instance leaf
    {
      const size_t q1 = Alloc4();
      const uint32_t rec = EmitInstanceRecord(top.prims[top.nodes[0].first].id);
      Node32 n = InvalidNode(); SetBox(n, top.nodes[0].box); n.leftOffsetAndLeaf = 0x80000000u | rec; n.escapeIndex = 1;
      nodes[q1] = n;
    }
    else
      EmitTopQuad(top, 0);            // allocates quad 1 first, as traversal expects

    // each visited node pushes at most 3 entries and pops one: bound of the traversal stack
    maxStack = 3*(top.depth + 1) + 3*(meshDepth + 1) + 2;
    committed = true;
    return HC_OK;
  }
};
} // namespace hcb

struct hc_bvh { hcb::Builder b; };

extern void hc_set_error(const char* msg);   // hc_api.cu

extern "C"
{
int hc_bvh_create(hc_bvh** out)
{
  if (!out) return HC_E_ARG;
  *out = new hc_bvh;
  return HC_OK;
}
void hc_bvh_destroy(hc_bvh* b) { delete b; }

int hc_bvh_add_mesh(hc_bvh* b, const float* vert4f, int numVert, const int32_t* indices, int numIndices, int* outMeshId)
{
  if (!b || !vert4f || !indices || numVert <= 0 || numIndices < 3 || numIndices % 3 != 0) return HC_E_ARG;
  for (int i = 0; i < numIndices; i++) if (indices[i] < 0 || indices[i] >= numVert) return HC_E_RANGE;
  hcb::Mesh m;
  m.vert.resize(numVert); std::memcpy(m.vert.data(), vert4f, size_t(numVert)*16);
  m.idx.assign(indices, indices + numIndices);
  b->b.meshes.push_back(std::move(m));
  b->b.committed = false;
  if (outMeshId) *outMeshId = int(b->b.meshes.size()) - 1;
  return HC_OK;
}

int hc_bvh_add_instance(hc_bvh* b, int meshId, const float* matrixRowMajor16, int* outInstId)
{
  if (!b || !matrixRowMajor16 || meshId < 0 || meshId >= int(b->b.meshes.size())) return HC_E_ARG;
  hcb::Instance I; I.meshId = meshId; std::memcpy(I.m, matrixRowMajor16, 64);
  b->b.insts.push_back(I);
  b->b.committed = false;
  if (outInstId) *outInstId = int(b->b.insts.size()) - 1;
  return HC_OK;
}

int hc_bvh_add_instance_id(hc_bvh* b, int meshId, const float* matrixRowMajor16, int realInstId)
{
  if (!b || realInstId < 0) return HC_E_ARG;
  const int rc = hc_bvh_add_instance(b, meshId, matrixRowMajor16, nullptr);
  if (rc == HC_OK) b->b.insts.back().realId = realInstId;
  return rc;
}

int hc_bvh_commit(hc_bvh* b) { return b ? b->b.Commit() : HC_E_ARG; }

int hc_bvh_result(hc_bvh* b, const void** nodes, int* nodesNum, const void** trif4, int* trif4Num,
                  const float** invMatrices16, int* numInst, int* maxStackDepth)
{
  if (!b || !b->b.committed) return HC_E_STATE;
  if (nodes) *nodes = b->b.nodes.data();
  if (nodesNum) *nodesNum = int(b->b.nodes.size());
  if (trif4) *trif4 = b->b.tris.data();
  if (trif4Num) *trif4Num = int(b->b.tris.size());
  if (invMatrices16) *invMatrices16 = b->b.invMatrices.data();
  if (numInst) *numInst = int(b->b.insts.size());
  if (maxStackDepth) *maxStackDepth = b->b.maxStack;
  return HC_OK;
}

int hc_bvh_bounds(hc_bvh* b, float bmin[3], float bmax[3])
{
  if (!b || !b->b.committed) return HC_E_STATE;
  bmin[0] = b->b.sceneBox.lo.x; bmin[1] = b->b.sceneBox.lo.y; bmin[2] = b->b.sceneBox.lo.z;
  bmax[0] = b->b.sceneBox.hi.x; bmax[1] = b->b.sceneBox.hi.y; bmax[2] = b->b.sceneBox.hi.z;
  return HC_OK;
}
} // extern "C"
