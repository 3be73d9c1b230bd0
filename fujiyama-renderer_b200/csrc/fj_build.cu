// fj_build.cu — linear BVH construction on the device (see fj_build.h).
//
//   k_vertex_flags      are all vertex coordinates FP32-representable (tri32 packets are then exact)?
//   k_face_bounds       exact FP64 bounds of the referenced vertices (ordered-integer atomics)
//   k_prims             padded FP32 box per triangle (the host builder's pad_box rule) + 63-bit Morton code of its centre
//   cub radix sort      (code, triangle)
//   k_radix_tree        binary radix tree over the sorted codes (Karras 2012; equal codes are split by index)
//   k_fit               bottom-up boxes: the second thread to reach a node merges its children; the same pass takes the SAH
//                       decision 'one leaf or keep the split' for every subtree of <= max_leaf triangles (cost 1 per node
//                       step, leaf_cost per triangle test — the host builder's constants)
//   k_binary_depth      depth of the binary tree the kernels will walk (subtrees of <= max_leaf triangles become leaves)
//   k_emit_binary       Node64 array (index = radix-tree node, root 0)
//   k_collapse_level    level-synchronous collapse to 4-wide Node128 nodes (largest-area child opened first), carrying the
//                       worst-case stack depth of a nearest-first walk
//   k_finish_wide       NodeQ64 (8-bit quantised, fj_quant.h) copy of every wide node
//   k_tris              triangle packets in leaf (= sorted) order
#include "fj_build.h"
#include "fj_bvh.h"
#include "fj_quant.h"

#include <cub/cub.cuh>
#include <algorithm>
#include <vector>

namespace {

using fjb::Node64; using fjb::Node128; using fjb::NodeQ64;

struct PBox { float lo[3], hi[3]; };

struct Ctl {
  unsigned long long bmin[3], bmax[3];     // ordered-integer encodings of doubles
  int not_f32;                             // some vertex coordinate is not FP32-representable
  int kept;                                // binary nodes with more than max_leaf triangles
  int depth2;                              // binary depth
  int nwide, next_count, depth4, stack_need;
  unsigned bmagq_bits; int quant_fail;
};

__device__ __forceinline__ unsigned long long enc(double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
inline double dec_host(unsigned long long e) {
  const unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
  double d; memcpy(&d, &b, 8); return d;
}

__global__ void k_init(Ctl *c) {
  for (int a = 0; a < 3; a++) { c->bmin[a] = ~0ull; c->bmax[a] = 0ull; }
  c->not_f32 = 0; c->kept = 0; c->depth2 = 0; c->nwide = 1; c->next_count = 0; c->depth4 = 0; c->stack_need = 0; c->bmagq_bits = 0; c->quant_fail = 0;
}

__global__ void k_vertex_flags(const double *P, size_t n3, Ctl *c) {
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (size_t)gridDim.x * blockDim.x) bad = bad || ((double)(float)P[i] != P[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(&c->not_f32, 1);
}

__global__ void k_face_bounds(const double *P, const int32_t *idx, int n, Ctl *c) {
  double lo[3] = {1.7976931348623157e308, 1.7976931348623157e308, 1.7976931348623157e308}, hi[3] = {-1.7976931348623157e308, -1.7976931348623157e308, -1.7976931348623157e308};
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x)
    for (int v = 0; v < 3; v++) {
      const double *p = P + 3 * (size_t)idx[3 * (size_t)f + v];
      for (int a = 0; a < 3; a++) { lo[a] = fmin(lo[a], p[a]); hi[a] = fmax(hi[a], p[a]); }
    }
  for (int a = 0; a < 3; a++) {
    for (int o = 16; o > 0; o >>= 1) { lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(&c->bmin[a], enc(lo[a])); atomicMax(&c->bmax[a], enc(hi[a])); }
  }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {      // 21 bits -> every third bit
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

__global__ void k_prims(const double *P, const int32_t *idx, int n, const double3 bmin, const double3 inv_ext, PBox *boxes, unsigned long long *keys, int32_t *vals) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const double *p0 = P + 3 * (size_t)idx[3 * (size_t)f], *p1 = P + 3 * (size_t)idx[3 * (size_t)f + 1], *p2 = P + 3 * (size_t)idx[3 * (size_t)f + 2];
  PBox b; float c[3];
  for (int a = 0; a < 3; a++) {
    const double lo = fmin(p0[a], fmin(p1[a], p2[a])), hi = fmax(p0[a], fmax(p1[a], p2[a]));
    // pad_box of fj_gpu.cu: rounded outward and padded by a few ulps
    const double m = fmax(fabs(lo), fabs(hi));
    const double pad = __dadd_rn(__dmul_rn(4e-7, m), 1e-30);
    b.lo[a] = __double2float_rd(__dsub_rn(lo, pad));
    b.hi[a] = __double2float_ru(__dadd_rn(hi, pad));
    c[a] = .5f * b.lo[a] + .5f * b.hi[a];
  }
  boxes[f] = b;
  const double bm[3] = {bmin.x, bmin.y, bmin.z}, ie[3] = {inv_ext.x, inv_ext.y, inv_ext.z};
  unsigned long long q[3];
  for (int a = 0; a < 3; a++) {
    double t = ((double)c[a] - bm[a]) * ie[a] * 2097152.0;
    t = fmin(fmax(t, 0.0), 2097151.0);
    q[a] = (unsigned long long)t;
  }
  keys[f] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
  vals[f] = f;
}

// Tree arrays.  Children: >= 0 internal node, < 0 leaf ~k (k = position in sorted order).
struct Tree {
  const unsigned long long *keys; const int32_t *vals; const PBox *pbox;
  int32_t *left, *right, *parent, *lparent, *first, *last, *flag; PBox *nbox;
  float *cost; int32_t *leafy;             // SAH cost of the subtree; 1 = the subtree is emitted as one leaf
  int n, max_leaf; float leaf_cost;
};

__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll((long long)(a ^ b));
}

__global__ void k_radix_tree(Tree t, Ctl *c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, n = t.n;
  if (i >= n - 1) return;
  const unsigned long long *keys = t.keys;
  const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int s = lmax / 2; s >= 1; s /= 2) if (delta(keys, n, i, i + (l + s) * d) > dmin) l += s;
  const int j = i + l * d;
  const int dnode = delta(keys, n, i, j);
  int s = 0, tt = l;
  do { tt = (tt + 1) / 2; if (delta(keys, n, i, i + (s + tt) * d) > dnode) s += tt; } while (tt > 1);
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const int lc = lo == gamma ? ~gamma : gamma, rc = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
  t.left[i] = lc; t.right[i] = rc; t.first[i] = lo; t.last[i] = hi; t.flag[i] = 0;
  if (lc >= 0) t.parent[lc] = i; else t.lparent[~lc] = i;
  if (rc >= 0) t.parent[rc] = i; else t.lparent[~rc] = i;
  if (i == 0) t.parent[0] = -1;
}

__device__ __forceinline__ PBox load_box(const PBox *p) {
  PBox b; const float *s = reinterpret_cast<const float *>(p);
  for (int a = 0; a < 3; a++) { b.lo[a] = __ldcg(s + a); b.hi[a] = __ldcg(s + 3 + a); }
  return b;
}
__device__ __forceinline__ PBox child_box(const Tree &t, int c) { return c < 0 ? load_box(t.pbox + t.vals[~c]) : load_box(t.nbox + c); }

__device__ __forceinline__ float area(const PBox &b) {
  const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  return (dx < 0 || dy < 0 || dz < 0) ? 0.f : 2.f * (dx * dy + dy * dz + dz * dx);
}

__global__ void k_fit(Tree t, Ctl *c) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= t.n) return;
  int cur = t.lparent[k];
  while (cur >= 0) {
    if (atomicAdd(&t.flag[cur], 1) == 0) return;            // the first child to arrive leaves; the second merges
    __threadfence();
    const int lc = t.left[cur], rc = t.right[cur];
    const PBox a = child_box(t, lc), b = child_box(t, rc);
    PBox u;
    for (int x = 0; x < 3; x++) { u.lo[x] = fminf(a.lo[x], b.lo[x]); u.hi[x] = fmaxf(a.hi[x], b.hi[x]); }
    t.nbox[cur] = u;
    // SAH: cost of walking this subtree as it is vs. as one leaf
    const float ca = lc < 0 ? t.leaf_cost * area(a) : __ldcg(t.cost + lc), cb = rc < 0 ? t.leaf_cost * area(b) : __ldcg(t.cost + rc);
    const float au = area(u);
    float cst = au + ca + cb;
    const int cnt = t.last[cur] - t.first[cur] + 1;
    int leafy = 0;
    if (cnt <= t.max_leaf && t.leaf_cost * cnt * au <= cst) { cst = t.leaf_cost * cnt * au; leafy = 1; }
    t.cost[cur] = cst; t.leafy[cur] = leafy;
    if (!leafy) atomicAdd(&c->kept, 1);
    __threadfence();
    cur = t.parent[cur];
  }
}

__device__ __forceinline__ bool is_inner(const Tree &t, int c) { return c >= 0 && !t.leafy[c]; }
__device__ __forceinline__ int32_t eff_ref(const Tree &t, int c) {      // child reference of the emitted trees
  if (c < 0) return ~((~c << 3) | 0);
  const int cnt = t.last[c] - t.first[c] + 1;
  return !t.leafy[c] ? c : ~((t.first[c] << 3) | (cnt - 1));
}

__global__ void k_binary_depth(Tree t, Ctl *c) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= t.n) return;
  int depth = 0;                      // inner nodes above the topmost collapsed ancestor
  for (int cur = t.lparent[k]; cur >= 0; cur = t.parent[cur]) depth = t.leafy[cur] ? 0 : depth + 1;
  for (int o = 16; o > 0; o >>= 1) depth = max(depth, __shfl_xor_sync(0xffffffffu, depth, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&c->depth2, depth);
}

__global__ void k_emit_binary(Tree t, Node64 *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.n - 1) return;
  Node64 nd; memset(&nd, 0, sizeof nd);
  if (is_inner(t, i)) {
    const int lc = t.left[i], rc = t.right[i];
    const PBox a = child_box(t, lc), b = child_box(t, rc);
    nd.f[0] = a.lo[0]; nd.f[1] = a.hi[0]; nd.f[2] = a.lo[1]; nd.f[3] = a.hi[1]; nd.f[8] = a.lo[2]; nd.f[9] = a.hi[2];
    nd.f[4] = b.lo[0]; nd.f[5] = b.hi[0]; nd.f[6] = b.lo[1]; nd.f[7] = b.hi[1]; nd.f[10] = b.lo[2]; nd.f[11] = b.hi[2];
    nd.c[0] = eff_ref(t, lc); nd.c[1] = eff_ref(t, rc);
  }
  out[i] = nd;
}

struct Front { int32_t bnode, widx, acc, depth; };

__global__ void k_collapse_level(Tree t, const Front *cur, int ncur, Front *next, Node128 *out, Ctl *c) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ncur) return;
  const Front fr = cur[e];
  int slots[4]; int ns = 0;
  slots[ns++] = t.left[fr.bnode]; slots[ns++] = t.right[fr.bnode];
  while (ns < 4) {            // open the inner child with the largest surface area
    int best = -1; float best_area = -1.f;
    for (int k = 0; k < ns; k++) {
      if (!is_inner(t, slots[k])) continue;
      const float ar = area(load_box(t.nbox + slots[k]));
      if (ar > best_area) { best_area = ar; best = k; }
    }
    if (best < 0) break;
    const int b = slots[best];
    slots[best] = t.left[b]; slots[ns++] = t.right[b];
  }
  Node128 w;
  for (int k = 0; k < 4; k++) { w.lox[k] = w.loy[k] = w.loz[k] = 3e38f; w.hix[k] = w.hiy[k] = w.hiz[k] = 3e38f; w.c[k] = ~0; w.pad[k] = 0; }
  int ninner = 0;
  for (int k = 0; k < ns; k++) ninner += is_inner(t, slots[k]) ? 1 : 0;
  const int wbase = ninner ? atomicAdd(&c->nwide, ninner) : 0;
  const int nbase = ninner ? atomicAdd(&c->next_count, ninner) : 0;
  const int acc = fr.acc + ns - 1;
  int ki = 0;
  for (int k = 0; k < ns; k++) {
    const PBox b = child_box(t, slots[k]);
    w.lox[k] = b.lo[0]; w.hix[k] = b.hi[0]; w.loy[k] = b.lo[1]; w.hiy[k] = b.hi[1]; w.loz[k] = b.lo[2]; w.hiz[k] = b.hi[2];
    if (is_inner(t, slots[k])) {
      w.c[k] = wbase + ki;
      Front nf; nf.bnode = slots[k]; nf.widx = wbase + ki; nf.acc = acc; nf.depth = fr.depth + 1;
      next[nbase + ki] = nf;
      ki++;
    } else w.c[k] = eff_ref(t, slots[k]);
  }
  out[fr.widx] = w;
  atomicMax(&c->stack_need, acc);
  atomicMax(&c->depth4, fr.depth);
}

__global__ void k_finish_wide(const Node128 *in, int n, NodeQ64 *qq, Ctl *c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Node128 w = in[i];
  NodeQ64 q; double mag = 0;
  if (!fjb::quantize_one(w, q, &mag)) { atomicOr(&c->quant_fail, 1); memset(&q, 0, sizeof q); }
  qq[i] = q;
  const float mf = __double2float_ru(mag * (1.0 + 1e-6));
  atomicMax(&c->bmagq_bits, __float_as_uint(mf));           // non-negative floats order like their bit patterns
}

__global__ void k_tris32(const double *P, const int32_t *idx, const int32_t *vals, int n, float4 *out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int f = vals[k];
  for (int v = 0; v < 3; v++) {
    const double *p = P + 3 * (size_t)idx[3 * (size_t)f + v];
    out[3 * (size_t)k + v] = make_float4((float)p[0], (float)p[1], (float)p[2], v == 0 ? __int_as_float(f) : 0.f);
  }
}
__global__ void k_tris64(const double *P, const int32_t *idx, const int32_t *vals, int n, double *out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int f = vals[k];
  for (int v = 0; v < 3; v++) {
    const double *p = P + 3 * (size_t)idx[3 * (size_t)f + v];
    for (int a = 0; a < 3; a++) out[10 * (size_t)k + 3 * v + a] = p[a];
  }
  out[10 * (size_t)k + 9] = __longlong_as_double((long long)f);
}

struct Scratch {
  std::vector<void *> ptrs;
  ~Scratch() { for (void *p : ptrs) cudaFree(p); }
  template <typename T> T *get(size_t count, cudaError_t *e) {
    void *p = nullptr;
    if (*e == cudaSuccess) *e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (p) ptrs.push_back(p);
    return (T *)p;
  }
};

}  // namespace

int fj_device_build(cudaStream_t st, const double *dP, int32_t nverts, const int32_t *didx, int32_t n, int max_leaf, float leaf_cost,
                    bool force_tri64, FjDeviceBuild *out, std::string *err) {
  auto fail = [&](const std::string &m) {
    if (err) *err = m;
    for (void **q : {&out->nodes, &out->nodes4, &out->nodesq, &out->tri}) { if (*q) cudaFree(*q); *q = nullptr; }
    return -1;
  };
  if (n < 2) return fail("device build needs at least two triangles");
  max_leaf = std::min(std::max(max_leaf, 1), 8);
  cudaError_t e = cudaSuccess;
  cudaEvent_t ev0, ev1; cudaEventCreate(&ev0); cudaEventCreate(&ev1);
  cudaEventRecord(ev0, st);
  Scratch S;
  Ctl *ctl = S.get<Ctl>(1, &e);
  PBox *pbox = S.get<PBox>(n, &e), *nbox = S.get<PBox>(n, &e);
  unsigned long long *keys0 = S.get<unsigned long long>(n, &e), *keys1 = S.get<unsigned long long>(n, &e);
  int32_t *vals0 = S.get<int32_t>(n, &e), *vals1 = S.get<int32_t>(n, &e);
  int32_t *left = S.get<int32_t>(n, &e), *right = S.get<int32_t>(n, &e), *parent = S.get<int32_t>(n, &e), *lparent = S.get<int32_t>(n, &e);
  int32_t *first = S.get<int32_t>(n, &e), *last = S.get<int32_t>(n, &e), *flag = S.get<int32_t>(n, &e), *leafy = S.get<int32_t>(n, &e);
  float *cost = S.get<float>(n, &e);
  if (e != cudaSuccess) return fail(std::string("device build scratch: ") + cudaGetErrorString(e));
  const int T = 256, G = (n + T - 1) / T;
  k_init<<<1, 1, 0, st>>>(ctl);
  k_vertex_flags<<<std::min<size_t>(((size_t)3 * nverts + T - 1) / T, 4096), T, 0, st>>>(dP, (size_t)3 * nverts, ctl);
  k_face_bounds<<<std::min(G, 4096), T, 0, st>>>(dP, didx, n, ctl);
  Ctl h;
  cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st);
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(std::string("device build (bounds): ") + cudaGetErrorString(e));
  double3 bmin, iext;
  double *bm = &bmin.x, *ie = &iext.x;
  for (int a = 0; a < 3; a++) {
    out->bmin[a] = dec_host(h.bmin[a]); out->bmax[a] = dec_host(h.bmax[a]);
    bm[a] = out->bmin[a];
    const double ext = out->bmax[a] - out->bmin[a];
    ie[a] = ext > 0 ? 1.0 / ext : 0.0;
  }
  out->tri64 = (force_tri64 || h.not_f32) ? 1 : 0;
  k_prims<<<G, T, 0, st>>>(dP, didx, n, bmin, iext, pbox, keys0, vals0);
  size_t temp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys0, keys1, vals0, vals1, n, 0, 63, st);
  void *temp = S.get<char>(temp_bytes, &e);
  if (e != cudaSuccess) return fail(std::string("device build sort scratch: ") + cudaGetErrorString(e));
  cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys0, keys1, vals0, vals1, n, 0, 63, st);
  Tree t;
  t.keys = keys1; t.vals = vals1; t.pbox = pbox; t.left = left; t.right = right; t.parent = parent; t.lparent = lparent;
  t.first = first; t.last = last; t.flag = flag; t.nbox = nbox; t.n = n; t.max_leaf = max_leaf;
  t.cost = cost; t.leafy = leafy; t.leaf_cost = leaf_cost;
  k_radix_tree<<<G, T, 0, st>>>(t, ctl);
  k_fit<<<G, T, 0, st>>>(t, ctl);
  k_binary_depth<<<G, T, 0, st>>>(t, ctl);
  cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st);
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(std::string("device build (tree): ") + cudaGetErrorString(e));
  int root_leafy = 1;
  cudaMemcpy(&root_leafy, leafy, 4, cudaMemcpyDeviceToHost);
  if (h.kept < 1 || root_leafy) return fail("device build: mesh fits one leaf");
  // outputs
  auto alloc = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, std::max<size_t>(bytes, 16)); };
  out->nnodes = n - 1; out->max_depth = h.depth2 + 1;
  out->nodes_bytes = (size_t)(n - 1) * sizeof(Node64);
  const size_t wide_cap = (size_t)h.kept;                    // every wide node is rooted at a distinct kept binary node
  Node128 *wide = nullptr; Front *f0 = nullptr, *f1 = nullptr;
  alloc(&out->nodes, out->nodes_bytes);
  alloc((void **)&wide, wide_cap * sizeof(Node128));
  f0 = S.get<Front>(wide_cap, &e); f1 = S.get<Front>(wide_cap, &e);
  out->tri_bytes = out->tri64 ? (size_t)n * 80 : (size_t)n * 48;
  alloc(&out->tri, out->tri_bytes);
  if (e != cudaSuccess) { cudaFree(wide); return fail(std::string("device build outputs: ") + cudaGetErrorString(e)); }
  k_emit_binary<<<G, T, 0, st>>>(t, (Node64 *)out->nodes);
  if (out->tri64) k_tris64<<<G, T, 0, st>>>(dP, didx, vals1, n, (double *)out->tri);
  else k_tris32<<<G, T, 0, st>>>(dP, didx, vals1, n, (float4 *)out->tri);
  // level-synchronous collapse to 4-wide nodes
  Front root; root.bnode = 0; root.widx = 0; root.acc = 0; root.depth = 1;
  cudaMemcpyAsync(f0, &root, sizeof root, cudaMemcpyHostToDevice, st);
  int ncur = 1;
  for (int level = 0; ncur > 0 && level < 4096; level++) {
    cudaMemsetAsync((char *)ctl + offsetof(Ctl, next_count), 0, 4, st);
    k_collapse_level<<<(ncur + 127) / 128, 128, 0, st>>>(t, f0, ncur, f1, wide, ctl);
    cudaMemcpyAsync(&ncur, (char *)ctl + offsetof(Ctl, next_count), 4, cudaMemcpyDeviceToHost, st);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { cudaFree(wide); return fail(std::string("device build (collapse): ") + cudaGetErrorString(e)); }
    std::swap(f0, f1);
  }
  cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  const int nw = h.nwide;
  out->nnodes4 = nw; out->max_depth4 = h.depth4; out->stack_need4 = h.stack_need;
  out->nodes4_bytes = (size_t)nw * sizeof(Node128); out->nodesq_bytes = (size_t)nw * sizeof(NodeQ64);
  alloc(&out->nodes4, out->nodes4_bytes); alloc(&out->nodesq, out->nodesq_bytes);
  if (e != cudaSuccess) { cudaFree(wide); return fail(std::string("device build outputs: ") + cudaGetErrorString(e)); }
  cudaMemcpyAsync(out->nodes4, wide, out->nodes4_bytes, cudaMemcpyDeviceToDevice, st);
  k_finish_wide<<<(nw + T - 1) / T, T, 0, st>>>((const Node128 *)out->nodes4, nw, (NodeQ64 *)out->nodesq, ctl);
  PBox rootbox;
  cudaMemcpyAsync(&rootbox, nbox, sizeof rootbox, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(&h, ctl, sizeof h, cudaMemcpyDeviceToHost, st);
  cudaEventRecord(ev1, st);
  e = cudaStreamSynchronize(st);
  cudaFree(wide);
  if (e != cudaSuccess) return fail(std::string("device build (finish): ") + cudaGetErrorString(e));
  double b = 0;
  for (int a = 0; a < 3; a++) b = std::max(b, std::max(std::fabs((double)rootbox.lo[a]), std::fabs((double)rootbox.hi[a])));
  out->bmag = fjb::round_up(b);
  memcpy(&out->bmagq, &h.bmagq_bits, 4);
  out->quant_ok = h.quant_fail ? 0 : 1;
  float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1);
  out->seconds = ms * 1e-3;
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
  return 0;
}
