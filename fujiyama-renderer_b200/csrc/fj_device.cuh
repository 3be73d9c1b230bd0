// fj_device.cuh — device-side data layout, exact FP64 math, RNG, traversal and shading of the
// B200-native Fujiyama hot path.  Everything here is __device__ code shared by the kernels in
// fj_gpu.cu.  Citations are file:line of tsubo164/Fujiyama-Renderer @ a451548.
//
// Exactness policy.  The reference takes every geometric DECISION in IEEE FP64 without FMA
// contraction (g++ -O3, baseline x86-64).  Wherever a decision or a value feeding one is computed
// (sample positions, camera rays, instance transforms, Moller-Trumbore, hit attributes, light
// geometry, filter weights) this file uses __dmul_rn/__dadd_rn/__dsub_rn/__ddiv_rn/__dsqrt_rn in the
// reference's operation order: those intrinsics are never fused, so the result is bit-identical to the
// CPU.  Bounding-box culling is NOT a decision of the reference (both of its accelerators return the
// closest hit) and runs on padded FP32 boxes.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#define FJ_MAX_SHADING_GROUPS 8
#define FJ_STACK 64              // binary-tree traversal stack entries (BLAS depth + TLAS depth + 1 sentinel)
#define FJ_STACK4 96             // 4-wide traversal: up to 3 pushes per level
#define FJ_PENDING 32            // megakernel: pending secondary rays per path (DFS of the reflect/refract/diffuse tree)
#define FJ_REAL_MAX DBL_MAX

namespace fj {

// ------------------------------------------------------------------------------------------ exact FP64
struct D3 { double x, y, z; };
__device__ __forceinline__ D3 mk(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ D3 operator+(const D3 &a, const D3 &b) { return mk(dadd(a.x, b.x), dadd(a.y, b.y), dadd(a.z, b.z)); }
__device__ __forceinline__ D3 operator-(const D3 &a, const D3 &b) { return mk(dsub(a.x, b.x), dsub(a.y, b.y), dsub(a.z, b.z)); }
__device__ __forceinline__ D3 operator*(const D3 &a, double s) { return mk(dmul(a.x, s), dmul(a.y, s), dmul(a.z, s)); }   // fj_vector.h:293-299
__device__ __forceinline__ D3 operator*(double s, const D3 &a) { return a * s; }
__device__ __forceinline__ double dot(const D3 &a, const D3 &b) { return dadd(dadd(dmul(a.x, b.x), dmul(a.y, b.y)), dmul(a.z, b.z)); }  // :318-324
__device__ __forceinline__ D3 cross(const D3 &a, const D3 &b) {                                                         // :326-332
  return mk(dsub(dmul(a.y, b.z), dmul(a.z, b.y)), dsub(dmul(a.z, b.x), dmul(a.x, b.z)), dsub(dmul(a.x, b.y), dmul(a.y, b.x)));
}
__device__ __forceinline__ double length(const D3 &a) { return __dsqrt_rn(dot(a, a)); }
__device__ __forceinline__ D3 normalize(const D3 &a) {            // fj_vector.h:339-345: a * (1/len)
  const double len = length(a);
  if (len == 0) return a;
  return a * ddiv(1., len);
}
// MatTransformPoint / MatTransformVector, src/fj_matrix.cc:209-223 (m = rows 0..2 of the 4x4, 12 doubles)
__device__ __forceinline__ D3 mat_point(const double *m, const D3 &p) {
  return mk(dadd(dadd(dadd(dmul(m[0], p.x), dmul(m[1], p.y)), dmul(m[2], p.z)), m[3]),
            dadd(dadd(dadd(dmul(m[4], p.x), dmul(m[5], p.y)), dmul(m[6], p.z)), m[7]),
            dadd(dadd(dadd(dmul(m[8], p.x), dmul(m[9], p.y)), dmul(m[10], p.z)), m[11]));
}
__device__ __forceinline__ D3 mat_vector(const double *m, const D3 &v) {
  return mk(dadd(dadd(dmul(m[0], v.x), dmul(m[1], v.y)), dmul(m[2], v.z)),
            dadd(dadd(dmul(m[4], v.x), dmul(m[5], v.y)), dmul(m[6], v.z)),
            dadd(dadd(dmul(m[8], v.x), dmul(m[9], v.y)), dmul(m[10], v.z)));
}

// ------------------------------------------------------------------------------------------ scene in HBM
struct DMesh {
  const float4 *nodes;      // BLAS, 4 x float4 per 64-B node (fj_bvh.h)
  const float4 *tri32;      // 3 x float4 per triangle in leaf order: (v0,prim_id) (v1,-) (v2,-); exact when P is FP32-representable
  const double *tri64;      // 10 doubles per triangle (v0 v1 v2, prim_id as bits) when P is not FP32-representable
  const double *N;          // vertex normals, 3 doubles per vertex (may be null)
  const int32_t *idx;       // 3 vertex indices per face (original order)
  const int32_t *group;     // shading group per face (may be null -> 0)
  int32_t top_count;        // leading BFS-ordered nodes (reserved for shared-memory staging)
  int32_t log2_tris;        // ceil(log2(triangle count)): root-to-leaf path length of the algorithmic-bytes model
  float bmag;               // max |coordinate| of the mesh bounds: scales the FP32 slab-test error bound (extend kernel)
  float pad1;
  const float4 *nodes4;     // the same tree as 4-wide 128-B nodes (fj_bvh.h Node128) for the wavefront's k_extend
  const float4 *nodesq;     // the 4-wide tree with 8-bit quantised child boxes (fj_bvh.h NodeQ64), null if not representable
  float bmagq, pad2;        // bound magnitude of the decoded planes
  const float *uv;          // per-vertex texture coordinates, 2 floats per vertex (null: uv = 0, fj_mesh.cc:292-297)
  const double *P;          // vertex positions by vertex index (FP64 as given), kept only for meshes with uv: dPdu / dPdv of bump maps
  const double *tri64v;     // meshes with per-vertex velocity (Mesh::velocity_, src/fj_mesh.h:210): 20 doubles per triangle in leaf
                            // order — v0 v1 v2, prim_id as bits, velocity0 velocity1 velocity2, pad; tri32 and tri64 are null then
  const double *vel;        // the velocities by vertex index (only kept next to P, for dPdu / dPdv)
};
struct DInstance {
  double inv[12];           // rows 0..2 of MatInverse(matrix): world -> object (fj_object_instance.cc:222-225)
  double fwd[12];           // object -> world
  int32_t mesh;
  int32_t shader_of_group[FJ_MAX_SHADING_GROUPS];
  int32_t reflect_target, refract_target, shadow_target;
  const double *motion;     // time-sampled transform (motion blur): 24 doubles per entry of the frame's time table, inv[12] then
                            // fwd[12], evaluated by the caller with the reference's own transform code (fjgpu.h); null = static
};
// The matrices of an instance for a ray whose sample drew entry `tidx` of the time table
// (XfmLerpTransformSample per ray, src/fj_object_instance.cc:219-220).
__device__ __forceinline__ const double *inst_inv(const DInstance &in, uint32_t tidx) { return in.motion ? in.motion + 24 * (size_t)tidx : in.inv; }
__device__ __forceinline__ const double *inst_fwd(const DInstance &in, uint32_t tidx) { return in.motion ? in.motion + 24 * (size_t)tidx + 12 : in.fwd; }
// Everything k_extend2 needs to enter an instance, in TLAS-leaf order (leaf `first` indexes this array directly): one
// 144-B record instead of the chain order[] -> DInstance -> DMesh (the kernel is bound by dependent-load latency).
struct DInstRec {
  double inv[12];           // world -> object
  const char *nodes4, *nodesq;
  const void *tri;          // tri32 or tri64 packets
  float bmag, bmagq;
  int32_t tri64, inst;      // packet format (0 = tri32, 1 = tri64, 2 = tri64v: moving triangles), instance index (DScene::inst)
  const double *motion;     // DInstance::motion
};
static_assert(sizeof(DInstRec) == 144, "instance record must be 144 bytes");
struct DGroup { const float4 *nodes; const int32_t *order; int32_t ninst; float bmag; const float4 *nodes4; const float4 *nodesq; float bmagq, pad2; const DInstRec *irec; };   // TLAS leaf (first,count) -> order[first..] = instance indices
struct DShader {
  int32_t kind, do_reflect, do_color_filter, texture;     // texture: 1 + index into DScene::textures, 0 = none
  float diffuse[3], reflect[3], refract[3], emission[3], transmit[3];
  float ior, opacity;
  int32_t bump_texture; float bump_amplitude;
};
struct DLight {
  int32_t kind, sample_count, double_sided, dome_count;
  float color[3], intensity;
  double translate[3];
  double fwd[12];
  const double *dome_dirs; const float *dome_colors;
};
// One `.mip` texture in HBM: the file's tiles in file order (src/fj_mipmap.cc:156-180).
struct DTexture { const float *tiles; int32_t width, height, nch, tilesize, xnt, ynt; };
struct DScene {
  const DTexture *textures;
  const DMesh *meshes; const DInstance *inst; const DGroup *groups; const DShader *shaders; const DLight *lights;
  int32_t nmeshes, ninst, ngroups, nshaders, nlights, pad;
  const double *time_tab;   // the frame's time table (fjgpu_time_table over the shutter): time VALUE of entry k, read only for
                            // meshes with vertex velocity (rays carry the entry's index)
};
struct DCamera { double fwd[12]; double uvx, uvy, znear, zfar; const double *motion; };   // motion: fwd[12] per time-table entry, null = static   // uv_size_ computed on the host (fj_camera.cc:97-101)
struct DFrame {
  int32_t xres, yres, xrate, yrate;
  int32_t mx, my;                 // margin samples, count_samples_in_margin (fj_fixed_grid_sampler.cc:131-136)
  double xfw, yfw, jitter, udelta, vdelta;
  int32_t max_diffuse, max_reflect, max_refract, cast_shadow;
  int32_t target_group; uint32_t seed; int32_t flags; int32_t max_ns;   // max_ns = sample slots per tile in the sample buffer
  const uint32_t *jitter_tab;     // first 2*max_ns draws of a default-seeded XorShift (fj_random.cc:10-43)
};
struct DTile { int32_t id, xmin, ymin, xmax, ymax; };
struct DCounters { unsigned long long rays[5]; unsigned long long samples; unsigned long long hits; unsigned long long levels; unsigned long long node_steps; unsigned long long tri_tests; unsigned long long leaf_phases, leaf_rounds; };

struct Hit { double t, u, v; int32_t prim, inst; };

// ------------------------------------------------------------------------------------------ RNG
// Philox-4x32-10 (Salmon et al., SC'11), keyed exactly like oracle/fj_oracle.cc ctr_rand so that the
// stochastic shaders can be checked sample-for-sample.
// One Philox block = the four draws of dimensions 4 (dim >> 2) .. 4 (dim >> 2) + 3.
__device__ __forceinline__ void ctr_block(uint32_t seed, uint32_t tile, uint32_t sample, unsigned long long node, uint32_t block, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)node, c1 = (uint32_t)(node >> 32), c2 = block, c3 = 0x46554a49u;
  uint32_t k0 = seed ^ (tile * 0x9E3779B1u), k1 = sample;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double ctr_rand(uint32_t seed, uint32_t tile, uint32_t sample, unsigned long long node, uint32_t dim) {
  uint32_t c[4];
  ctr_block(seed, tile, sample, node, dim >> 2, c);
  const uint32_t pick = (dim & 3) == 0 ? c[0] : ((dim & 3) == 1 ? c[1] : ((dim & 3) == 2 ? c[2] : c[3]));
  return ddiv((double)pick, 4294967295.0);     // XorShift::NextFloat01 mapping (fj_random.cc:40-43)
}
// The draws of dimensions `dim` and `dim + 1`, dim even: both lie in one block — one Philox evaluation instead of two
// (the same values as two ctr_rand calls).
__device__ __forceinline__ void ctr_rand2(uint32_t seed, uint32_t tile, uint32_t sample, unsigned long long node, uint32_t dim, double *a, double *b) {
  uint32_t c[4];
  ctr_block(seed, tile, sample, node, dim >> 2, c);
  const bool hi = (dim & 2u) != 0;
  *a = ddiv((double)(hi ? c[2] : c[0]), 4294967295.0);
  *b = ddiv((double)(hi ? c[3] : c[1]), 4294967295.0);
}

// ------------------------------------------------------------------------------------------ traversal
struct RayD { D3 o, d; double tmin, tmax; uint32_t tidx = 0; };   // tidx: entry of the frame's time table (TraceContext::time)

// TriRayIntersect, non-culling branch — src/fj_triangle.cc:81-153 (:127-151), EPSILON :12
__device__ __forceinline__ bool tri_intersect(const D3 &v0, const D3 &v1, const D3 &v2, const D3 &orig, const D3 &dir,
                                              double *t, double *u, double *v) {
  const D3 e1 = v1 - v0, e2 = v2 - v0;
  const D3 pvec = cross(dir, e2);
  const double det = dot(e1, pvec);
  if (det > -1e-6 && det < 1e-6) return false;
  const double inv_det = ddiv(1.0, det);
  const D3 tvec = orig - v0;
  *u = dmul(dot(tvec, pvec), inv_det);
  if (*u < 0.0 || *u > 1.0) return false;
  const D3 qvec = cross(tvec, e1);
  *v = dmul(dot(dir, qvec), inv_det);
  if (*v < 0.0 || dadd(*u, *v) > 1.0) return false;
  *t = dmul(dot(e2, qvec), inv_det);
  return true;
}

// Slab test of the two children of a node.  BOXF = float or double arithmetic.  The boxes were rounded
// outward and padded on the host; the FP32 variant additionally widens the interval by the rounding
// error of the FP32 copy of the ray so it stays conservative w.r.t. the FP64 ray.
template <typename T> struct BoxRay {
  T ox, oy, oz, ix, iy, iz;      // origin, 1/dir
  T ex, ey, ez;                  // |t| slack per axis (FP32 only): eps_o * |1/d|
};
template <typename T> __device__ __forceinline__ T tmin_(T a, T b);
template <typename T> __device__ __forceinline__ T tmax_(T a, T b);
template <> __device__ __forceinline__ float tmin_<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ float tmax_<float>(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ double tmin_<double>(double a, double b) { return fmin(a, b); }
template <> __device__ __forceinline__ double tmax_<double>(double a, double b) { return fmax(a, b); }

template <typename T>
__device__ __forceinline__ void make_box_ray(const D3 &o, const D3 &d, BoxRay<T> &r) {
  r.ox = (T)o.x; r.oy = (T)o.y; r.oz = (T)o.z;
  r.ix = (T)1 / (T)d.x; r.iy = (T)1 / (T)d.y; r.iz = (T)1 / (T)d.z;
  if (sizeof(T) == 4) {
    // |o32 - o64| <= 2^-24 |o|; subtraction and product add ~3 ulp relative (handled at the compare)
    const float k = 1.2e-7f;
    r.ex = (T)(k * fabsf((float)o.x) * fabsf((float)r.ix));
    r.ey = (T)(k * fabsf((float)o.y) * fabsf((float)r.iy));
    r.ez = (T)(k * fabsf((float)o.z) * fabsf((float)r.iz));
    if (!(r.ex == r.ex)) r.ex = 0; if (!(r.ey == r.ey)) r.ey = 0; if (!(r.ez == r.ez)) r.ez = 0;   // 0*inf
  } else { r.ex = r.ey = r.ez = 0; }
}

template <typename T>
__device__ __forceinline__ void slab2(const float4 &n0, const float4 &n1, const float4 &n2, const BoxRay<T> &r, T tmin, T tmax,
                                      T &near0, T &far0, T &near1, T &far1) {
  // child0: x,y in n0 (lo.x hi.x lo.y hi.y), z in n2.xy ; child1: x,y in n1, z in n2.zw
  T a, b;
  a = ((T)n0.x - r.ox) * r.ix; b = ((T)n0.y - r.ox) * r.ix; T nx0 = tmin_(a, b), fx0 = tmax_(a, b);
  a = ((T)n0.z - r.oy) * r.iy; b = ((T)n0.w - r.oy) * r.iy; T ny0 = tmin_(a, b), fy0 = tmax_(a, b);
  a = ((T)n2.x - r.oz) * r.iz; b = ((T)n2.y - r.oz) * r.iz; T nz0 = tmin_(a, b), fz0 = tmax_(a, b);
  a = ((T)n1.x - r.ox) * r.ix; b = ((T)n1.y - r.ox) * r.ix; T nx1 = tmin_(a, b), fx1 = tmax_(a, b);
  a = ((T)n1.z - r.oy) * r.iy; b = ((T)n1.w - r.oy) * r.iy; T ny1 = tmin_(a, b), fy1 = tmax_(a, b);
  a = ((T)n2.z - r.oz) * r.iz; b = ((T)n2.w - r.oz) * r.iz; T nz1 = tmin_(a, b), fz1 = tmax_(a, b);
  if (sizeof(T) == 4) {
    nx0 -= r.ex; fx0 += r.ex; ny0 -= r.ey; fy0 += r.ey; nz0 -= r.ez; fz0 += r.ez;
    nx1 -= r.ex; fx1 += r.ex; ny1 -= r.ey; fy1 += r.ey; nz1 -= r.ez; fz1 += r.ez;
  }
  near0 = tmax_(tmax_(nx0, ny0), tmax_(nz0, tmin)); far0 = tmin_(tmin_(fx0, fy0), tmin_(fz0, tmax));
  near1 = tmax_(tmax_(nx1, ny1), tmax_(nz1, tmin)); far1 = tmin_(tmin_(fx1, fy1), tmin_(fz1, tmax));
  if (sizeof(T) == 4) {   // relative slack for the FP32 subtract/multiply roundings
    far0 = far0 + fabsf((float)far0) * 1e-6f; near0 = near0 - fabsf((float)near0) * 1e-6f;
    far1 = far1 + fabsf((float)far1) * 1e-6f; near1 = near1 - fabsf((float)near1) * 1e-6f;
  }
}

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// Closest hit of a world-space ray in object group `g` (two-level BVH).  Semantics of
// Accelerator::Intersect on the group's surface accelerator (src/fj_shading.cc:538-541): the hit with
// the smallest t in [tmin, tmax] over all instances; the ray is taken to object space with the
// instance's inverse matrix WITHOUT renormalising dir, so t is shared (fj_object_instance.cc:221-225).
// Exact ties in t resolve to the lower instance and, inside one mesh, to the HIGHER face index — the order the
// reference's grid cell lists are walked in (new entries become the list head, src/fj_grid_accelerator.cc:128-133,
// strict `<` at :262) — so the result does not depend on the traversal order.
template <typename T>
__device__ __noinline__ bool trace_closest(const DScene &sc, int g, const RayD &ray, Hit *hit) {
  int stack[FJ_STACK];
  int sp = 0;
  double best_t = ray.tmax;       // inclusive upper bound while nothing is hit
  bool found = false;
  int best_prim = -1, best_inst = -1; double best_u = 0, best_v = 0;

  const DGroup grp = sc.groups[g];
  const float4 *nodes = grp.nodes;
  bool in_blas = false;
  int cur_inst = -1;
  D3 o = ray.o, d = ray.d;
  BoxRay<T> br; make_box_ray<T>(o, d, br);
  const DMesh *mesh = nullptr;
  int node = 0;
  const int SENTINEL = (int)0x80000000;   // negative like a leaf, decoded before leaves

  for (;;) {
    // ---- inner nodes: test both children, descend into the nearer, push the farther
    while (node >= 0) {
      const float4 *np = nodes + 4 * (size_t)node;
      const float4 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3];
      T near0, far0, near1, far1;
      slab2<T>(n0, n1, n2, br, (T)ray.tmin, (T)best_t, near0, far0, near1, far1);
      const bool h0 = near0 <= far0, h1 = near1 <= far1;
      const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
      if (h0 && h1) {
        const bool first0 = near0 <= near1;
        stack[sp++] = first0 ? c1 : c0;
        node = first0 ? c0 : c1;
      } else if (h0) node = c0;
      else if (h1) node = c1;
      else {
        if (sp == 0) goto done;
        node = stack[--sp];
      }
    }
    if (node == SENTINEL) {       // all of the instance's BLAS is done: back to world space
      in_blas = false; nodes = grp.nodes; o = ray.o; d = ray.d; make_box_ray<T>(o, d, br);
      if (sp == 0) goto done;
      node = stack[--sp];
      continue;
    }
    {
      const int ref = ~node;
      const int first = ref >> 3, count = (ref & 7) + 1;
      if (!in_blas) {
        // TLAS leaf: enter the instances one at a time (the others go back on the stack as 1-leaves)
        for (int k = count - 1; k >= 1; k--) stack[sp++] = ~(((first + k) << 3) | 0);
        cur_inst = grp.order[first];
        const DInstance &in = sc.inst[cur_inst];
        o = mat_point(inst_inv(in, ray.tidx), ray.o);
        d = mat_vector(inst_inv(in, ray.tidx), ray.d);
        make_box_ray<T>(o, d, br);
        mesh = &sc.meshes[in.mesh];
        nodes = mesh->nodes;
        in_blas = true;
        stack[sp++] = SENTINEL;
        node = 0;
        continue;
      }
      // BLAS leaf: exact FP64 triangle tests
      for (int k = 0; k < count; k++) {
        D3 v0, v1, v2; int prim;
        if (mesh->tri32) {
          const float4 *tp = mesh->tri32 + 3 * (size_t)(first + k);
          const float4 a = ldg4(tp), b = ldg4(tp + 1), c = ldg4(tp + 2);
          v0 = mk(a.x, a.y, a.z); v1 = mk(b.x, b.y, b.z); v2 = mk(c.x, c.y, c.z); prim = __float_as_int(a.w);
        } else if (mesh->tri64v) {                                       // `P0 += time * velocity0`, src/fj_mesh.cc:252-259
          const double *p = mesh->tri64v + 20 * (size_t)(first + k);
          const double tm = sc.time_tab[ray.tidx];
          v0 = mk(p[0], p[1], p[2]) + tm * mk(p[10], p[11], p[12]); v1 = mk(p[3], p[4], p[5]) + tm * mk(p[13], p[14], p[15]);
          v2 = mk(p[6], p[7], p[8]) + tm * mk(p[16], p[17], p[18]); prim = (int)__double_as_longlong(p[9]);
        } else {
          const double *p = mesh->tri64 + 10 * (size_t)(first + k);
          v0 = mk(p[0], p[1], p[2]); v1 = mk(p[3], p[4], p[5]); v2 = mk(p[6], p[7], p[8]); prim = (int)__double_as_longlong(p[9]);
        }
        double t, u, v;
        if (!tri_intersect(v0, v1, v2, o, d, &t, &u, &v)) continue;
        if (!(ray.tmin <= t && t <= ray.tmax)) continue;              // RayInRange, src/fj_ray.h:29-32
        const bool better = !found ? true : (t < best_t || (t == best_t && (cur_inst < best_inst || (cur_inst == best_inst && prim > best_prim))));
        if (better) { found = true; best_t = t; best_u = u; best_v = v; best_prim = prim; best_inst = cur_inst; }
      }
      if (sp == 0) goto done;
      node = stack[--sp];
    }
  }
done:
  hit->t = found ? best_t : FJ_REAL_MAX; hit->u = best_u; hit->v = best_v; hit->prim = best_prim; hit->inst = best_inst;
  return found;
}

// ------------------------------------------------------------------------------------------ shading
struct C3 { float r, g, b; };
__device__ __forceinline__ C3 c3(float r, float g, float b) { C3 c; c.r = r; c.g = g; c.b = b; return c; }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float luminance(const float *c) {       // fj_color.h:275-278 (double expression -> float)
  return (float)dadd(dadd(dmul(.298912, (double)c[0]), dmul(.586611, (double)c[1])), dmul(.114478, (double)c[2]));
}

__device__ __forceinline__ D3 sl_faceforward(const D3 &I, const D3 &N) { if (dot(I, N) < 0) return N; return mk(-N.x, -N.y, -N.z); }  // fj_shading.cc:42-52
__device__ __forceinline__ double sl_fresnel(const D3 &I, const D3 &N, double ior) {                  // :54-74
  double eta, c = dmul(-1., dot(I, N));
  if (c > 0) eta = ior; else { eta = ddiv(1., ior); c = dmul(c, -1.); }
  const double a = dsub(1., eta), b = dadd(1., eta);
  const double F0 = ddiv(dadd(dmul(a, a), .0), dadd(dmul(b, b), .0));
  return dadd(F0, dmul(dsub(1., F0), pow(dsub(1., c), 5.)));
}
__device__ __forceinline__ D3 sl_reflect(const D3 &I, const D3 &N) {                                   // :92-100
  const double c = dmul(-1., dot(I, N));
  return mk(dadd(I.x, dmul(dmul(2., c), N.x)), dadd(I.y, dmul(dmul(2., c), N.y)), dadd(I.z, dmul(dmul(2., c), N.z)));
}
__device__ __forceinline__ D3 sl_refract(const D3 &I, const D3 &N, double ior) {                       // :102-138
  D3 n; double eta, cos1 = dmul(-1., dot(I, N));
  if (cos1 < 0) { cos1 = dmul(cos1, -1.); eta = ddiv(1., ior); n = mk(-N.x, -N.y, -N.z); } else { eta = ior; n = N; }
  const double radicand = dsub(1., dmul(dmul(eta, eta), dsub(1., dmul(cos1, cos1))));
  if (radicand < 0.) return sl_reflect(I, N);
  const double ncoeff = dsub(dmul(eta, cos1), __dsqrt_rn(radicand));
  return mk(dadd(dmul(eta, I.x), dmul(ncoeff, n.x)), dadd(dmul(eta, I.y), dmul(ncoeff, n.y)), dadd(dmul(eta, I.z), dmul(ncoeff, n.z)));
}

enum { RAY_CAMERA = 0, RAY_SHADOW = 1, RAY_DIFFUSE = 2, RAY_REFLECT = 3, RAY_REFRACT = 4,      // enum RayContext, fj_shading.h:18-24
       RAY_SHADOW_HEAD = 5,     // not a ray: the header record of a block of shadow rays in the wavefront's queue (fj_kernels.cuh)
       RAY_DEAD = 6 };          // not a ray: filler of a queue chunk a warp reserved and did not use up (QueueSink, fj_kernels.cuh)

// One ray of the wavefront: an entry of the ray queues in HBM (and of the megakernel's per-path DFS stack).
// `thr` is the throughput that multiplies whatever the ray returns — the reference multiplies after the recursive
// SlTrace returns, and every use is linear in the child's radiance.  112 B = 7 x 16 B.
struct RayRec {
  double o[3], d[3];
  double tmin, tmax;
  float thr[3];
  uint32_t slot;                  // sample slot in the batch's accumulator buffer
  unsigned long long node;        // path-tree code keying the counter RNG
  int32_t target;                 // object group traced
  uint8_t type, dd, rd, fd;       // ray context and the three depth counters of TraceContext (fj_shading.h:26-47)
  int32_t filter_shader;          // >= 0: refracted child whose radiance is scaled by pow(transmit, t_hit) of that shader
  uint32_t key;                   // sort key of the ray (direction octant | Morton code of the origin cell), see k_shade; in
                                  // frames with time-sampled transforms (sorting off): the ray's entry of the time table
  int32_t pad2, pad3;
};
static_assert(sizeof(RayRec) == 112, "ray record must be 112 bytes");
struct HitRec { double t, u, v; int32_t prim, inst; };     // 32 B, inst < 0 = miss
static_assert(sizeof(HitRec) == 32, "hit record must be 32 bytes");
// Per-sample radiance accumulator: 32.32 fixed point so that the sum is independent of the order in which the rays of
// one sample are shaded (deterministic frames for any scheduling); alpha is written once by the camera ray.
struct Accum { long long r, g, b; float a; uint32_t pad; };
static_assert(sizeof(Accum) == 32, "accumulator must be 32 bytes");
#define FJ_FIX_SCALE 4294967296.0
#define FJ_FIX_INV 2.3283064365386963e-10
__device__ __forceinline__ long long to_fix(float v) {
  double d = dmul((double)v, FJ_FIX_SCALE);
  d = fmin(fmax(d, -4.0e18), 4.0e18);
  return __double2ll_rn(d);
}
__device__ __forceinline__ float from_fix(long long v) { return (float)dmul((double)v, FJ_FIX_INV); }

// Surface at a hit: world-space P and N exactly as ObjectInstance::RayIntersect returns them
// (fj_object_instance.cc:234-241) from Mesh::ray_intersect (fj_mesh.cc:277-305).
__device__ __forceinline__ void hit_surface(const DScene &sc, const RayD &ray, const Hit &h, D3 *P, D3 *N, int *shader_slot) {
  const DInstance &in = sc.inst[h.inst];
  const DMesh &m = sc.meshes[in.mesh];
  const double *inv = inst_inv(in, ray.tidx), *fwd = inst_fwd(in, ray.tidx);
  const D3 o = mat_point(inv, ray.o), d = mat_vector(inv, ray.d);
  const D3 Pobj = o + h.t * d;                                   // RayPointAt, fj_ray.h:24-27
  D3 Nobj = mk(0, 0, 0);
  const int i0 = m.idx[3 * (size_t)h.prim], i1 = m.idx[3 * (size_t)h.prim + 1], i2 = m.idx[3 * (size_t)h.prim + 2];
  if (m.N) {
    const D3 N0 = mk(m.N[3 * (size_t)i0], m.N[3 * (size_t)i0 + 1], m.N[3 * (size_t)i0 + 2]);
    const D3 N1 = mk(m.N[3 * (size_t)i1], m.N[3 * (size_t)i1 + 1], m.N[3 * (size_t)i1 + 2]);
    const D3 N2 = mk(m.N[3 * (size_t)i2], m.N[3 * (size_t)i2 + 1], m.N[3 * (size_t)i2 + 2]);
    Nobj = (dsub(dsub(1., h.u), h.v) * N0 + h.u * N1) + h.v * N2;   // TriComputeNormal, fj_triangle.cc:44-49
  }
  *P = mat_point(fwd, Pobj);
  *N = normalize(mat_vector(fwd, Nobj));
  int gid = m.group ? m.group[h.prim] : 0;                         // ObjectInstance::GetShader, fj_object_instance.cc:177-191
  int slot = (gid < 0 || gid >= FJ_MAX_SHADING_GROUPS) ? in.shader_of_group[0] : in.shader_of_group[gid];
  if (slot < 0) slot = in.shader_of_group[0];
  *shader_slot = slot;
}

// Texture coordinates of a hit: UV = (1-u-v) UV0 + u UV1 + v UV2 with the reference's mixed precision (`const float t =
// 1 - u - v`, float * float for the first term, double for the others, one rounding to float at the end;
// Mesh::ray_intersect, src/fj_mesh.cc:280-291).  Instances do not transform uv.
__device__ __forceinline__ void hit_uv(const DScene &sc, const Hit &h, float *tu, float *tv) {
  const DMesh &m = sc.meshes[sc.inst[h.inst].mesh];
  if (!m.uv) { *tu = 0.f; *tv = 0.f; return; }
  const int i0 = m.idx[3 * (size_t)h.prim], i1 = m.idx[3 * (size_t)h.prim + 1], i2 = m.idx[3 * (size_t)h.prim + 2];
  const float t = (float)dsub(dsub(1., h.u), h.v);
  const float2 a = reinterpret_cast<const float2 *>(m.uv)[i0], b = reinterpret_cast<const float2 *>(m.uv)[i1], c = reinterpret_cast<const float2 *>(m.uv)[i2];
  *tu = (float)dadd(dadd((double)fmul(t, a.x), dmul(h.u, (double)b.x)), dmul(h.v, (double)c.x));
  *tv = (float)dadd(dadd((double)fmul(t, a.y), dmul(h.u, (double)b.y)), dmul(h.v, (double)c.y));
}

// dPdu, dPdv of a hit in world space: TriComputeDerivatives (src/fj_triangle.cc:51-74: float uv differences and determinant,
// `const float invdet = 1. / determinant`, FP64 edges) inside Mesh::ray_intersect (fj_mesh.cc:287-290), then the instance's
// forward matrix (XfmTransformVector, fj_object_instance.cc:237-238).  Zero without uv or with a degenerate uv triangle.
__device__ __forceinline__ void hit_derivatives(const DScene &sc, const Hit &h, uint32_t tidx, D3 *dPdu, D3 *dPdv) {
  const DInstance &in = sc.inst[h.inst];
  const DMesh &m = sc.meshes[in.mesh];
  *dPdu = mk(0, 0, 0); *dPdv = mk(0, 0, 0);
  if (!m.uv || !m.P) return;
  const int i0 = m.idx[3 * (size_t)h.prim], i1 = m.idx[3 * (size_t)h.prim + 1], i2 = m.idx[3 * (size_t)h.prim + 2];
  D3 P0 = mk(m.P[3 * (size_t)i0], m.P[3 * (size_t)i0 + 1], m.P[3 * (size_t)i0 + 2]);
  D3 P1 = mk(m.P[3 * (size_t)i1], m.P[3 * (size_t)i1 + 1], m.P[3 * (size_t)i1 + 2]);
  D3 P2 = mk(m.P[3 * (size_t)i2], m.P[3 * (size_t)i2 + 1], m.P[3 * (size_t)i2 + 2]);
  if (m.vel) {                  // TriComputeDerivatives sees the vertices where the ray's time put them (fj_mesh.cc:252-259,287-290)
    const double tm = sc.time_tab[tidx];
    P0 = P0 + tm * mk(m.vel[3 * (size_t)i0], m.vel[3 * (size_t)i0 + 1], m.vel[3 * (size_t)i0 + 2]);
    P1 = P1 + tm * mk(m.vel[3 * (size_t)i1], m.vel[3 * (size_t)i1 + 1], m.vel[3 * (size_t)i1 + 2]);
    P2 = P2 + tm * mk(m.vel[3 * (size_t)i2], m.vel[3 * (size_t)i2 + 1], m.vel[3 * (size_t)i2 + 2]);
  }
  const float2 t0 = reinterpret_cast<const float2 *>(m.uv)[i0], t1 = reinterpret_cast<const float2 *>(m.uv)[i1], t2 = reinterpret_cast<const float2 *>(m.uv)[i2];
  const D3 dP1 = P1 - P0, dP2 = P2 - P0;
  const float du1 = __fsub_rn(t1.x, t0.x), du2 = __fsub_rn(t2.x, t0.x), dv1 = __fsub_rn(t1.y, t0.y), dv2 = __fsub_rn(t2.y, t0.y);
  const float det = __fsub_rn(fmul(du1, dv2), fmul(dv1, du2));
  if (det == 0.f) return;
  const float invdet = (float)ddiv(1., (double)det);
  const D3 a = ((double)dv2 * dP1 - (double)dv1 * dP2) * (double)invdet;
  const D3 b = ((double)(-du2) * dP1 + (double)du1 * dP2) * (double)invdet;
  *dPdu = mat_vector(inst_fwd(in, tidx), a); *dPdv = mat_vector(inst_fwd(in, tidx), b);
}

// TextureCache::LookupTexture, src/fj_texture.cc:51-78: wrap to [0,1), flip v, tile = floor(coordinate * tile count) clamped
// as MipInput::ReadTile does (src/fj_mipmap.cc:163-165), texel = (int)(fraction * 64) inside the tile; all in float as the
// reference compiles it.  Colour as FrameBuffer::GetColor (src/fj_framebuffer.cc:84-101).
__device__ __forceinline__ float4 tex_lookup(const DTexture &tx, float u, float v) {
  const float tsu = __fsub_rn(u, floorf(u)), tsv = __fsub_rn(v, floorf(v));
  const float tlu = fmul(tsu, (float)tx.xnt), tlv = fmul(__fsub_rn(1.f, tsv), (float)tx.ynt);
  const int xtile = (int)floorf(tlu), ytile = (int)floorf(tlv);
  const int xpxl = (int)fmul(__fsub_rn(tlu, floorf(tlu)), 64.f), ypxl = (int)fmul(__fsub_rn(tlv, floorf(tlv)), 64.f);
  const int tx_ = min(max(xtile, 0), tx.xnt - 1), ty_ = min(max(ytile, 0), tx.ynt - 1);
  const float *px = tx.tiles + ((size_t)(ty_ * tx.xnt + tx_) * tx.tilesize * tx.tilesize + (size_t)ypxl * tx.tilesize + xpxl) * tx.nch;
  if (tx.nch == 1) return make_float4(px[0], px[0], px[0], 1.f);
  if (tx.nch == 3) return make_float4(px[0], px[1], px[2], 1.f);
  if (tx.nch == 4) return make_float4(px[0], px[1], px[2], px[3]);
  return make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ float luminance4(const float4 &c) {       // Luminance4, fj_color.h:280-283
  return (float)dadd(dadd(dmul(.298912, (double)c.x), dmul(.586611, (double)c.y)), dmul(.114478, (double)c.z));
}
// SlBumpMapping, src/fj_shading.cc:418-465 (both cross products are scaled by du, as the reference does)
__device__ __forceinline__ D3 sl_bump_mapping(const DTexture &tx, const D3 &dPdu, const D3 &dPdv, float tu, float tv, double amplitude, const D3 &N) {
  if (tx.width == 0 || tx.height == 0) return N;
  const float du = (float)ddiv(1., (double)tx.width), dv = (float)ddiv(1., (double)tx.height);
  float val0 = luminance4(tex_lookup(tx, __fsub_rn(tu, du), tv)), val1 = luminance4(tex_lookup(tx, fadd(tu, du), tv));
  const float Bu = __fdiv_rn(__fsub_rn(val0, val1), fmul(2.f, du));
  val0 = luminance4(tex_lookup(tx, tu, __fsub_rn(tv, dv))); val1 = luminance4(tex_lookup(tx, tu, fadd(tv, dv)));
  const float Bv = __fdiv_rn(__fsub_rn(val0, val1), fmul(2.f, dv));
  const D3 NdPdu = cross(N, dPdu) * (double)du, NdPdv = cross(N, dPdv) * (double)du;
  const D3 nb = mk(dadd(N.x, dmul(amplitude, dsub(dmul((double)Bv, NdPdu.x), dmul((double)Bu, NdPdv.x)))),
                   dadd(N.y, dmul(amplitude, dsub(dmul((double)Bv, NdPdu.y), dmul((double)Bu, NdPdv.y)))),
                   dadd(N.z, dmul(amplitude, dsub(dmul((double)Bv, NdPdu.z), dmul((double)Bu, NdPdv.z)))));
  return normalize(nb);
}

// Opacity the shader of an occluder returns to a shadow ray (Os of evaluate(); the colour is unused and
// shadow contexts cannot spawn rays: fj_shading.cc:266-279,338-355).
__device__ __forceinline__ float occluder_opacity(const DScene &sc, const Hit &h) {
  const DInstance &in = sc.inst[h.inst];
  const DMesh &m = sc.meshes[in.mesh];
  int gid = m.group ? m.group[h.prim] : 0;
  int slot = (gid < 0 || gid >= FJ_MAX_SHADING_GROUPS) ? in.shader_of_group[0] : in.shader_of_group[gid];
  if (slot < 0) slot = in.shader_of_group[0];
  if (slot < 0) return 1.f;
  const DShader &s = sc.shaders[slot];
  float os = (s.kind == 2) ? s.opacity : 1.f;
  return fminf(fmaxf(os, 0.f), 1.f);
}

}  // namespace fj
