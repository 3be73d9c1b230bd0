// fj_extend.cuh — k_extend2: the closest-hit kernel of the wavefront with the exact (FP64) half of every ray's state
// and the traversal stack in shared memory.
//
// Same algorithm and results as k_extend (fj_kernels.cuh): persistent threads, one ray per lane, warp-synchronous
// phases (refill / 4-wide node steps with speculative leaf parking / exact FP64 triangle tests / transitions).
// What changed is where the state lives.  k_extend keeps everything in registers (112 unconstrained, 96 with spills
// at 5 CTAs/SM) and is bound by latency x occupancy: 5 warps per scheduler cannot cover the dependent-issue latency
// of the node step, let alone the node fetch.  Here the node loop only carries the FP32 box ray, the
// stack pointer and the node reference; object-space origin/direction, tmin, the best hit and the triangle pointer
// (84 B per lane) sit in shared memory, SoA so that a warp's accesses are conflict-free, and are touched only by the
// triangle and transition phases.
//
// Round 2: the traversal stack left local memory.  The round-1 kernel kept `int stack[96]` per lane in local memory; ncu
// counted 150 M local loads + 217 M local stores per launch against 533 M global loads, on an L1 data pipe at 65 % — every
// lane of a warp sits at a different stack depth, so one STL/LDL touches up to 32 different 128-B lines, and those lines
// compete with the BVH nodes for L1.  The first SD entries of every lane's stack now live in shared memory as
// sstack[entry][thread]: whatever the depths, a warp's push or pop is one conflict-free wavefront (bank = lane).  Deeper
// entries (rare: SD covers the depths a nearest-first walk of a SAH tree reaches) spill to a local array as before.
//
// TOP: the first `top_count` nodes of one tree (the BLAS with the most nodes, laid out breadth-first at the front of its
// node array by fj_bvh.cc) are staged in shared memory once per CTA by ONE bulk copy (cp.async.bulk + mbarrier, SASS
// UBLKCP): the first levels of every ray's walk through that tree are then LDS instead of L1/L2 round trips.
#pragma once

#include "fj_kernels.cuh"

namespace fj {

#define FJ_XT 128      // threads per CTA

struct ExtShared {
  double ox[FJ_XT], oy[FJ_XT], oz[FJ_XT], dx[FJ_XT], dy[FJ_XT], dz[FJ_XT];    // ray in the space being traversed (object space inside a BLAS)
  double tmin[FJ_XT], best_t[FJ_XT];
  const void *tri[FJ_XT];                                                     // triangle packets of the current mesh (tri32 or tri64)
  int cur_inst[FJ_XT], leaf[FJ_XT];
  unsigned ridx[FJ_XT];
};
// 84 B per lane = 10.75 KB per CTA: 7 (or 8) CTAs fit the 100 KB shared-memory carveout and leave 128 KB of L1.  The
// rest of the best hit (u, v, face, instance) goes straight to the ray's HitRec in global memory whenever it improves.

// 256-bit read-only global load (sm_100: LDG.E.ENL2.256.CONSTANT).  The closest-hit kernel is bound by L1 wavefronts —
// every lane reads its own node, so each load instruction costs one wavefront per lane whatever its width; a 128-B
// node is four 32-B loads instead of seven 16-B ones.  `p` must be 32-byte aligned.
struct F8 { float4 a, b; };
__device__ __forceinline__ F8 ldg256(const void *p) {
  F8 r;
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
  return r;
}

// mbarrier + bulk-copy primitives (TMA, non-tensor form) for the staged tree top
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned phase) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" :: "r"(bar), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Box ray of the min/max slab test: per axis 1/d and the two FMA constants of the lo and the hi plane,
//   t(lo plane) = fma(lo, inv, c_lo),  t(hi plane) = fma(hi, inv, c_hi),  near = min of the two, far = max of the two.
// The conservative widening of make_box_ray32 (near lowered, far raised by e) is kept by handing the lowered constant to
// whichever plane is the near one for this direction sign: (c_lo, c_hi) = inv >= 0 ? (cn, cf) : (cf, cn).  For a valid
// box (lo <= hi) this yields exactly the values of the address-selected form, with no per-step sign decoding.
struct BoxRayMM { float ix, iy, iz, lx, ly, lz, hx, hy, hz; };
__device__ __forceinline__ void make_box_ray_mm(const D3 &o, const D3 &d, float B, BoxRayMM &r) {
  float cn, cf;
  box_axis(o.x, d.x, B, &r.ix, &cn, &cf); r.lx = r.ix < 0.f ? cf : cn; r.hx = r.ix < 0.f ? cn : cf;
  box_axis(o.y, d.y, B, &r.iy, &cn, &cf); r.ly = r.iy < 0.f ? cf : cn; r.hy = r.iy < 0.f ? cn : cf;
  box_axis(o.z, d.z, B, &r.iz, &cn, &cf); r.lz = r.iz < 0.f ? cf : cn; r.hz = r.iz < 0.f ? cn : cf;
}

// TriRayIntersect (tri_intersect, fj_device.cuh) with the ray fetched from shared memory where it is used, so that origin
// and direction do not occupy twelve registers across the whole test.  Same operations in the same order: bit-identical.
__device__ __forceinline__ bool tri_intersect_smem(const D3 &v0, const D3 &v1, const D3 &v2, const volatile ExtShared &V, int tid,
                                                   double *t, double *u, double *v) {
  const D3 e1 = v1 - v0, e2 = v2 - v0;
  D3 pvec;
  { const D3 dir = mk(V.dx[tid], V.dy[tid], V.dz[tid]); pvec = cross(dir, e2); }
  const double det = dot(e1, pvec);
  if (det > -1e-6 && det < 1e-6) return false;
  const double inv_det = ddiv(1.0, det);
  const D3 tvec = mk(V.ox[tid], V.oy[tid], V.oz[tid]) - v0;
  *u = dmul(dot(tvec, pvec), inv_det);
  if (*u < 0.0 || *u > 1.0) return false;
  const D3 qvec = cross(tvec, e1);
  { const D3 dir = mk(V.dx[tid], V.dy[tid], V.dz[tid]); *v = dmul(dot(dir, qvec), inv_det); }
  if (*v < 0.0 || dadd(*u, *v) > 1.0) return false;
  *t = dmul(dot(e2, qvec), inv_det);
  return true;
}

// Lane state word: what the node loop has to know about the exact half of the state.
enum { XS_BLAS = 1, XS_WORLD = 2, XS_TRI64 = 4, XS_LEAF = 8, XS_FOUND = 16,
       XS_VEL = 32,      // moving triangles: 160-B packets with per-vertex velocity (Mesh::ray_intersect, src/fj_mesh.cc:252-259)
       XS_TOP = 64,      // the tree being walked is the one whose first nodes are staged in shared memory
       XS_ANY = 128 };   // a shadow ray whose every possible occluder is opaque: the first accepted hit ends the walk.  SlIlluminance
                         // scales the light by 1 - alpha of the CLOSEST occluder (src/fj_shading.cc:338-355); with alpha = 1 for every
                         // shader of the scene that is 0 whichever occluder is found, so any hit gives the reference's value

// One triangle of a leaf for the exact test: 48-B FP32 packet, 80-B FP64 packet, or the 160-B packet of a mesh with vertex
// velocity, moved to the ray's time as the reference does (`P0 += time * velocity0`, src/fj_mesh.cc:252-259).
__device__ __forceinline__ void load_leaf_triangle(const RenderArgs &a, const RayRec *rays, const void *tp_, unsigned fmt, unsigned ridx, int index,
                                                   D3 *v0, D3 *v1, D3 *v2, int *prim) {
  if (!(fmt & (XS_TRI64 | XS_VEL))) {
    // 48-B packet, 16-B aligned: one 32-B and one 16-B load, which comes first depends on the packet's parity
    const unsigned ti = (unsigned)index;
    const char *tb = (const char *)tp_ + 48 * (size_t)ti;
    const bool odd = ti & 1u;
    const float4 s4 = __ldg((const float4 *)(tb + (odd ? 0 : 32)));
    const F8 w8 = ldg256(tb + (odd ? 16 : 0));
    const float4 p0 = odd ? s4 : w8.a, p1 = odd ? w8.a : w8.b, p2 = odd ? w8.b : s4;
    *v0 = mk(p0.x, p0.y, p0.z); *v1 = mk(p1.x, p1.y, p1.z); *v2 = mk(p2.x, p2.y, p2.z); *prim = __float_as_int(p0.w);
  } else if (!(fmt & XS_VEL)) {
    const double *q = (const double *)tp_ + 10 * (size_t)index;
    *v0 = mk(__ldg(q), __ldg(q + 1), __ldg(q + 2)); *v1 = mk(__ldg(q + 3), __ldg(q + 4), __ldg(q + 5)); *v2 = mk(__ldg(q + 6), __ldg(q + 7), __ldg(q + 8));
    *prim = (int)__double_as_longlong(__ldg(q + 9));
  } else {
    const double *q = (const double *)tp_ + 20 * (size_t)index;
    const double tm = a.sc.time_tab[rays[ridx].key];                  // the ray's time (TraceContext::time): entry `key` of the frame's table
    *v0 = mk(__ldg(q), __ldg(q + 1), __ldg(q + 2)) + tm * mk(__ldg(q + 10), __ldg(q + 11), __ldg(q + 12));
    *v1 = mk(__ldg(q + 3), __ldg(q + 4), __ldg(q + 5)) + tm * mk(__ldg(q + 13), __ldg(q + 14), __ldg(q + 15));
    *v2 = mk(__ldg(q + 6), __ldg(q + 7), __ldg(q + 8)) + tm * mk(__ldg(q + 16), __ldg(q + 17), __ldg(q + 18));
    *prim = (int)__double_as_longlong(__ldg(q + 9));
  }
}

// pop: the shared-memory load is unconditional (clamped index), the local-memory entry replaces it in the rare deep case
template <int SD>
__device__ __forceinline__ int xpop_(const int (*sstack)[FJ_XT], const int *lstack, int &sp, int tid) {
  --sp;
  int v = sstack[min(sp, SD - 1)][tid];
  if (sp >= SD) v = lstack[sp - SD];
  return v;
}

template <int MINB, bool STATS, bool QUANT, bool COOP = false, int SD = 12, bool TOP = false>
__global__ void __launch_bounds__(FJ_XT, MINB) k_extend2(const RenderArgs a) {
  __shared__ ExtShared S;
  __shared__ int sstack[SD][FJ_XT];             // the first SD stack entries of every lane (entry-major: bank = lane)
  extern __shared__ __align__(128) unsigned char top_smem[];      // TOP: a.top_count staged nodes (64 B each)
  __shared__ unsigned long long top_bar;
  // COOP: (owner lane | triangle k << 5) of every (ray, triangle) pair of the warp's parked leaves, in owner order
  __shared__ unsigned char pairmap[COOP ? FJ_XT / 32 : 1][COOP ? 256 : 1];
  const unsigned FULL = 0xffffffffu;
  const int tid = threadIdx.x;
  const RayRec *rays = a.queue[a.cur];
  if (blockIdx.x == 0 && threadIdx.x == 0) a.ctl->count[a.cur ^ 1] = 0;      // the queue k_shade fills next: nobody touches it during this kernel
  const unsigned count = min(a.ctl->count[a.cur], a.capacity);
  const DScene &sc = a.sc;
  int lstack[FJ_STACK4 - SD];                   // entries beyond SD (local memory; rarely reached)
#define XPUSH(V) do { const int v_ = (V); if (sp < SD) sstack[sp][tid] = v_; else lstack[sp - SD] = v_; sp++; } while (0)
#define XPOP() xpop_<SD>(sstack, lstack, sp, tid)
  if (TOP) {                                    // stage the tree top: one bulk copy per CTA, completion on an mbarrier
    const unsigned bar = smem_u32(&top_bar);
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (tid == 0) {
      const unsigned bytes = (unsigned)a.top_count * 64u;
      mbar_expect_tx(bar, bytes);
      bulk_g2s(smem_u32(top_smem), a.top_src, bytes, bar);
    }
    mbar_wait(bar, 0);
  }
  // negative node references that are not leaves: end of a BLAS, end of the traversal, lane without a ray
  const int SENTINEL = (int)0x80000000, DONE = (int)0x80000001, IDLE = (int)0x80000002;
  const unsigned MISS = 0xffffffffu;

  bool drained = false;
  unsigned st = 0;
  int sp = 0, node = IDLE;
  float tn = 0, tf = 0;
  BoxRayMM br; br.ix = br.iy = br.iz = br.lx = br.ly = br.lz = br.hx = br.hy = br.hz = 0.f;
  const char *nodes = nullptr;                  // 4-wide nodes of the tree being walked: Node128, or NodeQ64 when QUANT
  unsigned n_steps = 0, n_tris = 0;             // warp totals (every lane carries the same value)
  unsigned n_phases = 0, n_rounds = 0;

#pragma unroll 1
  for (;;) {
    // ---- refill idle lanes from the queue head
    const unsigned idle = __ballot_sync(FULL, node == IDLE);
    if (idle == FULL && drained) break;
    if (!drained && __popc(idle) >= a.refill) {
      const int lane = tid & 31;
      const int n = __popc(idle), leader = __ffs(idle) - 1;
      unsigned base = 0;
      if (lane == leader) base = atomicAdd(&a.ctl->head, (unsigned)n);
      base = __shfl_sync(FULL, base, leader);
      if (base + n >= count) drained = true;
      if (node == IDLE) {
        const unsigned i = base + __popc(idle & ((1u << lane) - 1));
        if (i < count) {
          const unsigned ridx = a.perm ? a.perm[i] : i;
          const RayRec &r = rays[ridx];
          const double tmin = r.tmin, tmax = r.tmax;
          S.ridx[tid] = ridx; S.tmin[tid] = tmin; S.best_t[tid] = tmax; S.cur_inst[tid] = -1;
          tn = __double2float_rd(tmin); tf = __double2float_ru(tmax);
          const DGroup grp = sc.groups[r.target];
          nodes = (const char *)(QUANT ? grp.nodesq : grp.nodes4); sp = 0; node = 0;
          make_box_ray_mm(mk(r.o[0], r.o[1], r.o[2]), mk(r.d[0], r.d[1], r.d[2]), QUANT ? grp.bmagq : grp.bmag, br);
          st = XS_WORLD | ((a.shadow_anyhit && r.type == RAY_SHADOW) ? XS_ANY : 0);
        }
      }
      if (STATS && (n_steps | n_tris) > 0x40000000u) {      // keep the 32-bit warp totals from wrapping
        if (lane == 0) { atomicAdd(&a.counters->node_steps, (unsigned long long)n_steps); atomicAdd(&a.counters->tri_tests, (unsigned long long)n_tris); }
        n_steps = n_tris = 0;
      }
      if (__ballot_sync(FULL, node != IDLE) == 0) break;
    }

    // ---- phase A: 4-wide inner nodes, nearest hit child first
#pragma unroll 1
    for (;;) {
      const bool want = node >= 0;
      const unsigned wm = __ballot_sync(FULL, want);
      if (wm == 0) break;
      // few lanes left descending: switch to the parked leaves / transitions if there are any to work on
      if (__popc(wm) < a.phase_a_min && __any_sync(FULL, (st & XS_LEAF) || (node < 0 && node != IDLE))) break;
      if (STATS) n_steps += __popc(wm);
      if (want) {
        if (!(st & (XS_BLAS | XS_WORLD))) {      // back in the instance tree after a BLAS: world-space box ray and tree again
          const RayRec &r = rays[S.ridx[tid]];
          const DGroup grp = sc.groups[r.target];
          nodes = (const char *)(QUANT ? grp.nodesq : grp.nodes4);
          make_box_ray_mm(mk(r.o[0], r.o[1], r.o[2]), mk(r.d[0], r.d[1], r.d[2]), QUANT ? grp.bmagq : grp.bmag, br);
          st |= XS_WORLD;
        }
        unsigned key0, key1, key2, key3; int4 ch;
        if (QUANT) {
          // 64-B node with 8-bit planes (fj_bvh.h NodeQ64): plane = p + q s, so t(plane) = fma(q, s/d, fma(p, 1/d, c)) — two more
          // roundings of at most 2^-24 (|o| + B) |1/d| each, inside the 2^-20 widening c already carries
          F8 A, Q;
          if (TOP && (st & XS_TOP) && node < a.top_count) {       // staged: four 16-B shared-memory loads
            const float4 *sn = reinterpret_cast<const float4 *>(top_smem + 64 * (size_t)node);
            A.a = sn[0]; A.b = sn[1]; Q.a = sn[2]; Q.b = sn[3];
          } else {
            const char *np = nodes + 64 * (size_t)node;
            A = ldg256(np); Q = ldg256(np + 32);
          }
          const unsigned sw = __float_as_uint(A.a.w);
          float six = __fmul_rn(__uint_as_float(sw & 0xffff0000u), br.ix), siy = __fmul_rn(__uint_as_float(sw << 16), br.iy);
          float siz = __fmul_rn(Q.b.z, br.iz);
          // near / far plane by direction sign, resolved once per node on the packed words (not per child, and not by min/max:
          // unused slots hold an inverted box that must stay a miss).  br.l* / br.h* are the constants of the lo / hi plane.
          const bool gx = br.ix < 0.f, gy = br.iy < 0.f, gz = br.iz < 0.f;
          float bnx = fmaf(A.a.x, br.ix, gx ? br.hx : br.lx), bfx = fmaf(A.a.x, br.ix, gx ? br.lx : br.hx);
          float bny = fmaf(A.a.y, br.iy, gy ? br.hy : br.ly), bfy = fmaf(A.a.y, br.iy, gy ? br.ly : br.hy);
          float bnz = fmaf(A.a.z, br.iz, gz ? br.hz : br.lz), bfz = fmaf(A.a.z, br.iz, gz ? br.lz : br.hz);
          const unsigned qlx = __float_as_uint(A.b.x), qhx = __float_as_uint(A.b.y), qly = __float_as_uint(A.b.z), qhy = __float_as_uint(A.b.w);
          const unsigned qlz = __float_as_uint(Q.a.x), qhz = __float_as_uint(Q.a.y);
          const unsigned qnx = gx ? qhx : qlx, qfx = gx ? qlx : qhx, qny = gy ? qhy : qly, qfy = gy ? qly : qhy, qnz = gz ? qhz : qlz, qfz = gz ? qlz : qhz;
          ch = make_int4(__float_as_int(Q.a.z), __float_as_int(Q.a.w), __float_as_int(Q.b.x), __float_as_int(Q.b.y));
#define FJ_Q2F(W, K) ((float)(((W) >> (8 * K)) & 255u))
#define FJ_CHILD(KEY, K)                                                                                                        \
          {                                                                                                                        \
            const float nx_ = fmaf(FJ_Q2F(qnx, K), six, bnx), fx_ = fmaf(FJ_Q2F(qfx, K), six, bfx);                               \
            const float ny_ = fmaf(FJ_Q2F(qny, K), siy, bny), fy_ = fmaf(FJ_Q2F(qfy, K), siy, bfy);                               \
            const float nz_ = fmaf(FJ_Q2F(qnz, K), siz, bnz), fz_ = fmaf(FJ_Q2F(qfz, K), siz, bfz);                               \
            const float nr = fmaxf(fmaxf(nx_, ny_), fmaxf(nz_, tn));                                                              \
            const float fr_ = fminf(fminf(fx_, fy_), fminf(fz_, tf));                                                             \
            KEY = nr <= fr_ ? ((__float_as_uint(nr) & ~3u) | K##u) : MISS;                                                        \
          }
          FJ_CHILD(key0, 0) FJ_CHILD(key1, 1) FJ_CHILD(key2, 2) FJ_CHILD(key3, 3)
#undef FJ_CHILD
#undef FJ_Q2F
        } else {
        // 4-wide node: lo.x[4] hi.x[4] | lo.y[4] hi.y[4] | lo.z[4] hi.z[4] | child[4] — three 32-B loads and one 16-B load
        const char *np = nodes + 128 * (size_t)node;
        const F8 X = ldg256(np), Y = ldg256(np + 32), Z = ldg256(np + 64);
        ch = __ldg((const int4 *)(np + 96));
#define FJ_CHILD(KEY, K, C)                                                                                                     \
        {                                                                                                                          \
          const float x0 = fmaf(X.a.C, br.ix, br.lx), x1 = fmaf(X.b.C, br.ix, br.hx);                                            \
          const float y0 = fmaf(Y.a.C, br.iy, br.ly), y1 = fmaf(Y.b.C, br.iy, br.hy);                                            \
          const float z0 = fmaf(Z.a.C, br.iz, br.lz), z1 = fmaf(Z.b.C, br.iz, br.hz);                                            \
          const float nr = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tn));                                   \
          const float fr_ = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tf));                                  \
          KEY = nr <= fr_ ? ((__float_as_uint(nr) & ~3u) | K) : MISS;                                                              \
        }
        FJ_CHILD(key0, 0u, x) FJ_CHILD(key1, 1u, y) FJ_CHILD(key2, 2u, z) FJ_CHILD(key3, 3u, w)
#undef FJ_CHILD
        }
        // entry distances are positive (tn > 0), so their bit patterns order like unsigned integers.  The nearest hit child
        // is visited next; the other hit children are pushed (exact front-to-back order for up to two hits, slot order beyond)
        const unsigned kmin = min(min(key0, key1), min(key2, key3));
        if (kmin != MISS) {
          const unsigned w = kmin & 3u;
          const bool p0 = key0 != MISS && w != 0u, p1 = key1 != MISS && w != 1u, p2 = key2 != MISS && w != 2u, p3 = key3 != MISS && w != 3u;
          if (sp + 3 <= SD) {                    // (almost always) up to three predicated shared-memory stores, no branches
            if (p0) sstack[sp][tid] = ch.x;
            sp += p0;
            if (p1) sstack[sp][tid] = ch.y;
            sp += p1;
            if (p2) sstack[sp][tid] = ch.z;
            sp += p2;
            if (p3) sstack[sp][tid] = ch.w;
            sp += p3;
          } else {
            if (p0) XPUSH(ch.x);
            if (p1) XPUSH(ch.y);
            if (p2) XPUSH(ch.z);
            if (p3) XPUSH(ch.w);
          }
          node = (w & 2u) ? ((w & 1u) ? ch.w : ch.z) : ((w & 1u) ? ch.y : ch.x);
        } else { if (sp > 0) node = XPOP(); else node = DONE; }
        // speculative traversal: park the first triangle leaf and keep descending.  Inside a BLAS the bottom stack entry is
        // SENTINEL, so a negative reference there is SENTINEL or a triangle leaf and the pop below cannot underflow.
        if (node < 0 && (st & (XS_BLAS | XS_LEAF)) == XS_BLAS && node != SENTINEL) { S.leaf[tid] = node; st |= XS_LEAF; node = XPOP(); }
      }
    }

    // ---- phase B1 (COOP): the warp's parked leaves hold W (ray, triangle) pairs on typically 10-12 lanes; instead of every
    // owner looping over its own <= 8 triangles, the pairs are dealt out 32 at a time to ALL lanes (the FP64 ray of any lane
    // is in shared memory), and each owner then folds the hits among its pairs into its best hit in triangle order — the
    // same comparisons in the same order as the loop below, so the result is bit-identical.
    if (COOP && __any_sync(FULL, st & XS_LEAF)) {
      const int lane = tid & 31, wbase = tid & ~31;
      int cnt = 0;
      if (st & XS_LEAF) { cnt = ((~S.leaf[tid]) & 7) + 1; st &= ~XS_LEAF; }
      int off = cnt;                                           // inclusive, then exclusive prefix sum of the pair counts
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) { const int up = __shfl_up_sync(FULL, off, dlt); if (lane >= dlt) off += up; }
      const int total = __shfl_sync(FULL, off, 31);
      off -= cnt;
      unsigned char *pm = pairmap[tid >> 5];
      for (int k = 0; k < cnt; k++) pm[off + k] = (unsigned char)(lane | (k << 5));
      __syncwarp();
      if (STATS) { n_tris += (unsigned)total; n_phases++; n_rounds += (unsigned)(total + 31) >> 5; }
      const double tmin = S.tmin[tid];
      double best_t = S.best_t[tid]; bool found = (st & XS_FOUND) != 0;       // owner state (unused on lanes without a leaf)
      for (int R = 0; R < total; R += 32) {
        const int p = R + lane;
        const bool want = p < total;
        const unsigned e = want ? pm[p] : (unsigned)lane;
        const int otid = wbase + (int)(e & 31u), k = (int)(e >> 5);
        const unsigned ost = __shfl_sync(FULL, st, (int)(e & 31u));
        bool hit = false; double t = 0, u = 0, v = 0; int prim = 0;
        if (want) {
          const int first = (~S.leaf[otid]) >> 3;
          D3 v0, v1, v2;
          load_leaf_triangle(a, rays, S.tri[otid], ost, S.ridx[otid], first + k, &v0, &v1, &v2, &prim);
          hit = tri_intersect_smem(v0, v1, v2, S, otid, &t, &u, &v);
        }
        // owners fold the hits among their pairs of this round, lowest triangle first
        const unsigned hm = __ballot_sync(FULL, hit);
        const int lo = off - R, hi = off + cnt - R;             // this owner's pairs sit on lanes [lo, hi) of this round
        unsigned mine = 0;
        if (cnt > 0 && hi > 0 && lo < 32) {
          const unsigned upto = hi >= 32 ? FULL : ((1u << hi) - 1u);
          const unsigned from = lo <= 0 ? 0u : ((1u << lo) - 1u);
          mine = hm & upto & ~from;
        }
        while (__any_sync(FULL, mine != 0u)) {
          const int src = mine ? __ffs(mine) - 1 : lane;
          const double ht = __shfl_sync(FULL, t, src), hu = __shfl_sync(FULL, u, src), hv = __shfl_sync(FULL, v, src);
          const int hp = __shfl_sync(FULL, prim, src);
          if (mine) {
            mine &= mine - 1u;
            if (tmin <= ht && ht <= best_t) {
              bool better = !found || ht < best_t;
              if (!better) {                       // exact tie in t: lower instance, then higher face id (read back from the record)
                const HitRec *cur = a.hits + S.ridx[tid];
                const int ci = S.cur_inst[tid], bi = cur->inst;
                better = ci < bi || (ci == bi && hp > cur->prim);
              }
              if (better) {
                found = true; best_t = ht; st |= XS_FOUND;
                S.best_t[tid] = ht;
                HitRec hr; hr.t = ht; hr.u = hu; hr.v = hv; hr.prim = hp; hr.inst = S.cur_inst[tid];
                uint4 *dst = reinterpret_cast<uint4 *>(a.hits + S.ridx[tid]); const uint4 *srcp = reinterpret_cast<const uint4 *>(&hr);
                dst[0] = srcp[0]; dst[1] = srcp[1];
                tf = __double2float_ru(ht);
              }
            }
          }
        }
      }
      __syncwarp();                                            // pairmap is rewritten by the next leaf phase
    }

    // ---- phase B1: parked triangle leaves, exact FP64 tests, one triangle per lane per iteration
    if (!COOP && __any_sync(FULL, st & XS_LEAF)) {
      int first = 0, cnt = 0;
      double tmin = 0, best_t = 0; bool found = false;
      const void *tp_ = nullptr;
      if (st & XS_LEAF) {
        const int ref = ~S.leaf[tid];
        first = ref >> 3; cnt = (ref & 7) + 1;
        tmin = S.tmin[tid]; best_t = S.best_t[tid]; found = (st & XS_FOUND) != 0; tp_ = S.tri[tid];
        st &= ~XS_LEAF;
      }
      for (int k = 0;; k++) {
        const bool want = k < cnt;
        const unsigned wm = __ballot_sync(FULL, want);
        if (wm == 0) break;
        if (STATS) n_tris += __popc(wm);
        if (want) {
          D3 v0, v1, v2; int prim;
          load_leaf_triangle(a, rays, tp_, st, S.ridx[tid], first + k, &v0, &v1, &v2, &prim);
          double t, u, v;
          // RayInRange (src/fj_ray.h:29-32): tmin <= t <= tmax; best_t starts at tmax, so `t <= best_t` is the upper test
          if (tri_intersect_smem(v0, v1, v2, S, tid, &t, &u, &v) && tmin <= t && t <= best_t) {
            bool better = !found || t < best_t;
            if (!better) {                         // exact tie in t: lower instance, then higher face id (read back from the record)
              const HitRec *cur = a.hits + S.ridx[tid];
              const int ci = S.cur_inst[tid], bi = cur->inst;
              better = ci < bi || (ci == bi && prim > cur->prim);
            }
            if (better) {
              found = true; best_t = t; st |= XS_FOUND;
              S.best_t[tid] = t;
              HitRec hr; hr.t = t; hr.u = u; hr.v = v; hr.prim = prim; hr.inst = S.cur_inst[tid];
              uint4 *dst = reinterpret_cast<uint4 *>(a.hits + S.ridx[tid]); const uint4 *src = reinterpret_cast<const uint4 *>(&hr);
              dst[0] = src[0]; dst[1] = src[1];
              tf = __double2float_ru(t);
            }
          }
        }
      }
    }

    // ---- any-hit shadow rays: an accepted hit ends the walk (B2 below retires the lane; its HitRec is already written)
    if ((st & (XS_ANY | XS_FOUND)) == (XS_ANY | XS_FOUND)) { node = DONE; sp = 0; st &= ~XS_LEAF; }

    // ---- phase B2: transitions of lanes whose next stack entry is not an inner node
    const bool special = node < 0 && node != IDLE;
    if (__any_sync(FULL, special)) {
      if (special) {
        if (node == DONE) {                        // traversal finished: a hit is already in the ray's record, a miss is written now
          if (!(st & XS_FOUND)) {
            HitRec hr; hr.t = FJ_REAL_MAX; hr.u = 0; hr.v = 0; hr.prim = -1; hr.inst = -1;
            store_hit_cs(a.hits + S.ridx[tid], hr);
          }
          node = IDLE;
        } else if (node == SENTINEL) {             // the instance's BLAS is done: back to the instance tree (its nodes and box
          st &= ~(XS_BLAS | XS_WORLD | XS_TOP);    // ray are fetched again only if an inner node of that tree is still to be visited)
          if (sp > 0) node = XPOP(); else node = DONE;
        } else if (st & XS_BLAS) {                 // a second triangle leaf: park it now that the slot is free
          S.leaf[tid] = node; st |= XS_LEAF; node = XPOP();
        } else {                                   // TLAS leaf: enter the first instance, re-queue the others
          const int ref = ~node;
          const int first = ref >> 3, cnt = (ref & 7) + 1;
          for (int k = cnt - 1; k >= 1; k--) XPUSH(~(((first + k) << 3) | 0));
          const RayRec &r = rays[S.ridx[tid]];
          const DInstRec &in = sc.groups[r.target].irec[first];          // one record: matrix, tree, packets (no order[] -> instance -> mesh chain)
          S.cur_inst[tid] = in.inst;
          const double *inv = in.motion ? in.motion + 24 * (size_t)r.key : in.inv;     // time-sampled transform: the ray's entry of the time table
          const D3 o = mat_point(inv, mk(r.o[0], r.o[1], r.o[2]));
          const D3 d = mat_vector(inv, mk(r.d[0], r.d[1], r.d[2]));
          S.ox[tid] = o.x; S.oy[tid] = o.y; S.oz[tid] = o.z; S.dx[tid] = d.x; S.dy[tid] = d.y; S.dz[tid] = d.z;
          make_box_ray_mm(o, d, QUANT ? in.bmagq : in.bmag, br);
          nodes = QUANT ? in.nodesq : in.nodes4;
          S.tri[tid] = in.tri;
          st = (st & (XS_FOUND | XS_ANY)) | XS_BLAS | (in.tri64 == 1 ? XS_TRI64 : 0) | (in.tri64 == 2 ? XS_VEL : 0);
          if (TOP && QUANT && (const void *)in.nodesq == (const void *)a.top_src) st |= XS_TOP;
          XPUSH(SENTINEL);
          node = 0;
        }
      }
    }
  }
  // traversal statistics (4-wide node steps and exact triangle tests) for DESIGN.md / bench.py
  if (STATS && (tid & 31) == 0 && a.counters) {
    atomicAdd(&a.counters->node_steps, (unsigned long long)n_steps); atomicAdd(&a.counters->tri_tests, (unsigned long long)n_tris);
    atomicAdd(&a.counters->leaf_phases, (unsigned long long)n_phases); atomicAdd(&a.counters->leaf_rounds, (unsigned long long)n_rounds);
  }
#undef XPUSH
#undef XPOP
}

}  // namespace fj
