// fj_extend_ring.cuh — k_extend_ring: the closest-hit kernel that ships.  k_extend2 (fj_extend.cuh) with a cheaper node loop,
// its rare phases gated by the lanes that wait for them, and — as an option — a per-warp ring of PREPARED rays between the
// queue and the lanes.  Same contract and results (closest hit of Accelerator::Intersect, src/fj_shading.cc:538-541;
// bit-identical t / u / v / face / instance), quantised 64-B nodes and warp-cooperative leaves only — trees that cannot be
// quantised keep k_extend2.
//
// What the source-level profile of k_extend2 said (DESIGN.md 4.1 "Round 2, second half"): nothing is saturated — 0.65
// instructions per cycle and scheduler with 1.3 eligible warps of 7 — and on bounce rays every instruction runs on 17.5 of 32
// lanes: a warp left the node loop after 3 steps, tested 19 (ray, triangle) pairs in rounds of 32 and entered instances
// — 300 instructions — with 5 lanes.  Hence:
//   node loop   the slab constants are kept as (near, far) instead of (lo, hi) — k_extend2 selected near / far from lo / hi with
//               six FSEL per step, and for the quantised form the selection is the identity; the 24 plane distances of a step are
//               12 packed FMAs (FFMA2: each half an IEEE fma.rn, bit-identical keys); the stack is addressed by a running
//               shared-memory address with a DONE / SENTINEL entry at its bottom (a pop is SUB + LDS, no emptiness test; a
//               push is STS + ADD); ray records are read in 16-byte pieces (one L1 tag lookup per lane and instruction
//               whatever the width); the ray's target group travels in the state word;
//   transitions the cheap ones (park a leaf, leave a finished BLAS, retire a finished ray: FJ_TRANSIT) run once per outer
//               iteration and once after the leaf phase;
//   gates       the two heavy phases — exact triangle tests (B1), instance entry (E) — run when enough lanes wait for them
//               (RenderArgs::b1_min pairs, b2_min lanes), and whenever lanes are blocked at least one of the two runs;
//   ring        (RING = true, FJGPU_RING=1; off by default) when >= `refill` of the warp's 32 ring slots are free, that many
//               rays are taken from the queue head with one atomic and prepared side by side (box ray, tn / tf, state word)
//               into shared memory; every outer iteration each idle lane takes the next prepared ray.  Measured: the lanes
//               it fills were not the ones missing (bounce rays are BLOCKED, not idle) and its 6.5 KB per CTA come out of L1.
#pragma once

#include <type_traits>
#include "fj_extend.cuh"

namespace fj {

struct RingShared : ExtShared {
  // ring of prepared rays: slot s of warp w lives at index 32 w + s of every array (consecutive slots = consecutive banks)
  float r_ix[FJ_XT], r_iy[FJ_XT], r_iz[FJ_XT], r_nx[FJ_XT], r_ny[FJ_XT], r_nz[FJ_XT], r_fx[FJ_XT], r_fy[FJ_XT], r_fz[FJ_XT];
  float r_tn[FJ_XT], r_tf[FJ_XT];
  unsigned r_ridx[FJ_XT], r_st[FJ_XT];           // ray index in the queue; the lane's initial state word (target group << 12 | XS_ANY)
};

__device__ __forceinline__ void sts32(unsigned addr, int v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ int lds32(unsigned addr) { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }

// Two independent FP32 FMAs in one instruction (sm_100 FFMA2): each half is an IEEE fma.rn, so the result is bit-identical to
// two fmaf calls — the node step issues 12 of these instead of 24 FFMA.
__device__ __forceinline__ void ffma2(float a0, float a1, float b, float c0, float c1, float &d0, float &d1) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(rb) : "f"(b));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c0), "f"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(rd));
}

// A ray record (112 B, 16-B aligned) is read in 16-B pieces: every lane reads its own record, so each load instruction costs one
// L1 tag lookup per lane whatever its width — (o, d) is three lookups instead of six, (tmin, tmax) one instead of two.
__device__ __forceinline__ void ld_ray_od(const RayRec *r, D3 *o, D3 *d) {
  const double2 *q = reinterpret_cast<const double2 *>(r);
  const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  *o = mk(a.x, a.y, b.x); *d = mk(b.y, c.x, c.y);
}
__device__ __forceinline__ double2 ld_ray_range(const RayRec *r) { return __ldg(reinterpret_cast<const double2 *>(r) + 3); }   // (tmin, tmax)
// 32-B hit record in one store (sm_100: STG.E.ENL2.256); `p` is 32-byte aligned
__device__ __forceinline__ void stg256_cs(void *p, const HitRec &h) {
  const float *f = reinterpret_cast<const float *>(&h);
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7]) : "memory");
}
__device__ __forceinline__ void stg256(void *p, const HitRec &h) {
  const float *f = reinterpret_cast<const float *>(&h);
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7]) : "memory");
}
// One triangle of a leaf as load_leaf_triangle (fj_extend.cuh), the 48-B FP32 packet read as three 16-byte pieces: one tag
// lookup more than the 32 + 16 split of k_extend2, and none of the twelve selects that order the pieces by the packet's parity
// (measured: k_extend 279.2-280.2 -> 276.2 ms on the north star, 384.3 -> 380.2 on config 4).
__device__ __forceinline__ void load_leaf_triangle3(const RenderArgs &a, const RayRec *rays, const void *tp_, unsigned fmt, unsigned ridx, int index,
                                                    D3 *v0, D3 *v1, D3 *v2, int *prim) {
  if (!(fmt & (XS_TRI64 | XS_VEL))) {
    const float4 *tb = reinterpret_cast<const float4 *>(tp_) + 3 * (size_t)(unsigned)index;
    const float4 p0 = __ldg(tb), p1 = __ldg(tb + 1), p2 = __ldg(tb + 2);
    *v0 = mk(p0.x, p0.y, p0.z); *v1 = mk(p1.x, p1.y, p1.z); *v2 = mk(p2.x, p2.y, p2.z); *prim = __float_as_int(p0.w);
  } else load_leaf_triangle(a, rays, tp_, fmt, ridx, index, v0, v1, v2, prim);
}

// Lane state above the XS_* flags of fj_extend.cuh: the object group the ray is traced against (RayRec::target) from bit 12, so
// that the transitions and phase E do not read the record for it; bits 9-11 hold the size of the parked leaf
#define XS_INIT 256u            // S.tmin / S.best_t hold the ray's exact range (set when the first instance is entered)
#define XS_TARGET_SHIFT 12
#define XS_CNT_SHIFT 9           // triangles - 1 of the parked leaf (3 bits)

#define FJ_XSTRIDE (FJ_XT * 4)      // bytes between two stack entries of one lane

template <int MINB, int SD, bool F2 = true, bool RING = true>
__global__ void __launch_bounds__(FJ_XT, MINB) k_extend_ring(const RenderArgs a) {
  __shared__ typename std::conditional<RING, RingShared, ExtShared>::type S;      // (without the ring: its slots are not allocated)
  __shared__ int sstack[SD][FJ_XT];             // the first SD stack entries of every lane (entry-major: bank = lane)
  __shared__ unsigned char pairmap[FJ_XT / 32][256];
  const unsigned FULL = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
  const RayRec *rays = a.queue[a.cur];
  if (blockIdx.x == 0 && threadIdx.x == 0) a.ctl->count[a.cur ^ 1] = 0;      // the queue k_shade fills next: nobody touches it during this kernel
  const unsigned count = min(a.ctl->count[a.cur], a.capacity);
  const DScene &sc = a.sc;
  int lstack[FJ_STACK4 + 1 - SD];               // entries beyond SD (local memory; rarely reached)
  // The stack pointer is the shared-memory ADDRESS of the lane's next free entry (entry k of this lane sits at sa0 + FJ_XSTRIDE k):
  // an access is one LDS / STS, a push STS + ADD, a pop SUB + LDS.  sa0 passes through an opaque move so that the compiler
  // keeps it in a register instead of rebuilding the CTA's shared-window base (four instructions) at every use.  Entries
  // k >= SD are virtual: they live in lstack.
  unsigned sa0;
  asm volatile("mov.u32 %0, %1;" : "=r"(sa0) : "r"(smem_u32(&sstack[0][tid])));
  unsigned sa = sa0;
#define RPUSH(V) do { const int v_ = (V); if (sa - sa0 < SD * FJ_XSTRIDE) sts32(sa, v_); else lstack[(sa - sa0) / FJ_XSTRIDE - SD] = v_; sa += FJ_XSTRIDE; } while (0)
#define RPOP(DST) do { sa -= FJ_XSTRIDE; if (sa - sa0 < SD * FJ_XSTRIDE) DST = lds32(sa); else DST = lstack[(sa - sa0) / FJ_XSTRIDE - SD]; } while (0)
  // negative node references that are not leaves: end of a BLAS, end of the traversal, lane without a ray
  const int SENTINEL = (int)0x80000000, DONE = (int)0x80000001, IDLE = (int)0x80000002;
  const int FIXUP = (int)0x80000003;            // back in the instance tree with an inner node on the stack: the world-space box ray has to be rebuilt (phase E)
  const unsigned MISS = 0xffffffffu;
  // Cheap transitions of a lane whose next reference is not an inner node, done once per outer iteration after the node loop and
  // once after the leaf phase: park a triangle leaf when the slot is free, leave a finished BLAS, retire a finished ray.  What
  // stays blocked afterwards waits for a HEAVY phase: a second leaf (or the end of the walk) behind a parked one -> leaf phase
  // B1; a leaf of the instance tree or FIXUP -> entry phase E.  Inside a BLAS the bottom stack entry is SENTINEL and below it
  // lies the instance tree's DONE, so no pop underflows.
#define FJ_TRANSIT()                                                                                                              \
  if (node < 0 && node != IDLE) {                                                                                                 \
    if (st & XS_BLAS) {                                                                                                           \
      if (!(st & XS_LEAF) && node != SENTINEL) {                                                                                  \
        S.leaf[tid] = node; st |= XS_LEAF | (((unsigned)~node & 7u) << XS_CNT_SHIFT); RPOP(node);                                 \
      }                                                                                                                           \
      if (node == SENTINEL) {              /* (a parked leaf keeps S.tri / S.o / S.d / the packet format valid until the next entry) */ \
        st &= ~(unsigned)XS_BLAS; RPOP(node);                                                                                     \
        if (node >= 0) { RPUSH(node); node = FIXUP; }                                                                             \
      }                                                                                                                           \
    }                                                                                                                             \
    if (node == DONE && !(st & XS_LEAF)) {  /* a hit is already in the ray's record, a miss is written now */                    \
      if (!(st & XS_FOUND)) { HitRec hr_; hr_.t = FJ_REAL_MAX; hr_.u = 0; hr_.v = 0; hr_.prim = -1; hr_.inst = -1; stg256_cs(a.hits + S.ridx[tid], hr_); } \
      node = IDLE;                                                                                                                \
    }                                                                                                                             \
  }

  bool drained = false;
  unsigned st = 0;
  int node = IDLE;
  int rhead = 0, rtail = 0;                     // ring: slots consumed / produced so far (warp-uniform)
  float tn = 0, tf = 0;
  float ix = 0, iy = 0, iz = 0, cnx = 0, cny = 0, cnz = 0, cfx = 0, cfy = 0, cfz = 0;      // box ray: 1/d, near and far slab constants
  const char *nodes = nullptr;                  // NodeQ64 array of the tree being walked
  unsigned n_steps = 0, n_tris = 0;             // warp totals (every lane carries the same value)
  unsigned n_pr = 0;                           // leaf phases (low half) and rounds of 32 pairs (high half)

#pragma unroll 1
  for (;;) {
    // (the loop's exit test has to come first: with the test after the consume step the compiler no longer proves the warp
    // converged at the votes and shuffles below and guards every one of them with a divergence check)
    const unsigned idle = __ballot_sync(FULL, node == IDLE);
    if (idle == FULL && drained && rtail == rhead) break;
    if constexpr (!RING) {
      // ---- direct refill (FJGPU_RING=0, the scheme of k_extend2): when >= `refill` lanes are idle they take the next rays from the
      // queue head and prepare them in place — fewer shared-memory bytes (more L1) against idle lanes in the node loop
      if (!drained && __popc(idle) >= a.refill) {
        const int n = __popc(idle);
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&a.ctl->head, (unsigned)n);
        base = __shfl_sync(FULL, base, 0);
        const unsigned i = base + __popc(idle & ((1u << lane) - 1u));
        const bool valid = node == IDLE && i < count;
        if (valid) {
          const unsigned ridx = a.perm ? a.perm[i] : i;
          const RayRec *r = rays + ridx;
          D3 o, d; ld_ray_od(r, &o, &d);
          const double2 range = ld_ray_range(r);
          const int2 tt = __ldg(reinterpret_cast<const int2 *>(&r->target));      // target, (type, dd, rd, fd)
          const DGroup grp = sc.groups[tt.x];
          box_axis(o.x, d.x, grp.bmagq, &ix, &cnx, &cfx);
          box_axis(o.y, d.y, grp.bmagq, &iy, &cny, &cfy);
          box_axis(o.z, d.z, grp.bmagq, &iz, &cnz, &cfz);
          tn = __double2float_rd(range.x); tf = __double2float_ru(range.y);
          S.ridx[tid] = ridx; S.cur_inst[tid] = -1;
          nodes = (const char *)grp.nodesq;
          st = ((unsigned)tt.x << XS_TARGET_SHIFT) | ((a.shadow_anyhit && (tt.y & 255) == RAY_SHADOW) ? (unsigned)XS_ANY : 0u);
          sts32(sa0, DONE); sa = sa0 + FJ_XSTRIDE;
          node = 0;
        }
        if (__popc(__ballot_sync(FULL, valid)) < n) drained = true;
        if (((n_steps | n_tris) & 0xc0000000u) | (n_pr & 0x80008000u)) {
          if (lane == 0 && a.counters) {
            atomicAdd(&a.counters->node_steps, (unsigned long long)n_steps); atomicAdd(&a.counters->tri_tests, (unsigned long long)n_tris);
            atomicAdd(&a.counters->leaf_phases, (unsigned long long)(n_pr & 0xffffu)); atomicAdd(&a.counters->leaf_rounds, (unsigned long long)(n_pr >> 16));
          }
          n_steps = n_tris = n_pr = 0;
        }
      }
    } else {
      // ---- produce: prepare rays for the free ring slots
      if (!drained) {
        const int nfree = 32 - (rtail - rhead);
        if (nfree >= a.refill) {
          unsigned base = 0;
          if (lane == 0) base = atomicAdd(&a.ctl->head, (unsigned)nfree);
          base = __shfl_sync(FULL, base, 0);
          const unsigned i = base + lane;
          const bool valid = lane < nfree && i < count;
          if (valid) {
            const unsigned ridx = a.perm ? a.perm[i] : i;
            const RayRec *r = rays + ridx;
            D3 o, d; ld_ray_od(r, &o, &d);
            const double2 range = ld_ray_range(r);
            const int2 tt = __ldg(reinterpret_cast<const int2 *>(&r->target));      // target, (type, dd, rd, fd)
            const float B = sc.groups[tt.x].bmagq;
            const int s = wbase + ((rtail + lane) & 31);
            float inv, cn, cf;
            box_axis(o.x, d.x, B, &inv, &cn, &cf); S.r_ix[s] = inv; S.r_nx[s] = cn; S.r_fx[s] = cf;
            box_axis(o.y, d.y, B, &inv, &cn, &cf); S.r_iy[s] = inv; S.r_ny[s] = cn; S.r_fy[s] = cf;
            box_axis(o.z, d.z, B, &inv, &cn, &cf); S.r_iz[s] = inv; S.r_nz[s] = cn; S.r_fz[s] = cf;
            S.r_tn[s] = __double2float_rd(range.x); S.r_tf[s] = __double2float_ru(range.y);
            S.r_ridx[s] = ridx;
            S.r_st[s] = ((unsigned)tt.x << XS_TARGET_SHIFT) | ((a.shadow_anyhit && (tt.y & 255) == RAY_SHADOW) ? (unsigned)XS_ANY : 0u);
          }
          // the ring counters are derived from ballots only: the compiler then knows them (and every branch on them) to be
          // warp-uniform and emits no divergence checks around the loop's votes and shuffles
          const int got = __popc(__ballot_sync(FULL, valid));      // valid lanes are a prefix: the rays fill consecutive slots
          rtail += got;
          if (got < nfree) drained = true;                         // the queue gave fewer rays than asked for
          if (((n_steps | n_tris) & 0xc0000000u) | (n_pr & 0x80008000u)) {      // keep the warp totals from wrapping
            if (lane == 0 && a.counters) {
              atomicAdd(&a.counters->node_steps, (unsigned long long)n_steps); atomicAdd(&a.counters->tri_tests, (unsigned long long)n_tris);
              atomicAdd(&a.counters->leaf_phases, (unsigned long long)(n_pr & 0xffffu)); atomicAdd(&a.counters->leaf_rounds, (unsigned long long)(n_pr >> 16));
            }
            n_steps = n_tris = n_pr = 0;
          }
          __syncwarp();
        }
      }
      // ---- consume: idle lanes take the next prepared rays
      {
        const int avail = rtail - rhead;
        if (idle != 0u && avail > 0) {
          const int rank = __popc(idle & ((1u << lane) - 1u));
          if (node == IDLE && rank < avail) {
            const int s = wbase + ((rhead + rank) & 31);
            ix = S.r_ix[s]; iy = S.r_iy[s]; iz = S.r_iz[s];
            cnx = S.r_nx[s]; cny = S.r_ny[s]; cnz = S.r_nz[s]; cfx = S.r_fx[s]; cfy = S.r_fy[s]; cfz = S.r_fz[s];
            tn = S.r_tn[s]; tf = S.r_tf[s];
            st = S.r_st[s];
            S.ridx[tid] = S.r_ridx[s]; S.cur_inst[tid] = -1;
            nodes = (const char *)sc.groups[st >> XS_TARGET_SHIFT].nodesq;
            sts32(sa0, DONE); sa = sa0 + FJ_XSTRIDE;                // the bottom entry ends the walk: a pop never underflows
            node = 0;
          }
          rhead += min(__popc(idle), avail);
          __syncwarp();                                              // the slots just read may be rewritten by the next produce
        }
      }

    }

    // ---- phase A: 4-wide inner nodes, nearest hit child first
#pragma unroll 1
    for (;;) {
      const bool want = node >= 0;
      const unsigned wm = __ballot_sync(FULL, want);
      if (wm == 0) break;
      // few lanes left descending: switch to the parked leaves / transitions if there are any to work on
      if (__popc(wm) < a.phase_a_min && __any_sync(FULL, (st & XS_LEAF) || (node < 0 && node != IDLE))) break;
      n_steps += __popc(wm);
      if (want) {
        // 64-B node with 8-bit planes (fj_bvh.h NodeQ64): plane = p + q s, so t(plane) = fma(q, s/d, fma(p, 1/d, c)) — two more
        // roundings of at most 2^-24 (|o| + B) |1/d| each, inside the 2^-20 widening c already carries
        const char *np = nodes + 64 * (size_t)node;
        const F8 A = ldg256(np), Q = ldg256(np + 32);
        const unsigned sw = __float_as_uint(A.a.w);
        const float six = __fmul_rn(__uint_as_float(sw & 0xffff0000u), ix), siy = __fmul_rn(__uint_as_float(sw << 16), iy);
        const float siz = __fmul_rn(Q.b.z, iz);
        // near / far plane by direction sign, resolved once per node on the packed words (unused slots hold an inverted box
        // that must stay a miss, so not by min / max).  cn* / cf* are the constants of the near / far plane.
        const bool gx = ix < 0.f, gy = iy < 0.f, gz = iz < 0.f;
        const float bnx = fmaf(A.a.x, ix, cnx), bfx = fmaf(A.a.x, ix, cfx);
        const float bny = fmaf(A.a.y, iy, cny), bfy = fmaf(A.a.y, iy, cfy);
        const float bnz = fmaf(A.a.z, iz, cnz), bfz = fmaf(A.a.z, iz, cfz);
        const unsigned qlx = __float_as_uint(A.b.x), qhx = __float_as_uint(A.b.y), qly = __float_as_uint(A.b.z), qhy = __float_as_uint(A.b.w);
        const unsigned qlz = __float_as_uint(Q.a.x), qhz = __float_as_uint(Q.a.y);
        const unsigned qnx = gx ? qhx : qlx, qfx = gx ? qlx : qhx, qny = gy ? qhy : qly, qfy = gy ? qly : qhy, qnz = gz ? qhz : qlz, qfz = gz ? qlz : qhz;
        const int4 ch = make_int4(__float_as_int(Q.a.z), __float_as_int(Q.a.w), __float_as_int(Q.b.x), __float_as_int(Q.b.y));
        unsigned key0, key1, key2, key3;
#define FJ_Q2F(W, K) ((float)(((W) >> (8 * K)) & 255u))
#define FJ_CHILD(KEY, K)                                                                                                        \
        {                                                                                                                          \
          float nx_, fx_, ny_, fy_, nz_, fz_;                                                                                     \
          if (F2) {                                                                                                               \
            ffma2(FJ_Q2F(qnx, K), FJ_Q2F(qfx, K), six, bnx, bfx, nx_, fx_);                                                       \
            ffma2(FJ_Q2F(qny, K), FJ_Q2F(qfy, K), siy, bny, bfy, ny_, fy_);                                                       \
            ffma2(FJ_Q2F(qnz, K), FJ_Q2F(qfz, K), siz, bnz, bfz, nz_, fz_);                                                       \
          } else {                                                                                                                \
            nx_ = fmaf(FJ_Q2F(qnx, K), six, bnx); fx_ = fmaf(FJ_Q2F(qfx, K), six, bfx);                                           \
            ny_ = fmaf(FJ_Q2F(qny, K), siy, bny); fy_ = fmaf(FJ_Q2F(qfy, K), siy, bfy);                                           \
            nz_ = fmaf(FJ_Q2F(qnz, K), siz, bnz); fz_ = fmaf(FJ_Q2F(qfz, K), siz, bfz);                                           \
          }                                                                                                                       \
          const float nr = fmaxf(fmaxf(nx_, ny_), fmaxf(nz_, tn));                                                                \
          const float fr_ = fminf(fminf(fx_, fy_), fminf(fz_, tf));                                                               \
          KEY = nr <= fr_ ? ((__float_as_uint(nr) & ~3u) | K##u) : MISS;                                                          \
        }
        FJ_CHILD(key0, 0) FJ_CHILD(key1, 1) FJ_CHILD(key2, 2) FJ_CHILD(key3, 3)
#undef FJ_CHILD
#undef FJ_Q2F
        // entry distances are positive (tn > 0), so their bit patterns order like unsigned integers.  The nearest hit child
        // is visited next; the other hit children are pushed (exact front-to-back order for up to two hits, slot order beyond)
        const unsigned kmin = min(min(key0, key1), min(key2, key3));
        if (kmin != MISS) {
          const unsigned w = kmin & 3u;
          const bool p0 = key0 != MISS && w != 0u, p1 = key1 != MISS && w != 1u, p2 = key2 != MISS && w != 2u, p3 = key3 != MISS && w != 3u;
          if (sa - sa0 <= (SD - 3) * FJ_XSTRIDE) {  // k + 3 <= SD (almost always): up to three predicated shared-memory stores, no branches
            if (p0) { sts32(sa, ch.x); sa += FJ_XSTRIDE; }
            if (p1) { sts32(sa, ch.y); sa += FJ_XSTRIDE; }
            if (p2) { sts32(sa, ch.z); sa += FJ_XSTRIDE; }
            if (p3) { sts32(sa, ch.w); sa += FJ_XSTRIDE; }
          } else {
            if (p0) RPUSH(ch.x);
            if (p1) RPUSH(ch.y);
            if (p2) RPUSH(ch.z);
            if (p3) RPUSH(ch.w);
          }
          node = (w & 2u) ? ((w & 1u) ? ch.w : ch.z) : ((w & 1u) ? ch.y : ch.x);
        } else RPOP(node);
        // speculative traversal: park the first triangle leaf and keep descending (the other transitions wait for FJ_TRANSIT
        // below: doing them here, at the end of every node step, costs 10 % more instructions at 4-5 lanes — measured 292.6
        // against 285.4 ms)
        if (node < 0 && (st & (XS_BLAS | XS_LEAF)) == XS_BLAS && node != SENTINEL) {
          S.leaf[tid] = node; st |= XS_LEAF | (((unsigned)~node & 7u) << XS_CNT_SHIFT); RPOP(node);
        }
      }
    }
    FJ_TRANSIT()                                      // once per outer iteration: leave finished BLASes, retire finished rays, park second leaves

    // ---- phase B1: the warp's parked leaves hold W (ray, triangle) pairs on typically 10-12 lanes; the pairs are dealt out 32
    // at a time to ALL lanes (the FP64 ray of any lane is in shared memory), and each owner then folds the hits among its pairs
    // into its best hit in triangle order — the same comparisons in the same order as a per-owner loop, so the result is
    // bit-identical (k_extend2 with FJGPU_COOP=0 keeps that loop as the cross-check).
    // Which heavy phases run is decided by how many lanes wait for them (FJGPU_B1_MIN pairs, FJGPU_B2_MIN lanes): a phase costs
    // the same ~300-450 instructions for 3 lanes as for 30.  Whenever lanes are blocked at least one of the two runs.
    const unsigned lm = __ballot_sync(FULL, st & XS_LEAF);
    const bool want_e = node < 0 && node != IDLE && node != DONE && !(st & (XS_BLAS | XS_LEAF));      // leaf of the instance tree or FIXUP, nothing parked
    const unsigned em = __ballot_sync(FULL, want_e);
    const int pairs_waiting = __reduce_add_sync(FULL, (st & XS_LEAF) ? (int)((st >> XS_CNT_SHIFT) & 7u) + 1 : 0);
    const bool run_b1 = lm != 0u && (pairs_waiting >= a.b1_min || __popc(em) < a.b2_min);
    const bool run_e = em != 0u && (__popc(em) >= a.b2_min || !run_b1);
    if (run_b1) {
      int cnt = 0;
      if (st & XS_LEAF) { cnt = (int)((st >> XS_CNT_SHIFT) & 7u) + 1; st &= ~(unsigned)(XS_LEAF | (7u << XS_CNT_SHIFT)); }
      int off = cnt;                                           // inclusive, then exclusive prefix sum of the pair counts
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) { const int up = __shfl_up_sync(FULL, off, dlt); if (lane >= dlt) off += up; }
      const int total = __shfl_sync(FULL, off, 31);
      off -= cnt;
      unsigned char *pm = pairmap[tid >> 5];
      for (int k = 0; k < cnt; k++) pm[off + k] = (unsigned char)(lane | (k << 5));
      __syncwarp();
      n_tris += (unsigned)total; n_pr += 1u + (((unsigned)(total + 31) >> 5) << 16);
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
          load_leaf_triangle3(a, rays, S.tri[otid], ost, S.ridx[otid], first + k, &v0, &v1, &v2, &prim);
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
            // RayInRange (src/fj_ray.h:29-32): tmin <= t <= tmax; best_t starts at tmax, so `t <= best_t` is the upper test
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
                stg256(a.hits + S.ridx[tid], hr);
                tf = __double2float_ru(ht);
              }
            }
          }
        }
      }
      __syncwarp();                                            // pairmap is rewritten by the next leaf phase
      // ---- any-hit shadow rays: an accepted hit ends the walk; then the transitions of the lanes the leaf phase has unblocked
      if ((st & (XS_ANY | XS_FOUND)) == (XS_ANY | XS_FOUND)) { node = DONE; st &= ~(unsigned)XS_BLAS; }
      FJ_TRANSIT()
    }

    // ---- phase E: enter an instance (or rebuild the world-space box ray on the way back into the instance tree)
    if (run_e) {
      if (want_e) {
        if (node == FIXUP) {
          D3 o, d; ld_ray_od(rays + S.ridx[tid], &o, &d);
          const DGroup grp = sc.groups[st >> XS_TARGET_SHIFT];
          nodes = (const char *)grp.nodesq;
          box_axis(o.x, d.x, grp.bmagq, &ix, &cnx, &cfx);
          box_axis(o.y, d.y, grp.bmagq, &iy, &cny, &cfy);
          box_axis(o.z, d.z, grp.bmagq, &iz, &cnz, &cfz);
          RPOP(node);                              // the inner node FJ_TRANSIT pushed back
        } else {                                   // TLAS leaf: enter the first instance, re-queue the others
          const int ref = ~node;
          const int first = ref >> 3, cnt = (ref & 7) + 1;
          for (int k = cnt - 1; k >= 1; k--) RPUSH(~(((first + k) << 3) | 0));
          const RayRec *r = rays + S.ridx[tid];
          const DInstRec &in = sc.groups[st >> XS_TARGET_SHIFT].irec[first];      // one record: matrix, tree, packets (no order[] -> instance -> mesh chain)
          S.cur_inst[tid] = in.inst;
          if (!(st & XS_INIT)) {                                         // first instance of this ray: its exact range (the lane carries only the FP32 roundings)
            const double2 range = ld_ray_range(r);
            S.tmin[tid] = range.x; S.best_t[tid] = range.y;
          }
          const double *inv = in.inv;
          if (in.motion) inv = in.motion + 24 * (size_t)r->key;           // time-sampled transform: the ray's entry of the time table
          D3 wo, wd; ld_ray_od(r, &wo, &wd);
          const D3 o = mat_point(inv, wo);
          const D3 d = mat_vector(inv, wd);
          S.ox[tid] = o.x; S.oy[tid] = o.y; S.oz[tid] = o.z; S.dx[tid] = d.x; S.dy[tid] = d.y; S.dz[tid] = d.z;
          box_axis(o.x, d.x, in.bmagq, &ix, &cnx, &cfx);
          box_axis(o.y, d.y, in.bmagq, &iy, &cny, &cfy);
          box_axis(o.z, d.z, in.bmagq, &iz, &cnz, &cfz);
          nodes = in.nodesq;
          S.tri[tid] = in.tri;
          st = (st & ~(unsigned)(XS_TRI64 | XS_VEL)) | XS_BLAS | XS_INIT | (in.tri64 == 1 ? XS_TRI64 : 0) | (in.tri64 == 2 ? XS_VEL : 0);
          RPUSH(SENTINEL);
          node = 0;
        }
      }
    }
  }
  // traversal statistics (4-wide node steps and exact triangle tests) for DESIGN.md / bench.py
  if (lane == 0 && a.counters) {
    atomicAdd(&a.counters->node_steps, (unsigned long long)n_steps); atomicAdd(&a.counters->tri_tests, (unsigned long long)n_tris);
    atomicAdd(&a.counters->leaf_phases, (unsigned long long)(n_pr & 0xffffu)); atomicAdd(&a.counters->leaf_rounds, (unsigned long long)(n_pr >> 16));
  }
#undef RPUSH
#undef RPOP
#undef FJ_TRANSIT
}

}  // namespace fj
