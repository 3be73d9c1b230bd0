// fj_extend_quad.cuh — k_extend3: the closest-hit kernel of the wavefront, one ray per QUAD of lanes.
//
// Why.  k_extend / k_extend2 give every lane its own ray, so every lane fetches its own 128-B node: the L1 data pipe
// serves 16 B per lane per pass, i.e. 7-8 passes ("wavefronts") per node step, and ncu shows that pipe at 87-91 % of
// peak while the issue slots are half idle (profiles/r1_k_extend2_ncu_full.txt).  tools/micro/l1_wavefronts.cu
// measures the ceiling of that access pattern at 0.14 node fetches / cycle / SM however small the working set, and
// 0.54-0.91 when four adjacent lanes read the four 32-B quarters of one line (one pass serves all four).
//
// So here a warp walks 8 rays.  The four lanes of a quad own the four children of the current node:
//   node step   lane s loads child s (one 32-B load, Node4Q layout) and slab-tests it; two xor-shuffles find the nearest
//               hit child, one ballot gives the quad's hit mask; every lane whose child is hit but not nearest stores
//               its child reference straight into the ray's stack in shared memory at (sp + rank) — one predicated
//               store instead of four ordered pushes; the winner's reference comes back with one shuffle
//   leaf        lane s tests triangle s of the leaf in exact FP64 (the whole leaf in one pass), a two-step quad
//               reduction picks the closest (ties: higher face id), the winning lane updates the ray's best hit
//   state       per ray, in shared memory: FP64 ray in the traversed space, tmin, best hit, stack.  Per lane, in
//               registers: the FP32 box ray (replicated across the quad), node, sp, a state word.
// The phases are warp-synchronous exactly as in k_extend (refill / node steps with one parked leaf / leaves /
// transitions) and the walk visits the same nodes in the same order, so hits AND traversal counters are identical.
#pragma once

#include "fj_extend.cuh"

namespace fj {

#define FJ_QT 128                 // threads per CTA
#define FJ_QR (FJ_QT / 4)         // rays per CTA

struct QuadShared {
  double ox[FJ_QR], oy[FJ_QR], oz[FJ_QR], dx[FJ_QR], dy[FJ_QR], dz[FJ_QR];    // ray in the space being traversed
  double tmin[FJ_QR], best_t[FJ_QR], best_u[FJ_QR], best_v[FJ_QR];
  const void *tri[FJ_QR];
  int cur_inst[FJ_QR], best_inst[FJ_QR], best_prim[FJ_QR];
  unsigned ridx[FJ_QR];
};

// Moller-Trumbore of fj_device.cuh tri_intersect with the ray read from the quad's shared-memory slot (broadcast reads).
__device__ __forceinline__ bool tri_intersect_q(const D3 &v0, const D3 &v1, const D3 &v2, const volatile QuadShared &V, int rs,
                                                double *t, double *u, double *v) {
  const D3 e1 = v1 - v0, e2 = v2 - v0;
  D3 pvec;
  { const D3 dir = mk(V.dx[rs], V.dy[rs], V.dz[rs]); pvec = cross(dir, e2); }
  const double det = dot(e1, pvec);
  if (det > -1e-6 && det < 1e-6) return false;
  const double inv_det = ddiv(1.0, det);
  const D3 tvec = mk(V.ox[rs], V.oy[rs], V.oz[rs]) - v0;
  *u = dmul(dot(tvec, pvec), inv_det);
  if (*u < 0.0 || *u > 1.0) return false;
  const D3 qvec = cross(tvec, e1);
  { const D3 dir = mk(V.dx[rs], V.dy[rs], V.dz[rs]); *v = dmul(dot(dir, qvec), inv_det); }
  if (*v < 0.0 || dadd(*u, *v) > 1.0) return false;
  *t = dmul(dot(e2, qvec), inv_det);
  return true;
}

// stack_stride: words per ray stack (odd, >= the scene's worst-case need: fjgpu_context::stack_need)
template <int MINB, bool STATS>
__global__ void __launch_bounds__(FJ_QT, MINB) k_extend3(const RenderArgs a, const int stack_stride) {
  __shared__ QuadShared S;
  extern __shared__ int stack_mem[];
  const unsigned FULL = 0xffffffffu, QUADS = 0x11111111u;
  const int tid = threadIdx.x, lane = tid & 31, s = tid & 3, rs = tid >> 2, qbase = lane & 28;
  int *const stack = stack_mem + rs * stack_stride;
  const RayRec *rays = a.queue[a.cur];
  if (blockIdx.x == 0 && threadIdx.x == 0) a.ctl->count[a.cur ^ 1] = 0;      // the queue k_shade fills next
  const unsigned count = min(a.ctl->count[a.cur], a.capacity);
  const DScene &sc = a.sc;
  const int SENTINEL = (int)0x80000000, DONE = (int)0x80000001, IDLE = (int)0x80000002;
  const unsigned MISS = 0xffffffffu;
  const int refill_q = max(1, a.refill >> 2), phase_a_min_q = max(1, a.phase_a_min >> 2);

  bool drained = false;
  unsigned st = 0;
  int sp = 0, node = IDLE, leaf = 0;
  float tn = 0, tf = 0;
  BoxRayMM br; br.ix = br.iy = br.iz = br.lx = br.ly = br.lz = br.hx = br.hy = br.hz = 0.f;
  const char *nodes = nullptr;
  unsigned n_steps = 0, n_tris = 0;

  for (;;) {
    // ---- refill idle quads from the queue head
    const unsigned idle = __ballot_sync(FULL, node == IDLE) & QUADS;
    if (idle == QUADS && drained) break;
    if (!drained && __popc(idle) >= refill_q) {
      const int n = __popc(idle);
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(&a.ctl->head, (unsigned)n);
      base = __shfl_sync(FULL, base, 0);
      if (base + n >= count) drained = true;
      if (node == IDLE) {
        const unsigned i = base + __popc(idle & ((1u << qbase) - 1));
        if (i < count) {
          const unsigned ridx = a.perm ? a.perm[i] : i;
          const RayRec &r = rays[ridx];
          const double tmin = r.tmin, tmax = r.tmax;
          S.ridx[rs] = ridx; S.tmin[rs] = tmin; S.best_t[rs] = tmax; S.best_u[rs] = 0; S.best_v[rs] = 0;
          S.best_inst[rs] = -1; S.best_prim[rs] = -1; S.cur_inst[rs] = -1;
          tn = __double2float_rd(tmin); tf = __double2float_ru(tmax);
          const DGroup grp = sc.groups[r.target];
          nodes = (const char *)grp.nodes4q; sp = 0; node = 0; leaf = 0;
          make_box_ray_mm(mk(r.o[0], r.o[1], r.o[2]), mk(r.d[0], r.d[1], r.d[2]), grp.bmag, br);
          st = XS_WORLD;
        }
      }
      if (STATS && (n_steps | n_tris) > 0x40000000u) {      // keep the 32-bit warp totals from wrapping
        if (lane == 0) { atomicAdd(&a.counters->node_steps, (unsigned long long)n_steps); atomicAdd(&a.counters->tri_tests, (unsigned long long)n_tris); }
        n_steps = n_tris = 0;
      }
      if (__ballot_sync(FULL, node != IDLE) == 0) break;
    }

    // ---- phase A: 4-wide inner nodes, one child per lane, nearest hit child first
    for (;;) {
      const bool want = node >= 0;
      const unsigned wm = __ballot_sync(FULL, want) & QUADS;
      if (wm == 0) break;
      // few quads left descending: switch to the parked leaves / transitions if there are any to work on
      if (__popc(wm) < phase_a_min_q && __any_sync(FULL, (st & XS_LEAF) || (node < 0 && node != IDLE))) break;
      if (STATS) n_steps += __popc(wm);
      if (want && !(st & (XS_BLAS | XS_WORLD))) {      // back in the instance tree after a BLAS: world-space box ray and tree again
        const RayRec &r = rays[S.ridx[rs]];
        const DGroup grp = sc.groups[r.target];
        nodes = (const char *)grp.nodes4q;
        make_box_ray_mm(mk(r.o[0], r.o[1], r.o[2]), mk(r.d[0], r.d[1], r.d[2]), grp.bmag, br);
        st |= XS_WORLD;
      }
      // child s of the node: lo.x lo.y lo.z hi.x | hi.y hi.z ref pad
      F8 c; c.a = make_float4(0.f, 0.f, 0.f, 0.f); c.b = c.a;
      if (want) c = ldg256(nodes + 128 * (size_t)node + 32 * s);
      const float x0 = fmaf(c.a.x, br.ix, br.lx), x1 = fmaf(c.a.w, br.ix, br.hx);
      const float y0 = fmaf(c.a.y, br.iy, br.ly), y1 = fmaf(c.b.x, br.iy, br.hy);
      const float z0 = fmaf(c.a.z, br.iz, br.lz), z1 = fmaf(c.b.y, br.iz, br.hz);
      const float nr = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tn));
      const float fr_ = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tf));
      const bool hit = want && nr <= fr_;
      // entry distances are positive (tn > 0), so their bit patterns order like unsigned integers
      const unsigned key = hit ? ((__float_as_uint(nr) & ~3u) | (unsigned)s) : MISS;
      const unsigned hm = (__ballot_sync(FULL, hit) >> qbase) & 15u;        // hit mask of this quad
      const unsigned k1 = min(key, __shfl_xor_sync(FULL, key, 1));
      const unsigned kmin = min(k1, __shfl_xor_sync(FULL, k1, 2));
      const int ref = __float_as_int(c.b.z);
      const int wref = __shfl_sync(FULL, ref, qbase | (int)(kmin & 3u));    // the nearest hit child
      if (want) {
        if (kmin != MISS) {
          // the other hit children go on the stack in slot order (the order k_extend pushes them in)
          const unsigned others = hm & ~(1u << (kmin & 3u));
          if ((others >> s) & 1u) stack[sp + __popc(others & ((1u << s) - 1u))] = ref;
          sp += __popc(others);
          node = wref;
        }
      }
      __syncwarp();                                     // the quad's pushes are visible to its pops
      if (want) {
        if (kmin == MISS) node = sp > 0 ? stack[--sp] : DONE;
        // speculative traversal: park the first triangle leaf and keep descending.  Inside a BLAS the bottom stack entry is
        // SENTINEL, so a negative reference there is SENTINEL or a triangle leaf and the pop below cannot underflow.
        if (node < 0 && (st & (XS_BLAS | XS_LEAF)) == XS_BLAS && node != SENTINEL) { leaf = node; st |= XS_LEAF; node = stack[--sp]; }
      }
    }

    // ---- phase B1: parked triangle leaves, exact FP64 tests, one triangle per lane: a leaf of up to 4 in one pass
    if (__any_sync(FULL, st & XS_LEAF)) {
      const bool has = (st & XS_LEAF) != 0;
      int first = 0, cnt = 0;
      if (has) { const int ref = ~leaf; first = ref >> 3; cnt = (ref & 7) + 1; st &= ~XS_LEAF; }
      for (int j = 0;; j += 4) {
        const bool mine = has && j + s < cnt;
        const unsigned mm = __ballot_sync(FULL, mine);
        if (mm == 0) break;
        if (STATS) n_tris += __popc(mm);
        double t = FJ_REAL_MAX, u = 0, v = 0; int prim = -1;
        if (mine) {
          D3 v0, v1, v2;
          const void *tp_ = S.tri[rs];
          if (!(st & XS_TRI64)) {
            const float4 *tp = (const float4 *)tp_ + 3 * (size_t)(first + j + s);
            const float4 p0 = __ldg(tp), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
            v0 = mk(p0.x, p0.y, p0.z); v1 = mk(p1.x, p1.y, p1.z); v2 = mk(p2.x, p2.y, p2.z); prim = __float_as_int(p0.w);
          } else {
            const double *p = (const double *)tp_ + 10 * (size_t)(first + j + s);
            v0 = mk(__ldg(p), __ldg(p + 1), __ldg(p + 2)); v1 = mk(__ldg(p + 3), __ldg(p + 4), __ldg(p + 5)); v2 = mk(__ldg(p + 6), __ldg(p + 7), __ldg(p + 8));
            prim = (int)__double_as_longlong(__ldg(p + 9));
          }
          double tt;
          // RayInRange (src/fj_ray.h:29-32): tmin <= t <= tmax; best_t starts at tmax and only decreases
          if (tri_intersect_q(v0, v1, v2, S, rs, &tt, &u, &v) && S.tmin[rs] <= tt && tt <= S.best_t[rs]) t = tt; else prim = -1;
        }
        // quad reduction: smallest t, exact ties to the higher face id; lanes without a candidate carry prim = -1
        int wl = s;
#pragma unroll
        for (int m = 1; m <= 2; m <<= 1) {
          const double ot = __shfl_xor_sync(FULL, t, m);
          const int op = __shfl_xor_sync(FULL, prim, m), ol = __shfl_xor_sync(FULL, wl, m);
          if (op >= 0 && (prim < 0 || ot < t || (ot == t && op > prim))) { t = ot; prim = op; wl = ol; }
        }
        bool better = false;
        if (has && prim >= 0) {                          // quad-uniform: the leaf's best candidate against the ray's best hit
          better = S.best_inst[rs] < 0 || t < S.best_t[rs];
          if (!better) {                                 // t == best_t: lower instance, then higher face id
            const int ci = S.cur_inst[rs], bi = S.best_inst[rs];
            better = ci < bi || (ci == bi && prim > S.best_prim[rs]);
          }
        }
        __syncwarp();                                    // every lane of the quad has read the old best hit
        if (better) {
          if (s == wl) { S.best_t[rs] = t; S.best_u[rs] = u; S.best_v[rs] = v; S.best_prim[rs] = prim; S.best_inst[rs] = S.cur_inst[rs]; }
          tf = __double2float_ru(t);
        }
        __syncwarp();
      }
    }

    // ---- phase B2: transitions of quads whose next stack entry is not an inner node
    const bool special = node < 0 && node != IDLE;
    if (__any_sync(FULL, special)) {
      if (special) {
        if (node == DONE) {                        // traversal finished: the quad writes the 32-B hit record, 8 B per lane
          const int bi = S.best_inst[rs];
          double w;
          if (s == 0) w = bi >= 0 ? S.best_t[rs] : FJ_REAL_MAX;
          else if (s == 1) w = S.best_u[rs];
          else if (s == 2) w = S.best_v[rs];
          else w = __hiloint2double(bi, S.best_prim[rs]);       // {int prim, int inst}: prim in the low word
          __stcs(reinterpret_cast<double *>(a.hits + S.ridx[rs]) + s, w);
          node = IDLE;
        } else if (node == SENTINEL) {             // the instance's BLAS is done: back to the instance tree (its nodes and box
          st &= ~(XS_BLAS | XS_WORLD);             // ray are fetched again only if an inner node of that tree is still to be visited)
          node = sp > 0 ? stack[--sp] : DONE;
        } else if (st & XS_BLAS) {                 // a second triangle leaf: park it now that the slot is free
          leaf = node; st |= XS_LEAF; node = stack[--sp];
        } else {                                   // TLAS leaf: enter the first instance, re-queue the others
          const int ref = ~node;
          const int first = ref >> 3, cnt = (ref & 7) + 1;
          if (s == 0) for (int k = cnt - 1; k >= 1; k--) stack[sp + (cnt - 1 - k)] = ~(((first + k) << 3) | 0);
          sp += cnt - 1;
          const RayRec &r = rays[S.ridx[rs]];
          const int ci = sc.groups[r.target].order[first];
          const DInstance &in = sc.inst[ci];
          const double *inv = inst_inv(in, r.key);
          const D3 o = mat_point(inv, mk(r.o[0], r.o[1], r.o[2]));
          const D3 d = mat_vector(inv, mk(r.d[0], r.d[1], r.d[2]));
          const DMesh &m = sc.meshes[in.mesh];
          make_box_ray_mm(o, d, m.bmag, br);
          nodes = (const char *)m.nodes4q;
          const bool t64 = m.tri32 == nullptr;
          if (s == 0) {
            S.cur_inst[rs] = ci;
            S.ox[rs] = o.x; S.oy[rs] = o.y; S.oz[rs] = o.z; S.dx[rs] = d.x; S.dy[rs] = d.y; S.dz[rs] = d.z;
            S.tri[rs] = t64 ? (const void *)m.tri64 : (const void *)m.tri32;
            stack[sp] = SENTINEL;
          }
          sp++;
          st = XS_BLAS | (t64 ? XS_TRI64 : 0);
          node = 0;
        }
      }
      __syncwarp();                                // shared-memory state written by lane 0 of a quad is visible to the quad
    }
  }
  // traversal statistics (4-wide node steps and exact triangle tests) for DESIGN.md / bench.py
  if (STATS && lane == 0 && a.counters) { atomicAdd(&a.counters->node_steps, (unsigned long long)n_steps); atomicAdd(&a.counters->tri_tests, (unsigned long long)n_tris); }
}

}  // namespace fj
