// Microbenchmark: cost of divergent node fetches on the L1 data pipe.
//   mode 0: one 128-B node per LANE, 7 x LDG.128        (k_extend / k_extend2 before the 256-bit loads)
//   mode 1: one 128-B node per LANE, 3 x LDG.256 + 1 x LDG.128
//   mode 2: one 128-B node per QUAD of lanes, 1 x LDG.256 per lane (lane s reads bytes 32s..32s+31)
//   mode 3: one 128-B node per QUAD, 2 x LDG.128 per lane
//   mode 4: one 64-B node per LANE, 2 x LDG.256
//   mode 5: one 64-B node per LANE, fetched by a per-lane TMA bulk copy (cp.async.bulk, SASS UBLKCP) into shared memory,
//           one mbarrier per warp, then 4 x LDS.128 — does the async proxy get around the 16 B / lane / pass of the L1 pipe?
//   mode 6: as 5, double-buffered (the next node's copy is in flight while the current one is read)
// Reports node fetches per cycle per SM for a working set of `mb` MiB (L1-resident, L2-resident, ...).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
struct F8 { float4 a, b; };
__device__ __forceinline__ F8 ldg256(const void *p) {
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float sum4(float4 v) { return v.x + v.y + v.z + v.w; }
template <int MODE>
__global__ void __launch_bounds__(128) k(const char *nodes, unsigned nnodes, int iters, float *out) {
  const int lane = threadIdx.x & 31;
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  if (MODE == 2 || MODE == 3) s = (blockIdx.x * blockDim.x + (threadIdx.x & ~3)) * 2654435761u + 12345u;   // same stream within a quad
  float acc = 0;
  for (int i = 0; i < iters; i++) {
    s = s * 1664525u + 1013904223u;
    const unsigned n = (s >> 8) % nnodes;
    const char *p = nodes + (size_t)n * (MODE == 4 ? 64 : 128);
    if (MODE == 0) {
      const float4 *q = (const float4 *)p;
      acc += sum4(__ldg(q)) + sum4(__ldg(q + 1)) + sum4(__ldg(q + 2)) + sum4(__ldg(q + 3)) + sum4(__ldg(q + 4)) + sum4(__ldg(q + 5)) + sum4(__ldg(q + 6));
    } else if (MODE == 1) {
      F8 a = ldg256(p), b = ldg256(p + 32), c = ldg256(p + 64); float4 d = __ldg((const float4 *)(p + 96));
      acc += sum4(a.a) + sum4(a.b) + sum4(b.a) + sum4(b.b) + sum4(c.a) + sum4(c.b) + sum4(d);
    } else if (MODE == 2) {
      F8 a = ldg256(p + 32 * (lane & 3));
      acc += sum4(a.a) + sum4(a.b);
    } else if (MODE == 3) {
      const float4 *q = (const float4 *)(p + 32 * (lane & 3));
      acc += sum4(__ldg(q)) + sum4(__ldg(q + 1));
    } else {
      F8 a = ldg256(p), b = ldg256(p + 32);
      acc += sum4(a.a) + sum4(a.b) + sum4(b.a) + sum4(b.b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned phase) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" :: "r"(bar), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <int NBUF>
__global__ void __launch_bounds__(128) k_tma(const char *nodes, unsigned nnodes, int iters, float *out) {
  __shared__ __align__(16) char buf[NBUF][128][80];        // 64-B node per lane, 80-B pitch: conflict-free LDS.128
  __shared__ __align__(8) unsigned long long bars[NBUF][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  if (lane == 0) for (int b = 0; b < NBUF; b++) mbar_init(smem_u32(&bars[b][warp]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  float acc = 0;
  unsigned phase[NBUF]; for (int b = 0; b < NBUF; b++) phase[b] = 0;
  auto issue = [&](int b) {
    s = s * 1664525u + 1013904223u;
    const unsigned n = (s >> 8) % nnodes;
    if (lane == 0) mbar_expect(smem_u32(&bars[b][warp]), 32 * 64);
    bulk_g2s(smem_u32(&buf[b][threadIdx.x][0]), nodes + (size_t)n * 64, 64, smem_u32(&bars[b][warp]));
  };
  if (NBUF == 2) issue(0);
  for (int i = 0; i < iters; i++) {
    const int b = NBUF == 2 ? (i & 1) : 0;
    if (NBUF == 2) { if (i + 1 < iters) issue(b ^ 1); } else issue(0);
    mbar_wait(smem_u32(&bars[b][warp]), phase[b]); phase[b] ^= 1;
    const float4 *q = (const float4 *)&buf[b][threadIdx.x][0];
    acc += sum4(q[0]) + sum4(q[1]) + sum4(q[2]) + sum4(q[3]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic reads before the async proxy overwrites the slot
    __syncwarp();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int NBUF> void run_tma(const char *d, size_t bytes, float *out, int sms, double mhz) {
  const unsigned nnodes = (unsigned)(bytes / 64);
  const int iters = 4000, grid = sms * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_tma<NBUF><<<grid, 128>>>(d, nnodes, 200, out);
  cudaEventRecord(e0);
  k_tma<NBUF><<<grid, 128>>>(d, nnodes, iters, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fetches = (double)grid * 128 * iters, cycles = ms * 1e-3 * mhz * 1e6;
  printf("  mode %d: %8.3f ms  %6.3f node fetches/cycle/SM  (%.1f GB/s of node bytes)  [%s]\n", 4 + NBUF, ms, fetches / cycles / sms,
         fetches * 64 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}
template <int MODE> void run(const char *d, size_t bytes, float *out, int sms, double mhz) {
  const unsigned nnodes = (unsigned)(bytes / (MODE == 4 ? 64 : 128));
  const int iters = 4000, grid = sms * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 128>>>(d, nnodes, 200, out);
  cudaEventRecord(e0);
  k<MODE><<<grid, 128>>>(d, nnodes, iters, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double lanes = (double)grid * 128 * iters;
  const double fetches = (MODE == 2 || MODE == 3) ? lanes / 4 : lanes;
  const double cycles = ms * 1e-3 * mhz * 1e6;
  printf("  mode %d: %8.3f ms  %6.3f node fetches/cycle/SM  (%.1f GB/s of node bytes)\n", MODE, ms, fetches / cycles / sms,
         fetches * (MODE == 4 ? 64 : (MODE == 0 || MODE == 1 ? 112 : 128)) / (ms * 1e-3) / 1e9);
}
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount; const double mhz = pr.clockRate / 1e3;
  float *out; cudaMalloc(&out, sizeof(float) * sms * 8 * 128);
  for (size_t kb : {64, 4096, 32768, 65536, 1048576}) {
    char *d; cudaMalloc(&d, kb << 10); cudaMemset(d, 0, kb << 10);
    printf("working set %zu KiB (SMs %d, %.0f MHz nominal)\n", kb, sms, mhz);
    run<0>(d, kb << 10, out, sms, mhz); run<1>(d, kb << 10, out, sms, mhz); run<2>(d, kb << 10, out, sms, mhz);
    run<3>(d, kb << 10, out, sms, mhz); run<4>(d, kb << 10, out, sms, mhz);
    run_tma<1>(d, kb << 10, out, sms, mhz); run_tma<2>(d, kb << 10, out, sms, mhz);
    cudaFree(d);
  }
  return 0;
}
