// fj_quant.h — quantisation of ONE 4-wide node (Node128 -> NodeQ64, layout in fj_bvh.h), shared by the host builder
// (fj_bvh.cc) and the device builder (fj_build.cu).  Decoded boxes contain the FP32 boxes they come from: qlo is rounded
// down, qhi up, and every plane is verified in FP64.
#ifndef FJ_QUANT_H
#define FJ_QUANT_H

#include <math.h>
#include <string.h>
#include "fj_bvh.h"

#ifdef __CUDACC__
#define FJ_HD __host__ __device__
#else
#define FJ_HD
#endif

namespace fjb {

// Returns false when the node cannot be represented (extent beyond 254 * 2^60).  *mag is raised to the largest
// |coordinate| of any decoded plane.
FJ_HD inline bool quantize_one(const Node128 &w, NodeQ64 &q, double *mag) {
  for (int k = 0; k < 16; k++) q.w[k] = 0;
  const float *lo[3] = {w.lox, w.loy, w.loz}, *hi[3] = {w.hix, w.hiy, w.hiz};
  bool valid[4]; int nvalid = 0;
  for (int k = 0; k < 4; k++) { valid[k] = !(w.lox[k] > w.hix[k]) && w.lox[k] < 3e38f; nvalid += valid[k]; }
  uint32_t sbits[3];
  for (int a = 0; a < 3; a++) {
    double plo = 1e300, phi = -1e300;
    for (int k = 0; k < 4; k++) if (valid[k]) { plo = fmin(plo, (double)lo[a][k]); phi = fmax(phi, (double)hi[a][k]); }
    if (nvalid == 0) { plo = phi = 0; }
    const float p = (float)plo;                         // exact: plo is one of the float planes
    // smallest power of two s with extent <= 254 s (one step of headroom for the outward rounding below)
    const double extent = phi - plo;
    int e = -100;
    if (extent > 0) { int ex; frexp(extent / 254.0, &ex); e = ex; if (ldexp(1.0, e - 1) * 254.0 >= extent) e--; }
    if (e < -60) e = -60;                               // keeps s / d inside the normal FP32 range for any direction
    uint32_t qlo = 0, qhi = 0;
    double sc = 0;
    for (;; e++) {                                      // (one more step only if the outward rounding ran out of range)
      if (e > 60) return false;
      sc = ldexp(1.0, e);
      qlo = qhi = 0;
      bool fits = true;
      for (int k = 0; k < 4 && fits; k++) {
        int l = 255, h = 0;                             // unused slot: inverted box
        if (valid[k]) {
          l = (int)floor(((double)lo[a][k] - plo) / sc); h = (int)ceil(((double)hi[a][k] - plo) / sc);
          if (l > 255) l = 255;
          if (h < 0) h = 0;
          while (l > 0 && plo + l * sc > (double)lo[a][k]) l--;
          while (h < 255 && plo + h * sc < (double)hi[a][k]) h++;
          if (l < 0) l = 0;
          if (plo + h * sc < (double)hi[a][k]) fits = false;
        }
        qlo |= (uint32_t)l << (8 * k); qhi |= (uint32_t)h << (8 * k);
      }
      if (fits) break;
    }
    for (int k = 0; k < 4; k++) if (valid[k])
      *mag = fmax(*mag, fmax(fabs(plo + ((qlo >> (8 * k)) & 255) * sc), fabs(plo + ((qhi >> (8 * k)) & 255) * sc)));
    memcpy(&q.w[a], &p, 4);
    const float sf = (float)sc; memcpy(&sbits[a], &sf, 4);
    q.w[a == 0 ? 4 : (a == 1 ? 6 : 8)] = qlo; q.w[a == 0 ? 5 : (a == 1 ? 7 : 9)] = qhi;
  }
  q.w[3] = (sbits[0] & 0xffff0000u) | (sbits[1] >> 16);
  q.w[14] = sbits[2] & 0xffff0000u;
  for (int k = 0; k < 4; k++) q.w[10 + k] = (uint32_t)w.c[k];
  return true;
}

}  // namespace fjb
#endif
