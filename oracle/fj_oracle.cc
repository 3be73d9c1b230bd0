// fj_oracle.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (plain C++, FP64, flat arrays) of the reference's hot path
//   Renderer::execute_rendering -> render_tile -> integrate_samples -> SlTrace ->
//   BVH-of-grids closest hit -> Moller-Trumbore -> Shader::Evaluate -> recursive SlTrace
// of tsubo164/Fujiyama-Renderer @ a451548.  Every function cites the reference file:line it
// follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; the product (libfjgpu.so, libfjscene.so) never links or calls it.
//
// Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
//   (1) per-function vectors dumped from the reference's own libscene.so by oracle/ref_probe.cc
//       (tests/golden/ref_vectors.json), and
//   (2) whole-frame .fb images rendered by the unmodified reference binary (tests/golden/*.npz),
// and, when oracle/_ref is built, against the reference binary run live.
//
// Build: g++ -O2 -ffp-contract=off -fPIC -shared -pthread -Iinclude oracle/fj_oracle.cc -o oracle/_build/libfjoracle.so
// (-ffp-contract=off: the reference is built for baseline x86-64 without FMA; keep it that way.)

#include "fjgpu.h"   // struct layouts of the scene description only (shared with the C-ABI)

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <limits>
#include <vector>
#include <array>
#include <algorithm>
#include <thread>
#include <atomic>
#include <map>

namespace {

typedef double Real;
const Real PI = 3.14159265358979323846;            // src/fj_numeric.h:13
const Real REAL_MAX = std::numeric_limits<Real>::max(); // src/fj_numeric.h:14

// ---------------------------------------------------------------- vector (src/fj_vector.h)
struct V3 { Real x, y, z; V3() : x(0), y(0), z(0) {} V3(Real a, Real b, Real c) : x(a), y(b), z(c) {}
  Real operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  Real &at(int i) { return i == 0 ? x : (i == 1 ? y : z); } };
inline V3 operator+(const V3 &a, const V3 &b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3 &a, const V3 &b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(const V3 &a, Real s) { return V3(a.x * s, a.y * s, a.z * s); }      // :293-299
inline V3 operator*(Real s, const V3 &a) { return a * s; }
inline V3 operator*(const V3 &a, const V3 &b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator/(const V3 &a, const V3 &b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline V3 operator/(const V3 &a, Real s) { const Real inv = 1. / s; return a * inv; }   // :306-311
inline Real Dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; } // :318-324
inline V3 Cross(const V3 &a, const V3 &b) {                                            // :326-332
  return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline Real Length(const V3 &a) { return std::sqrt(Dot(a, a)); }                        // :334
inline V3 Normalize(const V3 &a) { const Real len = Length(a); if (len == 0) return a; return a / len; } // :339-345
inline Real Min(Real x, Real y) { return x < y ? x : y; }                               // fj_numeric.h:41-49
inline Real Max(Real x, Real y) { return x > y ? x : y; }
inline Real Clamp(Real x, Real a, Real b) { return x < a ? a : (x > b ? b : x); }
inline Real Radian(Real deg) { return deg * PI / 180.; }                                // :56-59

struct Col { float r, g, b; Col() : r(0), g(0), b(0) {} Col(float a, float c, float d) : r(a), g(c), b(d) {} };
struct Col4 { float r, g, b, a; Col4() : r(0), g(0), b(0), a(0) {} };
inline Col operator*(const Col &A, const Col &B) { return Col(A.r * B.r, A.g * B.g, A.b * B.b); }  // fj_color.h:100-106
inline Col operator*(const Col &A, float s) { return Col(A.r * s, A.g * s, A.b * s); }            // :108-114
inline Col operator*(float s, const Col &A) { return A * s; }
inline Col operator+(const Col &A, const Col &B) { return Col(A.r + B.r, A.g + B.g, A.b + B.b); }
inline float Luminance(const Col &A) { return .298912 * A.r + .586611 * A.g + .114478 * A.b; }    // :275-278

// ---------------------------------------------------------------- RNG (src/fj_random.cc:10-43)
struct XorShift {
  uint32_t s[4];
  XorShift() { s[0] = 123456789; s[1] = 362436069; s[2] = 521288629; s[3] = 88675123; }
  uint32_t NextInteger() {
    uint32_t t = (s[0] ^ (s[0] << 11));
    s[0] = s[1]; s[1] = s[2]; s[2] = s[3];
    s[3] = (s[3] ^ (s[3] >> 19)) ^ (t ^ (t >> 8));
    return s[3];
  }
  double NextFloat01() { return static_cast<double>(NextInteger()) / UINT32_MAX; }
  V3 HollowSphereRand() {                                  // src/fj_random.cc:75-92
    double dot = 0; V3 out;
    for (;;) {
      out.x = 2 * NextFloat01() - 1; out.y = 2 * NextFloat01() - 1; out.z = 2 * NextFloat01() - 1;
      dot = Dot(out, out);
      if (dot > 0 && dot <= 1) break;
    }
    return out / std::sqrt(dot);
  }
};

// Counter RNG shared with the device path (fujiyama-renderer_b200/csrc/fj_rng.cuh restates the
// same published Philox-4x32-10 algorithm, Salmon et al. SC'11).  It exists so the stochastic
// shaders can be compared sample-for-sample between oracle and GPU; the reference itself draws
// from a per-thread sequential XorShift (pathtracing_shader.cc:50,186-188) which no parallel
// renderer can reproduce.
inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// ---------------------------------------------------------------- matrix (src/fj_matrix.cc)
struct Mat { Real e[16]; };
inline void MatIdentity(Mat *m) { for (int i = 0; i < 16; i++) m->e[i] = (i % 5 == 0) ? 1. : 0.; }
inline void MatMultiply(Mat *dst, const Mat &a, const Mat &b) {         // :104-117
  Mat c;
  for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) {
    c.e[4 * j + i] = 0.;
    for (int k = 0; k < 4; k++) c.e[4 * j + i] += a.e[4 * j + k] * b.e[4 * k + i];
  }
  *dst = c;
}
void MatInverse(Mat *dst, const Mat &a) {                               // :119-207 (Cramer's rule)
  Real tmp[12], src[16], det;
  for (int i = 0; i < 4; i++) { src[i] = a.e[i * 4]; src[i + 4] = a.e[i * 4 + 1]; src[i + 8] = a.e[i * 4 + 2]; src[i + 12] = a.e[i * 4 + 3]; }
  tmp[0] = src[10] * src[15]; tmp[1] = src[11] * src[14]; tmp[2] = src[9] * src[15]; tmp[3] = src[11] * src[13];
  tmp[4] = src[9] * src[14]; tmp[5] = src[10] * src[13]; tmp[6] = src[8] * src[15]; tmp[7] = src[11] * src[12];
  tmp[8] = src[8] * src[14]; tmp[9] = src[10] * src[12]; tmp[10] = src[8] * src[13]; tmp[11] = src[9] * src[12];
  Real *d = dst->e;
  d[0] = tmp[0] * src[5] + tmp[3] * src[6] + tmp[4] * src[7];   d[0] -= tmp[1] * src[5] + tmp[2] * src[6] + tmp[5] * src[7];
  d[1] = tmp[1] * src[4] + tmp[6] * src[6] + tmp[9] * src[7];   d[1] -= tmp[0] * src[4] + tmp[7] * src[6] + tmp[8] * src[7];
  d[2] = tmp[2] * src[4] + tmp[7] * src[5] + tmp[10] * src[7];  d[2] -= tmp[3] * src[4] + tmp[6] * src[5] + tmp[11] * src[7];
  d[3] = tmp[5] * src[4] + tmp[8] * src[5] + tmp[11] * src[6];  d[3] -= tmp[4] * src[4] + tmp[9] * src[5] + tmp[10] * src[6];
  d[4] = tmp[1] * src[1] + tmp[2] * src[2] + tmp[5] * src[3];   d[4] -= tmp[0] * src[1] + tmp[3] * src[2] + tmp[4] * src[3];
  d[5] = tmp[0] * src[0] + tmp[7] * src[2] + tmp[8] * src[3];   d[5] -= tmp[1] * src[0] + tmp[6] * src[2] + tmp[9] * src[3];
  d[6] = tmp[3] * src[0] + tmp[6] * src[1] + tmp[11] * src[3];  d[6] -= tmp[2] * src[0] + tmp[7] * src[1] + tmp[10] * src[3];
  d[7] = tmp[4] * src[0] + tmp[9] * src[1] + tmp[10] * src[2];  d[7] -= tmp[5] * src[0] + tmp[8] * src[1] + tmp[11] * src[2];
  tmp[0] = src[2] * src[7]; tmp[1] = src[3] * src[6]; tmp[2] = src[1] * src[7]; tmp[3] = src[3] * src[5];
  tmp[4] = src[1] * src[6]; tmp[5] = src[2] * src[5]; tmp[6] = src[0] * src[7]; tmp[7] = src[3] * src[4];
  tmp[8] = src[0] * src[6]; tmp[9] = src[2] * src[4]; tmp[10] = src[0] * src[5]; tmp[11] = src[1] * src[4];
  d[8] = tmp[0] * src[13] + tmp[3] * src[14] + tmp[4] * src[15];    d[8] -= tmp[1] * src[13] + tmp[2] * src[14] + tmp[5] * src[15];
  d[9] = tmp[1] * src[12] + tmp[6] * src[14] + tmp[9] * src[15];    d[9] -= tmp[0] * src[12] + tmp[7] * src[14] + tmp[8] * src[15];
  d[10] = tmp[2] * src[12] + tmp[7] * src[13] + tmp[10] * src[15];  d[10] -= tmp[3] * src[12] + tmp[6] * src[13] + tmp[11] * src[15];
  d[11] = tmp[5] * src[12] + tmp[8] * src[13] + tmp[11] * src[14];  d[11] -= tmp[4] * src[12] + tmp[9] * src[13] + tmp[10] * src[14];
  d[12] = tmp[2] * src[10] + tmp[5] * src[11] + tmp[1] * src[9];    d[12] -= tmp[4] * src[11] + tmp[0] * src[9] + tmp[3] * src[10];
  d[13] = tmp[8] * src[11] + tmp[0] * src[8] + tmp[7] * src[10];    d[13] -= tmp[6] * src[10] + tmp[9] * src[11] + tmp[1] * src[8];
  d[14] = tmp[6] * src[9] + tmp[11] * src[11] + tmp[3] * src[8];    d[14] -= tmp[10] * src[11] + tmp[2] * src[8] + tmp[7] * src[9];
  d[15] = tmp[10] * src[10] + tmp[4] * src[8] + tmp[9] * src[9];    d[15] -= tmp[8] * src[9] + tmp[11] * src[10] + tmp[5] * src[8];
  det = src[0] * d[0] + src[1] * d[1] + src[2] * d[2] + src[3] * d[3];
  det = 1. / det;
  for (int j = 0; j < 16; j++) d[j] *= det;
}
inline V3 MatPoint(const Mat &m, const V3 &p) {                         // :209-215
  return V3(m.e[0] * p.x + m.e[1] * p.y + m.e[2] * p.z + m.e[3],
            m.e[4] * p.x + m.e[5] * p.y + m.e[6] * p.z + m.e[7],
            m.e[8] * p.x + m.e[9] * p.y + m.e[10] * p.z + m.e[11]);
}
inline V3 MatVector(const Mat &m, const V3 &v) {                        // :217-223
  return V3(m.e[0] * v.x + m.e[1] * v.y + m.e[2] * v.z,
            m.e[4] * v.x + m.e[5] * v.y + m.e[6] * v.z,
            m.e[8] * v.x + m.e[9] * v.y + m.e[10] * v.z);
}

// make_transform_matrix, src/fj_transform.cc:335-391; builders src/fj_matrix.cc:50-102
void make_transform_matrix(int transform_order, int rotate_order,
                           Real tx, Real ty, Real tz, Real rx, Real ry, Real rz,
                           Real sx, Real sy, Real sz, Mat *out) {
  Mat T, R, S, RX, RY, RZ;
  MatIdentity(&T); T.e[3] = tx; T.e[7] = ty; T.e[11] = tz;
  MatIdentity(&S); S.e[0] = sx; S.e[5] = sy; S.e[10] = sz;
  { const Real s = std::sin(Radian(rx)), c = std::cos(Radian(rx)); MatIdentity(&RX); RX.e[5] = c; RX.e[6] = -s; RX.e[9] = s; RX.e[10] = c; }
  { const Real s = std::sin(Radian(ry)), c = std::cos(Radian(ry)); MatIdentity(&RY); RY.e[0] = c; RY.e[2] = s; RY.e[8] = -s; RY.e[10] = c; }
  { const Real s = std::sin(Radian(rz)), c = std::cos(Radian(rz)); MatIdentity(&RZ); RZ.e[0] = c; RZ.e[1] = -s; RZ.e[4] = s; RZ.e[5] = c; }
  Mat *q[3] = {0, 0, 0};
  switch (rotate_order) {          // enum TransformOrder, src/fj_transform.h:15-28
    case 6: q[0] = &RX; q[1] = &RY; q[2] = &RZ; break;   // XYZ
    case 7: q[0] = &RX; q[1] = &RZ; q[2] = &RY; break;   // XZY
    case 8: q[0] = &RY; q[1] = &RX; q[2] = &RZ; break;   // YXZ
    case 9: q[0] = &RY; q[1] = &RZ; q[2] = &RX; break;   // YZX
    case 10: q[0] = &RZ; q[1] = &RX; q[2] = &RY; break;  // ZXY
    default: q[0] = &RZ; q[1] = &RY; q[2] = &RX; break;  // ZYX
  }
  MatIdentity(&R);
  for (int i = 0; i < 3; i++) MatMultiply(&R, *q[i], R);
  switch (transform_order) {
    case 0: q[0] = &S; q[1] = &R; q[2] = &T; break;      // SRT
    case 1: q[0] = &S; q[1] = &T; q[2] = &R; break;      // STR
    case 2: q[0] = &R; q[1] = &S; q[2] = &T; break;      // RST
    case 3: q[0] = &R; q[1] = &T; q[2] = &S; break;      // RTS
    case 4: q[0] = &T; q[1] = &R; q[2] = &S; break;      // TRS
    default: q[0] = &T; q[1] = &S; q[2] = &R; break;     // TSR
  }
  MatIdentity(out);
  for (int i = 0; i < 3; i++) MatMultiply(out, *q[i], *out);
}

// ---------------------------------------------------------------- box (src/fj_box.cc)
struct Box { V3 min, max;
  void ReverseInfinite() { min = V3(REAL_MAX, REAL_MAX, REAL_MAX); max = V3(-REAL_MAX, -REAL_MAX, -REAL_MAX); }
  void Expand(Real d) { min = min - V3(d, d, d); max = max + V3(d, d, d); }
  void AddPoint(const V3 &p) { min.x = Min(min.x, p.x); min.y = Min(min.y, p.y); min.z = Min(min.z, p.z);
                               max.x = Max(max.x, p.x); max.y = Max(max.y, p.y); max.z = Max(max.z, p.z); }
  void AddBox(const Box &o) { min.x = Min(min.x, o.min.x); min.y = Min(min.y, o.min.y); min.z = Min(min.z, o.min.z);
                              max.x = Max(max.x, o.max.x); max.y = Max(max.y, o.max.y); max.z = Max(max.z, o.max.z); }
  bool ContainsPoint(const V3 &p) const {                              // :34-41
    if ((p.x < min.x) || (max.x < p.x)) return false;
    if ((p.y < min.y) || (max.y < p.y)) return false;
    if ((p.z < min.z) || (max.z < p.z)) return false;
    return true; }
  V3 Centroid() const { return .5 * (min + max); }
  V3 Diagonal() const { return max - min; }
};
bool BoxBoxIntersect(const Box &a, const Box &b) {                      // :140-152
  if (a.max.x < b.min.x || a.min.x > b.max.x || a.max.y < b.min.y || a.min.y > b.max.y ||
      a.max.z < b.min.z || a.min.z > b.max.z) return false;
  return true;
}
// BoxRayIntersect, src/fj_box.cc:73-138 (slab test with IEEE divides)
bool BoxRayIntersect(const Box &box, const V3 &o, const V3 &d, Real ray_tmin, Real ray_tmax,
                     Real *hit_tmin, Real *hit_tmax) {
  Real tmin, tmax, tymin, tymax, tzmin, tzmax;
  if (d.x >= 0) { tmin = (box.min.x - o.x) / d.x; tmax = (box.max.x - o.x) / d.x; }
  else          { tmin = (box.max.x - o.x) / d.x; tmax = (box.min.x - o.x) / d.x; }
  if (d.y >= 0) { tymin = (box.min.y - o.y) / d.y; tymax = (box.max.y - o.y) / d.y; }
  else          { tymin = (box.max.y - o.y) / d.y; tymax = (box.min.y - o.y) / d.y; }
  if ((tmin > tymax) || (tymin > tmax)) return false;
  if (tymin > tmin) tmin = tymin;
  if (tymax < tmax) tmax = tymax;
  if (d.z >= 0) { tzmin = (box.min.z - o.z) / d.z; tzmax = (box.max.z - o.z) / d.z; }
  else          { tzmin = (box.max.z - o.z) / d.z; tzmax = (box.min.z - o.z) / d.z; }
  if ((tmin > tzmax) || (tzmin > tmax)) return false;
  if (tzmin > tmin) tmin = tzmin;
  if (tzmax < tmax) tmax = tzmax;
  const bool hit = ((tmin < ray_tmax) && (tmax > ray_tmin));
  if (hit) { *hit_tmin = tmin; *hit_tmax = tmax; }
  return hit;
}
void MatTransformBounds(const Mat &m, Box *b) {                         // src/fj_matrix.cc:225-252
  Box box; box.ReverseInfinite();
  const V3 lo = b->min, hi = b->max;
  const V3 c[8] = { V3(lo.x, lo.y, lo.z), V3(hi.x, lo.y, lo.z), V3(lo.x, hi.y, lo.z), V3(lo.x, lo.y, hi.z),
                    V3(lo.x, hi.y, hi.z), V3(hi.x, lo.y, hi.z), V3(hi.x, hi.y, lo.z), V3(hi.x, hi.y, hi.z) };
  for (int i = 0; i < 8; i++) box.AddPoint(MatPoint(m, c[i]));
  *b = box;
}

// ---------------------------------------------------------------- triangle (src/fj_triangle.cc)
const Real EPSILON = 1e-6;                                              // :12
// TriRayIntersect non-culling branch, src/fj_triangle.cc:81-153 (:127-151)
bool TriRayIntersect(const V3 &v0, const V3 &v1, const V3 &v2, const V3 &orig, const V3 &dir,
                     Real *t, Real *u, Real *v) {
  const V3 edge1 = v1 - v0, edge2 = v2 - v0;
  const V3 pvec = Cross(dir, edge2);
  const Real det = Dot(edge1, pvec);
  if (det > -EPSILON && det < EPSILON) return false;
  const Real inv_det = 1.0 / det;
  const V3 tvec = orig - v0;
  *u = Dot(tvec, pvec) * inv_det;
  if (*u < 0.0 || *u > 1.0) return false;
  const V3 qvec = Cross(tvec, edge1);
  *v = Dot(dir, qvec) * inv_det;
  if (*v < 0.0 || *u + *v > 1.0) return false;
  *t = Dot(edge2, qvec) * inv_det;
  return true;
}

// ---------------------------------------------------------------- scene data
struct Ray { V3 orig, dir; Real tmin, tmax; };
inline V3 RayPointAt(const Ray &r, Real t) { return r.orig + t * r.dir; }          // src/fj_ray.h:24-27
inline bool RayInRange(const Ray &r, Real t) { return r.tmin <= t && t <= r.tmax; } // :29-32

struct Isect {                       // src/fj_intersection.h:21-55
  V3 P, N; int object, prim_id, shading_group_id; Real t_hit; Real u, v;
  float tu, tv;                      // TexCoord uv
  V3 dPdu, dPdv;
  Isect() : object(-1), prim_id(0), shading_group_id(0), t_hit(REAL_MAX), u(0), v(0), tu(0), tv(0) {}
};

struct Mesh {                        // src/fj_mesh.h:200-216 (subset on the path)
  std::vector<V3> P, N; std::vector<int32_t> idx; std::vector<int32_t> group; int nfaces;
  std::vector<float> uv;             // 2 floats per vertex (empty: no point texture)
  std::vector<V3> vel;               // per-vertex velocity (empty: none), Mesh::velocity_, src/fj_mesh.h:210
  Box bounds;                        // Mesh::ComputeBounds, src/fj_mesh.cc:235-244
  // GridAccelerator state, src/fj_grid_accelerator.h
  Box acc_bounds;                    // Accelerator::bounds_ (padded), src/fj_accelerator.cc:60-64
  Box grid_bounds; int nc[3]; V3 cellsize;
  std::vector<int32_t> cell_head;    // cells_[id] -> first list node (-1 = NULL)
  std::vector<int32_t> node_prim, node_next;  // the singly linked Cell lists
  void tri(int f, V3 &a, V3 &b, V3 &c) const { a = P[idx[3 * f]]; b = P[idx[3 * f + 1]]; c = P[idx[3 * f + 2]]; }
  // Mesh::get_primitive_bounds, src/fj_mesh.cc:420-439: with velocity the bounds also hold the vertices at time 1
  void prim_bounds(int f, Box *b) const {
    V3 a, bb, c; tri(f, a, bb, c); b->ReverseInfinite(); b->AddPoint(a); b->AddPoint(bb); b->AddPoint(c);
    if (!vel.empty()) { b->AddPoint(a + vel[idx[3 * f]]); b->AddPoint(bb + vel[idx[3 * f + 1]]); b->AddPoint(c + vel[idx[3 * f + 2]]); }
  }
};

const Real PADDING = .0001;          // src/fj_accelerator.cc:13

// ---------------------------------------------------------------- vertex velocity generator
// Ken Perlin's "improved noise" (SIGGRAPH 2002) as src/fj_noise.cc:33-128 evaluates it: the published 256-entry permutation
// (repeated once), quintic fade, 12 gradient directions from the low 4 hash bits; PerlinNoise sums `octaves` of it with
// amplitude *= persistence and position *= lacunarity; PerlinNoise3d offsets the position for the y and z components.
const unsigned char PERLIN_P[256] = {
  151,160,137,91,90,15,131,13,201,95,96,53,194,233,7,225,140,36,103,30,69,142,8,99,37,240,21,10,23,190,6,148,
  247,120,234,75,0,26,197,62,94,252,219,203,117,35,11,32,57,177,33,88,237,149,56,87,174,20,125,136,171,168,
  68,175,74,165,71,134,139,48,27,166,77,146,158,231,83,111,229,122,60,211,133,230,220,105,92,41,55,46,245,40,
  244,102,143,54,65,25,63,161,1,216,80,73,209,76,132,187,208,89,18,169,200,196,135,130,116,188,159,86,164,100,
  109,198,173,186,3,64,52,217,226,250,124,123,5,202,38,147,118,126,255,82,85,212,207,206,59,227,47,16,58,17,
  182,189,28,42,223,183,170,213,119,248,152,2,44,154,163,70,221,153,101,155,167,43,172,9,129,22,39,253,19,98,
  108,110,79,113,224,232,178,185,112,104,218,246,97,228,251,34,242,193,238,210,144,12,191,179,162,241,81,51,
  145,235,249,14,239,107,49,192,214,31,181,199,106,157,184,84,204,176,115,121,50,45,127,4,150,254,138,236,
  205,93,222,114,67,29,24,72,243,141,128,195,78,66,215,61,156,180};
inline int perlin_perm(int i) { return PERLIN_P[i & 255]; }            // perm[] holds the table twice: indices reach 511
inline Real perlin_fade(Real t) { return t * t * t * (t * (t * 6 - 15) + 10); }
inline Real perlin_lerp(Real t, Real a, Real b) { return a + t * (b - a); }
inline Real perlin_grad(int hash, Real x, Real y, Real z) {
  const int h = hash & 15;
  const Real u = h < 8 ? x : y;
  const Real v = h < 4 ? y : (h == 12 || h == 14 ? x : z);
  return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
}
Real PeriodicNoise3d(Real x, Real y, Real z) {                          // src/fj_noise.cc:90-128
  const int X = (int)std::floor(x) & 255, Y = (int)std::floor(y) & 255, Z = (int)std::floor(z) & 255;
  const Real xx = x - std::floor(x), yy = y - std::floor(y), zz = z - std::floor(z);
  const Real u = perlin_fade(xx), v = perlin_fade(yy), w = perlin_fade(zz);
  const int A = perlin_perm(X) + Y, AA = perlin_perm(A) + Z, AB = perlin_perm(A + 1) + Z;
  const int B = perlin_perm(X + 1) + Y, BA = perlin_perm(B) + Z, BB = perlin_perm(B + 1) + Z;
  return perlin_lerp(w,
      perlin_lerp(v, perlin_lerp(u, perlin_grad(perlin_perm(AA), xx, yy, zz), perlin_grad(perlin_perm(BA), xx - 1, yy, zz)),
                     perlin_lerp(u, perlin_grad(perlin_perm(AB), xx, yy - 1, zz), perlin_grad(perlin_perm(BB), xx - 1, yy - 1, zz))),
      perlin_lerp(v, perlin_lerp(u, perlin_grad(perlin_perm(AA + 1), xx, yy, zz - 1), perlin_grad(perlin_perm(BA + 1), xx - 1, yy, zz - 1)),
                     perlin_lerp(u, perlin_grad(perlin_perm(AB + 1), xx, yy - 1, zz - 1), perlin_grad(perlin_perm(BB + 1), xx - 1, yy - 1, zz - 1))));
}
Real PerlinNoise(const V3 &position, Real lacunarity, Real persistence, int octaves) {     // :33-49
  V3 P = position; Real value = 0, amp = 1;
  for (int i = 0; i < octaves; i++) { value += amp * PeriodicNoise3d(P.x, P.y, P.z); amp *= persistence; P = P * lacunarity; }
  return value;
}
V3 PerlinNoise3d(const V3 &position, Real lacunarity, Real persistence, int octaves) {     // :51-68
  V3 out;
  out.x = PerlinNoise(position, lacunarity, persistence, octaves);
  out.y = PerlinNoise(position + V3(131.977, 21.1823, 71.0231), lacunarity, persistence, octaves);
  out.z = PerlinNoise(position + V3(237.492, 11.1312, 133.129), lacunarity, persistence, octaves);
  return out;
}
inline Real SmoothStep(Real a, Real b, Real x) {                        // src/fj_numeric.h:66-77
  const Real t = (x - a) / (b - a);
  if (t <= 0) return 0;
  if (t >= 1) return 1;
  return t * t * (3 - 2 * t);
}
// generate_velocity of VelocityGeneratorProcedure (procedures/velocity_generator_procedure/velocity_generator_procedure.cc:101-124):
// velocity = .2 (1 - smoothstep(.2, .7, z normalised over the mesh bounds)) * PerlinNoise3d(.2 P, 2, .5, 1), then the
// mesh bounds are recomputed with it (ComputeNormals leaves the normals as they are: positions do not change).
void generate_velocity(Mesh &m) {
  const Real zmin = m.bounds.min.z, zmax = m.bounds.max.z;
  m.vel.assign(m.P.size(), V3(0, 0, 0));
  for (size_t i = 0; i < m.P.size(); i++) {
    const V3 pos = m.P[i];
    const Real znml = (pos.z - zmin) / (zmax - zmin);
    const Real vscale = .2 * (1 - SmoothStep(.2, .7, znml));
    const V3 noise_vec = PerlinNoise3d(.2 * pos, 2, .5, 1);
    m.vel[i] = vscale * noise_vec;
  }
  m.bounds.ReverseInfinite();
  for (int i = 0; i < m.nfaces; i++) { Box b; m.prim_bounds(i, &b); m.bounds.AddBox(b); }
}

// Mesh::ComputeNormals, src/fj_mesh.cc:195-233
void compute_normals(Mesh &m) {
  const int nv = (int)m.P.size();
  m.N.assign(nv, V3(0, 0, 0));
  for (int i = 0; i < m.nfaces; i++) {
    V3 P0, P1, P2; m.tri(i, P0, P1, P2);
    const int i0 = m.idx[3 * i], i1 = m.idx[3 * i + 1], i2 = m.idx[3 * i + 2];
    const V3 N0 = m.N[i0], N1 = m.N[i1], N2 = m.N[i2];
    const V3 Ng = Normalize(Cross(P1 - P0, P2 - P0));   // TriComputeFaceNormal, fj_triangle.cc:35-42
    m.N[i0] = N0 + Ng; m.N[i1] = N1 + Ng; m.N[i2] = N2 + Ng;
  }
  for (int i = 0; i < nv; i++) m.N[i] = Normalize(m.N[i]);
}

// Mesh::ray_intersect, src/fj_mesh.cc:246-308 wrapped by PrimitiveSet::RayIntersect, src/fj_primitive_set.cc:10-26.
// With per-vertex velocity the triangle is tested where it is at the ray's time (:252-259); the shading normal stays the
// mesh's (:273-276).
bool mesh_ray_intersect(const Mesh &m, int prim, const Ray &ray, Real time, Isect *is) {
  V3 P0, P1, P2; m.tri(prim, P0, P1, P2);
  if (!m.vel.empty()) {
    P0 = P0 + time * m.vel[m.idx[3 * prim]]; P1 = P1 + time * m.vel[m.idx[3 * prim + 1]]; P2 = P2 + time * m.vel[m.idx[3 * prim + 2]];
  }
  Real t, u, v;
  if (!TriRayIntersect(P0, P1, P2, ray.orig, ray.dir, &t, &u, &v)) { is->t_hit = REAL_MAX; return false; }
  if (!m.N.empty()) {
    const V3 N0 = m.N[m.idx[3 * prim]], N1 = m.N[m.idx[3 * prim + 1]], N2 = m.N[m.idx[3 * prim + 2]];
    is->N = (1 - u - v) * N0 + u * N1 + v * N2;        // TriComputeNormal, fj_triangle.cc:44-49
  } else is->N = V3(0, 0, 0);
  if (!m.uv.empty()) {                                   // UV = (1-u-v) UV0 + u UV1 + v UV2, fj_mesh.cc:280-291
    const int i0 = m.idx[3 * prim], i1 = m.idx[3 * prim + 1], i2 = m.idx[3 * prim + 2];
    const float tt = 1 - u - v;
    is->tu = tt * m.uv[2 * i0] + u * m.uv[2 * i1] + v * m.uv[2 * i2];
    is->tv = tt * m.uv[2 * i0 + 1] + u * m.uv[2 * i1 + 1] + v * m.uv[2 * i2 + 1];
    // TriComputeDerivatives, src/fj_triangle.cc:51-74
    const V3 dP1 = P1 - P0, dP2 = P2 - P0;
    const float du1 = m.uv[2 * i1] - m.uv[2 * i0], du2 = m.uv[2 * i2] - m.uv[2 * i0];
    const float dv1 = m.uv[2 * i1 + 1] - m.uv[2 * i0 + 1], dv2 = m.uv[2 * i2 + 1] - m.uv[2 * i0 + 1];
    const float determinant = du1 * dv2 - dv1 * du2;
    if (determinant == 0) { is->dPdu = V3(0, 0, 0); is->dPdv = V3(0, 0, 0); }
    else {
      const float invdet = 1. / determinant;
      is->dPdu = (dv2 * dP1 - dv1 * dP2) * invdet;
      is->dPdv = (-du2 * dP1 + du1 * dP2) * invdet;
    }
  } else { is->tu = 0; is->tv = 0; is->dPdu = V3(0, 0, 0); is->dPdv = V3(0, 0, 0); }
  is->P = RayPointAt(ray, t);
  is->object = -1; is->prim_id = prim;
  is->shading_group_id = m.group.empty() ? 0 : m.group[prim];
  is->t_hit = t; is->u = u; is->v = v;
  if (!RayInRange(ray, is->t_hit)) { is->t_hit = REAL_MAX; return false; }
  return true;
}

// GridAccelerator::build, src/fj_grid_accelerator.cc:69-160 ; compute_grid_cellsizes :318-332
void grid_build(Mesh &m) {
  Box b = m.bounds; b.Expand(PADDING);
  m.acc_bounds = b;
  const Real HALF_PADDING = .5 * PADDING;
  const V3 gs = b.Diagonal();
  const Real max_width = Max(Max(gs.x, gs.y), gs.z);
  const Real cube_root = 3 * std::pow(m.nfaces, 1. / 3);
  const Real per = cube_root / max_width;
  for (int a = 0; a < 3; a++) {
    const int n = static_cast<int>(std::floor(gs[a] * per + .5));
    m.nc[a] = (int)Clamp(n, 1, 512);
  }
  const int X = m.nc[0], Y = m.nc[1], Z = m.nc[2];
  m.cell_head.assign((size_t)X * Y * Z, -1);
  m.node_prim.clear(); m.node_next.clear();
  const V3 cs = (b.max - b.min) / V3(X, Y, Z);
  for (int i = 0; i < m.nfaces; i++) {
    Box pb; m.prim_bounds(i, &pb);
    Box tri_bounds = pb;                         // Mesh::box_intersect uses the unpadded triangle bounds (fj_mesh.cc:316-334,405-418)
    pb.Expand(HALF_PADDING);
    int X0 = static_cast<int>(std::floor((pb.min.x - b.min.x) / cs.x));
    int X1 = static_cast<int>(std::floor((pb.max.x - b.min.x) / cs.x) + 1);
    int Y0 = static_cast<int>(std::floor((pb.min.y - b.min.y) / cs.y));
    int Y1 = static_cast<int>(std::floor((pb.max.y - b.min.y) / cs.y) + 1);
    int Z0 = static_cast<int>(std::floor((pb.min.z - b.min.z) / cs.z));
    int Z1 = static_cast<int>(std::floor((pb.max.z - b.min.z) / cs.z) + 1);
    X0 = (int)Clamp(X0, 0, X); X1 = (int)Clamp(X1, 0, X);
    Y0 = (int)Clamp(Y0, 0, Y); Y1 = (int)Clamp(Y1, 0, Y);
    Z0 = (int)Clamp(Z0, 0, Z); Z1 = (int)Clamp(Z1, 0, Z);
    for (int z = Z0; z < Z1; z++) for (int y = Y0; y < Y1; y++) for (int x = X0; x < X1; x++) {
      Box cell; cell.min = b.min + V3(x, y, z) * cs; cell.max = cell.min + cs;   // get_grid_cell :334-343
      if (m.vel.empty()) { if (!BoxBoxIntersect(tri_bounds, cell)) continue; }
      else {                                     // box_tri_intersect, fj_mesh.cc:316-340: the sweep in 8 steps, each bounded by its two ends
        V3 A, B, C; m.tri(i, A, B, C);
        const int N_STEPS = 8;
        const V3 s0 = m.vel[m.idx[3 * i]] / N_STEPS, s1 = m.vel[m.idx[3 * i + 1]] / N_STEPS, s2 = m.vel[m.idx[3 * i + 2]] / N_STEPS;
        bool any = false;
        for (int k = 0; k < N_STEPS && !any; k++) {
          const V3 Q0 = A + k * s0, Q1 = B + k * s1, Q2 = C + k * s2;
          Box sb; sb.ReverseInfinite(); sb.AddPoint(Q0); sb.AddPoint(Q1); sb.AddPoint(Q2);
          sb.AddPoint(Q0 + s0); sb.AddPoint(Q1 + s1); sb.AddPoint(Q2 + s2);
          any = BoxBoxIntersect(sb, cell);
        }
        if (!any) continue;
      }
      const size_t cid = (size_t)z * Y * X + (size_t)y * X + x;
      m.node_prim.push_back(i); m.node_next.push_back(m.cell_head[cid]);         // new cell becomes the list head
      m.cell_head[cid] = (int32_t)m.node_prim.size() - 1;
    }
  }
  m.cellsize = cs; m.grid_bounds = b;
}

// Accelerator::Intersect + GridAccelerator::intersect, src/fj_accelerator.cc:94-113,
// src/fj_grid_accelerator.cc:162-306 (3D-DDA)
bool grid_intersect(const Mesh &m, const Ray &ray, Real time, Isect *isect) {
  Real bt0 = 0, bt1 = 0;
  if (!BoxRayIntersect(m.acc_bounds, ray.orig, ray.dir, ray.tmin, ray.tmax, &bt0, &bt1)) return false;
  Real boxhit_tmin = REAL_MAX, boxhit_tmax = REAL_MAX;
  if (!BoxRayIntersect(m.grid_bounds, ray.orig, ray.dir, ray.tmin, ray.tmax, &boxhit_tmin, &boxhit_tmax)) return false;
  V3 start; Real t_start = REAL_MAX, t_end = REAL_MAX;
  if (m.grid_bounds.ContainsPoint(ray.orig)) { start = ray.orig; t_start = 0; }
  else { t_start = boxhit_tmin; t_end = boxhit_tmax; start = RayPointAt(ray, t_start); }
  t_end = Min(t_end, ray.tmax);
  int cell_id[3], cell_step[3], cell_end[3];
  Real t_next[3], t_delta[3];
  const V3 &gmin = m.grid_bounds.min; const V3 &dir = ray.dir;
  for (int i = 0; i < 3; i++) {
    cell_id[i] = static_cast<int>(std::floor((start[i] - gmin[i]) / m.cellsize[i]));
    cell_id[i] = (int)Clamp(cell_id[i], 0, m.nc[i] - 1);
    if (dir[i] > 0) {
      t_next[i] = t_start + (((cell_id[i] + 1) * m.cellsize[i] + gmin[i]) - start[i]) / dir[i];
      t_delta[i] = m.cellsize[i] / dir[i]; cell_step[i] = +1; cell_end[i] = m.nc[i];
    } else if (dir[i] < 0) {
      t_next[i] = t_start + ((cell_id[i] * m.cellsize[i] + gmin[i]) - start[i]) / dir[i];
      t_delta[i] = -1 * m.cellsize[i] / dir[i]; cell_step[i] = -1; cell_end[i] = -1;
    } else { t_next[i] = REAL_MAX; t_delta[i] = 0; cell_step[i] = 0; cell_end[i] = -1; }
  }
  bool hit = false;
  for (;;) {
    Isect cand[2]; Isect *imin = &cand[0], *itmp = &cand[1];
    imin->t_hit = REAL_MAX;
    const size_t id = (size_t)m.nc[0] * m.nc[1] * cell_id[2] + (size_t)m.nc[0] * cell_id[1] + cell_id[0];
    for (int c = m.cell_head[id]; c != -1; c = m.node_next[c]) {
      const bool hittmp = mesh_ray_intersect(m, m.node_prim[c], ray, time, itmp);
      if (!hittmp) continue;
      Box cell; cell.min = m.grid_bounds.min + V3(cell_id[0], cell_id[1], cell_id[2]) * m.cellsize; cell.max = cell.min + m.cellsize;
      const V3 P_hit = RayPointAt(ray, itmp->t_hit);
      if (!cell.ContainsPoint(P_hit)) continue;
      if (itmp->t_hit < imin->t_hit) { std::swap(imin, itmp); hit = hittmp; }
    }
    if (hit) { *isect = *imin; break; }
    if ((t_next[0] < t_next[1]) && (t_next[0] < t_next[2])) {
      if (t_end < t_next[0]) break;
      cell_id[0] += cell_step[0]; if (cell_id[0] == cell_end[0]) break; t_next[0] += t_delta[0];
    } else if ((t_next[2] < t_next[1])) {
      if (t_end < t_next[2]) break;
      cell_id[2] += cell_step[2]; if (cell_id[2] == cell_end[2]) break; t_next[2] += t_delta[2];
    } else {
      if (t_end < t_next[1]) break;
      cell_id[1] += cell_step[1]; if (cell_id[1] == cell_end[1]) break; t_next[1] += t_delta[1];
    }
  }
  return hit;
}

// PropertySampleList / TransformSampleList (src/fj_property.h:117-137, src/fj_transform.h): up to 8 (x, y, z, time) samples
// per channel, kept sorted by time.
struct TimeSamples { std::vector<std::array<Real, 4>> v; };
struct XfSamples { int torder, rorder; TimeSamples T, R, S; bool moving; XfSamples() : torder(0), rorder(0), moving(false) {} };
inline Real Fit(Real x, Real src0, Real src1, Real dst0, Real dst1) {    // src/fj_numeric.h:84-93
  if (x <= src0) return dst0;
  if (x >= src1) return dst1;
  return dst0 + (dst1 - dst0) * ((x - src0) / (src1 - src0));
}
// PropLerpSamples, src/fj_property.cc:317-345 (VEC4_LERP :18-23)
void lerp_samples(const TimeSamples &l, Real time, Real out[3]) {
  const size_t n = l.v.size();
  if (l.v[0][3] >= time || n == 1) { for (int k = 0; k < 3; k++) out[k] = l.v[0][k]; return; }
  if (l.v[n - 1][3] <= time) { for (int k = 0; k < 3; k++) out[k] = l.v[n - 1][k]; return; }
  for (size_t i = 0; i < n; i++) {
    if (l.v[i][3] == time) { for (int k = 0; k < 3; k++) out[k] = l.v[i][k]; return; }
    if (l.v[i][3] > time) {
      const Real t = Fit(time, l.v[i - 1][3], l.v[i][3], 0, 1);
      for (int k = 0; k < 3; k++) out[k] = (1 - t) * l.v[i - 1][k] + t * l.v[i][k];
      return;
    }
  }
}
// XfmLerpTransformSample, src/fj_transform.cc:306-322: lerp T, R, S at `time`, rebuild the matrix and its inverse
void lerp_transform(const XfSamples &x, Real time, Mat *fwd, Mat *inv) {
  Real T[3], R[3], S[3];
  lerp_samples(x.T, time, T); lerp_samples(x.R, time, R); lerp_samples(x.S, time, S);
  make_transform_matrix(x.torder, x.rorder, T[0], T[1], T[2], R[0], R[1], R[2], S[0], S[1], S[2], fwd);
  MatInverse(inv, *fwd);
}

struct Instance { int mesh; Mat fwd, inv; Box bounds; int shader_of_group[FJGPU_MAX_SHADING_GROUPS];
                  int reflect_target, refract_target, shadow_target; XfSamples xs; };

// BVHAccelerator over an ObjectSet, src/fj_bvh_accelerator.cc:39-56,79-107,253-334
struct BvhNode { int left, right; Box bounds; int prim_id; BvhNode() : left(-1), right(-1), prim_id(-1) {} };
struct Group { std::vector<int> inst; Box acc_bounds; std::vector<BvhNode> nodes; int root; Group() : root(-1) {} };
struct BvhPrim { Box bounds; V3 centroid; int index; };

int find_median(BvhPrim **prims, int begin, int end, int axis) {        // :314-334
  int low = begin, high = end - 1, mid = -1;
  const Real key = (prims[low]->centroid[axis] + prims[high]->centroid[axis]) / 2;
  while (low != mid) {
    mid = (low + high) / 2;
    if (key < prims[mid]->centroid[axis]) high = mid;
    else if (prims[mid]->centroid[axis] < key) low = mid;
    else break;
  }
  return mid + 1;
}
int build_bvh(Group &g, BvhPrim **p, int begin, int end, int axis) {    // :253-296
  const int me = (int)g.nodes.size(); g.nodes.push_back(BvhNode());
  if (end - begin == 1) { g.nodes[me].prim_id = p[begin]->index; g.nodes[me].bounds = p[begin]->bounds; return me; }
  std::sort(p + begin, p + end, [axis](BvhPrim *a, BvhPrim *b) { return a->centroid[axis] < b->centroid[axis]; });
  const int median = find_median(p, begin, end, axis);
  const int na = (axis + 1) % 3;
  const int l = build_bvh(g, p, begin, median, na);
  const int r = build_bvh(g, p, median, end, na);
  g.nodes[me].left = l; g.nodes[me].right = r;
  g.nodes[me].bounds = g.nodes[l].bounds; g.nodes[me].bounds.AddBox(g.nodes[r].bounds);
  return me;
}

struct Shader { fjgpu_shader d; };
struct Light { fjgpu_light d; Mat fwd; std::vector<V3> dome_dir; std::vector<Col> dome_col; XorShift rng; };

// Texture over a `.mip` file's tiles (src/fj_texture.cc, src/fj_mipmap.cc:124-180)
struct Tex { int width, height, nch, tilesize, xnt, ynt; std::vector<float> tiles; };
struct Scene {
  std::vector<Tex> textures;
  std::map<int, Mesh> meshes; std::vector<Instance> inst; std::vector<Group> groups;
  std::vector<Shader> shaders; std::vector<Light> lights; fjgpu_camera cam; bool built;
  XfSamples cam_xs; Real time_range[2];                 // Renderer::SetSampleTimeRange(0, 1), src/fj_renderer.cc:381
  Scene() : built(false) { memset(&cam, 0, sizeof cam); time_range[0] = 0; time_range[1] = 1; }
};

// ObjectInstance::RayIntersect, src/fj_object_instance.cc:213-243 (with a single sample per channel
// XfmLerpTransformSample rebuilds the same matrix for every ray; with several it is rebuilt at the ray's time)
bool instance_ray_intersect(const Scene &s, int iid, const Ray &ray, Real time, Isect *isect) {
  const Instance &in = s.inst[iid];
  Mat fwd_t, inv_t;
  if (in.xs.moving) lerp_transform(in.xs, time, &fwd_t, &inv_t);
  const Mat &fwd = in.xs.moving ? fwd_t : in.fwd, &inv = in.xs.moving ? inv_t : in.inv;
  Ray ro = ray;
  ro.orig = MatPoint(inv, ray.orig);
  ro.dir = MatVector(inv, ray.dir);
  const Mesh &m = s.meshes.find(in.mesh)->second;
  if (!grid_intersect(m, ro, time, isect)) return false;
  isect->P = MatPoint(fwd, isect->P);
  isect->N = Normalize(MatVector(fwd, isect->N));
  isect->dPdu = MatVector(fwd, isect->dPdu); isect->dPdv = MatVector(fwd, isect->dPdv);     // fj_object_instance.cc:237-238
  isect->object = iid;
  return true;
}

// Accelerator::Intersect + intersect_bvh_loop, src/fj_accelerator.cc:94-113, src/fj_bvh_accelerator.cc:164-241
bool group_intersect(const Scene &s, const Group &g, const Ray &ray, Real time, Isect *isect) {
  Real t0, t1;
  if (!BoxRayIntersect(g.acc_bounds, ray.orig, ray.dir, ray.tmin, ray.tmax, &t0, &t1)) return false;
  if (g.root < 0) return false;
  bool hit = false;
  int node = g.root; std::vector<int> stack;
  Isect cand[2]; Isect *imin = &cand[0], *itmp = &cand[1];
  for (;;) {
    const BvhNode &n = g.nodes[node];
    if (n.left == -1 && n.right == -1 && n.prim_id != -1) {
      bool hittmp = instance_ray_intersect(s, g.inst[n.prim_id], ray, time, itmp);
      if (!hittmp) itmp->t_hit = REAL_MAX;                      // PrimitiveSet::RayIntersect
      else if (!RayInRange(ray, itmp->t_hit)) { itmp->t_hit = REAL_MAX; hittmp = false; }
      if (hittmp && itmp->t_hit < imin->t_hit) { std::swap(imin, itmp); hit = hittmp; }
      if (stack.empty()) break;
      node = stack.back(); stack.pop_back();
      continue;
    }
    const bool hl = BoxRayIntersect(g.nodes[n.left].bounds, ray.orig, ray.dir, ray.tmin, ray.tmax, &t0, &t1);
    const bool hr = BoxRayIntersect(g.nodes[n.right].bounds, ray.orig, ray.dir, ray.tmin, ray.tmax, &t0, &t1);
    if (!hl && !hr) { if (stack.empty()) break; node = stack.back(); stack.pop_back(); }
    else if (hl && !hr) node = n.left;
    else if (!hl && hr) node = n.right;
    else { stack.push_back(n.right); node = n.left; }
  }
  if (hit) *isect = *imin;
  return hit;
}

// ---------------------------------------------------------------- shading (src/fj_shading.cc)
enum { CXT_CAMERA_RAY = 0, CXT_SHADOW_RAY, CXT_DIFFUSE_RAY, CXT_REFLECT_RAY, CXT_REFRACT_RAY };
struct Cxt { int ray_context, diffuse_depth, reflect_depth, refract_depth, max_diffuse_depth, max_reflect_depth,
             max_refract_depth, cast_shadow; float opacity_threshold; int trace_target;
             uint64_t node;      // node = path-tree code for the counter RNG (not in the reference)
             Real time; };       // TraceContext::time: every context is a copy of its parent's (src/fj_shading.cc:226-289)

struct RenderState {
  const Scene *s; int rng_mode;      // 0 = counter (Philox), 1 = reference sequential XorShift (single thread)
  uint32_t seed; uint32_t tile_id; uint32_t sample_id;
  std::vector<XorShift> *pt_rng;     // pathtracing_shader.cc:50 `mutable XorShift rng[64]` is a member of EACH shader
                                     // instance: one sequential stream per shader slot (thread id 0)
  int cur_shader;
  std::vector<Light> *lights;        // mutable light RNGs (fj_rectangle_light.cc:36)
  uint64_t rays[5];
};

// counter-RNG draw: dimension `dim` of path node `node` of the current sample
inline double ctr_rand(const RenderState &rs, uint64_t node, uint32_t dim) {
  uint32_t c[4] = { (uint32_t)node, (uint32_t)(node >> 32), dim >> 2, 0x46554a49u };
  philox4x32_10(c, rs.seed ^ (rs.tile_id * 0x9E3779B1u), rs.sample_id);
  return static_cast<double>(c[dim & 3]) / UINT32_MAX;
}

inline V3 SlFaceforward(const V3 &I, const V3 &N) { if (Dot(I, N) < 0) return N; return V3(-N.x, -N.y, -N.z); } // :42-52
double SlFresnel(const V3 &I, const V3 &N, double ior) {                 // :54-74
  double eta, cosv = -1 * Dot(I, N);
  if (cosv > 0) eta = ior; else { eta = 1. / ior; cosv *= -1; }
  const double k2 = .0;
  const double F0 = ((1. - eta) * (1. - eta) + k2) / ((1. + eta) * (1. + eta) + k2);
  return F0 + (1. - F0) * std::pow(1. - cosv, 5.);
}
inline V3 SlReflect(const V3 &I, const V3 &N) {                          // :92-100
  const double c = -1 * Dot(I, N);
  return V3(I.x + 2 * c * N.x, I.y + 2 * c * N.y, I.z + 2 * c * N.z);
}
V3 SlRefract(const V3 &I, const V3 &N, double ior) {                     // :102-138
  V3 n; double eta, cos1 = -1 * Dot(I, N);
  if (cos1 < 0) { cos1 *= -1; eta = 1 / ior; n = V3(-N.x, -N.y, -N.z); } else { eta = ior; n = N; }
  const double radicand = 1 - eta * eta * (1 - cos1 * cos1);
  if (radicand < 0.) return SlReflect(I, N);
  const double ncoeff = eta * cos1 - std::sqrt(radicand);
  return V3(eta * I.x + ncoeff * n.x, eta * I.y + ncoeff * n.y, eta * I.z + ncoeff * n.z);
}

int SlTrace(RenderState &rs, const Cxt &cxt, const V3 &o, const V3 &d, double tmin, double tmax, Col4 *out, double *t_hit);

struct LightSample { int light; V3 P, N; Col color; };   // src/fj_light.h:20-29

// Light::GetSamples for every light of the object, SlNewLightSamples src/fj_shading.cc:380-404
void new_light_samples(RenderState &rs, const Cxt &cxt, std::vector<LightSample> &out) {
  std::vector<Light> &L = *rs.lights;
  uint32_t dim = 16;
  for (size_t li = 0; li < L.size(); li++) {
    Light &lt = L[li];
    if (lt.d.kind == FJGPU_LIGHT_POINT) {                // src/fj_point_light.cc:21-39
      LightSample s; s.light = (int)li; s.P = V3(lt.d.translate[0], lt.d.translate[1], lt.d.translate[2]); s.N = V3(0, 0, 0);
      out.push_back(s);
    } else if (lt.d.kind == FJGPU_LIGHT_GRID) {          // src/fj_rectangle_light.cc:26-47
      const V3 N_sample = Normalize(MatVector(lt.fwd, V3(0, 1, 0)));
      for (int i = 0; i < lt.d.sample_count; i++) {
        Real x, z;
        if (rs.rng_mode == 1) { x = lt.rng.NextFloat01() - .5; z = lt.rng.NextFloat01() - .5; }
        else { x = ctr_rand(rs, cxt.node, dim) - .5; z = ctr_rand(rs, cxt.node, dim + 1) - .5; dim += 2; }
        LightSample s; s.light = (int)li; s.P = MatPoint(lt.fwd, V3(x, 0, z)); s.N = N_sample; out.push_back(s);
      }
    } else if (lt.d.kind == FJGPU_LIGHT_SPHERE) {        // src/fj_sphere_light.cc:22-46
      for (int i = 0; i < lt.d.sample_count; i++) {
        V3 P;
        if (rs.rng_mode == 1) P = lt.rng.HollowSphereRand();
        else {   // same rejection loop, counter-driven
          double dot = 0;
          for (;;) { P.x = 2 * ctr_rand(rs, cxt.node, dim) - 1; P.y = 2 * ctr_rand(rs, cxt.node, dim + 1) - 1; P.z = 2 * ctr_rand(rs, cxt.node, dim + 2) - 1; dim += 3;
                     dot = Dot(P, P); if (dot > 0 && dot <= 1) break; }
          P = P / std::sqrt(dot);
        }
        LightSample s; s.light = (int)li; s.P = MatPoint(lt.fwd, P); s.N = Normalize(MatVector(lt.fwd, P)); out.push_back(s);
      }
    } else {                                             // src/fj_dome_light.cc:25-52
      for (int i = 0; i < lt.d.sample_count && i < (int)lt.dome_dir.size(); i++) {
        LightSample s; s.light = (int)li;
        s.P = MatPoint(lt.fwd, lt.dome_dir[i] * (Real)FLT_MAX);
        s.N = MatVector(lt.fwd, -1 * lt.dome_dir[i]);
        s.color = lt.dome_col[i]; out.push_back(s);
      }
    }
  }
}

Col light_illuminate(const Light &lt, const LightSample &s, const V3 &Ps) {
  const Col color(lt.d.color[0], lt.d.color[1], lt.d.color[2]);
  const float sample_intensity = lt.d.intensity / std::max(lt.d.sample_count, 1);  // Light::SetIntensity, src/fj_light.cc:38-50
  switch (lt.d.kind) {
    case FJGPU_LIGHT_POINT: return lt.d.intensity * color;                 // fj_point_light.cc:41-44
    case FJGPU_LIGHT_GRID: {                                               // fj_rectangle_light.cc:49-61
      const V3 Ln = Normalize(Ps - s.P);
      Real dot = Dot(Ln, s.N);
      if (lt.d.double_sided) dot = std::abs(dot); else dot = Max(dot, 0.);
      return dot * sample_intensity * color;                               // Real*float -> double, then Color*float
    }
    case FJGPU_LIGHT_SPHERE: {                                             // fj_sphere_light.cc:48-60
      const V3 Ln = Normalize(Ps - s.P);
      Col Cl; if (Dot(Ln, s.N) > 0) Cl = sample_intensity * color; return Cl;
    }
    default: return sample_intensity * s.color;                            // fj_dome_light.cc:54-57
  }
}

// SlIlluminance, src/fj_shading.cc:296-359
struct LightOut { Col Cl; V3 Ln; double distance; };
int SlIlluminance(RenderState &rs, const Cxt &cxt, const LightSample &sample, const V3 &Ps, const V3 &axis, double angle,
                  int shaded_object, LightOut *out) {
  out->Cl = Col();
  out->Ln = V3(sample.P.x - Ps.x, sample.P.y - Ps.y, sample.P.z - Ps.z);
  out->distance = Length(out->Ln);
  if (out->distance > 0) { const double inv = 1. / out->distance; out->Ln.x *= inv; out->Ln.y *= inv; out->Ln.z *= inv; }
  const V3 nml_axis = Normalize(axis);
  const double cosangle = Dot(nml_axis, out->Ln);
  if (cosangle < std::cos(angle)) return 0;
  Col lc = light_illuminate((*rs.lights)[sample.light], sample, Ps);
  if (lc.r < .0001 && lc.g < .0001 && lc.b < .0001) return 0;
  if (cxt.ray_context == CXT_SHADOW_RAY) return 0;
  if (cxt.cast_shadow) {
    Cxt sc = cxt;                                       // SlShadowContext :266-279
    sc.ray_context = CXT_SHADOW_RAY; sc.max_diffuse_depth = 0; sc.max_reflect_depth = 0; sc.max_refract_depth = 0;
    sc.trace_target = rs.s->inst[shaded_object].shadow_target;
    sc.node = cxt.node;   // occluder shaders cannot spawn rays, their draws are unused
    Col4 C_occl; double t_hit = FLT_MAX;
    const int hit = SlTrace(rs, sc, Ps, out->Ln, .0001, out->distance, &C_occl, &t_hit);
    if (hit) { const float ac = 1 - C_occl.a; lc.r *= ac; lc.g *= ac; lc.b *= ac; }
  }
  out->Cl = lc;
  return 1;
}

struct SurfIn { V3 P, N, I; Col Cd; int shaded_object; float tu, tv; V3 dPdu, dPdv; };

// TextureCache::LookupTexture, src/fj_texture.cc:51-78 (float arithmetic, as compiled) + MipInput::ReadTile's clamp
// (src/fj_mipmap.cc:163-165) + FrameBuffer::GetColor (src/fj_framebuffer.cc:84-101)
Col4 tex_lookup(const Tex &tx, float u, float v) {
  const float tsu = u - std::floor(u), tsv = v - std::floor(v);
  const float tlu = tsu * tx.xnt, tlv = (1 - tsv) * tx.ynt;
  const int xtile = (int)std::floor(tlu), ytile = (int)std::floor(tlv);
  const int xpxl = (int)((tlu - std::floor(tlu)) * 64), ypxl = (int)((tlv - std::floor(tlv)) * 64);
  const int x = Clamp(xtile, 0, tx.xnt - 1), y = Clamp(ytile, 0, tx.ynt - 1);
  const float *px = &tx.tiles[((size_t)(y * tx.xnt + x) * tx.tilesize * tx.tilesize + (size_t)ypxl * tx.tilesize + xpxl) * tx.nch];
  Col4 c;
  if (tx.nch == 1) { c.r = c.g = c.b = px[0]; c.a = 1; }
  else if (tx.nch == 3) { c.r = px[0]; c.g = px[1]; c.b = px[2]; c.a = 1; }
  else if (tx.nch == 4) { c.r = px[0]; c.g = px[1]; c.b = px[2]; c.a = px[3]; }
  return c;
}

inline float Luminance4(const Col4 &A) { return .298912 * A.r + .586611 * A.g + .114478 * A.b; }     // fj_color.h:280-283
// SlBumpMapping, src/fj_shading.cc:418-465
V3 SlBumpMapping(const Tex &bump, const V3 &dPdu, const V3 &dPdv, float tu, float tv, double amplitude, const V3 &N) {
  if (bump.width == 0 || bump.height == 0) return N;
  const float du = 1. / bump.width, dv = 1. / bump.height;
  float val0 = Luminance4(tex_lookup(bump, tu - du, tv)), val1 = Luminance4(tex_lookup(bump, tu + du, tv));
  const float Bu = (val0 - val1) / (2 * du);
  val0 = Luminance4(tex_lookup(bump, tu, tv - dv)); val1 = Luminance4(tex_lookup(bump, tu, tv + dv));
  const float Bv = (val0 - val1) / (2 * dv);
  V3 N_dPdu = Cross(N, dPdu), N_dPdv = Cross(N, dPdv);
  N_dPdu.x *= du; N_dPdu.y *= du; N_dPdu.z *= du;
  N_dPdv.x *= du; N_dPdv.y *= du; N_dPdv.z *= du;          // (du, not dv: as the reference)
  V3 Nb;
  Nb.x = N.x + amplitude * (Bv * N_dPdu.x - Bu * N_dPdv.x);
  Nb.y = N.y + amplitude * (Bv * N_dPdu.y - Bu * N_dPdv.y);
  Nb.z = N.z + amplitude * (Bv * N_dPdu.z - Bu * N_dPdv.z);
  return Normalize(Nb);
}

// ConstantShader::evaluate, shaders/constant_shader/constant_shader.cc:72-94
void eval_constant(const Scene &s, const fjgpu_shader &sh, const SurfIn &in, Col *Cs, float *Os) {
  if (sh.texture) {
    Col4 C_tex = tex_lookup(s.textures[sh.texture - 1], in.tu, in.tv);
    C_tex.r *= sh.diffuse[0]; C_tex.g *= sh.diffuse[1]; C_tex.b *= sh.diffuse[2];
    *Cs = Col(C_tex.r, C_tex.g, C_tex.b);
  } else *Cs = Col(sh.diffuse[0], sh.diffuse[1], sh.diffuse[2]);
  *Os = 1;
}

// PlasticShader::evaluate, shaders/plastic_shader/plastic_shader.cc:101-179 (diffuse_map; no bump map)
void eval_plastic(RenderState &rs, const Cxt &cxt, const fjgpu_shader &sh, const SurfIn &in, Col *Cs, float *Os) {
  Col diff, spec;
  V3 Nf = SlFaceforward(in.I, in.N);
  if (sh.bump_texture) Nf = SlBumpMapping(rs.s->textures[sh.bump_texture - 1], in.dPdu, in.dPdv, in.tu, in.tv, sh.bump_amplitude, Nf);   // :115-123
  std::vector<LightSample> samples; new_light_samples(rs, cxt, samples);
  for (size_t i = 0; i < samples.size(); i++) {
    LightOut Lout; Lout.Ln = V3(); Lout.distance = 0;
    SlIlluminance(rs, cxt, samples[i], in.P, Nf, PI / 2., in.shaded_object, &Lout);
    float Kd = Dot(Nf, Lout.Ln);
    Kd = Max(0, Kd);
    diff.r += Kd * Lout.Cl.r; diff.g += Kd * Lout.Cl.g; diff.b += Kd * Lout.Cl.b;
  }
  Col4 diff_map; diff_map.r = diff_map.g = diff_map.b = diff_map.a = 1;
  if (sh.texture) diff_map = tex_lookup(rs.s->textures[sh.texture - 1], in.tu, in.tv);
  Cs->r = diff.r * sh.diffuse[0] * diff_map.r + spec.r;
  Cs->g = diff.g * sh.diffuse[1] * diff_map.g + spec.g;
  Cs->b = diff.b * sh.diffuse[2] * diff_map.b + spec.b;
  if (sh.do_reflect) {
    Col4 C_refl; double t_hit = REAL_MAX;
    Cxt rc = cxt; rc.reflect_depth++; rc.ray_context = CXT_REFLECT_RAY;          // SlReflectContext :242-252
    rc.trace_target = rs.s->inst[in.shaded_object].reflect_target; rc.node = cxt.node * 4 + 2;
    const V3 R = Normalize(SlReflect(in.I, Nf));
    SlTrace(rs, rc, in.P, R, .001, 1000, &C_refl, &t_hit);
    const double Kr = SlFresnel(in.I, Nf, 1 / sh.ior);
    Cs->r += Kr * C_refl.r * sh.reflect[0];
    Cs->g += Kr * C_refl.g * sh.reflect[1];
    Cs->b += Kr * C_refl.b * sh.reflect[2];
  }
  *Os = sh.opacity;
}

// PathtracingShader::evaluate + integrate_*, shaders/pathtracing_shader/pathtracing_shader.cc:125-257 (diffuse_map :132-135,
// bump_map :136-144: every integrator sees the bumped normal)
void eval_pathtracing(RenderState &rs, const Cxt &cxt, const fjgpu_shader &sh, const SurfIn &in_, Col *Cs, float *Os) {
  SurfIn in = in_;
  if (sh.bump_texture) in.N = SlBumpMapping(rs.s->textures[sh.bump_texture - 1], in_.dPdu, in_.dPdv, in_.tu, in_.tv, sh.bump_amplitude, in_.N);
  const Col diffuse(sh.diffuse[0], sh.diffuse[1], sh.diffuse[2]);
  const Col reflect(sh.reflect[0], sh.reflect[1], sh.reflect[2]);
  const Col refract(sh.refract[0], sh.refract[1], sh.refract[2]);
  const Col Le(sh.emission[0], sh.emission[1], sh.emission[2]);
  Col L_diffuse, L_reflect, L_refract;
  Col Cd = in.Cd;
  if (sh.texture) { const Col4 m = tex_lookup(rs.s->textures[sh.texture - 1], in.tu, in.tv); Cd = Cd * Col(m.r, m.g, m.b); }    // diffuse_map :132-135
  if (Luminance(diffuse) > 0.) {                                        // integrate_diffuse :176-208
    V3 u, v, w = in.N;
    u = std::abs(w.x) > .001 ? V3(0, 1, 0) : V3(1, 0, 0);
    u = Normalize(Cross(u, w));
    v = Cross(w, u);
    Real x1, x2;
    if (rs.rng_mode == 1) { XorShift &r = (*rs.pt_rng)[rs.cur_shader]; x1 = r.NextFloat01(); x2 = r.NextFloat01(); }
    else { x1 = ctr_rand(rs, cxt.node, 0); x2 = ctr_rand(rs, cxt.node, 1); }
    const Real r1 = 2. * PI * x1, r2 = x2, r2sqrt = std::sqrt(r2);
    const V3 D = Normalize(u * std::cos(r1) * r2sqrt + v * std::sin(r1) * r2sqrt + w * std::sqrt(1. - r2));
    const Real Kd = Dot(in.N, D);
    Col4 C; double t_hit = REAL_MAX;
    Cxt dc = cxt; dc.diffuse_depth++; dc.ray_context = CXT_DIFFUSE_RAY;          // SlDiffuseContext :230-240
    dc.trace_target = rs.s->inst[in.shaded_object].reflect_target; dc.node = cxt.node * 4 + 1;
    SlTrace(rs, dc, in.P, D, .001, 1000, &C, &t_hit);
    L_diffuse = Cd * (float)Kd * diffuse * Col(C.r, C.g, C.b);
  }
  if (Luminance(reflect) > 0.) {                                        // integrate_reflect :210-229
    const V3 R = Normalize(SlReflect(in.I, in.N));
    const Real Kr = SlFresnel(in.I, in.N, 1. / sh.ior);
    Col4 C; double t_hit = REAL_MAX;
    Cxt rc = cxt; rc.reflect_depth++; rc.ray_context = CXT_REFLECT_RAY;
    rc.trace_target = rs.s->inst[in.shaded_object].reflect_target; rc.node = cxt.node * 4 + 2;
    SlTrace(rs, rc, in.P, R, .001, 1000, &C, &t_hit);
    L_reflect = (float)Kr * reflect * Col(C.r, C.g, C.b);
  }
  if (Luminance(refract) > 0.) {                                        // integrate_refract :231-257
    const V3 T = Normalize(SlRefract(in.I, in.N, 1. / sh.ior));
    const Real Kr = SlFresnel(in.I, in.N, 1 / sh.ior);
    const Real Kt = 1 - Kr;
    Col4 C; double t_hit = REAL_MAX;
    Cxt rc = cxt; rc.refract_depth++; rc.ray_context = CXT_REFRACT_RAY;           // SlRefractContext :254-264
    rc.trace_target = rs.s->inst[in.shaded_object].refract_target; rc.node = cxt.node * 4 + 3;
    SlTrace(rs, rc, in.P, T, .0001, 1000, &C, &t_hit);
    if (sh.do_color_filter && Dot(in.I, in.N) < 0) {
      C.r *= std::pow(sh.transmit[0], t_hit); C.g *= std::pow(sh.transmit[1], t_hit); C.b *= std::pow(sh.transmit[2], t_hit);
    }
    L_refract = (float)Kt * refract * Col(C.r, C.g, C.b);
  }
  *Cs = Le + L_diffuse + L_reflect + L_refract;
  *Os = 1;
}

// GlassShader::evaluate, shaders/glass_shader/glass_shader.cc:88-133
void eval_glass(RenderState &rs, const Cxt &cxt, const fjgpu_shader &sh, const SurfIn &in, Col *Cs, float *Os) {
  *Cs = Col();
  const double Kr = SlFresnel(in.I, in.N, 1 / sh.ior), Kt = 1 - Kr;
  Col4 C_refl, C_refr; double t_hit = REAL_MAX;
  Cxt rc = cxt; rc.reflect_depth++; rc.ray_context = CXT_REFLECT_RAY;          // SlReflectContext :242-252
  rc.trace_target = rs.s->inst[in.shaded_object].reflect_target; rc.node = cxt.node * 4 + 2;
  const V3 R = Normalize(SlReflect(in.I, in.N));
  SlTrace(rs, rc, in.P, R, .0001, 1000, &C_refl, &t_hit);
  Cs->r += Kr * C_refl.r; Cs->g += Kr * C_refl.g; Cs->b += Kr * C_refl.b;
  Cxt fc = cxt; fc.refract_depth++; fc.ray_context = CXT_REFRACT_RAY;          // SlRefractContext :254-264
  fc.trace_target = rs.s->inst[in.shaded_object].refract_target; fc.node = cxt.node * 4 + 3;
  const V3 T = Normalize(SlRefract(in.I, in.N, 1 / sh.ior));
  SlTrace(rs, fc, in.P, T, .0001, 1000, &C_refr, &t_hit);
  if (sh.do_color_filter && Dot(in.I, in.N) < 0) {                              // filter_color is carried in `transmit`
    C_refr.r *= std::pow(sh.transmit[0], t_hit); C_refr.g *= std::pow(sh.transmit[1], t_hit); C_refr.b *= std::pow(sh.transmit[2], t_hit);
  }
  Cs->r += Kt * C_refr.r; Cs->g += Kt * C_refr.g; Cs->b += Kt * C_refr.b;
  *Os = 1;
}

// SlTrace + trace_surface, src/fj_shading.cc:140-179, :527-572 ; bounce gate :467-499
int SlTrace(RenderState &rs, const Cxt &cxt, const V3 &o, const V3 &d, double tmin, double tmax, Col4 *out, double *t_hit) {
  *out = Col4();
  int cur = 0, mx = 1;
  switch (cxt.ray_context) {
    case CXT_DIFFUSE_RAY: cur = cxt.diffuse_depth; mx = cxt.max_diffuse_depth; break;
    case CXT_REFLECT_RAY: cur = cxt.reflect_depth; mx = cxt.max_reflect_depth; break;
    case CXT_REFRACT_RAY: cur = cxt.refract_depth; mx = cxt.max_refract_depth; break;
    default: break;
  }
  if (cur > mx) return 0;
  rs.rays[cxt.ray_context]++;
  Ray ray; ray.orig = o; ray.dir = d; ray.tmin = tmin; ray.tmax = tmax;
  Isect isect;
  const bool hit = group_intersect(*rs.s, rs.s->groups[cxt.trace_target], ray, cxt.time, &isect);
  Col4 surf;
  if (hit) {
    SurfIn in; in.shaded_object = isect.object; in.P = isect.P; in.N = isect.N; in.Cd = Col(1, 1, 1); in.I = ray.dir;
    in.tu = isect.tu; in.tv = isect.tv; in.dPdu = isect.dPdu; in.dPdv = isect.dPdv;
    const Instance &inst = rs.s->inst[isect.object];
    int g = isect.shading_group_id;                                  // ObjectInstance::GetShader :177-191
    int slot = (g < 0 || g >= FJGPU_MAX_SHADING_GROUPS) ? inst.shader_of_group[0] : inst.shader_of_group[g];
    if (slot < 0) slot = inst.shader_of_group[0];
    Col Cs; float Os = 1;
    if (slot < 0 || rs.s->shaders[slot].d.kind == FJGPU_SHADER_NONE) { Cs = Col(.5, 1., 0.); Os = 1; }
    else {
      const fjgpu_shader &sh = rs.s->shaders[slot].d;
      if (sh.kind == FJGPU_SHADER_CONSTANT) eval_constant(*rs.s, sh, in, &Cs, &Os);
      else if (sh.kind == FJGPU_SHADER_PLASTIC) eval_plastic(rs, cxt, sh, in, &Cs, &Os);
      else if (sh.kind == FJGPU_SHADER_GLASS) eval_glass(rs, cxt, sh, in, &Cs, &Os);
      else { rs.cur_shader = slot; eval_pathtracing(rs, cxt, sh, in, &Cs, &Os); }
    }
    Os = (float)Clamp(Os, 0, 1);
    surf.r = Cs.r; surf.g = Cs.g; surf.b = Cs.b; surf.a = Os;
    *t_hit = isect.t_hit;
  }
  // shadow early-out (:162-165) and the no-volume composite (:173-176) both reduce to `surf`
  *out = surf;
  return hit ? 1 : 0;
}

// ---------------------------------------------------------------- sampler / camera / filter
struct Sample { Real u, v; Real data[4]; Real time; };

inline void sampler_counts(const fjgpu_render_params &p, int m[2]) {   // count_samples_in_margin, fj_fixed_grid_sampler.cc:131-136
  m[0] = static_cast<int>(std::ceil(((p.xfwidth - 1) * p.xrate) * .5));
  m[1] = static_cast<int>(std::ceil(((p.yfwidth - 1) * p.yrate) * .5));
}
// FixedGridSampler::generate_samples, src/fj_fixed_grid_sampler.cc:33-84
void generate_samples(const fjgpu_render_params &p, const fjgpu_tile &t, std::vector<Sample> &out, int ns[2], const Real *time_range = nullptr) {
  int m[2]; sampler_counts(p, m);
  ns[0] = p.xrate * (t.xmax - t.xmin) + 2 * m[0];
  ns[1] = p.yrate * (t.ymax - t.ymin) + 2 * m[1];
  out.resize((size_t)ns[0] * ns[1]);
  XorShift rng;
  XorShift rng_time;                                                   // :42: a second, identically seeded stream
  const Real udelta = 1. / (p.xrate * p.xres), vdelta = 1. / (p.yrate * p.yres);
  const int xoffset = t.xmin * p.xrate - m[0], yoffset = t.ymin * p.yrate - m[1];
  Sample *s = out.data();
  for (int y = 0; y < ns[1]; y++) for (int x = 0; x < ns[0]; x++) {
    s->u = (.5 + x + xoffset) * udelta;
    s->v = 1 - (.5 + y + yoffset) * vdelta;
    if (p.jitter > 0) {
      const Real uj = rng.NextFloat01() * p.jitter, vj = rng.NextFloat01() * p.jitter;
      s->u += udelta * (uj - .5); s->v += vdelta * (vj - .5);
    }
    // :72-77 (IsSamplingTime() is always true: Renderer's constructor calls SetSampleTimeRange(0, 1))
    if (time_range) { const Real rnd = rng_time.NextFloat01(); s->time = Fit(rnd, 0, 1, time_range[0], time_range[1]); }
    else s->time = 0;
    s->data[0] = s->data[1] = s->data[2] = s->data[3] = 0;
    s++;
  }
}
// Camera::GetRay, src/fj_camera.cc:79-110
void camera_ray(const fjgpu_camera &c, int xres, int yres, Real u, Real v, Ray *ray, const XfSamples *xs = nullptr, Real time = 0) {
  Mat m; memcpy(m.e, c.fwd, sizeof m.e);
  if (xs && xs->moving) { Mat inv; lerp_transform(*xs, time, &m, &inv); }      // :81-82
  const Real aspect = xres / (double)yres;
  const Real uvy = 2 * std::tan(Radian(c.fov / 2.)), uvx = uvy * aspect;
  const V3 target = MatPoint(m, V3((u - .5) * uvx, (v - .5) * uvy, -1));
  const V3 eye = MatPoint(m, V3(0, 0, 0));
  ray->dir = Normalize(target - eye); ray->orig = eye; ray->tmin = c.znear; ray->tmax = c.zfar;
}
inline Real eval_gaussian(Real xw, Real yw, Real x, Real y) {          // src/fj_filter.cc:49-58
  const Real xx = 2 * x / xw, yy = 2 * y / yw; return std::exp(-2 * (xx * xx + yy * yy)); }

// render_tile: integrate_samples + reconstruct_image, src/fj_renderer.cc:1061-1121, :939-995
void render_tile(RenderState &rs, const fjgpu_render_params &p, const fjgpu_tile &t, float *rgba,
                 double *dump_uv, float *dump_rgba) {
  std::vector<Sample> smp; int ns[2]; generate_samples(p, t, smp, ns, rs.s->time_range);
  Cxt cxt; memset(&cxt, 0, sizeof cxt);                                // SlCameraContext + init_worker :896-905
  cxt.ray_context = CXT_CAMERA_RAY; cxt.max_diffuse_depth = p.max_diffuse_depth; cxt.max_reflect_depth = p.max_reflect_depth;
  cxt.max_refract_depth = p.max_refract_depth; cxt.cast_shadow = p.cast_shadow; cxt.opacity_threshold = .995f;
  cxt.trace_target = p.target_group; cxt.node = 1;
  rs.tile_id = (uint32_t)t.id;
  for (size_t i = 0; i < smp.size(); i++) {
    Ray ray; camera_ray(rs.s->cam, p.xres, p.yres, smp[i].u, smp[i].v, &ray, &rs.s->cam_xs, smp[i].time);
    cxt.time = smp[i].time;                                            // src/fj_renderer.cc:1073-1074
    Col4 C; double t_hit = FLT_MAX;
    rs.sample_id = (uint32_t)i;
    const int hit = SlTrace(rs, cxt, ray.orig, ray.dir, ray.tmin, ray.tmax, &C, &t_hit);
    if (hit) { smp[i].data[0] = C.r; smp[i].data[1] = C.g; smp[i].data[2] = C.b; smp[i].data[3] = C.a; }
    if (dump_uv) { dump_uv[2 * i] = smp[i].u; dump_uv[2 * i + 1] = smp[i].v; }
    if (dump_rgba) for (int c = 0; c < 4; c++) dump_rgba[4 * i + c] = (float)smp[i].data[c];
  }
  if (!rgba) return;
  int m[2]; sampler_counts(p, m);
  const int npx = p.xrate + 2 * m[0], npy = p.yrate + 2 * m[1];
  for (int y = t.ymin; y < t.ymax; y++) for (int x = t.xmin; x < t.xmax; x++) {
    const int off = (y - t.ymin) * p.yrate * ns[0] + (x - t.xmin) * p.xrate;   // get_sampleset_in_pixel :97-124
    float pr = 0, pg = 0, pb = 0, pa = 0, wsum = 0;
    for (int sy = 0; sy < npy; sy++) for (int sx = 0; sx < npx; sx++) {
      const Sample &s = smp[off + sy * ns[0] + sx];
      const double fx = p.xres * s.u - (x + .5), fy = p.yres * (1 - s.v) - (y + .5);
      const double w = eval_gaussian(p.xfwidth, p.yfwidth, fx, fy);
      pr += w * s.data[0]; pg += w * s.data[1]; pb += w * s.data[2]; pa += w * s.data[3]; wsum += w;
    }
    const float inv = 1.f / wsum;
    float *px = rgba + ((size_t)y * p.xres + x) * 4;
    px[0] = pr * inv; px[1] = pg * inv; px[2] = pb * inv; px[3] = pa * inv;
  }
}

}  // namespace

// ============================================================================ C API (ctypes)
extern "C" {

struct fjo_scene { Scene s; };

fjo_scene *fjo_scene_new() { return new fjo_scene(); }
void fjo_scene_free(fjo_scene *sc) { delete sc; }

int fjo_mesh(fjo_scene *sc, int mesh_id, const double *P, const double *N, int nverts, const int32_t *idx,
             const int32_t *group, int nfaces) {
  Mesh &m = sc->s.meshes[mesh_id]; m = Mesh();
  m.P.resize(nverts); for (int i = 0; i < nverts; i++) m.P[i] = V3(P[3 * i], P[3 * i + 1], P[3 * i + 2]);
  if (N) { m.N.resize(nverts); for (int i = 0; i < nverts; i++) m.N[i] = V3(N[3 * i], N[3 * i + 1], N[3 * i + 2]); }
  m.idx.assign(idx, idx + 3 * (size_t)nfaces); m.nfaces = nfaces;
  if (group) m.group.assign(group, group + nfaces);
  m.bounds.ReverseInfinite();
  for (int i = 0; i < nfaces; i++) { Box b; m.prim_bounds(i, &b); m.bounds.AddBox(b); }
  sc->s.built = false;
  return 0;
}
// Runs the velocity generator on a mesh (second half); vel3_out (may be null) receives the velocities.
int fjo_mesh_generate_velocity(fjo_scene *sc, int mesh_id, double *vel3_out) {
  auto it = sc->s.meshes.find(mesh_id); if (it == sc->s.meshes.end()) return -1;
  generate_velocity(it->second);
  if (vel3_out) for (size_t i = 0; i < it->second.vel.size(); i++) { vel3_out[3 * i] = it->second.vel[i].x; vel3_out[3 * i + 1] = it->second.vel[i].y; vel3_out[3 * i + 2] = it->second.vel[i].z; }
  sc->s.built = false;
  return 0;
}
int fjo_mesh_set_uv(fjo_scene *sc, int mesh_id, const float *uv2, int nverts) {
  auto it = sc->s.meshes.find(mesh_id);
  if (it == sc->s.meshes.end() || (uv2 && nverts != (int)it->second.P.size())) return -1;
  if (uv2) it->second.uv.assign(uv2, uv2 + 2 * (size_t)nverts); else it->second.uv.clear();
  return 0;
}
int fjo_textures(fjo_scene *sc, int n, const fjgpu_texture *t) {
  sc->s.textures.resize(n);
  for (int i = 0; i < n; i++) {
    Tex &x = sc->s.textures[i];
    x.width = t[i].width; x.height = t[i].height; x.nch = t[i].nchannels; x.tilesize = t[i].tilesize;
    x.xnt = x.width / x.tilesize; x.ynt = x.height / x.tilesize;
    x.tiles.assign(t[i].tiles, t[i].tiles + (size_t)x.xnt * x.ynt * x.tilesize * x.tilesize * x.nch);
  }
  return 0;
}
// DomeLight::preprocess with an environment map (src/fj_dome_light.cc:78-96: the map sampled at 1/8 of its resolution)
// -> StratifiedImportanceSampling (src/fj_importance_sampling.cc:102-150) with make_histgram / lookup_histgram /
// index_to_uv / uv_to_dir (:222-277).  Writes `nsamples` directions (3 doubles) and colours (3 floats).
int fjo_dome_samples(const fjgpu_texture *t, int nsamples, double *dirs3, float *cols3) {
  Tex tx; tx.width = t->width; tx.height = t->height; tx.nch = t->nchannels; tx.tilesize = t->tilesize;
  tx.xnt = tx.width / tx.tilesize; tx.ynt = tx.height / tx.tilesize;
  tx.tiles.assign(t->tiles, t->tiles + (size_t)tx.xnt * tx.ynt * tx.tilesize * tx.tilesize * tx.nch);
  const int xres = tx.width / 8, yres = tx.height / 8, NPIXELS = xres * yres;
  if (NPIXELS <= 0) return -1;
  auto index_to_uv = [&](int index, float *u, float *v) {
    const int x = index % xres, y = index / xres;
    *u = (.5 + x) / xres; *v = 1. - ((.5 + y) / yres);
  };
  std::vector<double> hist(NPIXELS);
  double sum = 0;
  for (int i = 0; i < NPIXELS; i++) {
    float u, v; index_to_uv(i, &u, &v);
    const Col4 c = tex_lookup(tx, u, v);
    sum += (float)(.298912 * c.r + .586611 * c.g + .114478 * c.b);      // Luminance4 returns float
    hist[i] = sum;
  }
  sum = hist[NPIXELS - 1];
  XorShift rng;
  for (int i = 0; i < nsamples; i++) {
    const double rnd = sum * ((i + rng.NextFloat01()) / nsamples);
    int index = -1;
    for (int k = 0; k < NPIXELS; k++) if (rnd < hist[k]) { index = k; break; }
    float u, v; index_to_uv(index, &u, &v);
    const double phi = 2 * PI * u, theta = PI * (v - .5), r = std::cos(theta);
    dirs3[3 * i] = r * std::sin(phi); dirs3[3 * i + 1] = std::sin(theta); dirs3[3 * i + 2] = r * std::cos(phi);
    const Col4 c = tex_lookup(tx, u, v);
    cols3[3 * i] = c.r; cols3[3 * i + 1] = c.g; cols3[3 * i + 2] = c.b;
  }
  return 0;
}
// Mesh::ComputeNormals restatement exposed for the host-side parity test
void fjo_compute_normals(const double *P, int nverts, const int32_t *idx, int nfaces, double *N_out) {
  Mesh m; m.P.resize(nverts); for (int i = 0; i < nverts; i++) m.P[i] = V3(P[3 * i], P[3 * i + 1], P[3 * i + 2]);
  m.idx.assign(idx, idx + 3 * (size_t)nfaces); m.nfaces = nfaces; compute_normals(m);
  for (int i = 0; i < nverts; i++) { N_out[3 * i] = m.N[i].x; N_out[3 * i + 1] = m.N[i].y; N_out[3 * i + 2] = m.N[i].z; }
}
int fjo_instances(fjo_scene *sc, int n, const fjgpu_instance *in) {
  sc->s.inst.resize(n);
  for (int i = 0; i < n; i++) {
    Instance &o = sc->s.inst[i]; o.mesh = in[i].mesh_id;
    memcpy(o.fwd.e, in[i].fwd, sizeof o.fwd.e); memcpy(o.inv.e, in[i].inv, sizeof o.inv.e);
    memcpy(o.shader_of_group, in[i].shader_of_group, sizeof o.shader_of_group);
    o.reflect_target = in[i].reflect_target; o.refract_target = in[i].refract_target; o.shadow_target = in[i].shadow_target;
  }
  sc->s.built = false; return 0;
}
int fjo_groups(fjo_scene *sc, int ng, const int32_t *off, const int32_t *ids) {
  sc->s.groups.assign(ng, Group());
  for (int g = 0; g < ng; g++) sc->s.groups[g].inst.assign(ids + off[g], ids + off[g + 1]);
  sc->s.built = false; return 0;
}
int fjo_shaders(fjo_scene *sc, int n, const fjgpu_shader *sh) { sc->s.shaders.resize(n); for (int i = 0; i < n; i++) sc->s.shaders[i].d = sh[i]; return 0; }
int fjo_lights(fjo_scene *sc, int n, const fjgpu_light *l) {
  sc->s.lights.assign(n, Light());
  for (int i = 0; i < n; i++) {
    Light &o = sc->s.lights[i]; o.d = l[i]; memcpy(o.fwd.e, l[i].fwd, sizeof o.fwd.e);
    for (int k = 0; k < l[i].dome_sample_count; k++) {
      o.dome_dir.push_back(V3(l[i].dome_dirs[3 * k], l[i].dome_dirs[3 * k + 1], l[i].dome_dirs[3 * k + 2]));
      o.dome_col.push_back(Col(l[i].dome_colors[3 * k], l[i].dome_colors[3 * k + 1], l[i].dome_colors[3 * k + 2]));
    }
    o.d.dome_dirs = 0; o.d.dome_colors = 0;
  }
  return 0;
}
int fjo_camera(fjo_scene *sc, const fjgpu_camera *c) { sc->s.cam = *c; sc->s.cam_xs = XfSamples(); return 0; }

// Time-sampled transforms (SiSetSampleProperty3 on translate / rotate / scale): samples as (x, y, z, time) rows sorted by
// time; target < 0 = the camera, else the instance index.  More than one sample in any channel makes the transform
// time dependent.  fjo_time_range = Renderer sample_time_range.
static void load_samples(TimeSamples *l, int n, const double *v4) { l->v.resize(n); for (int i = 0; i < n; i++) for (int k = 0; k < 4; k++) l->v[i][k] = v4[4 * i + k]; }
int fjo_transform_samples(fjo_scene *sc, int target, int torder, int rorder, int nT, const double *T4, int nR, const double *R4, int nS, const double *S4) {
  if (nT < 1 || nR < 1 || nS < 1 || nT > 8 || nR > 8 || nS > 8) return -1;
  XfSamples xs; xs.torder = torder; xs.rorder = rorder;
  load_samples(&xs.T, nT, T4); load_samples(&xs.R, nR, R4); load_samples(&xs.S, nS, S4);
  xs.moving = nT > 1 || nR > 1 || nS > 1;
  if (target < 0) sc->s.cam_xs = xs;
  else { if (target >= (int)sc->s.inst.size()) return -1; sc->s.inst[target].xs = xs; }
  sc->s.built = false; return 0;
}
// probes: PerlinNoise3d / SmoothStep
void fjo_perlin3d(const double *p3, double lacunarity, double persistence, int octaves, double *out3) {
  const V3 n = PerlinNoise3d(V3(p3[0], p3[1], p3[2]), lacunarity, persistence, octaves); out3[0] = n.x; out3[1] = n.y; out3[2] = n.z;
}
double fjo_smoothstep(double a, double b, double x) { return SmoothStep(a, b, x); }
// probe: XfmLerpTransformSample of a sample list at one time
int fjo_lerp_transform(int torder, int rorder, int nT, const double *T4, int nR, const double *R4, int nS, const double *S4, double time, double *fwd16, double *inv16) {
  if (nT < 1 || nR < 1 || nS < 1) return -1;
  XfSamples xs; xs.torder = torder; xs.rorder = rorder;
  load_samples(&xs.T, nT, T4); load_samples(&xs.R, nR, R4); load_samples(&xs.S, nS, S4);
  Mat f, i; lerp_transform(xs, time, &f, &i);
  memcpy(fwd16, f.e, sizeof f.e); memcpy(inv16, i.e, sizeof i.e);
  return 0;
}
void fjo_time_range(fjo_scene *sc, double t0, double t1) { sc->s.time_range[0] = t0; sc->s.time_range[1] = t1; }
// the frame's time table as fjgpu_time_table defines it: time of the k-th sample of a tile
int fjo_sample_times(const fjgpu_render_params *p, const fjgpu_tile *t, double t0, double t1, double *times, int cap) {
  const Real tr[2] = {t0, t1};
  std::vector<Sample> s; int ns[2]; generate_samples(*p, *t, s, ns, tr);
  for (size_t i = 0; i < s.size() && (int)i < cap; i++) times[i] = s[i].time;
  return (int)s.size();
}

// compute_objects_bounds + build_accelerators, src/fj_scene_interface.cc:1137-1202
int fjo_build(fjo_scene *sc) {
  Scene &s = sc->s;
  for (auto &kv : s.meshes) grid_build(kv.second);
  for (auto &in : s.inst) {                               // ObjectInstance::update_bounds :299-360 (single sample)
    auto it = s.meshes.find(in.mesh); if (it == s.meshes.end()) return -1;
    in.bounds = it->second.acc_bounds; MatTransformBounds(in.fwd, &in.bounds);
    if (in.xs.moving) {                                   // merge_sampled_bounds :313-360, index quirk (rotate sample i of translate sample i) kept
      Box ob = it->second.acc_bounds;
      if (in.xs.R.v.size() > 1) {
        const V3 dg = ob.max - ob.min; const Real half = .5 * Length(dg);
        const V3 c = ob.Centroid(); ob.min = c; ob.max = c; ob.Expand(half);
      }
      Real S[3] = {0, 0, 0};
      for (auto &e : in.xs.S.v) for (int k = 0; k < 3; k++) S[k] = std::max(S[k], std::fabs(e[k]));
      Box merged; merged.ReverseInfinite();
      for (size_t i = 0; i < in.xs.T.v.size(); i++) {
        const Real *T = in.xs.T.v[i].data();
        const Real zero[4] = {0, 0, 0, 0};
        const Real *R = i < in.xs.R.v.size() ? in.xs.R.v[i].data() : zero;     // unused slots of the 8-entry array are zero
        Mat m; make_transform_matrix(in.xs.torder, in.xs.rorder, T[0], T[1], T[2], R[0], R[1], R[2], S[0], S[1], S[2], &m);
        Box sb = ob; MatTransformBounds(m, &sb); merged.AddBox(sb);
      }
      in.bounds = merged;
    }
  }
  for (auto &g : s.groups) {
    Box b; b.ReverseInfinite();
    for (int i : g.inst) b.AddBox(s.inst[i].bounds);        // ObjectSet::AddObject :30-38
    b.Expand(PADDING); g.acc_bounds = b;                    // Accelerator::ComputeBounds
    g.nodes.clear(); g.root = -1;
    const int n = (int)g.inst.size(); if (n == 0) continue;
    std::vector<BvhPrim> prims(n); std::vector<BvhPrim *> ptr(n);
    for (int i = 0; i < n; i++) { prims[i].bounds = s.inst[g.inst[i]].bounds; prims[i].centroid = prims[i].bounds.Centroid(); prims[i].index = i; ptr[i] = &prims[i]; }
    g.root = build_bvh(g, ptr.data(), 0, n, 0);
  }
  s.built = true; return 0;
}

// rng_mode 0: counter RNG, tiles rendered by `nthreads` threads (deterministic for any count)
// rng_mode 1: the reference's sequential XorShift streams, single thread, tiles in order
int fjo_render(fjo_scene *sc, const fjgpu_render_params *p, const fjgpu_tile *tiles, int ntiles, float *rgba,
               int rng_mode, int nthreads, fjgpu_stats *stats) {
  if (!sc->s.built && fjo_build(sc)) return -1;
  uint64_t rays[5] = {0, 0, 0, 0, 0}; uint64_t nsmp = 0;
  if (rng_mode == 1 || nthreads <= 1) {
    std::vector<Light> lights = sc->s.lights; std::vector<XorShift> pt(sc->s.shaders.size() + 1);
    RenderState rs; rs.s = &sc->s; rs.rng_mode = rng_mode; rs.seed = p->seed; rs.pt_rng = &pt; rs.lights = &lights;
    memset(rs.rays, 0, sizeof rs.rays);
    for (int i = 0; i < ntiles; i++) render_tile(rs, *p, tiles[i], rgba, 0, 0);
    for (int k = 0; k < 5; k++) rays[k] = rs.rays[k];
  } else {
    std::atomic<int> next(0); std::vector<std::thread> th; std::vector<std::vector<uint64_t>> acc(nthreads, std::vector<uint64_t>(5, 0));
    for (int w = 0; w < nthreads; w++) th.emplace_back([&, w]() {
      std::vector<Light> lights = sc->s.lights; std::vector<XorShift> pt(sc->s.shaders.size() + 1);
      RenderState rs; rs.s = &sc->s; rs.rng_mode = 0; rs.seed = p->seed; rs.pt_rng = &pt; rs.lights = &lights;
      memset(rs.rays, 0, sizeof rs.rays);
      for (;;) { const int i = next.fetch_add(1); if (i >= ntiles) break; render_tile(rs, *p, tiles[i], rgba, 0, 0); }
      for (int k = 0; k < 5; k++) acc[w][k] = rs.rays[k];
    });
    for (auto &t : th) t.join();
    for (int w = 0; w < nthreads; w++) for (int k = 0; k < 5; k++) rays[k] += acc[w][k];
  }
  int m[2]; sampler_counts(*p, m);
  for (int i = 0; i < ntiles; i++) nsmp += (uint64_t)(p->xrate * (tiles[i].xmax - tiles[i].xmin) + 2 * m[0]) * (p->yrate * (tiles[i].ymax - tiles[i].ymin) + 2 * m[1]);
  if (stats) { memset(stats, 0, sizeof *stats); stats->rays_camera = rays[0]; stats->rays_shadow = rays[1]; stats->rays_diffuse = rays[2];
               stats->rays_reflect = rays[3]; stats->rays_refract = rays[4]; stats->camera_samples = nsmp; }
  return 0;
}

int fjo_render_tile_samples(fjo_scene *sc, const fjgpu_render_params *p, const fjgpu_tile *tile, int rng_mode,
                            double *out_uv, float *out_rgba) {
  if (!sc->s.built && fjo_build(sc)) return -1;
  std::vector<Light> lights = sc->s.lights; std::vector<XorShift> pt(sc->s.shaders.size() + 1);
  RenderState rs; rs.s = &sc->s; rs.rng_mode = rng_mode; rs.seed = p->seed; rs.pt_rng = &pt; rs.lights = &lights;
  memset(rs.rays, 0, sizeof rs.rays);
  render_tile(rs, *p, *tile, 0, out_uv, out_rgba);
  return 0;
}

int fjo_trace_closest(fjo_scene *sc, int group, int n, const double *o, const double *d, const double *tmin, const double *tmax,
                      double *out_t, double *out_u, double *out_v, int32_t *out_prim, int32_t *out_inst) {
  if (!sc->s.built && fjo_build(sc)) return -1;
  for (int i = 0; i < n; i++) {
    Ray r; r.orig = V3(o[3 * i], o[3 * i + 1], o[3 * i + 2]); r.dir = V3(d[3 * i], d[3 * i + 1], d[3 * i + 2]); r.tmin = tmin[i]; r.tmax = tmax[i];
    Isect is; const bool hit = group_intersect(sc->s, sc->s.groups[group], r, 0, &is);
    out_t[i] = hit ? is.t_hit : REAL_MAX; out_u[i] = hit ? is.u : 0; out_v[i] = hit ? is.v : 0;
    out_prim[i] = hit ? is.prim_id : -1; out_inst[i] = hit ? is.object : -1;
  }
  return 0;
}

// ---- per-function probes (compared with tests/golden/ref_vectors.json) ----
void fjo_xorshift_u32(uint32_t *out, int n) { XorShift r; for (int i = 0; i < n; i++) out[i] = r.NextInteger(); }
void fjo_xorshift_f01(double *out, int n) { XorShift r; for (int i = 0; i < n; i++) out[i] = r.NextFloat01(); }
int fjo_tri_intersect(const double *v0, const double *v1, const double *v2, const double *o, const double *d, double *tuv) {
  Real t = 0, u = 0, v = 0;
  const bool h = TriRayIntersect(V3(v0[0], v0[1], v0[2]), V3(v1[0], v1[1], v1[2]), V3(v2[0], v2[1], v2[2]), V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]), &t, &u, &v);
  tuv[0] = t; tuv[1] = u; tuv[2] = v; return h ? 1 : 0;
}
int fjo_box_intersect(const double *bmin, const double *bmax, const double *o, const double *d, double tmin, double tmax, double *t01) {
  Box b; b.min = V3(bmin[0], bmin[1], bmin[2]); b.max = V3(bmax[0], bmax[1], bmax[2]);
  Real a = 0, c = 0; const bool h = BoxRayIntersect(b, V3(o[0], o[1], o[2]), V3(d[0], d[1], d[2]), tmin, tmax, &a, &c);
  if (h) { t01[0] = a; t01[1] = c; }   // a miss leaves the outputs untouched (tests/box_test.cc)
  return h ? 1 : 0;
}
void fjo_make_transform(int torder, int rorder, const double *T, const double *R, const double *S, double *fwd, double *inv) {
  Mat m, i; make_transform_matrix(torder, rorder, T[0], T[1], T[2], R[0], R[1], R[2], S[0], S[1], S[2], &m); MatInverse(&i, m);
  memcpy(fwd, m.e, sizeof m.e); memcpy(inv, i.e, sizeof i.e);
}
void fjo_camera_ray(const fjgpu_camera *c, int xres, int yres, double u, double v, double *o3, double *d3) {
  Ray r; camera_ray(*c, xres, yres, u, v, &r); o3[0] = r.orig.x; o3[1] = r.orig.y; o3[2] = r.orig.z; d3[0] = r.dir.x; d3[1] = r.dir.y; d3[2] = r.dir.z;
}
int fjo_generate_samples(const fjgpu_render_params *p, const fjgpu_tile *t, double *uv, int max_samples) {
  std::vector<Sample> s; int ns[2]; generate_samples(*p, *t, s, ns);
  const int n = std::min<int>((int)s.size(), max_samples);
  for (int i = 0; i < n; i++) { uv[2 * i] = s[i].u; uv[2 * i + 1] = s[i].v; }
  return (int)s.size();
}
double fjo_gaussian(double xw, double yw, double x, double y) { return eval_gaussian(xw, yw, x, y); }
double fjo_fresnel(const double *I, const double *N, double ior) { return SlFresnel(V3(I[0], I[1], I[2]), V3(N[0], N[1], N[2]), ior); }
void fjo_reflect(const double *I, const double *N, double *R) { const V3 r = SlReflect(V3(I[0], I[1], I[2]), V3(N[0], N[1], N[2])); R[0] = r.x; R[1] = r.y; R[2] = r.z; }
void fjo_refract(const double *I, const double *N, double ior, double *T) { const V3 r = SlRefract(V3(I[0], I[1], I[2]), V3(N[0], N[1], N[2]), ior); T[0] = r.x; T[1] = r.y; T[2] = r.z; }
void fjo_philox(uint32_t seed, uint32_t tile, uint32_t sample, uint64_t node, uint32_t dim, double *out) {
  RenderState rs; rs.seed = seed; rs.tile_id = tile; rs.sample_id = sample; *out = ctr_rand(rs, node, dim);
}

}  // extern "C"
