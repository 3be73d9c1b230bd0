// fj_scene_host.cc — libfjscene.so: C++ host mirror of the reference's scene interface (see include/fjscene.h).
//
// Keeps the scene the way the reference's `Scene` does (entities addressed by type-tagged IDs,
// src/fj_scene_interface.cc:40-60,1043-1075), records every property as it is set, and at SiRenderScene
// flattens it into the structs of include/fjgpu.h: instance matrices with the reference's own arithmetic
// (make_transform_matrix + MatInverse, src/fj_transform.cc:335-391, src/fj_matrix.cc:119-207), implicit object
// groups (create_implicit_groups, src/fj_scene_interface.cc:1077-1135), shader parameters after the clamping
// the plugin setters apply.  The frame itself runs in libfjgpu.so — this file has no renderer and no CPU
// fallback.  Compiled with -ffp-contract=off so the FP64 matrices are bit-identical to the reference's.
#include "fjscene.h"

#include <algorithm>
#include <array>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace fj {
namespace {

const double PI = 3.14159265358979323846;
const ID TYPE_ID_OFFSET = 10000000;            // src/fj_scene_interface.cc:42
enum EntryType {                               // :44-62 (same numbering, so IDs agree with the reference)
  Type_ObjectInstance = 1, Type_Accelerator, Type_FrameBuffer, Type_ObjectGroup, Type_PointCloud, Type_Turbulence,
  Type_Procedure, Type_Renderer, Type_Texture, Type_Camera, Type_Plugin, Type_Shader, Type_Volume, Type_Curve,
  Type_Light, Type_Mesh, Type_End
};
inline ID encode_id(int type, int index) { return TYPE_ID_OFFSET * type + index; }
inline bool decode_id(ID id, int *type, int *index) {
  const int t = (int)(id / TYPE_ID_OFFSET);
  if (id < 0 || t <= 0 || t >= Type_End) return false;
  *type = t; *index = (int)(id - (ID)t * TYPE_ID_OFFSET);
  return true;
}

// ---------------------------------------------------------------- 4x4 matrices (row-major, src/fj_matrix.cc)
struct M4 { double e[16]; };
M4 identity() { M4 m; for (int i = 0; i < 16; i++) m.e[i] = (i % 5 == 0) ? 1. : 0.; return m; }
M4 mul(const M4 &a, const M4 &b) {             // MatMultiply :104-117
  M4 c;
  for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) {
    double s = 0.;
    for (int k = 0; k < 4; k++) s += a.e[4 * j + k] * b.e[4 * k + i];
    c.e[4 * j + i] = s;
  }
  return c;
}
inline double radian(double deg) { return deg * PI / 180.; }   // src/fj_numeric.h:56-59
M4 rot(int axis, double deg) {                 // MatRotateX/Y/Z :66-100
  const double s = std::sin(radian(deg)), c = std::cos(radian(deg));
  M4 m = identity();
  if (axis == 0) { m.e[5] = c; m.e[6] = -s; m.e[9] = s; m.e[10] = c; }
  else if (axis == 1) { m.e[0] = c; m.e[2] = s; m.e[8] = -s; m.e[10] = c; }
  else { m.e[0] = c; m.e[1] = -s; m.e[4] = s; m.e[5] = c; }
  return m;
}
// make_transform_matrix, src/fj_transform.cc:335-391: every stage is `out = stage * out`
M4 compose(int torder, int rorder, const double T[3], const double R[3], const double S[3]) {
  M4 t = identity(); t.e[3] = T[0]; t.e[7] = T[1]; t.e[11] = T[2];
  M4 s = identity(); s.e[0] = S[0]; s.e[5] = S[1]; s.e[10] = S[2];
  const M4 rx = rot(0, R[0]), ry = rot(1, R[1]), rz = rot(2, R[2]);
  static const int ROT[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};   // XYZ XZY YXZ YZX ZXY ZYX
  const int ri = (rorder >= SI_ORDER_XYZ && rorder <= SI_ORDER_ZYX) ? rorder - SI_ORDER_XYZ : 5;
  const M4 *axes[3] = {&rx, &ry, &rz};
  M4 r = identity();
  for (int i = 0; i < 3; i++) r = mul(*axes[ROT[ri][i]], r);
  static const char ORD[6][4] = {"SRT", "STR", "RST", "RTS", "TRS", "TSR"};
  const int ti = (torder >= SI_ORDER_SRT && torder <= SI_ORDER_TSR) ? torder : 5;
  M4 out = identity();
  for (int i = 0; i < 3; i++) {
    const char c = ORD[ti][i];
    out = mul(c == 'S' ? s : (c == 'R' ? r : t), out);
  }
  return out;
}
// MatInverse, src/fj_matrix.cc:119-207: Cramer's rule in the arrangement of Intel AP-928 ("Streaming SIMD
// Extensions - Inverse of 4x4 Matrix"): transpose, 12 pair products per half, cofactors, one reciprocal.
M4 inverse(const M4 &a) {
  double s[16], p[12]; M4 out; double *d = out.e;
  for (int i = 0; i < 4; i++) { s[i] = a.e[4 * i]; s[i + 4] = a.e[4 * i + 1]; s[i + 8] = a.e[4 * i + 2]; s[i + 12] = a.e[4 * i + 3]; }
  p[0] = s[10] * s[15]; p[1] = s[11] * s[14]; p[2] = s[9] * s[15]; p[3] = s[11] * s[13]; p[4] = s[9] * s[14]; p[5] = s[10] * s[13];
  p[6] = s[8] * s[15]; p[7] = s[11] * s[12]; p[8] = s[8] * s[14]; p[9] = s[10] * s[12]; p[10] = s[8] * s[13]; p[11] = s[9] * s[12];
  d[0] = p[0] * s[5] + p[3] * s[6] + p[4] * s[7];   d[0] -= p[1] * s[5] + p[2] * s[6] + p[5] * s[7];
  d[1] = p[1] * s[4] + p[6] * s[6] + p[9] * s[7];   d[1] -= p[0] * s[4] + p[7] * s[6] + p[8] * s[7];
  d[2] = p[2] * s[4] + p[7] * s[5] + p[10] * s[7];  d[2] -= p[3] * s[4] + p[6] * s[5] + p[11] * s[7];
  d[3] = p[5] * s[4] + p[8] * s[5] + p[11] * s[6];  d[3] -= p[4] * s[4] + p[9] * s[5] + p[10] * s[6];
  d[4] = p[1] * s[1] + p[2] * s[2] + p[5] * s[3];   d[4] -= p[0] * s[1] + p[3] * s[2] + p[4] * s[3];
  d[5] = p[0] * s[0] + p[7] * s[2] + p[8] * s[3];   d[5] -= p[1] * s[0] + p[6] * s[2] + p[9] * s[3];
  d[6] = p[3] * s[0] + p[6] * s[1] + p[11] * s[3];  d[6] -= p[2] * s[0] + p[7] * s[1] + p[10] * s[3];
  d[7] = p[4] * s[0] + p[9] * s[1] + p[10] * s[2];  d[7] -= p[5] * s[0] + p[8] * s[1] + p[11] * s[2];
  p[0] = s[2] * s[7]; p[1] = s[3] * s[6]; p[2] = s[1] * s[7]; p[3] = s[3] * s[5]; p[4] = s[1] * s[6]; p[5] = s[2] * s[5];
  p[6] = s[0] * s[7]; p[7] = s[3] * s[4]; p[8] = s[0] * s[6]; p[9] = s[2] * s[4]; p[10] = s[0] * s[5]; p[11] = s[1] * s[4];
  d[8] = p[0] * s[13] + p[3] * s[14] + p[4] * s[15];    d[8] -= p[1] * s[13] + p[2] * s[14] + p[5] * s[15];
  d[9] = p[1] * s[12] + p[6] * s[14] + p[9] * s[15];    d[9] -= p[0] * s[12] + p[7] * s[14] + p[8] * s[15];
  d[10] = p[2] * s[12] + p[7] * s[13] + p[10] * s[15];  d[10] -= p[3] * s[12] + p[6] * s[13] + p[11] * s[15];
  d[11] = p[5] * s[12] + p[8] * s[13] + p[11] * s[14];  d[11] -= p[4] * s[12] + p[9] * s[13] + p[10] * s[14];
  d[12] = p[2] * s[10] + p[5] * s[11] + p[1] * s[9];    d[12] -= p[4] * s[11] + p[0] * s[9] + p[3] * s[10];
  d[13] = p[8] * s[11] + p[0] * s[8] + p[7] * s[10];    d[13] -= p[6] * s[10] + p[9] * s[11] + p[1] * s[8];
  d[14] = p[6] * s[9] + p[11] * s[11] + p[3] * s[8];    d[14] -= p[10] * s[11] + p[2] * s[8] + p[7] * s[9];
  d[15] = p[10] * s[10] + p[4] * s[8] + p[9] * s[9];    d[15] -= p[8] * s[9] + p[11] * s[10] + p[5] * s[8];
  const double det = 1. / (s[0] * d[0] + s[1] * d[1] + s[2] * d[2] + s[3] * d[3]);
  for (int j = 0; j < 16; j++) d[j] *= det;
  return out;
}

// ---------------------------------------------------------------- entities
struct TimeSamples {                           // PropertySampleList, src/fj_property.h:117-137 (<= 8 samples)
  std::vector<std::pair<double, std::array<double, 3>>> s;
  explicit TimeSamples(double v) { s.push_back({0., {v, v, v}}); }
  bool push(double x, double y, double z, double time) {          // PropPushSample, src/fj_property.cc:294-312
    if (s.size() >= 8) return false;                               // checked before the equal-time replacement, as the reference does
    for (auto &e : s) if (e.first == time) { e.second = {x, y, z}; return true; }
    s.push_back({time, {x, y, z}});
    std::stable_sort(s.begin(), s.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    return true;
  }
  bool is_static() const { return s.size() == 1; }
  // PropLerpSamples, src/fj_property.cc:317-345 (Fit src/fj_numeric.h:84-93, VEC4_LERP :18-23)
  std::array<double, 3> at(double time) const {
    if (s.front().first >= time || s.size() == 1) return s.front().second;
    if (s.back().first <= time) return s.back().second;
    for (size_t i = 0; i < s.size(); i++) {
      if (s[i].first == time) return s[i].second;
      if (s[i].first > time) {
        const double t0 = s[i - 1].first, t1 = s[i].first;
        const double t = time <= t0 ? 0. : (time >= t1 ? 1. : 0. + (1. - 0.) * ((time - t0) / (t1 - t0)));
        const std::array<double, 3> &a = s[i - 1].second, &b = s[i].second;
        return {(1 - t) * a[0] + t * b[0], (1 - t) * a[1] + t * b[1], (1 - t) * a[2] + t * b[2]};
      }
    }
    return s.back().second;
  }
};
struct Xform {
  int torder = SI_ORDER_SRT, rorder = SI_ORDER_ZXY;               // XfmInitTransformSampleList, fj_transform.cc:240-254
  TimeSamples T{0.}, R{0.}, S{1.};
  bool is_static() const { return T.is_static() && R.is_static() && S.is_static(); }
  M4 matrix() const { return compose(torder, rorder, T.s[0].second.data(), R.s[0].second.data(), S.s[0].second.data()); }
  // XfmLerpTransformSample, src/fj_transform.cc:306-322: the transform the reference rebuilds for a ray of time `time`
  M4 matrix_at(double time) const { const auto t = T.at(time), r = R.at(time), sc = S.at(time); return compose(torder, rorder, t.data(), r.data(), sc.data()); }
};

struct Plugin { std::string name; int kind; };   // kind: FJGPU_SHADER_* for shaders, 100 = StanfordPlyProcedure
struct Mesh { std::vector<double> P, N; std::vector<int32_t> idx; std::vector<float> uv; bool dirty = true;
              std::vector<double> vel; };     // per-vertex velocity (VelocityGeneratorProcedure); empty = none
struct Texture { int width = 0, height = 0, nch = 0, tilesize = 0; std::vector<float> tiles; };   // a `.mip` file's header and tiles (src/fj_mipmap.cc:124-180)
struct Shader { int plugin; std::map<std::string, std::array<double, 4>> props; ID texture = SI_BADID, bump = SI_BADID; };   // texture: the `texture` / `diffuse_map` property
struct Procedure { int plugin; ID mesh = SI_BADID; std::string filepath, io_mode; };
struct Instance { ID mesh; Xform x; std::map<std::string, ID> shaders; ID reflect = SI_BADID, refract = SI_BADID, shadow = SI_BADID; };
struct Group { std::vector<int> members; };
struct Light { int type; Xform x; double intensity = 1; double color[3] = {1, 1, 1}; int sample_count = 16; int double_sided = 0; ID envmap = SI_BADID; };
struct Camera { Xform x; double fov = 30, znear = .01, zfar = 1000; };
struct FrameBuf { int w = 0, h = 0, c = 4; std::vector<float> px; };
struct Renderer {
  ID camera = SI_BADID, fb = SI_BADID;
  int res[2] = {320, 240}, tile[2] = {32, 32}, rate[2] = {3, 3}, region[4] = {0, 0, 320, 240};
  double fw[2] = {2, 2}, jitter = 1, time_range[2] = {0, 1};
  int cast_shadow = 1, max_diffuse = 3, max_reflect = 3, max_refract = 3, sampler_type = 0, use_max_thread = 1, thread_count = 8;
  void *frame_data = nullptr; FrameStartCallback frame_start = nullptr; FrameAbortCallback frame_abort = nullptr; FrameDoneCallback frame_done = nullptr;
  void *tile_data = nullptr; TileStartCallback tile_start = nullptr; TileDoneCallback tile_done = nullptr;
};

struct Scene {
  std::vector<Texture> textures; bool textures_dirty = true;
  std::vector<Plugin> plugins; std::vector<Mesh> meshes; std::vector<Shader> shaders; std::vector<Procedure> procedures;
  std::vector<Instance> instances; std::vector<Group> groups; std::vector<Light> lights; std::vector<Camera> cameras;
  std::vector<FrameBuf> framebuffers; std::vector<Renderer> renderers;
  fjgpu_context *gpu = nullptr; int gpu_device = -1;
  std::vector<fjgpu_context *> more_gpus;              // FJ_GPU_COUNT > 1: contexts on the devices after gpu_device, same scene on each
  std::vector<fjgpu_instance> flat_inst;
  std::vector<fjgpu_tile> last_tiles; int last_res[2] = {0, 0};      // the whole tile list of the last frame (fjscene_assemble_gathered)
  std::map<int, std::vector<double>> motion_key;      // what the motion table held by libfjgpu for instance i (-1: camera) was built from
  ~Scene() { for (fjgpu_context *c : more_gpus) fjgpu_destroy(c); if (gpu) fjgpu_destroy(gpu); }
};

Scene *the_scene = nullptr;
int si_errno = SI_ERR_NONE;
std::string last_message;
int g_device = 0, g_rank = 0, g_world = 1, g_resident = 0, g_resend = 0;
int g_gpu_count = 0;       // GPUs one process renders a frame on (fjscene_set_gpu_count / FJ_GPU_COUNT); 0 = not set
uint64_t g_resend_bytes = 0;
void *g_dev_blocks = nullptr; int g_dev_bw = 0, g_dev_bh = 0;
fjgpu_stats g_stats; fjgpu_scene_info g_info; double g_upload_seconds = 0; int32_t g_frame_id = 0;

Status ok() { si_errno = SI_ERR_NONE; return SI_SUCCESS; }
ID bad(int err) { si_errno = err; return SI_BADID; }
Status failmsg(const std::string &m) { last_message = m; fprintf(stderr, "fjscene: %s\n", m.c_str()); return SI_FAIL; }

template <typename V> V *get(std::vector<V> &v, ID id, int type) {
  int t, i;
  if (!the_scene || !decode_id(id, &t, &i) || t != type || i < 0 || i >= (int)v.size()) return nullptr;
  return &v[i];
}

// Mesh::ComputeNormals, src/fj_mesh.cc:195-233: unweighted sum of unit face normals, then normalise.
void compute_normals(Mesh &m) {
  const size_t nv = m.P.size() / 3, nf = m.idx.size() / 3;
  m.N.assign(3 * nv, 0.);
  auto nrm = [](double v[3]) { const double len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); if (len == 0) return; const double inv = 1. / len; v[0] *= inv; v[1] *= inv; v[2] *= inv; };
  for (size_t f = 0; f < nf; f++) {
    const int i0 = m.idx[3 * f], i1 = m.idx[3 * f + 1], i2 = m.idx[3 * f + 2];
    const double *p0 = &m.P[3 * (size_t)i0], *p1 = &m.P[3 * (size_t)i1], *p2 = &m.P[3 * (size_t)i2];
    const double a[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, b[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
    double g[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    nrm(g);
    double n0[3], n1[3], n2[3];     // the three sums are formed from the values read BEFORE any is written back
    for (int k = 0; k < 3; k++) { n0[k] = m.N[3 * (size_t)i0 + k] + g[k]; n1[k] = m.N[3 * (size_t)i1 + k] + g[k]; n2[k] = m.N[3 * (size_t)i2 + k] + g[k]; }
    for (int k = 0; k < 3; k++) m.N[3 * (size_t)i0 + k] = n0[k];
    for (int k = 0; k < 3; k++) m.N[3 * (size_t)i1 + k] = n1[k];
    for (int k = 0; k < 3; k++) m.N[3 * (size_t)i2 + k] = n2[k];
  }
  for (size_t v = 0; v < nv; v++) nrm(&m.N[3 * v]);
}

// ---------------------------------------------------------------- PLY (the layouts ply2mesh.cc:32-49 accepts)
struct PlyProp { std::string name, type, ctype; bool list; };
int ply_size(const std::string &t) {
  if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
  if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
  if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
  if (t == "double" || t == "float64") return 8;
  return 0;
}
double ply_scalar(const unsigned char *p, const std::string &t, bool swap) {
  unsigned char b[8]; const int n = ply_size(t);
  for (int i = 0; i < n; i++) b[i] = swap ? p[n - 1 - i] : p[i];
  if (t == "char" || t == "int8") return (double)*(int8_t *)b;
  if (t == "uchar" || t == "uint8") return (double)*(uint8_t *)b;
  if (t == "short" || t == "int16") { int16_t v; memcpy(&v, b, 2); return v; }
  if (t == "ushort" || t == "uint16") { uint16_t v; memcpy(&v, b, 2); return v; }
  if (t == "int" || t == "int32") { int32_t v; memcpy(&v, b, 4); return v; }
  if (t == "uint" || t == "uint32") { uint32_t v; memcpy(&v, b, 4); return v; }
  if (t == "float" || t == "float32") { float v; memcpy(&v, b, 4); return v; }
  double v; memcpy(&v, b, 8); return v;
}
// ---------------------------------------------------------------- VelocityGeneratorProcedure
// procedures/velocity_generator_procedure/velocity_generator_procedure.cc:101-124 over Ken Perlin's improved noise as
// src/fj_noise.cc:33-128 evaluates it (published permutation, quintic fade, 12 gradients; octaves with amplitude *= persistence,
// position *= lacunarity; fixed offsets for the y and z components) and SmoothStep (src/fj_numeric.h:66-77).
const unsigned char kPerlin[256] = {
  151,160,137,91,90,15,131,13,201,95,96,53,194,233,7,225,140,36,103,30,69,142,8,99,37,240,21,10,23,190,6,148,
  247,120,234,75,0,26,197,62,94,252,219,203,117,35,11,32,57,177,33,88,237,149,56,87,174,20,125,136,171,168,
  68,175,74,165,71,134,139,48,27,166,77,146,158,231,83,111,229,122,60,211,133,230,220,105,92,41,55,46,245,40,
  244,102,143,54,65,25,63,161,1,216,80,73,209,76,132,187,208,89,18,169,200,196,135,130,116,188,159,86,164,100,
  109,198,173,186,3,64,52,217,226,250,124,123,5,202,38,147,118,126,255,82,85,212,207,206,59,227,47,16,58,17,
  182,189,28,42,223,183,170,213,119,248,152,2,44,154,163,70,221,153,101,155,167,43,172,9,129,22,39,253,19,98,
  108,110,79,113,224,232,178,185,112,104,218,246,97,228,251,34,242,193,238,210,144,12,191,179,162,241,81,51,
  145,235,249,14,239,107,49,192,214,31,181,199,106,157,184,84,204,176,115,121,50,45,127,4,150,254,138,236,
  205,93,222,114,67,29,24,72,243,141,128,195,78,66,215,61,156,180};
inline int pperm(int i) { return kPerlin[i & 255]; }
inline double pfade(double t) { return t * t * t * (t * (t * 6 - 15) + 10); }
inline double plerp(double t, double a, double b) { return a + t * (b - a); }
inline double pgrad(int hash, double x, double y, double z) {
  const int h = hash & 15;
  const double u = h < 8 ? x : y, v = h < 4 ? y : (h == 12 || h == 14 ? x : z);
  return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
}
double periodic_noise(double x, double y, double z) {
  const int X = (int)std::floor(x) & 255, Y = (int)std::floor(y) & 255, Z = (int)std::floor(z) & 255;
  const double xx = x - std::floor(x), yy = y - std::floor(y), zz = z - std::floor(z);
  const double u = pfade(xx), v = pfade(yy), w = pfade(zz);
  const int A = pperm(X) + Y, AA = pperm(A) + Z, AB = pperm(A + 1) + Z, B = pperm(X + 1) + Y, BA = pperm(B) + Z, BB = pperm(B + 1) + Z;
  return plerp(w, plerp(v, plerp(u, pgrad(pperm(AA), xx, yy, zz), pgrad(pperm(BA), xx - 1, yy, zz)),
                           plerp(u, pgrad(pperm(AB), xx, yy - 1, zz), pgrad(pperm(BB), xx - 1, yy - 1, zz))),
                  plerp(v, plerp(u, pgrad(pperm(AA + 1), xx, yy, zz - 1), pgrad(pperm(BA + 1), xx - 1, yy, zz - 1)),
                           plerp(u, pgrad(pperm(AB + 1), xx, yy - 1, zz - 1), pgrad(pperm(BB + 1), xx - 1, yy - 1, zz - 1))));
}
double perlin_noise(double x, double y, double z, double lacunarity, double persistence, int octaves) {
  double value = 0, amp = 1;
  for (int i = 0; i < octaves; i++) { value += amp * periodic_noise(x, y, z); amp *= persistence; x *= lacunarity; y *= lacunarity; z *= lacunarity; }
  return value;
}
inline double smooth_step(double a, double b, double x) { const double t = (x - a) / (b - a); return t <= 0 ? 0 : (t >= 1 ? 1 : t * t * (3 - 2 * t)); }
int generate_velocity(Mesh *m) {
  if (m->idx.empty()) return -1;
  double zmin = DBL_MAX, zmax = -DBL_MAX;              // Mesh::GetBounds: over the vertices the faces refer to (fj_mesh.cc:235-244)
  for (int32_t v : m->idx) { zmin = std::min(zmin, m->P[3 * (size_t)v + 2]); zmax = std::max(zmax, m->P[3 * (size_t)v + 2]); }
  const size_t n = m->P.size() / 3;
  m->vel.assign(3 * n, 0.);
  for (size_t i = 0; i < n; i++) {
    const double px = m->P[3 * i], py = m->P[3 * i + 1], pz = m->P[3 * i + 2];
    const double znml = (pz - zmin) / (zmax - zmin);
    const double vscale = .2 * (1 - smooth_step(.2, .7, znml));
    const double qx = px * .2, qy = py * .2, qz = pz * .2;                       // `.2 * pos`
    const double nx = perlin_noise(qx, qy, qz, 2, .5, 1);
    const double ny = perlin_noise(qx + 131.977, qy + 21.1823, qz + 71.0231, 2, .5, 1);
    const double nz = perlin_noise(qx + 237.492, qy + 11.1312, qz + 133.129, 2, .5, 1);
    m->vel[3 * i] = nx * vscale; m->vel[3 * i + 1] = ny * vscale; m->vel[3 * i + 2] = nz * vscale;    // `vscale * noise_vec`
  }
  m->dirty = true;
  return 0;
}

int read_ply(const std::string &path, Mesh *mesh) {
  std::ifstream f(path.c_str(), std::ios::binary);
  if (!f) { fprintf(stderr, "error: couldn't open input file: %s\n", path.c_str()); return -1; }
  std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const size_t hend = data.find("end_header");
  if (data.compare(0, 3, "ply") != 0 || hend == std::string::npos) return -1;
  size_t body = data.find('\n', hend); if (body == std::string::npos) return -1; body++;
  std::istringstream hdr(data.substr(0, hend));
  std::string line, fmt;
  struct Elem { std::string name; long count; std::vector<PlyProp> props; };
  std::vector<Elem> elems;
  while (std::getline(hdr, line)) {
    std::istringstream ls(line); std::string w; ls >> w;
    if (w == "format") ls >> fmt;
    else if (w == "element") { Elem e; ls >> e.name >> e.count; elems.push_back(e); }
    else if (w == "property" && !elems.empty()) {
      PlyProp p; std::string t; ls >> t;
      if (t == "list") { p.list = true; ls >> p.ctype >> p.type >> p.name; } else { p.list = false; p.type = t; ls >> p.name; }
      elems.back().props.push_back(p);
    }
  }
  const bool ascii = fmt == "ascii", swap = fmt == "binary_big_endian";
  if (!ascii && !swap && fmt != "binary_little_endian") return -1;
  const unsigned char *p = (const unsigned char *)data.data() + body, *end = (const unsigned char *)data.data() + data.size();
  std::istringstream as; if (ascii) as.str(data.substr(body));
  mesh->P.clear(); mesh->idx.clear(); mesh->uv.clear();
  bool has_uv = false;
  for (const Elem &e : elems) if (e.name == "vertex") for (const PlyProp &pr : e.props) has_uv = has_uv || pr.name == "uv1" || pr.name == "uv2";
  for (const Elem &e : elems) {
    for (long i = 0; i < e.count; i++) {
      double xyz[3] = {0, 0, 0}, tuv[2] = {0, 0}; std::vector<int32_t> poly;
      for (const PlyProp &pr : e.props) {
        if (!pr.list) {
          double v;
          if (ascii) { if (!(as >> v)) return -1; } else { const int n = ply_size(pr.type); if (!n || p + n > end) return -1; v = ply_scalar(p, pr.type, swap); p += n; }
          if (pr.name == "x") xyz[0] = v; else if (pr.name == "y") xyz[1] = v; else if (pr.name == "z") xyz[2] = v;
          else if (pr.name == "uv1") tuv[0] = v; else if (pr.name == "uv2") tuv[1] = v;     // ply2mesh.cc:42-43,91-98
        } else {
          long cnt;
          if (ascii) { double c; if (!(as >> c)) return -1; cnt = (long)c; } else { const int n = ply_size(pr.ctype); if (!n || p + n > end) return -1; cnt = (long)ply_scalar(p, pr.ctype, swap); p += n; }
          for (long k = 0; k < cnt; k++) {
            double v;
            if (ascii) { if (!(as >> v)) return -1; } else { const int n = ply_size(pr.type); if (!n || p + n > end) return -1; v = ply_scalar(p, pr.type, swap); p += n; }
            if (e.name == "face" && pr.name == "vertex_indices") poly.push_back((int32_t)v);
          }
        }
      }
      if (e.name == "vertex") { mesh->P.push_back(xyz[0]); mesh->P.push_back(xyz[1]); mesh->P.push_back(xyz[2]); if (has_uv) { mesh->uv.push_back((float)tuv[0]); mesh->uv.push_back((float)tuv[1]); } }
      else if (e.name == "face")          // fan triangulation, ply2mesh.cc:129-136
        for (size_t k = 0; k + 2 < poly.size(); k++) { mesh->idx.push_back(poly[0]); mesh->idx.push_back(poly[k + 1]); mesh->idx.push_back(poly[k + 2]); }
    }
  }
  const int32_t nv = (int32_t)(mesh->P.size() / 3);
  for (int32_t i : mesh->idx) if (i < 0 || i >= nv) return -1;
  compute_normals(*mesh);
  mesh->dirty = true;
  return 0;
}

// ---------------------------------------------------------------- environment-map dome lights (host side of Light::Preprocess)
// TextureCache::LookupTexture (src/fj_texture.cc:51-78) on the tiles SiNewTexture read: float arithmetic as compiled.
void host_tex_lookup(const Texture &t, float u, float v, float rgba[4]) {
  const int xnt = t.width / t.tilesize, ynt = t.height / t.tilesize;
  const float tsu = u - std::floor(u), tsv = v - std::floor(v);
  const float tlu = tsu * xnt, tlv = (1 - tsv) * ynt;
  int xtile = (int)std::floor(tlu), ytile = (int)std::floor(tlv);
  const int xpxl = (int)((tlu - std::floor(tlu)) * 64), ypxl = (int)((tlv - std::floor(tlv)) * 64);
  xtile = std::min(std::max(xtile, 0), xnt - 1); ytile = std::min(std::max(ytile, 0), ynt - 1);      // MipInput::ReadTile :163-165
  const float *px = &t.tiles[((size_t)(ytile * xnt + xtile) * t.tilesize * t.tilesize + (size_t)ypxl * t.tilesize + xpxl) * t.nch];
  if (t.nch == 1) { rgba[0] = rgba[1] = rgba[2] = px[0]; rgba[3] = 1; }
  else if (t.nch == 3) { rgba[0] = px[0]; rgba[1] = px[1]; rgba[2] = px[2]; rgba[3] = 1; }
  else { rgba[0] = px[0]; rgba[1] = px[1]; rgba[2] = px[2]; rgba[3] = px[3]; }
}
// DomeLight::preprocess with an environment map (src/fj_dome_light.cc:78-96) -> StratifiedImportanceSampling
// (src/fj_importance_sampling.cc:102-150, helpers :222-277): the map sampled at 1/8 of its resolution, one pick per stratum.
bool dome_samples_from_envmap(const Texture &t, int nsamples, std::vector<double> *dirs, std::vector<float> *cols) {
  const int xres = t.width / 8, yres = t.height / 8, npixels = xres * yres;
  if (npixels <= 0 || t.tilesize < 64) return false;
  const double PI = 3.14159265358979323846;
  auto index_to_uv = [&](int index, float *u, float *v) { const int x = index % xres, y = index / xres; *u = (.5 + x) / xres; *v = 1. - ((.5 + y) / yres); };
  std::vector<double> hist(npixels);
  double sum = 0;
  for (int i = 0; i < npixels; i++) {
    float u, v, c[4]; index_to_uv(i, &u, &v); host_tex_lookup(t, u, v, c);
    sum += (float)(.298912 * c[0] + .586611 * c[1] + .114478 * c[2]);
    hist[i] = sum;
  }
  uint32_t s[4] = {123456789u, 362436069u, 521288629u, 88675123u};                                      // XorShift, src/fj_random.cc:10-43
  dirs->clear(); cols->clear();
  for (int i = 0; i < nsamples; i++) {
    const uint32_t tt = s[0] ^ (s[0] << 11);
    s[0] = s[1]; s[1] = s[2]; s[2] = s[3]; s[3] = (s[3] ^ (s[3] >> 19)) ^ (tt ^ (tt >> 8));
    const double f01 = (double)s[3] / 4294967295.0;
    const double rnd = sum * ((i + f01) / nsamples);
    int index = -1;
    for (int k = 0; k < npixels; k++) if (rnd < hist[k]) { index = k; break; }
    float u, v, c[4]; index_to_uv(index, &u, &v);
    const double phi = 2 * PI * u, theta = PI * (v - .5), r = std::cos(theta);
    dirs->insert(dirs->end(), {r * std::sin(phi), std::sin(theta), r * std::cos(phi)});
    host_tex_lookup(t, u, v, c);
    cols->insert(cols->end(), {c[0], c[1], c[2]});
  }
  return true;
}

// ---------------------------------------------------------------- shader property tables
// name -> default; the first three components are used for vectors.  constant_shader.cc:29-33,
// plastic_shader.cc:50-62, pathtracing_shader.cc:71-86, glass_shader.cc:48-56.
typedef std::map<std::string, std::array<double, 4>> PropMap;
PropMap shader_defaults(int kind) {
  PropMap m;
  if (kind == FJGPU_SHADER_CONSTANT) { m["diffuse"] = {1, 1, 1, 0}; m["texture"] = {0, 0, 0, 0}; }
  else if (kind == FJGPU_SHADER_PLASTIC) {
    m["diffuse"] = {.8, .8, .8, 0}; m["specular"] = {1, 1, 1, 0}; m["ambient"] = {1, 1, 1, 0}; m["roughness"] = {.1, 0, 0, 0};
    m["reflect"] = {1, 1, 1, 0}; m["ior"] = {1.4, 0, 0, 0}; m["opacity"] = {1, 0, 0, 0}; m["bump_amplitude"] = {1, 0, 0, 0};
    m["diffuse_map"] = {0, 0, 0, 0}; m["bump_map"] = {0, 0, 0, 0};
  } else if (kind == FJGPU_SHADER_GLASS) {                                  // glass_shader.cc:48-56
    m["diffuse"] = {0, 0, 0, 0}; m["specular"] = {1, 1, 1, 0}; m["ambient"] = {1, 1, 1, 0}; m["filter_color"] = {1, 1, 1, 0};
    m["roughness"] = {.1, 0, 0, 0}; m["ior"] = {1.4, 0, 0, 0};
  } else {
    m["emission"] = {0, 0, 0, 0}; m["diffuse"] = {.8, .8, .8, 0}; m["specular"] = {0, 0, 0, 0}; m["ambient"] = {1, 1, 1, 0};
    m["transmit"] = {1, 1, 1, 0}; m["roughness"] = {.1, 0, 0, 0}; m["reflect"] = {0, 0, 0, 0}; m["refract"] = {0, 0, 0, 0};
    m["ior"] = {1.4, 0, 0, 0}; m["opacity"] = {1, 0, 0, 0}; m["bump_amplitude"] = {1, 0, 0, 0};
    m["diffuse_map"] = {0, 0, 0, 0}; m["bump_map"] = {0, 0, 0, 0};
  }
  return m;
}
inline float clamp0(double v) { return (float)(v > 0 ? v : 0); }          // Max(0, value.vector[i]) stored in a float
fjgpu_shader flatten_shader(const Scene &sc, const Shader &s) {
  fjgpu_shader o; memset(&o, 0, sizeof o);
  const int kind = sc.plugins[s.plugin].kind;
  o.kind = kind;
  auto P = [&](const char *n) { return s.props.find(n)->second; };
  { int tt, ti; if (s.texture != SI_BADID && decode_id(s.texture, &tt, &ti) && tt == Type_Texture && kind != FJGPU_SHADER_GLASS) o.texture = ti + 1; }
  { int tt, ti; if (s.bump != SI_BADID && decode_id(s.bump, &tt, &ti) && tt == Type_Texture && (kind == FJGPU_SHADER_PLASTIC || kind == FJGPU_SHADER_PATHTRACING)) o.bump_texture = ti + 1; }
  if (kind == FJGPU_SHADER_PLASTIC || kind == FJGPU_SHADER_PATHTRACING) o.bump_amplitude = (float)P("bump_amplitude")[0];     // set_bump_amplitude, default 1 (used only with a bump_map)
  if (kind == FJGPU_SHADER_CONSTANT) {                                      // constant_shader.cc:96-107
    for (int k = 0; k < 3; k++) o.diffuse[k] = clamp0(P("diffuse")[k]);
  } else if (kind == FJGPU_SHADER_PLASTIC) {                                // plastic_shader.cc:183-273
    for (int k = 0; k < 3; k++) { o.diffuse[k] = clamp0(P("diffuse")[k]); o.reflect[k] = clamp0(P("reflect")[k]); }
    o.do_reflect = (o.reflect[0] > 0 || o.reflect[1] > 0 || o.reflect[2] > 0) ? 1 : 0;
    float ior = (float)P("ior")[0]; o.ior = (float)std::max(.001, (double)ior);
    float op = (float)P("opacity")[0]; o.opacity = op < 0 ? 0 : (op > 1 ? 1 : op);
  } else if (kind == FJGPU_SHADER_GLASS) {                                  // glass_shader.cc:135-213
    for (int k = 0; k < 3; k++) o.transmit[k] = (float)std::max(.001, P("filter_color")[k]);
    o.do_color_filter = (o.transmit[0] == 1 && o.transmit[1] == 1 && o.transmit[2] == 1) ? 0 : 1;
    float ior = (float)P("ior")[0]; o.ior = ior > 0 ? ior : 0;
    o.opacity = 1;
  } else {                                                                  // pathtracing_shader.cc:305-420
    for (int k = 0; k < 3; k++) {
      o.emission[k] = clamp0(P("emission")[k]); o.diffuse[k] = clamp0(P("diffuse")[k]);
      o.reflect[k] = clamp0(P("reflect")[k]); o.refract[k] = clamp0(P("refract")[k]);
      o.transmit[k] = (float)std::max(.001, P("transmit")[k]);
    }
    o.do_color_filter = (o.transmit[0] == 1 && o.transmit[1] == 1 && o.transmit[2] == 1) ? 0 : 1;
    float ior = (float)P("ior")[0]; o.ior = (float)std::max(.001, (double)ior);
    o.opacity = 1;
  }
  return o;
}

// ---------------------------------------------------------------- property setter
Status set_xform_prop(Xform &x, const std::string &name, const double v[4], double time, bool *found) {
  *found = true;
  if (name == "transform_order") { x.torder = (int)v[0]; return SI_SUCCESS; }
  if (name == "rotate_order") { x.rorder = (int)v[0]; return SI_SUCCESS; }
  if (name == "translate") return x.T.push(v[0], v[1], v[2], time) ? SI_SUCCESS : SI_FAIL;
  if (name == "rotate") return x.R.push(v[0], v[1], v[2], time) ? SI_SUCCESS : SI_FAIL;
  if (name == "scale") return x.S.push(v[0], v[1], v[2], time) ? SI_SUCCESS : SI_FAIL;
  *found = false;
  return SI_FAIL;
}

// set_property, src/fj_scene_interface.cc:1250-1275: unknown (type, name) pairs fail.
Status set_property(ID id, const char *name_c, const double v[4], int ncomp, double time) {
  int type, index;
  if (!the_scene || !name_c || !decode_id(id, &type, &index)) return SI_FAIL;
  const std::string name(name_c);
  Scene &sc = *the_scene;
  bool found = false;
  switch (type) {
    case Type_ObjectInstance: { Instance *o = get(sc.instances, id, type); if (!o) return SI_FAIL;
      const Status st = set_xform_prop(o->x, name, v, time, &found); if (!found) return SI_FAIL; if (st) return st; return ok(); }
    case Type_Camera: { Camera *c = get(sc.cameras, id, type); if (!c) return SI_FAIL;
      if (name == "fov") { c->fov = v[0]; return ok(); } if (name == "znear") { c->znear = v[0]; return ok(); } if (name == "zfar") { c->zfar = v[0]; return ok(); }
      if (name == "scale") return SI_FAIL;
      const Status st = set_xform_prop(c->x, name, v, time, &found); if (!found) return SI_FAIL; if (st) return st; return ok(); }
    case Type_Light: { Light *l = get(sc.lights, id, type); if (!l) return SI_FAIL;
      if (name == "intensity") { l->intensity = v[0]; return ok(); }
      if (name == "color") { l->color[0] = v[0]; l->color[1] = v[1]; l->color[2] = v[2]; return ok(); }
      if (name == "sample_count") { l->sample_count = std::max(1, (int)v[0]); return ok(); }     // Light::SetSampleCount, fj_light.cc:52-56
      if (name == "double_sided") { l->double_sided = ((int)v[0]) != 0; return ok(); }
      const Status st = set_xform_prop(l->x, name, v, time, &found); if (!found) return SI_FAIL; if (st) return st; return ok(); }
    case Type_Shader: { Shader *s = get(sc.shaders, id, type); if (!s) return SI_FAIL;
      auto it = s->props.find(name); if (it == s->props.end()) return SI_FAIL;
      for (int k = 0; k < 4; k++) it->second[k] = k < ncomp ? v[k] : 0.;
      return ok(); }
    case Type_Renderer: { Renderer *r = get(sc.renderers, id, type); if (!r) return SI_FAIL;
      // setters of src/internal/fj_property_list_include.cc:13-250 (values are truncated to int like the (int) casts there)
      if (name == "sample_jitter") r->jitter = v[0];
      else if (name == "cast_shadow") r->cast_shadow = (int)v[0];
      else if (name == "max_diffuse_depth") r->max_diffuse = (int)v[0];
      else if (name == "max_reflect_depth") r->max_reflect = (int)v[0];
      else if (name == "max_refract_depth") r->max_refract = (int)v[0];
      else if (name == "raymarch_step" || name == "raymarch_shadow_step" || name == "raymarch_diffuse_step" ||
               name == "raymarch_reflect_step" || name == "raymarch_refract_step" || name == "adaptive_max_subdivision" ||
               name == "adaptive_subdivision_threshold") { /* volume / adaptive-sampler settings: recorded nowhere, no volumes on this path */ }
      else if (name == "sample_time_range") { r->time_range[0] = v[0]; r->time_range[1] = v[1]; }
      else if (name == "resolution") { if ((int)v[0] <= 0 || (int)v[1] <= 0) return SI_FAIL; r->res[0] = (int)v[0]; r->res[1] = (int)v[1];
                                       r->region[0] = r->region[1] = 0; r->region[2] = r->res[0]; r->region[3] = r->res[1]; }   // Renderer::SetResolution resets the region
      else if (name == "tilesize") { if ((int)v[0] <= 0 || (int)v[1] <= 0) return SI_FAIL; r->tile[0] = (int)v[0]; r->tile[1] = (int)v[1]; }
      else if (name == "filterwidth") { if (!(v[0] > 0) || !(v[1] > 0)) return SI_FAIL; r->fw[0] = (float)v[0]; r->fw[1] = (float)v[1]; }   // float members, fj_renderer.cc:479-485
      else if (name == "sampler_type") r->sampler_type = (int)v[0];
      else if (name == "pixelsamples") { if ((int)v[0] <= 0 || (int)v[1] <= 0) return SI_FAIL; r->rate[0] = (int)v[0]; r->rate[1] = (int)v[1]; }
      else if (name == "render_region") { if (!((int)v[0] < (int)v[2] && (int)v[1] < (int)v[3]) || v[0] < 0 || v[1] < 0) return SI_FAIL; for (int k = 0; k < 4; k++) r->region[k] = (int)v[k]; }
      else if (name == "use_max_thread") r->use_max_thread = (int)v[0];
      else if (name == "thread_count") r->thread_count = (int)v[0];
      else return SI_FAIL;
      return ok(); }
    default: return SI_FAIL;
  }
}

// ---------------------------------------------------------------- render
struct Flat {
  std::vector<fjgpu_instance> inst; std::vector<int32_t> goff, gids; std::vector<fjgpu_shader> shaders;
  std::vector<fjgpu_light> lights; std::vector<std::vector<double>> dome_dirs; std::vector<std::vector<float>> dome_cols;
  fjgpu_camera cam; fjgpu_render_params params; std::vector<fjgpu_tile> tiles;
};

int group_index_of(ID gid) { int t, i; if (gid == SI_BADID || !decode_id(gid, &t, &i) || t != Type_ObjectGroup) return -1; return i; }

Status flatten(Scene &sc, const Renderer &r, Flat *f) {
  if (r.sampler_type != SI_FIXED_GRID_SAMPLER) return failmsg("sampler_type 1 (adaptive grid sampler) has no device implementation");
  const Camera *cam = get(sc.cameras, r.camera, Type_Camera);
  if (!cam) return failmsg("renderer has no camera");
  memset(&f->cam, 0, sizeof f->cam);
  const M4 cm = cam->x.matrix_at(0.);
  memcpy(f->cam.fwd, cm.e, sizeof cm.e); f->cam.fov = cam->fov; f->cam.znear = cam->znear; f->cam.zfar = cam->zfar;

  // create_implicit_groups, src/fj_scene_interface.cc:1077-1135: device group 0 = all objects, user groups follow
  const int ni = (int)sc.instances.size(), ng = (int)sc.groups.size();
  f->goff.assign(1, 0);
  for (int i = 0; i < ni; i++) f->gids.push_back(i);
  f->goff.push_back((int32_t)f->gids.size());
  for (int g = 0; g < ng; g++) { for (int m : sc.groups[g].members) f->gids.push_back(m); f->goff.push_back((int32_t)f->gids.size()); }

  f->shaders.clear();
  for (const Shader &s : sc.shaders) f->shaders.push_back(flatten_shader(sc, s));
  f->inst.resize(ni);
  for (int i = 0; i < ni; i++) {
    const Instance &o = sc.instances[i]; fjgpu_instance &d = f->inst[i];
    memset(&d, 0, sizeof d);
    int t, mi; decode_id(o.mesh, &t, &mi); d.mesh_id = mi;
    for (int g = 0; g < FJGPU_MAX_SHADING_GROUPS; g++) d.shader_of_group[g] = -1;
    // ObjectInstance::AddShader / GetShader, src/fj_object_instance.cc:160-191: the mesh of this path has one
    // shading group ("" = DEFAULT_SHADING_GROUP -> slot 0)
    for (auto &kv : o.shaders) { int st, si; if (decode_id(kv.second, &st, &si) && kv.first.empty()) d.shader_of_group[0] = si; }
    if (d.shader_of_group[0] < 0 && !o.shaders.empty()) { int st, si; decode_id(o.shaders.begin()->second, &st, &si); d.shader_of_group[0] = si; }
    const int gr = group_index_of(o.reflect), gf = group_index_of(o.refract), gs = group_index_of(o.shadow);
    d.reflect_target = gr < 0 ? 0 : gr + 1; d.refract_target = gf < 0 ? 0 : gf + 1; d.shadow_target = gs < 0 ? 0 : gs + 1;
    const M4 m = o.x.matrix_at(0.), inv = inverse(m);      // moving instances: replaced per ray by the motion table (render)
    memcpy(d.fwd, m.e, sizeof m.e); memcpy(d.inv, inv.e, sizeof inv.e);
  }
  const int nl = (int)sc.lights.size();
  f->lights.resize(nl); f->dome_dirs.assign(nl, {}); f->dome_cols.assign(nl, {});
  for (int i = 0; i < nl; i++) {
    const Light &l = sc.lights[i]; fjgpu_light &d = f->lights[i];
    memset(&d, 0, sizeof d);
    // lights are sampled at time 0 (`const float time = 0`, fj_point_light.cc:27-29 and the other three light types)
    d.kind = l.type; d.sample_count = l.sample_count; d.double_sided = l.double_sided;
    for (int k = 0; k < 3; k++) { d.color[k] = (float)l.color[k]; d.translate[k] = l.x.T.at(0.)[k]; }
    d.intensity = (float)l.intensity;
    const M4 m = l.x.matrix_at(0.); memcpy(d.fwd, m.e, sizeof m.e);
    int et, ei;
    if (l.type == SI_DOME_LIGHT && l.envmap != SI_BADID && decode_id(l.envmap, &et, &ei) && et == Type_Texture && ei < (int)sc.textures.size()) {
      if (!dome_samples_from_envmap(sc.textures[ei], l.sample_count, &f->dome_dirs[i], &f->dome_cols[i])) return failmsg("environment map too small to sample");
      d.dome_sample_count = l.sample_count;
    } else if (l.type == SI_DOME_LIGHT) {   // DomeLight::preprocess without an environment map, src/fj_dome_light.cc:58-76
      const int n = l.sample_count; const double a = 1. / n;
      const double len = std::sqrt(a * a + 1. * 1. + a * a), inv = 1. / len;
      for (int k = 0; k < n; k++) { f->dome_dirs[i].insert(f->dome_dirs[i].end(), {a * inv, 1. * inv, a * inv}); f->dome_cols[i].insert(f->dome_cols[i].end(), {1.f, .63f, .63f}); }
      d.dome_sample_count = n;
    }
  }
  for (int i = 0; i < nl; i++) if (f->lights[i].dome_sample_count) { f->lights[i].dome_dirs = f->dome_dirs[i].data(); f->lights[i].dome_colors = f->dome_cols[i].data(); }

  fjgpu_render_params &p = f->params; memset(&p, 0, sizeof p);
  p.xres = r.res[0]; p.yres = r.res[1]; p.xrate = r.rate[0]; p.yrate = r.rate[1]; p.xfwidth = r.fw[0]; p.yfwidth = r.fw[1];
  p.jitter = r.jitter; p.max_diffuse_depth = r.max_diffuse; p.max_reflect_depth = r.max_reflect; p.max_refract_depth = r.max_refract;
  p.cast_shadow = r.cast_shadow; p.target_group = 0;
  const char *seed = getenv("FJ_SEED"); p.seed = seed ? (uint32_t)strtoul(seed, nullptr, 10) : 1u;
  // Tiler::GenerateTiles, src/fj_tiler.cc:56-113
  const int xmin = r.region[0], ymin = r.region[1], xmax = r.region[2], ymax = r.region[3];
  const int X0 = (int)std::floor(std::max(0, xmin) / (double)r.tile[0]), Y0 = (int)std::floor(std::max(0, ymin) / (double)r.tile[1]);
  const int X1 = (int)std::ceil(std::min(r.res[0], xmax) / (double)r.tile[0]), Y1 = (int)std::ceil(std::min(r.res[1], ymax) / (double)r.tile[1]);
  int id = 0;
  for (int y = Y0; y < Y1; y++) for (int x = X0; x < X1; x++) {
    fjgpu_tile t; t.id = id++;
    t.xmin = std::max(x * r.tile[0], xmin); t.ymin = std::max(y * r.tile[1], ymin);
    t.xmax = std::min((x + 1) * r.tile[0], xmax); t.ymax = std::min((y + 1) * r.tile[1], ymax);
    f->tiles.push_back(t);
  }
  return SI_SUCCESS;
}

Status render(Scene &sc, Renderer &r) {
  FrameBuf *fb = get(sc.framebuffers, r.fb, Type_FrameBuffer);
  if (!fb) return failmsg("renderer has no framebuffer");
  Flat f;
  if (flatten(sc, r, &f) != SI_SUCCESS) return SI_FAIL;
  // preprocess_framebuffer, src/fj_renderer.cc:805-815: Resize clears the buffer
  // (the host frame is neither written nor read while the tiles stay on the device — device blocks for the host's all-gather,
  // resident frame of the bench leg: no 33-MB clear per frame on those paths)
  fb->w = r.res[0]; fb->h = r.res[1]; fb->c = 4;
  if (g_dev_blocks || g_resident) fb->px.resize((size_t)fb->w * fb->h * 4); else fb->px.assign((size_t)fb->w * fb->h * 4, 0.f);
  // GPUs of this process: FJ_GPU_COUNT (or fjscene_set_gpu_count) contexts on consecutive devices from g_device on, the scene
  // replicated on each; tiles are dealt to them by fjgpu_render_frame_multi (one all-gather ends the frame)
  int ngpu = g_gpu_count > 0 ? g_gpu_count : 1;
  if (g_gpu_count <= 0) { const char *e = getenv("FJ_GPU_COUNT"); if (e && atoi(e) > 1) ngpu = atoi(e); }
  if (g_world > 1 || g_dev_blocks || g_resident) ngpu = 1;      // (one process per GPU under torchrun: the host does the gather)
  if (sc.gpu && (sc.gpu_device != g_device || (int)sc.more_gpus.size() != ngpu - 1)) {
    for (fjgpu_context *c : sc.more_gpus) fjgpu_destroy(c);
    sc.more_gpus.clear();
    fjgpu_destroy(sc.gpu); sc.gpu = nullptr; sc.motion_key.clear(); for (Mesh &m : sc.meshes) m.dirty = true; sc.textures_dirty = true;
  }
  if (!sc.gpu) {
    if (fjgpu_create(g_device, &sc.gpu) != FJGPU_OK) return failmsg(std::string("fjgpu_create: ") + fjgpu_last_error(nullptr));
    sc.gpu_device = g_device; sc.motion_key.clear();
    for (int k = 1; k < ngpu; k++) {
      fjgpu_context *c = nullptr;
      if (fjgpu_create(g_device + k, &c) != FJGPU_OK) return failmsg(std::string("fjgpu_create (FJ_GPU_COUNT): ") + fjgpu_last_error(nullptr));
      sc.more_gpus.push_back(c);
    }
  }
  std::vector<fjgpu_context *> gpus{sc.gpu};
  gpus.insert(gpus.end(), sc.more_gpus.begin(), sc.more_gpus.end());
  const auto t0 = std::chrono::steady_clock::now();
  printf("# Building Accelerators\n");                                  // src/fj_scene_interface.cc:1172-1201
  for (size_t i = 0; i < sc.meshes.size(); i++) {
    Mesh &m = sc.meshes[i];
    if (!m.dirty) continue;
    for (fjgpu_context *g : gpus) {
      // a mesh VelocityGeneratorProcedure ran on carries its velocities: moving triangles (Mesh::ray_intersect, src/fj_mesh.cc:252-259)
      if (fjgpu_mesh_upload_velocity(g, (int32_t)i, m.P.data(), m.N.empty() ? nullptr : m.N.data(), (int32_t)(m.P.size() / 3), m.idx.data(), nullptr,
                                     (int32_t)(m.idx.size() / 3), m.vel.empty() ? nullptr : m.vel.data()) != FJGPU_OK)
        return failmsg(std::string("fjgpu_mesh_upload: ") + fjgpu_last_error(g));
      if (!m.uv.empty() && fjgpu_mesh_set_uv(g, (int32_t)i, m.uv.data(), (int32_t)(m.uv.size() / 2)) != FJGPU_OK)
        return failmsg(std::string("fjgpu_mesh_set_uv: ") + fjgpu_last_error(g));
    }
    m.dirty = false;
  }
  if (sc.textures_dirty) {
    std::vector<fjgpu_texture> ft(sc.textures.size());
    for (size_t i = 0; i < ft.size(); i++) { const Texture &t = sc.textures[i]; ft[i].width = t.width; ft[i].height = t.height; ft[i].nchannels = t.nch; ft[i].tilesize = t.tilesize; ft[i].tiles = t.tiles.data(); }
    for (fjgpu_context *g : gpus)
      if (fjgpu_textures_set(g, (int32_t)ft.size(), ft.data()) != FJGPU_OK) return failmsg(std::string("fjgpu_textures_set: ") + fjgpu_last_error(g));
    sc.textures_dirty = false;
  }
  int rc = 0;
  for (fjgpu_context *g : gpus) {
    if (!rc) rc = fjgpu_shaders_set(g, (int32_t)f.shaders.size(), f.shaders.data());
    if (!rc) rc = fjgpu_shutter_set(g, r.time_range[0], r.time_range[1]);      // Renderer::SetSampleTimeRange
    if (!rc) rc = fjgpu_groups_set(g, (int32_t)f.goff.size() - 1, f.goff.data(), f.gids.data());
    if (!rc) rc = fjgpu_instances_set(g, (int32_t)f.inst.size(), f.inst.data());
    if (!rc) rc = fjgpu_lights_set(g, (int32_t)f.lights.size(), f.lights.data());
    if (!rc) rc = fjgpu_camera_set(g, &f.cam);
    if (rc) return failmsg(std::string("scene upload: ") + fjgpu_last_error(g));
  }
  sc.flat_inst = f.inst;
  sc.last_tiles = f.tiles; sc.last_res[0] = r.res[0]; sc.last_res[1] = r.res[1];

  std::vector<fjgpu_tile> mine;
  for (size_t i = 0; i < f.tiles.size(); i++) if ((int)(i % (size_t)g_world) == g_rank) mine.push_back(f.tiles[i]);      // (one process per GPU: this rank's share)
  {
    // Motion blur: the frame's time table (one entry per sample of a tile: src/fj_fixed_grid_sampler.cc:42,72-77) and the
    // transform XfmLerpTransformSample would rebuild for a ray of each entry's time, for every time-sampled instance / camera
    const Camera *cam = get(sc.cameras, r.camera, Type_Camera);
    bool moving = !cam->x.is_static();
    for (const Instance &o : sc.instances) moving = moving || !o.x.is_static();
    std::vector<double> times, fwd, inv;
    int n = 0;
    if (moving && !mine.empty()) {
      n = fjgpu_time_table(&f.params, mine.data(), (int32_t)mine.size(), r.time_range[0], r.time_range[1], nullptr, 0);
      if (n <= 0) return failmsg("fjgpu_time_table failed");
      times.resize(n);
      fjgpu_time_table(&f.params, mine.data(), (int32_t)mine.size(), r.time_range[0], r.time_range[1], times.data(), n);
    }
    // a table is rebuilt only when what it was built from changed: the time table (count, range) or the entry's samples
    auto key_of = [&](const Xform &x) {
      std::vector<double> k = {(double)n, r.time_range[0], r.time_range[1], (double)x.torder, (double)x.rorder};
      for (const TimeSamples *ts : {&x.T, &x.R, &x.S}) { k.push_back((double)ts->s.size()); for (auto &e : ts->s) { k.push_back(e.first); k.insert(k.end(), e.second.begin(), e.second.end()); } }
      return k;
    };
    for (int i = -1; i < (int)sc.instances.size() && !rc; i++) {                       // -1: the camera
      const Xform &x = i < 0 ? cam->x : sc.instances[i].x;
      if (x.is_static() || n == 0) {
        for (fjgpu_context *g : gpus) if (!rc) rc = i < 0 ? fjgpu_camera_motion_set(g, 0, nullptr) : fjgpu_instance_motion_set(g, i, 0, nullptr, nullptr);
        sc.motion_key.erase(i);
        continue;
      }
      std::vector<double> key = key_of(x);
      auto it = sc.motion_key.find(i);
      if (it != sc.motion_key.end() && it->second == key) continue;
      fwd.resize((size_t)n * 16); inv.resize((size_t)n * 16);
      for (int k = 0; k < n; k++) { const M4 m = x.matrix_at(times[k]), mi = inverse(m); memcpy(&fwd[16 * (size_t)k], m.e, 128); memcpy(&inv[16 * (size_t)k], mi.e, 128); }
      for (fjgpu_context *g : gpus) if (!rc) rc = i < 0 ? fjgpu_camera_motion_set(g, n, fwd.data()) : fjgpu_instance_motion_set(g, i, n, fwd.data(), inv.data());
      if (!rc) sc.motion_key[i].swap(key);
    }
    if (rc) return failmsg(std::string("motion tables: ") + fjgpu_last_error(sc.gpu));
  }
  g_upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  FrameInfo info; memset(&info, 0, sizeof info);
  info.frame_id = ++g_frame_id; info.worker_count = 1; info.tile_count = (int)mine.size(); info.xres = r.res[0]; info.yres = r.res[1];
  info.frame_region.min[0] = r.region[0]; info.frame_region.min[1] = r.region[1]; info.frame_region.max[0] = r.region[2]; info.frame_region.max[1] = r.region[3];
  info.framebuffer = reinterpret_cast<const FrameBuffer *>(fb);
  if (ngpu > 1) printf("# Rendering Frame\n#   Tile Count: %d\n#   Devices: cuda:%d..%d (tiles dealt round-robin, one all-gather)\n", (int)mine.size(), g_device, g_device + ngpu - 1);
  else printf("# Rendering Frame\n#   Tile Count: %d\n#   Device: cuda:%d (rank %d of %d)\n", (int)mine.size(), g_device, g_rank, g_world);
  if (r.frame_start && r.frame_start(r.frame_data, &info) == CALLBACK_INTERRUPT) {
    if (r.frame_abort) r.frame_abort(r.frame_data, &info);
    return SI_FAIL;
  }
  memset(&g_stats, 0, sizeof g_stats);
  g_resend_bytes = 0;
  if (g_resend) for (fjgpu_context *g : gpus) if (fjgpu_scene_resend(g, &g_resend_bytes) != FJGPU_OK) return failmsg(std::string("fjgpu_scene_resend: ") + fjgpu_last_error(g));
  if (ngpu > 1) {
    std::vector<fjgpu_stats> per(ngpu);
    rc = fjgpu_render_frame_multi(gpus.data(), ngpu, &f.params, mine.data(), (int32_t)mine.size(), fb->px.data(), per.data());
    for (const fjgpu_stats &q : per) {      // the frame's totals: counts add up, device times are the slowest rank's
      g_stats.rays_camera += q.rays_camera; g_stats.rays_shadow += q.rays_shadow; g_stats.rays_diffuse += q.rays_diffuse; g_stats.rays_reflect += q.rays_reflect;
      g_stats.rays_refract += q.rays_refract; g_stats.camera_samples += q.camera_samples; g_stats.rays_hit += q.rays_hit; g_stats.hit_mesh_levels += q.hit_mesh_levels;
      g_stats.node_steps += q.node_steps; g_stats.tri_tests += q.tri_tests; g_stats.kernel_launches += q.kernel_launches; g_stats.trace_launches += q.trace_launches;
      g_stats.ms_trace = std::max(g_stats.ms_trace, q.ms_trace); g_stats.ms_shade = std::max(g_stats.ms_shade, q.ms_shade);
      g_stats.ms_resolve = std::max(g_stats.ms_resolve, q.ms_resolve); g_stats.ms_total = std::max(g_stats.ms_total, q.ms_total);
    }
  } else if (g_dev_blocks) rc = fjgpu_render_tiles_device(sc.gpu, &f.params, mine.data(), (int32_t)mine.size(), g_dev_bw, g_dev_bh, g_dev_blocks, &g_stats);
  else if (g_resident) rc = fjgpu_render_tiles_resident(sc.gpu, &f.params, mine.data(), (int32_t)mine.size(), &g_stats);
  else rc = fjgpu_render_tiles(sc.gpu, &f.params, mine.data(), (int32_t)mine.size(), fb->px.data(), &g_stats);
  if (rc) { if (r.frame_abort) r.frame_abort(r.frame_data, &info); return failmsg(std::string("fjgpu_render_tiles: ") + fjgpu_last_error(sc.gpu)); }
  fjgpu_scene_info_get(sc.gpu, &g_info);
  if (r.tile_done || r.tile_start) {       // tile callbacks are honoured per tile after the device pass (the per-sample hook cannot be)
    for (size_t i = 0; i < mine.size(); i++) {
      TileInfo ti; memset(&ti, 0, sizeof ti);
      ti.frame_id = info.frame_id; ti.region_id = mine[i].id; ti.total_region_count = (int)f.tiles.size();
      ti.tile_region.min[0] = mine[i].xmin; ti.tile_region.min[1] = mine[i].ymin; ti.tile_region.max[0] = mine[i].xmax; ti.tile_region.max[1] = mine[i].ymax;
      ti.framebuffer = info.framebuffer;
      if (r.tile_start) r.tile_start(r.tile_data, &ti);
      if (r.tile_done) r.tile_done(r.tile_data, &ti);
    }
  }
  const unsigned long long rays = g_stats.rays_camera + g_stats.rays_shadow + g_stats.rays_diffuse + g_stats.rays_reflect + g_stats.rays_refract;
  printf("# Frame Done\n#   %.3f ms on device, %llu rays, %.2f Mrays/s\n", g_stats.ms_total, rays, g_stats.ms_total > 0 ? rays / (g_stats.ms_total * 1e3) : 0.);
  if (r.frame_done) r.frame_done(r.frame_data, &info);
  return SI_SUCCESS;
}

}  // namespace

// ================================================================================== the Si* API
int SiGetErrorNo(void) { return si_errno; }

ID SiOpenPlugin(const char *filename) {
  if (!the_scene || !filename) return bad(SI_ERR_PLUGIN_NOT_FOUND);
  std::string base(filename);
  const size_t slash = base.find_last_of("/\\"); if (slash != std::string::npos) base = base.substr(slash + 1);
  for (const char *ext : {".so", ".dll", ".dylib"}) { const size_t n = strlen(ext); if (base.size() > n && base.compare(base.size() - n, n, ext) == 0) base.resize(base.size() - n); }
  int kind = -1;
  if (base == "ConstantShader") kind = FJGPU_SHADER_CONSTANT;
  else if (base == "PlasticShader") kind = FJGPU_SHADER_PLASTIC;
  else if (base == "PathtracingShader") kind = FJGPU_SHADER_PATHTRACING;
  else if (base == "GlassShader") kind = FJGPU_SHADER_GLASS;
  else if (base == "StanfordPlyProcedure") kind = 100;
  else if (base == "VelocityGeneratorProcedure") kind = 101;
  if (kind < 0) { last_message = "plugin '" + base + "' has no device implementation"; return bad(SI_ERR_PLUGIN_NOT_FOUND); }
  the_scene->plugins.push_back({base, kind});
  si_errno = SI_ERR_NONE;
  return encode_id(Type_Plugin, (int)the_scene->plugins.size() - 1);
}

Status SiOpenScene(void) { delete the_scene; the_scene = new Scene(); return ok(); }
Status SiCloseScene(void) { delete the_scene; the_scene = nullptr; return ok(); }

Status SiRenderScene(ID renderer) {
  Renderer *r = the_scene ? get(the_scene->renderers, renderer, Type_Renderer) : nullptr;
  if (!r) return SI_FAIL;
  if (render(*the_scene, *r) != SI_SUCCESS) return SI_FAIL;
  return ok();
}

Status SiSaveFrameBuffer(ID framebuffer, const char *filename) {
  FrameBuf *fb = the_scene ? get(the_scene->framebuffers, framebuffer, Type_FrameBuffer) : nullptr;
  if (!fb || !filename) return SI_FAIL;
  // WriteFrameBuffer, src/fj_framebuffer_io.cc:46-68: PTO text, default ostream precision (6 significant digits)
  std::ofstream strm(filename);
  if (!strm) return SI_FAIL;
  strm << "#PTO Plain Text Object\n#Fujiyama Renderer FrameBuffer\n";
  strm << "resolution " << fb->w << " " << fb->h << '\n' << "channel_count " << fb->c << '\n' << "begin pixels\n";
  const float *p = fb->px.data();
  for (long i = 0; i < (long)fb->w * fb->h; i++, p += 4) strm << p[0] << " " << p[1] << " " << p[2] << " " << p[3] << '\n';
  strm << "end pixels\n";
  return ok();
}

Status SiRunProcedure(ID procedure) {
  Procedure *p = the_scene ? get(the_scene->procedures, procedure, Type_Procedure) : nullptr;
  if (!p) return SI_FAIL;
  Mesh *m = get(the_scene->meshes, p->mesh, Type_Mesh);
  if (!m) return SI_FAIL;                                        // StanfordPlyProcedure::run: no mesh assigned
  if (the_scene->plugins[p->plugin].kind == 101) {               // VelocityGeneratorProcedure::run
    printf("Point Count: %d\n", (int)(m->P.size() / 3));
    return generate_velocity(m) ? SI_FAIL : ok();
  }
  if (p->io_mode == "w") return failmsg("StanfordPlyProcedure io_mode w is outside the device path");
  if (read_ply(p->filepath, m)) return SI_FAIL;
  return ok();
}

Status SiAddObjectToGroup(ID group, ID object) {
  Group *g = the_scene ? get(the_scene->groups, group, Type_ObjectGroup) : nullptr;
  int t, i;
  if (!g || !decode_id(object, &t, &i) || t != Type_ObjectInstance || i >= (int)the_scene->instances.size()) return SI_FAIL;
  g->members.push_back(i);
  return ok();
}

ID SiNewObjectInstance(ID primset) {
  if (!the_scene) return bad(SI_ERR_BADTYPE);
  if (!get(the_scene->meshes, primset, Type_Mesh)) return bad(SI_ERR_BADTYPE);
  Instance o; o.mesh = primset;
  the_scene->instances.push_back(o);
  si_errno = SI_ERR_NONE;
  return encode_id(Type_ObjectInstance, (int)the_scene->instances.size() - 1);
}
ID SiNewFrameBuffer(const char *) { if (!the_scene) return bad(SI_ERR_NO_MEMORY); the_scene->framebuffers.push_back(FrameBuf()); si_errno = SI_ERR_NONE; return encode_id(Type_FrameBuffer, (int)the_scene->framebuffers.size() - 1); }
ID SiNewObjectGroup(void) { if (!the_scene) return bad(SI_ERR_NO_MEMORY); the_scene->groups.push_back(Group()); si_errno = SI_ERR_NONE; return encode_id(Type_ObjectGroup, (int)the_scene->groups.size() - 1); }
ID SiNewPointCloud(void) { return bad(SI_ERR_FAILNEW); }
ID SiNewTurbulence(void) { return bad(SI_ERR_FAILNEW); }
// SiNewTexture, src/fj_scene_interface.cc:529-545 -> Texture::LoadFile -> MipInput::ReadHeader (src/fj_mipmap.cc:124-154)
ID SiNewTexture(const char *filename) {
  if (!the_scene) return bad(SI_ERR_NO_MEMORY);
  Texture t;
  FILE *f = filename ? fopen(filename, "rb") : nullptr;
  if (!f) return bad(SI_ERR_FAILLOAD);
  char magic[4]; int32_t hdr[5];
  bool good = fread(magic, 1, 4, f) == 4 && memcmp(magic, "MIPM", 4) == 0 && fread(hdr, 4, 5, f) == 5 && hdr[0] == 1;
  if (good) {
    t.width = hdr[1]; t.height = hdr[2]; t.nch = hdr[3]; t.tilesize = hdr[4];
    good = t.width > 0 && t.height > 0 && t.tilesize > 0 && (t.nch == 1 || t.nch == 3 || t.nch == 4) && t.width / t.tilesize > 0 && t.height / t.tilesize > 0;
  }
  if (good) {
    const size_t n = (size_t)(t.width / t.tilesize) * (t.height / t.tilesize) * t.tilesize * t.tilesize * t.nch;
    t.tiles.resize(n);
    good = fread(t.tiles.data(), sizeof(float), n, f) == n;
  }
  fclose(f);
  if (!good) return bad(SI_ERR_FAILLOAD);
  the_scene->textures.push_back(std::move(t)); the_scene->textures_dirty = true;
  si_errno = SI_ERR_NONE;
  return encode_id(Type_Texture, (int)the_scene->textures.size() - 1);
}
ID SiNewVolume(void) { return bad(SI_ERR_FAILNEW); }
ID SiNewCurve(void) { return bad(SI_ERR_FAILNEW); }
ID SiNewProcedure(ID plugin) {
  int t, i;
  if (!the_scene || !decode_id(plugin, &t, &i) || t != Type_Plugin || i >= (int)the_scene->plugins.size()) return bad(SI_ERR_BADTYPE);
  if (the_scene->plugins[i].kind != 100 && the_scene->plugins[i].kind != 101) return bad(SI_ERR_FAILNEW);
  Procedure p; p.plugin = i; p.io_mode = "r";
  the_scene->procedures.push_back(p);
  si_errno = SI_ERR_NONE;
  return encode_id(Type_Procedure, (int)the_scene->procedures.size() - 1);
}
ID SiNewRenderer(void) { if (!the_scene) return bad(SI_ERR_NO_MEMORY); the_scene->renderers.push_back(Renderer()); si_errno = SI_ERR_NONE; return encode_id(Type_Renderer, (int)the_scene->renderers.size() - 1); }
ID SiNewCamera(const char *) { if (!the_scene) return bad(SI_ERR_NO_MEMORY); the_scene->cameras.push_back(Camera()); si_errno = SI_ERR_NONE; return encode_id(Type_Camera, (int)the_scene->cameras.size() - 1); }
ID SiNewShader(ID plugin) {
  int t, i;
  if (!the_scene || !decode_id(plugin, &t, &i) || t != Type_Plugin || i >= (int)the_scene->plugins.size()) return bad(SI_ERR_BADTYPE);
  if (the_scene->plugins[i].kind > FJGPU_SHADER_GLASS) return bad(SI_ERR_FAILNEW);
  Shader s; s.plugin = i; s.props = shader_defaults(the_scene->plugins[i].kind);      // PropSetAllDefaultValues
  the_scene->shaders.push_back(s);
  si_errno = SI_ERR_NONE;
  return encode_id(Type_Shader, (int)the_scene->shaders.size() - 1);
}
ID SiNewLight(int light_type) {
  if (!the_scene) return bad(SI_ERR_NO_MEMORY);
  if (light_type < SI_POINT_LIGHT || light_type > SI_DOME_LIGHT) return bad(SI_ERR_FAILNEW);
  Light l; l.type = light_type;
  the_scene->lights.push_back(l);
  si_errno = SI_ERR_NONE;
  return encode_id(Type_Light, (int)the_scene->lights.size() - 1);
}
ID SiNewMesh(void) { if (!the_scene) return bad(SI_ERR_NO_MEMORY); the_scene->meshes.push_back(Mesh()); si_errno = SI_ERR_NONE; return encode_id(Type_Mesh, (int)the_scene->meshes.size() - 1); }

Status SiAssignFrameBuffer(ID renderer, ID framebuffer) {
  Renderer *r = the_scene ? get(the_scene->renderers, renderer, Type_Renderer) : nullptr;
  if (!r || !get(the_scene->framebuffers, framebuffer, Type_FrameBuffer)) return SI_FAIL;
  r->fb = framebuffer; return ok();
}
Status SiAssignCamera(ID renderer, ID camera) {
  Renderer *r = the_scene ? get(the_scene->renderers, renderer, Type_Renderer) : nullptr;
  if (!r || !get(the_scene->cameras, camera, Type_Camera)) return SI_FAIL;
  r->camera = camera; return ok();
}
Status SiAssignObjectGroup(ID id, const char *name, ID group) {
  Instance *o = the_scene ? get(the_scene->instances, id, Type_ObjectInstance) : nullptr;
  if (!o || !name || !get(the_scene->groups, group, Type_ObjectGroup)) return SI_FAIL;
  const std::string n(name);
  if (n == "reflect_target") o->reflect = group; else if (n == "refract_target") o->refract = group; else if (n == "shadow_target") o->shadow = group; else return SI_FAIL;
  return ok();
}
Status SiAssignPointCloud(ID, const char *, ID) { return SI_FAIL; }
Status SiAssignTurbulence(ID, const char *, ID) { return SI_FAIL; }
// SiAssignTexture, src/fj_scene_interface.cc:786-805: the PropTexture properties of the device shaders
// (constant_shader `texture`, plastic_shader / pathtracing_shader `diffuse_map`, plastic_shader `bump_map`)
Status SiAssignTexture(ID id, const char *name, ID texture) {
  if (Light *lt = the_scene ? get(the_scene->lights, id, Type_Light) : nullptr) {       // Light::SetEnvironmentMap, src/fj_light.cc:54-57
    if (!name || std::string(name) != "environment_map" || !get(the_scene->textures, texture, Type_Texture)) return SI_FAIL;
    lt->envmap = texture;
    return ok();
  }
  Shader *sh = the_scene ? get(the_scene->shaders, id, Type_Shader) : nullptr;
  if (!sh || !name || !get(the_scene->textures, texture, Type_Texture)) return SI_FAIL;
  const int kind = the_scene->plugins[sh->plugin].kind;
  const std::string n(name);
  const bool okname = (kind == FJGPU_SHADER_CONSTANT && n == "texture") || ((kind == FJGPU_SHADER_PLASTIC || kind == FJGPU_SHADER_PATHTRACING) && n == "diffuse_map");
  if ((kind == FJGPU_SHADER_PLASTIC || kind == FJGPU_SHADER_PATHTRACING) && n == "bump_map") { sh->bump = texture; return ok(); }      // SlBumpMapping, src/fj_shading.cc:418-465
  if (!okname) return failmsg("AssignTexture " + n + ": no device implementation of this texture property");
  sh->texture = texture;
  return ok();
}
Status SiAssignVolume(ID, const char *, ID) { return SI_FAIL; }
Status SiAssignCurve(ID, const char *, ID) { return SI_FAIL; }
Status SiAssignShader(ID object, const char *shading_group, ID shader) {
  Instance *o = the_scene ? get(the_scene->instances, object, Type_ObjectInstance) : nullptr;
  if (!o || !get(the_scene->shaders, shader, Type_Shader)) return SI_FAIL;
  o->shaders[shading_group ? shading_group : ""] = shader;
  return ok();
}
Status SiAssignMesh(ID id, const char *name, ID mesh) {
  Procedure *p = the_scene ? get(the_scene->procedures, id, Type_Procedure) : nullptr;
  if (!p || !name || std::string(name) != "mesh" || !get(the_scene->meshes, mesh, Type_Mesh)) return SI_FAIL;
  p->mesh = mesh; return ok();
}

Status SiSetProperty1(ID id, const char *name, double v0) { const double v[4] = {v0, 0, 0, 0}; return set_property(id, name, v, 1, 0.); }
Status SiSetProperty2(ID id, const char *name, double v0, double v1) { const double v[4] = {v0, v1, 0, 0}; return set_property(id, name, v, 2, 0.); }
Status SiSetProperty3(ID id, const char *name, double v0, double v1, double v2) { const double v[4] = {v0, v1, v2, 0}; return set_property(id, name, v, 3, 0.); }
Status SiSetProperty4(ID id, const char *name, double v0, double v1, double v2, double v3) { const double v[4] = {v0, v1, v2, v3}; return set_property(id, name, v, 4, 0.); }
Status SiSetSampleProperty3(ID id, const char *name, double v0, double v1, double v2, double time) { const double v[4] = {v0, v1, v2, 0}; return set_property(id, name, v, 3, time); }
Status SiSetStringProperty(ID id, const char *name, const char *string) {
  Procedure *p = the_scene ? get(the_scene->procedures, id, Type_Procedure) : nullptr;
  if (!p || !name || !string) return SI_FAIL;
  const std::string n(name);
  if (n == "filepath") p->filepath = string; else if (n == "io_mode") p->io_mode = string; else return SI_FAIL;
  return ok();
}

Status SiSetFrameReportCallback(ID id, void *data, FrameStartCallback frame_start, FrameAbortCallback frame_abort, FrameDoneCallback frame_done) {
  Renderer *r = the_scene ? get(the_scene->renderers, id, Type_Renderer) : nullptr;
  if (!r) return SI_FAIL;
  r->frame_data = data; r->frame_start = frame_start; r->frame_abort = frame_abort; r->frame_done = frame_done;
  return ok();
}
Status SiSetTileReportCallback(ID id, void *data, TileStartCallback tile_start, SampleDoneCallback, TileDoneCallback tile_done) {
  Renderer *r = the_scene ? get(the_scene->renderers, id, Type_Renderer) : nullptr;
  if (!r) return SI_FAIL;
  r->tile_data = data; r->tile_start = tile_start; r->tile_done = tile_done;
  return ok();
}

}  // namespace fj

// ================================================================================== `.scn` interpreter + C entry points
using namespace fj;

struct fjscene_parser { std::map<std::string, ID> names; int line_no = 0; bool echo = true; std::string file = "<text>"; };

namespace {
enum Arg { A_CMD, A_NEW, A_ID, A_NUM, A_LIGHT, A_PROP, A_GROUP, A_PATH, A_STR };
struct Cmd { const char *name; std::vector<Arg> args; };
// the command table of tools/scene_parser/command.cc:502-543 (argument kinds of its *_args arrays)
const std::vector<Cmd> &commands() {
  static const std::vector<Cmd> c = {
    {"OpenPlugin", {A_CMD, A_NEW, A_PATH}}, {"RenderScene", {A_CMD, A_ID}}, {"RunProcedure", {A_CMD, A_ID}},
    {"SaveFrameBuffer", {A_CMD, A_ID, A_PATH}}, {"AddObjectToGroup", {A_CMD, A_ID, A_ID}},
    {"NewObjectInstance", {A_CMD, A_NEW, A_ID}}, {"NewFrameBuffer", {A_CMD, A_NEW, A_STR}}, {"NewObjectGroup", {A_CMD, A_NEW}},
    {"NewPointCloud", {A_CMD, A_NEW}}, {"NewTurbulence", {A_CMD, A_NEW}}, {"NewProcedure", {A_CMD, A_NEW, A_ID}},
    {"NewRenderer", {A_CMD, A_NEW}}, {"NewTexture", {A_CMD, A_NEW, A_PATH}}, {"NewShader", {A_CMD, A_NEW, A_ID}},
    {"NewCamera", {A_CMD, A_NEW, A_STR}}, {"NewVolume", {A_CMD, A_NEW}}, {"NewCurve", {A_CMD, A_NEW}},
    {"NewLight", {A_CMD, A_NEW, A_LIGHT}}, {"NewMesh", {A_CMD, A_NEW}},
    {"AssignFrameBuffer", {A_CMD, A_ID, A_ID}}, {"AssignObjectGroup", {A_CMD, A_ID, A_PROP, A_ID}},
    {"AssignPointCloud", {A_CMD, A_ID, A_PROP, A_ID}}, {"AssignTurbulence", {A_CMD, A_ID, A_PROP, A_ID}},
    {"AssignTexture", {A_CMD, A_ID, A_PROP, A_ID}}, {"AssignCamera", {A_CMD, A_ID, A_ID}},
    {"AssignShader", {A_CMD, A_ID, A_GROUP, A_ID}}, {"AssignVolume", {A_CMD, A_ID, A_PROP, A_ID}},
    {"AssignCurve", {A_CMD, A_ID, A_PROP, A_ID}}, {"AssignMesh", {A_CMD, A_ID, A_PROP, A_ID}},
    {"SetProperty1", {A_CMD, A_ID, A_PROP, A_NUM}}, {"SetProperty2", {A_CMD, A_ID, A_PROP, A_NUM, A_NUM}},
    {"SetProperty3", {A_CMD, A_ID, A_PROP, A_NUM, A_NUM, A_NUM}}, {"SetProperty4", {A_CMD, A_ID, A_PROP, A_NUM, A_NUM, A_NUM, A_NUM}},
    {"SetStringProperty", {A_CMD, A_ID, A_PROP, A_STR}}, {"SetSampleProperty3", {A_CMD, A_ID, A_PROP, A_NUM, A_NUM, A_NUM, A_NUM}},
    {"ShowPropertyList", {A_CMD, A_STR}},
  };
  return c;
}
bool symbol_number(const std::string &s, double *out) {
  static const char *names[] = {"ORDER_SRT", "ORDER_STR", "ORDER_RST", "ORDER_RTS", "ORDER_TRS", "ORDER_TSR",
                                "ORDER_XYZ", "ORDER_XZY", "ORDER_YXZ", "ORDER_YZX", "ORDER_ZXY", "ORDER_ZYX"};
  for (int i = 0; i < 12; i++) if (s == names[i]) { *out = i; return true; }
  return false;
}
int perr(fjscene_parser *p, const char *msg) {
  last_message = std::string(msg);
  fprintf(stderr, "error: %s:%d: %s\n", p->file.c_str(), p->line_no, msg);
  return -1;
}
const char *si_error_text(int e) {
  switch (e) {
    case SI_ERR_PLUGIN_NOT_FOUND: return "plugin not found";
    case SI_ERR_BADTYPE: return "invalid entry type";
    case SI_ERR_FAILLOAD: return "load file failed";
    case SI_ERR_FAILNEW: return "new entry failed";
    case SI_ERR_NO_MEMORY: return "no memory";
    default: return "command failed";
  }
}
}  // namespace

extern "C" {

fjscene_parser *fjscene_parser_new(void) { SiOpenScene(); return new fjscene_parser(); }
void fjscene_parser_free(fjscene_parser *p) { if (!p) return; SiCloseScene(); delete p; }
void fjscene_set_echo(fjscene_parser *p, int echo) { if (p) p->echo = echo != 0; }
long fjscene_lookup(fjscene_parser *p, const char *name) { auto it = p->names.find(name); return it == p->names.end() ? -1 : it->second; }

int fjscene_parse_line(fjscene_parser *p, const char *line) {
  p->line_no++;
  std::istringstream iss(line);
  std::vector<std::string> tok; std::string s;
  while (iss >> s) tok.push_back(s);
  if (tok.empty() || tok[0][0] == '#') return 0;
  const Cmd *cmd = nullptr;
  for (const Cmd &c : commands()) if (tok[0] == c.name) cmd = &c;
  if (!cmd) return perr(p, "unknown command");
  if (tok.size() < cmd->args.size()) return perr(p, "too few arguments");
  if (tok.size() > cmd->args.size()) return perr(p, "too many arguments");
  std::vector<ID> ids(tok.size(), SI_BADID); std::vector<double> num(tok.size(), 0.);
  for (size_t i = 1; i < tok.size(); i++) {
    switch (cmd->args[i]) {
      case A_NEW: if (p->names.count(tok[i])) return perr(p, "entry name already exists"); break;
      case A_ID: { auto it = p->names.find(tok[i]); if (it == p->names.end()) return perr(p, "entry name not found"); ids[i] = it->second; break; }
      case A_NUM: { if (symbol_number(tok[i], &num[i])) break; char *end = nullptr; num[i] = strtod(tok[i].c_str(), &end); if (*end != '\0') return perr(p, "bad number arguments"); break; }
      case A_LIGHT: if (tok[i] == "PointLight") num[i] = SI_POINT_LIGHT; else if (tok[i] == "GridLight") num[i] = SI_GRID_LIGHT;
                    else if (tok[i] == "SphereLight") num[i] = SI_SPHERE_LIGHT; else if (tok[i] == "DomeLight") num[i] = SI_DOME_LIGHT; else return perr(p, "bad enum arguments"); break;
      case A_GROUP: if (tok[i] == "DEFAULT_SHADING_GROUP") tok[i] = ""; break;
      default: break;
    }
  }
  if (p->echo) {      // print_command, parser.cc:274-285
    printf("-- %s: ", tok[0].c_str());
    for (size_t i = 1; i < tok.size(); i++) printf("[%s]%s", tok[i].c_str(), i + 1 == tok.size() ? "\n" : " ");
    if (tok.size() == 1) printf("\n");
  }
  const std::string &c = tok[0];
  ID nid = SI_BADID; Status st = SI_SUCCESS; bool makes = false;
  si_errno = SI_ERR_UNDEFINED;      // failure paths that leave the error number untouched report a generic failure
  if (c == "OpenPlugin") { nid = SiOpenPlugin(tok[2].c_str()); makes = true; }
  else if (c == "RenderScene") st = SiRenderScene(ids[1]);
  else if (c == "RunProcedure") st = SiRunProcedure(ids[1]);
  else if (c == "SaveFrameBuffer") st = SiSaveFrameBuffer(ids[1], tok[2].c_str());
  else if (c == "AddObjectToGroup") st = SiAddObjectToGroup(ids[1], ids[2]);
  else if (c == "NewObjectInstance") { nid = SiNewObjectInstance(ids[2]); makes = true; }
  else if (c == "NewFrameBuffer") { nid = SiNewFrameBuffer(tok[2].c_str()); makes = true; }
  else if (c == "NewObjectGroup") { nid = SiNewObjectGroup(); makes = true; }
  else if (c == "NewPointCloud") { nid = SiNewPointCloud(); makes = true; }
  else if (c == "NewTurbulence") { nid = SiNewTurbulence(); makes = true; }
  else if (c == "NewProcedure") { nid = SiNewProcedure(ids[2]); makes = true; }
  else if (c == "NewRenderer") { nid = SiNewRenderer(); makes = true; }
  else if (c == "NewTexture") { nid = SiNewTexture(tok[2].c_str()); makes = true; }
  else if (c == "NewShader") { nid = SiNewShader(ids[2]); makes = true; }
  else if (c == "NewCamera") { nid = SiNewCamera(tok[2].c_str()); makes = true; }
  else if (c == "NewVolume") { nid = SiNewVolume(); makes = true; }
  else if (c == "NewCurve") { nid = SiNewCurve(); makes = true; }
  else if (c == "NewLight") { nid = SiNewLight((int)num[2]); makes = true; }
  else if (c == "NewMesh") { nid = SiNewMesh(); makes = true; }
  else if (c == "AssignFrameBuffer") st = SiAssignFrameBuffer(ids[1], ids[2]);
  else if (c == "AssignObjectGroup") st = SiAssignObjectGroup(ids[1], tok[2].c_str(), ids[3]);
  else if (c == "AssignPointCloud") st = SiAssignPointCloud(ids[1], tok[2].c_str(), ids[3]);
  else if (c == "AssignTurbulence") st = SiAssignTurbulence(ids[1], tok[2].c_str(), ids[3]);
  else if (c == "AssignTexture") st = SiAssignTexture(ids[1], tok[2].c_str(), ids[3]);
  else if (c == "AssignCamera") st = SiAssignCamera(ids[1], ids[2]);
  else if (c == "AssignShader") st = SiAssignShader(ids[1], tok[2].c_str(), ids[3]);
  else if (c == "AssignVolume") st = SiAssignVolume(ids[1], tok[2].c_str(), ids[3]);
  else if (c == "AssignCurve") st = SiAssignCurve(ids[1], tok[2].c_str(), ids[3]);
  else if (c == "AssignMesh") st = SiAssignMesh(ids[1], tok[2].c_str(), ids[3]);
  else if (c == "SetProperty1") st = SiSetProperty1(ids[1], tok[2].c_str(), num[3]);
  else if (c == "SetProperty2") st = SiSetProperty2(ids[1], tok[2].c_str(), num[3], num[4]);
  else if (c == "SetProperty3") st = SiSetProperty3(ids[1], tok[2].c_str(), num[3], num[4], num[5]);
  else if (c == "SetProperty4") st = SiSetProperty4(ids[1], tok[2].c_str(), num[3], num[4], num[5], num[6]);
  else if (c == "SetStringProperty") st = SiSetStringProperty(ids[1], tok[2].c_str(), tok[3].c_str());
  else if (c == "SetSampleProperty3") st = SiSetSampleProperty3(ids[1], tok[2].c_str(), num[3], num[4], num[5], num[6]);
  else if (c == "ShowPropertyList") { printf("# ShowPropertyList is not mirrored\n"); }
  if (makes) {
    if (nid == SI_BADID) return perr(p, si_error_text(SiGetErrorNo()));
    p->names[tok[1]] = nid;
  } else if (st != SI_SUCCESS) return perr(p, si_error_text(SiGetErrorNo()));
  return 0;
}

int fjscene_parse_text(fjscene_parser *p, const char *text) {
  std::istringstream in(text ? text : "");
  std::string line;
  while (std::getline(in, line)) if (fjscene_parse_line(p, line.c_str())) return -1;
  return 0;
}

int fjscene_parse_file(fjscene_parser *p, const char *path) {
  std::ifstream f(path);
  if (!f) { fprintf(stderr, "error: couldn't open %s\n", path); return -1; }
  p->file = path; p->line_no = 0;
  std::string line;
  while (std::getline(f, line)) if (fjscene_parse_line(p, line.c_str())) return -1;
  return 0;
}

int fjscene_mesh_set(long mesh_id, const double *P, int32_t nverts, const int32_t *idx3, int32_t nfaces) {
  Mesh *m = the_scene ? get(the_scene->meshes, mesh_id, Type_Mesh) : nullptr;
  if (!m || nverts < 0 || nfaces < 0 || (nverts && !P) || (nfaces && !idx3)) return -1;
  for (long i = 0; i < 3l * nfaces; i++) if (idx3[i] < 0 || idx3[i] >= nverts) return -1;
  m->P.assign(P, P + 3 * (size_t)nverts); m->idx.assign(idx3, idx3 + 3 * (size_t)nfaces);
  compute_normals(*m); m->dirty = true;
  return 0;
}

const float *fjscene_framebuffer(long framebuffer_id, int32_t *width, int32_t *height, int32_t *channels) {
  FrameBuf *fb = the_scene ? get(the_scene->framebuffers, framebuffer_id, Type_FrameBuffer) : nullptr;
  if (!fb) return nullptr;
  if (width) *width = fb->w; if (height) *height = fb->h; if (channels) *channels = fb->c;
  return fb->px.data();
}

int fjscene_last_stats(fjgpu_stats *stats, fjgpu_scene_info *info, double *upload_seconds) {
  if (stats) *stats = g_stats; if (info) *info = g_info; if (upload_seconds) *upload_seconds = g_upload_seconds;
  return 0;
}

void fjscene_set_device(int device_ordinal, int rank, int world_size) { g_device = device_ordinal; g_rank = rank; g_world = world_size > 0 ? world_size : 1; }
void fjscene_set_resident(int resident) { g_resident = resident; }
void fjscene_set_resend(int resend) { g_resend = resend; }
void fjscene_set_gpu_count(int count) { g_gpu_count = count; }
int fjscene_assemble_gathered(const void *d_gathered_blocks, int nranks, int tile_w_max, int tile_h_max, float *rgba_frame) {
  if (!the_scene || !the_scene->gpu || the_scene->last_tiles.empty()) return -1;
  Scene &sc = *the_scene;
  if (fjgpu_assemble_frame(sc.gpu, d_gathered_blocks, nranks, tile_w_max, tile_h_max, sc.last_tiles.data(), (int32_t)sc.last_tiles.size(),
                           sc.last_res[0], sc.last_res[1], rgba_frame) != FJGPU_OK) { failmsg(std::string("fjgpu_assemble_frame: ") + fjgpu_last_error(sc.gpu)); return -1; }
  return 0;
}
void fjscene_set_device_blocks(void *d_tile_blocks, int tile_w_max, int tile_h_max) { g_dev_blocks = d_tile_blocks; g_dev_bw = tile_w_max; g_dev_bh = tile_h_max; }
uint64_t fjscene_last_resend_bytes(void) { return g_resend_bytes; }

int fjscene_instance_matrices(int32_t index, double *fwd16, double *inv16) {
  if (!the_scene || index < 0 || index >= (int)the_scene->flat_inst.size()) return -1;
  memcpy(fwd16, the_scene->flat_inst[index].fwd, 128); memcpy(inv16, the_scene->flat_inst[index].inv, 128);
  return 0;
}
int fjscene_mesh_normals(long mesh_id, double *N_out, int32_t nverts) {
  Mesh *m = the_scene ? get(the_scene->meshes, mesh_id, Type_Mesh) : nullptr;
  if (!m || (size_t)nverts * 3 != m->N.size()) return -1;
  memcpy(N_out, m->N.data(), m->N.size() * 8);
  return 0;
}
int fjscene_mesh_velocity(long mesh_id, double *vel_out, int32_t nverts) {
  Mesh *m = the_scene ? get(the_scene->meshes, mesh_id, Type_Mesh) : nullptr;
  if (!m || m->vel.empty() || (size_t)nverts * 3 != m->vel.size()) return -1;
  memcpy(vel_out, m->vel.data(), m->vel.size() * 8);
  return 0;
}
const char *fjscene_last_message(void) { return last_message.c_str(); }
// The flat scene description SiRenderScene would hand to libfjgpu for renderer `renderer_id`, without touching a device
// (host-logic tests compare it with an independent flattening of the same scene).  Pointers inside the returned
// structs (dome sample tables) stay valid until the next fjscene_flatten call.
static Flat g_flat;
int fjscene_flatten(long renderer_id, int32_t *ninst, int32_t *nlights, int32_t *nshaders, int32_t *ntiles) {
  Renderer *r = the_scene ? get(the_scene->renderers, renderer_id, Type_Renderer) : nullptr;
  if (!r) return -1;
  g_flat = Flat();
  if (flatten(*the_scene, *r, &g_flat) != SI_SUCCESS) return -1;
  if (ninst) *ninst = (int32_t)g_flat.inst.size();
  if (nlights) *nlights = (int32_t)g_flat.lights.size();
  if (nshaders) *nshaders = (int32_t)g_flat.shaders.size();
  if (ntiles) *ntiles = (int32_t)g_flat.tiles.size();
  return 0;
}
int fjscene_flat_instance(int32_t i, fjgpu_instance *out) { if (i < 0 || i >= (int)g_flat.inst.size() || !out) return -1; *out = g_flat.inst[i]; return 0; }
int fjscene_flat_light(int32_t i, fjgpu_light *out) { if (i < 0 || i >= (int)g_flat.lights.size() || !out) return -1; *out = g_flat.lights[i]; return 0; }
int fjscene_flat_shader(int32_t i, fjgpu_shader *out) { if (i < 0 || i >= (int)g_flat.shaders.size() || !out) return -1; *out = g_flat.shaders[i]; return 0; }
int fjscene_flat_tile(int32_t i, fjgpu_tile *out) { if (i < 0 || i >= (int)g_flat.tiles.size() || !out) return -1; *out = g_flat.tiles[i]; return 0; }
int fjscene_flat_frame(fjgpu_camera *cam, fjgpu_render_params *params) { if (cam) *cam = g_flat.cam; if (params) *params = g_flat.params; return 0; }

int fjscene_lerp_transform(long id, double time, double *fwd16, double *inv16) {
  const Xform *x = nullptr;
  if (the_scene) {
    if (const Instance *o = get(the_scene->instances, id, Type_ObjectInstance)) x = &o->x;
    else if (const Camera *c = get(the_scene->cameras, id, Type_Camera)) x = &c->x;
    else if (const Light *l = get(the_scene->lights, id, Type_Light)) x = &l->x;
  }
  if (!x) return -1;
  const M4 m = x->matrix_at(time), mi = inverse(m);
  memcpy(fwd16, m.e, 128); memcpy(inv16, mi.e, 128);
  return 0;
}
void fjscene_make_transform(int transform_order, int rotate_order, const double *T, const double *R, const double *S, double *fwd16, double *inv16) {
  const M4 m = compose(transform_order, rotate_order, T, R, S), inv = inverse(m);
  memcpy(fwd16, m.e, 128); memcpy(inv16, inv.e, 128);
}

}  // extern "C"
