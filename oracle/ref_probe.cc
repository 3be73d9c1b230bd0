// ref_probe.cc — TEST INFRASTRUCTURE.  Links against the UNMODIFIED reference libscene.so
// (built by oracle/Makefile.ref from /root/reference) and
//   ref_probe vectors            dumps per-function golden vectors as JSON on stdout
//                                (committed as tests/golden/ref_vectors.json by oracle/gen_golden.py)
//   ref_probe run <file.scn>     executes a scene-description file exactly like bin/scene
//                                (tools/scene_parser/main.cc:9-52) but times every frame with
//                                std::chrono through SiSetFrameReportCallback
//                                (src/fj_scene_interface.h:119-122) and prints
//                                "FJ_FRAME_SECONDS <s>" / "FJ_RENDER_CALL_SECONDS <s>".
// It includes the reference headers from where they lie; no reference source is copied.

#include "fj_scene_interface.h"
#include "fj_fixed_grid_sampler.h"
#include "fj_rectangle.h"
#include "fj_triangle.h"
#include "fj_transform.h"
#include "fj_shading.h"
#include "fj_camera.h"
#include "fj_filter.h"
#include "fj_random.h"
#include "fj_matrix.h"
#include "fj_vector.h"
#include "fj_mesh.h"
#include "fj_box.h"
#include "fj_ray.h"
#include "fj_noise.h"
#include "fj_numeric.h"
#include "fj_transform.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

using namespace fj;

// ------------------------------------------------------------------ tiny deterministic LCG for inputs
static uint64_t lcg_state = 0x243F6A8885A308D3ull;
static double urand() { lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull; return (double)(lcg_state >> 11) / 9007199254740992.0; }
static double srand11() { return 2 * urand() - 1; }
static void pv(const char *name, const Vector &v, bool comma = true) { printf("\"%s\":[%.17g,%.17g,%.17g]%s", name, v.x, v.y, v.z, comma ? "," : ""); }

static void dump_vectors()
{
  printf("{\n");
  // XorShift (src/fj_random.cc:10-43)
  { XorShift r; printf("\"xorshift_u32\":["); for (int i = 0; i < 64; i++) printf("%u%s", r.NextInteger(), i < 63 ? "," : ""); printf("],\n");
    XorShift q; printf("\"xorshift_f01\":["); for (int i = 0; i < 16; i++) printf("%.17g%s", q.NextFloat01(), i < 15 ? "," : ""); printf("],\n"); }
  // TriRayIntersect (src/fj_triangle.cc:81-153)
  printf("\"tri\":[\n");
  const int NTRI = 400;
  for (int i = 0; i < NTRI; i++) {
    const double s = (i % 4 == 0) ? 1e-3 : 1.0;   // small triangles exercise the EPSILON cut
    Vector v0(srand11(), srand11(), srand11()), v1 = v0 + s * Vector(srand11(), srand11(), srand11()), v2 = v0 + s * Vector(srand11(), srand11(), srand11());
    Vector o(3 * srand11(), 3 * srand11(), 3 * srand11());
    const double a = urand(), b = urand() * (1 - a);
    Vector tgt = (1 - a - b) * v0 + a * v1 + b * v2;
    if (i % 3 == 0) tgt = tgt + s * Vector(srand11(), srand11(), srand11());
    Vector d = tgt - o; if (i % 2) d = Normalize(d);
    if (i % 17 == 0) { d = v1 - v0; }             // ray in the plane
    double t = 0, u = 0, v = 0;
    const bool hit = TriRayIntersect(v0, v1, v2, o, d, DO_NOT_CULL_BACKFACES, &t, &u, &v);
    printf("{"); pv("v0", v0); pv("v1", v1); pv("v2", v2); pv("o", o); pv("d", d);
    printf("\"hit\":%d,\"t\":%.17g,\"u\":%.17g,\"v\":%.17g}%s\n", hit ? 1 : 0, hit ? t : 0., hit ? u : 0., hit ? v : 0., i < NTRI - 1 ? "," : "");
  }
  printf("],\n");
  // BoxRayIntersect (src/fj_box.cc:73-138) incl. the cases of tests/box_test.cc:14-110
  printf("\"box\":[\n");
  const int NBOX = 200;
  for (int i = 0; i < NBOX; i++) {
    Box b(Vector(srand11(), srand11(), srand11()), Vector(srand11(), srand11(), srand11()));
    Vector o(2 * srand11(), 2 * srand11(), 2 * srand11()), d(srand11(), srand11(), srand11());
    double tmin = .001, tmax = 1000;
    if (i == 0) { b = Box(Vector(-1, -1, -1), Vector(1, 1, 1)); o = Vector(0, 0, 0); d = Vector(0, 0, 1); }           // box_test.cc: from inside
    if (i == 1) { b = Box(Vector(-1, -1, -1), Vector(1, 1, 1)); o = Vector(0, 0, -2); d = Vector(0, 0, 1); }          // from outside
    if (i == 2) { b = Box(Vector(-1, -1, -1), Vector(1, 1, 1)); o = Vector(0, 0, -2); d = Vector(0, 0, 1); tmax = 2; } // tmax clipping
    if (i == 3) { b = Box(Vector(-1, -1, -1), Vector(1, 1, 1)); o = Vector(0, 0, -2); d = Vector(0, 1, 0); }          // miss
    if (i == 4) { b.ReverseInfinite(); o = Vector(0, 0, 0); d = Vector(0, 0, 1); }                                   // ReverseInfinite never hits
    if (i == 5) { d.x = 0; }
    if (i % 5 == 4) { d = Normalize(b.Centroid() - o); }
    double t0 = 0, t1 = 0;
    const bool hit = BoxRayIntersect(b, o, d, tmin, tmax, &t0, &t1);
    printf("{"); pv("min", b.min); pv("max", b.max); pv("o", o); pv("d", d);
    printf("\"tmin\":%.17g,\"tmax\":%.17g,\"hit\":%d,\"t0\":%.17g,\"t1\":%.17g}%s\n", tmin, tmax, hit ? 1 : 0, hit ? t0 : 0., hit ? t1 : 0., i < NBOX - 1 ? "," : "");
  }
  printf("],\n");
  // make_transform_matrix + MatInverse (src/fj_transform.cc:335-391, src/fj_matrix.cc:119-207)
  printf("\"xfm\":[\n");
  const int NX = 40;
  for (int i = 0; i < NX; i++) {
    const int to = i % 6, ro = 6 + (i / 2) % 6;
    double T[3] = {3 * srand11(), 3 * srand11(), 3 * srand11()}, R[3] = {180 * srand11(), 180 * srand11(), 180 * srand11()}, S[3] = {.1 + 2 * urand(), .1 + 2 * urand(), .1 + 2 * urand()};
    if (i == 0) { T[0] = T[1] = T[2] = 3; R[0] = -35.264389682754654; R[1] = 45; R[2] = 0; S[0] = S[1] = S[2] = 1; }
    Transform x; XfmSetTransform(&x, i == 0 ? ORDER_SRT : to, i == 0 ? ORDER_ZXY : ro, T[0], T[1], T[2], R[0], R[1], R[2], S[0], S[1], S[2]);
    printf("{\"torder\":%d,\"rorder\":%d,\"T\":[%.17g,%.17g,%.17g],\"R\":[%.17g,%.17g,%.17g],\"S\":[%.17g,%.17g,%.17g],\"fwd\":[",
           i == 0 ? 0 : to, i == 0 ? 10 : ro, T[0], T[1], T[2], R[0], R[1], R[2], S[0], S[1], S[2]);
    for (int k = 0; k < 16; k++) printf("%.17g%s", x.matrix.e[k], k < 15 ? "," : "");
    printf("],\"inv\":[");
    for (int k = 0; k < 16; k++) printf("%.17g%s", x.inverse.e[k], k < 15 ? "," : "");
    printf("]}%s\n", i < NX - 1 ? "," : "");
  }
  printf("],\n");
  // Camera::GetRay (src/fj_camera.cc:79-110)
  printf("\"camera\":[\n");
  const int NC = 24;
  for (int i = 0; i < NC; i++) {
    Camera cam; const double fov = (i < 12) ? 30. : 20 + 60 * urand();
    const int xres = (i % 2) ? 1920 : 256, yres = (i % 2) ? 1080 : 256;
    double T[3] = {3, 3, 3}, R[3] = {-35.264389682754654, 45, 0};
    if (i >= 6) { T[0] = 5 * srand11(); T[1] = 5 * srand11(); T[2] = 5 * srand11(); R[0] = 90 * srand11(); R[1] = 180 * srand11(); R[2] = 30 * srand11(); }
    cam.SetTranslate(T[0], T[1], T[2], 0); cam.SetRotate(R[0], R[1], R[2], 0); cam.SetFov(fov); cam.SetAspect(xres / (double)yres);
    const Vector2 uv(urand(), urand()); Ray ray; cam.GetRay(uv, 0.37, &ray);
    printf("{\"T\":[%.17g,%.17g,%.17g],\"R\":[%.17g,%.17g,%.17g],\"fov\":%.17g,\"xres\":%d,\"yres\":%d,\"u\":%.17g,\"v\":%.17g,",
           T[0], T[1], T[2], R[0], R[1], R[2], fov, xres, yres, uv[0], uv[1]);
    pv("o", ray.orig); pv("d", ray.dir); printf("\"tmin\":%.17g,\"tmax\":%.17g}%s\n", ray.tmin, ray.tmax, i < NC - 1 ? "," : "");
  }
  printf("],\n");
  // FixedGridSampler (src/fj_fixed_grid_sampler.cc:33-84)
  printf("\"sampler\":[\n");
  struct SC { int xres, yres, xr, yr; double fw, jit; int x0, y0, x1, y1; } sc[] = {
    {256, 256, 1, 1, 2, 1, 0, 0, 32, 32}, {256, 256, 1, 1, 2, 1, 224, 224, 256, 256}, {1280, 720, 4, 4, 2, 1, 1248, 704, 1280, 720},
    {1920, 1080, 8, 8, 2, 1, 64, 1056, 96, 1080}, {320, 240, 3, 3, 2, 1, 32, 64, 64, 96}, {320, 240, 3, 2, 3, .5, 0, 0, 32, 32}, {64, 48, 2, 2, 1, 0, 32, 32, 64, 48} };
  const int NS = sizeof(sc) / sizeof(sc[0]);
  for (int i = 0; i < NS; i++) {
    FixedGridSampler s; s.SetResolution(Int2(sc[i].xres, sc[i].yres)); s.SetPixelSamples(Int2(sc[i].xr, sc[i].yr));
    s.SetFilterWidth(Vector2(sc[i].fw, sc[i].fw)); s.SetJitter(sc[i].jit); s.SetSampleTimeRange(0, 1);
    Rectangle r; r.min = Int2(sc[i].x0, sc[i].y0); r.max = Int2(sc[i].x1, sc[i].y1);
    s.GenerateSamples(r);
    std::vector<Sample> all; Sample *p; while ((p = s.GetNextSample()) != NULL) all.push_back(*p);
    printf("{\"xres\":%d,\"yres\":%d,\"xrate\":%d,\"yrate\":%d,\"fw\":%.17g,\"jitter\":%.17g,\"tile\":[%d,%d,%d,%d],\"count\":%zu,\"idx\":[",
           sc[i].xres, sc[i].yres, sc[i].xr, sc[i].yr, sc[i].fw, sc[i].jit, sc[i].x0, sc[i].y0, sc[i].x1, sc[i].y1, all.size());
    std::vector<size_t> idx; for (size_t k = 0; k < 8 && k < all.size(); k++) idx.push_back(k);
    for (int k = 0; k < 8; k++) idx.push_back((size_t)(urand() * all.size())); idx.push_back(all.size() - 1);
    for (size_t k = 0; k < idx.size(); k++) printf("%zu%s", idx[k], k + 1 < idx.size() ? "," : "");
    printf("],\"uv\":[");
    for (size_t k = 0; k < idx.size(); k++) printf("[%.17g,%.17g]%s", all[idx[k]].uv[0], all[idx[k]].uv[1], k + 1 < idx.size() ? "," : "");
    printf("]}%s\n", i < NS - 1 ? "," : "");
  }
  printf("],\n");
  // Gaussian filter (src/fj_filter.cc:49-58)
  printf("\"filter\":[\n");
  for (int i = 0; i < 32; i++) {
    Filter f; const double w = (i % 3 == 0) ? 2 : 1 + 3 * urand(); f.SetFilterType(FLT_GAUSSIAN, w, w);
    const double x = 2 * srand11(), y = 2 * srand11();
    printf("{\"w\":%.17g,\"x\":%.17g,\"y\":%.17g,\"wgt\":%.17g}%s\n", w, x, y, f.Evaluate(x, y), i < 31 ? "," : "");
  }
  printf("],\n");
  // SlFresnel / SlReflect / SlRefract / SlFaceforward (src/fj_shading.cc:42-138)
  printf("\"optics\":[\n");
  for (int i = 0; i < 64; i++) {
    Vector I = Normalize(Vector(srand11(), srand11(), srand11())), N = Normalize(Vector(srand11(), srand11(), srand11()));
    const double ior = (i % 2) ? 1 / 1.4 : .3 + 2 * urand();
    Vector R, T, Nf; SlReflect(&I, &N, &R); SlRefract(&I, &N, ior, &T); SlFaceforward(&I, &N, &Nf);
    printf("{"); pv("I", I); pv("N", N); pv("R", R); pv("T", T); pv("Nf", Nf);
    printf("\"ior\":%.17g,\"F\":%.17g}%s\n", ior, SlFresnel(&I, &N, ior), i < 63 ? "," : "");
  }
  printf("],\n");
  // Mesh::ComputeNormals + ComputeBounds (src/fj_mesh.cc:195-244) on a small bumpy grid
  {
    const int n = 6; Mesh m; m.SetPointCount((n + 1) * (n + 1)); m.AddPointPosition();
    std::vector<Vector> P;
    for (int j = 0; j <= n; j++) for (int i = 0; i <= n; i++) { const Vector p((float)(i / (double)n), (float)(.2 * srand11()), (float)(j / (double)n)); P.push_back(p); m.SetPointPosition(j * (n + 1) + i, p); }
    m.SetFaceCount(2 * n * n); m.AddFaceIndices(); std::vector<int> idx;
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
      const int a = j * (n + 1) + i, b = a + 1, c = a + n + 1, d = c + 1;
      Index3 f0; f0.i0 = a; f0.i1 = c; f0.i2 = b; Index3 f1; f1.i0 = b; f1.i1 = c; f1.i2 = d;
      m.SetFaceIndices(2 * (j * n + i), f0); m.SetFaceIndices(2 * (j * n + i) + 1, f1);
      idx.push_back(a); idx.push_back(c); idx.push_back(b); idx.push_back(b); idx.push_back(c); idx.push_back(d);
    }
    m.ComputeNormals(); m.ComputeBounds();
    printf("\"normals\":{\"P\":[");
    for (size_t k = 0; k < P.size(); k++) printf("[%.17g,%.17g,%.17g]%s", P[k].x, P[k].y, P[k].z, k + 1 < P.size() ? "," : "");
    printf("],\"idx\":[");
    for (size_t k = 0; k < idx.size(); k++) printf("%d%s", idx[k], k + 1 < idx.size() ? "," : "");
    printf("],\"N\":[");
    for (size_t k = 0; k < P.size(); k++) { const Vector N = m.GetPointNormal(k); printf("[%.17g,%.17g,%.17g]%s", N.x, N.y, N.z, k + 1 < P.size() ? "," : ""); }
    printf("]},\n");
  }
  // PerlinNoise3d / SmoothStep (src/fj_noise.cc:33-128, src/fj_numeric.h:66-77): what VelocityGeneratorProcedure evaluates
  printf("\"perlin\":[\n");
  for (int i = 0; i < 48; i++) {
    const double sc = i < 40 ? 3. : 300.;
    const Vector p(sc * srand11(), sc * srand11(), sc * srand11());
    const Vector n = PerlinNoise3d(p, 2, .5, 1 + i % 3);
    printf("{"); pv("p", p); printf("\"octaves\":%d,", 1 + i % 3); pv("n", n);
    const double x = 1.5 * urand() - .25; printf("\"x\":%.17g,\"smooth\":%.17g}%s\n", x, SmoothStep(.2, .7, x), i < 47 ? "," : "");
  }
  printf("],\n");
  // XfmLerpTransformSample (src/fj_transform.cc:306-322) over pushed sample lists (PropPushSample keeps them sorted by time
  // on top of the initial time-0 sample, src/fj_property.cc:284-312)
  printf("\"lerp\":[\n");
  for (int i = 0; i < 24; i++) {
    TransformSampleList list; XfmInitTransformSampleList(&list);
    const int nt = 1 + i % 3, nr = 1 + (i / 3) % 3, ns = 1 + (i / 9) % 2;
    printf("{\"T\":[");
    for (int k = 0; k < nt; k++) { const double v[4] = {3 * srand11(), 3 * srand11(), 3 * srand11(), k == 0 && i % 2 ? 0. : 2 * urand() - .5}; XfmPushTranslateSample(&list, v[0], v[1], v[2], v[3]); printf("[%.17g,%.17g,%.17g,%.17g]%s", v[0], v[1], v[2], v[3], k < nt - 1 ? "," : ""); }
    printf("],\"R\":[");
    for (int k = 0; k < nr; k++) { const double v[4] = {180 * srand11(), 180 * srand11(), 180 * srand11(), k == 0 ? 0. : 2 * urand() - .5}; XfmPushRotateSample(&list, v[0], v[1], v[2], v[3]); printf("[%.17g,%.17g,%.17g,%.17g]%s", v[0], v[1], v[2], v[3], k < nr - 1 ? "," : ""); }
    printf("],\"S\":[");
    for (int k = 0; k < ns; k++) { const double v[4] = {.2 + 2 * urand(), .2 + 2 * urand(), .2 + 2 * urand(), k == 0 ? 0. : 2 * urand() - .5}; XfmPushScaleSample(&list, v[0], v[1], v[2], v[3]); printf("[%.17g,%.17g,%.17g,%.17g]%s", v[0], v[1], v[2], v[3], k < ns - 1 ? "," : ""); }
    printf("],\"at\":[");
    for (int k = 0; k < 5; k++) {
      const double time = k == 0 ? 0. : (k == 1 ? 1. : 2.4 * urand() - .7);
      Transform x; XfmLerpTransformSample(&list, time, &x);
      printf("{\"time\":%.17g,\"fwd\":[", time);
      for (int e = 0; e < 16; e++) printf("%.17g%s", x.matrix.e[e], e < 15 ? "," : "");
      printf("],\"inv\":[");
      for (int e = 0; e < 16; e++) printf("%.17g%s", x.inverse.e[e], e < 15 ? "," : "");
      printf("]}%s", k < 4 ? "," : "");
    }
    printf("]}%s\n", i < 23 ? "," : "");
  }
  printf("]\n");
  printf("}\n");
}

// ------------------------------------------------------------------ timed scene runner
#include "parser.h"   // tools/scene_parser/parser.h via -I (Makefile.ref)

static std::chrono::steady_clock::time_point t_frame_start;
static double frame_seconds = 0;
static Interrupt my_frame_start(void *, const FrameInfo *info) {
  printf("# Rendering Frame %d x %d, %d threads, %d tiles\n", info->xres, info->yres, info->worker_count, info->tile_count);
  fflush(stdout);
  t_frame_start = std::chrono::steady_clock::now(); return CALLBACK_CONTINUE; }
static Interrupt my_frame_done(void *, const FrameInfo *) {
  frame_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_frame_start).count();
  printf("FJ_FRAME_SECONDS %.6f\n", frame_seconds); fflush(stdout); return CALLBACK_CONTINUE; }
static Interrupt my_tile(void *, const TileInfo *) { return CALLBACK_CONTINUE; }
static Interrupt my_sample(void *) { return CALLBACK_CONTINUE; }

static int run_scene(const char *path)
{
  std::ifstream file(path);
  if (!file) { std::cerr << "error: Could not open file: " << path << std::endl; return -1; }
  Parser parser; std::string line; int nrenderers = 0;
  const ID RENDERER_ID0 = 10000000L * 8;   // encode_id(Type_Renderer, 0), src/fj_scene_interface.cc:44-62,1053-1061
  while (getline(file, line)) {
    const bool is_render = line.compare(0, 11, "RenderScene") == 0;
    const auto t0 = std::chrono::steady_clock::now();
    const int err = parser.ParseLine(line);
    if (err) { std::cerr << "error: " << parser.GetErrorMessage() << ": " << parser.GetLineNumber() << ": " << line << std::endl; return -1; }
    if (is_render) { printf("FJ_RENDER_CALL_SECONDS %.6f\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()); fflush(stdout); }
    if (line.compare(0, 11, "NewRenderer") == 0) {
      const ID id = RENDERER_ID0 + nrenderers++;
      SiSetFrameReportCallback(id, NULL, my_frame_start, NULL, my_frame_done);
      SiSetTileReportCallback(id, NULL, my_tile, my_sample, my_tile);
    }
  }
  return 0;
}

int main(int argc, char **argv)
{
  if (argc >= 2 && strcmp(argv[1], "vectors") == 0) { dump_vectors(); return 0; }
  if (argc >= 3 && strcmp(argv[1], "run") == 0) return run_scene(argv[2]);
  fprintf(stderr, "usage: ref_probe vectors | ref_probe run file.scn\n");
  return 1;
}
