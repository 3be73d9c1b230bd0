// fj_gpu.cu — libfjgpu.so: the extern "C" ABI of include/fjgpu.h over the sm_100a kernels.
//
// Host side of the device path: keeps the scene description the caller hands over (meshes, instances,
// object groups, shaders, textures, lights, camera), builds the two BVH levels (fj_bvh.cc on the host or
// fj_build.cu on the device), lays everything out in HBM (DESIGN.md "Data layout"), and runs the frame:
// batches of tiles -> k_generate -> (k_extend2 -> k_shade) x rounds -> k_resolve_tiles -> packed tile blocks ->
// host frame / device frame / caller's device buffer.
// There is no CPU fallback anywhere in this file: without a CUDA device every entry point fails.
#include "fjgpu.h"
#include "fj_bvh.h"
#include "fj_build.h"
#include "fj_kernels.cuh"
#include "fj_extend.cuh"
#include "fj_extend_ring.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

std::string g_last_error;

struct DevBuf {
  void *p = nullptr; size_t bytes = 0;
  void *h = nullptr; size_t used = 0;     // retained pinned host copy of a scene array (fjgpu_scene_resend)
  void release() { if (p) cudaFree(p); if (h) cudaFreeHost(h); p = nullptr; h = nullptr; bytes = 0; used = 0; }
};

struct MeshRec {
  DevBuf nodes, nodes4, nodesq, tri, N, idx, group, uv, P, vel;
  std::vector<double> hostP;     // vertex positions kept until fjgpu_mesh_set_uv decides whether the device needs them
  fj::DMesh d;
  double bmin[3], bmax[3];      // exact FP64 bounds of the mesh (Mesh::ComputeBounds, fj_mesh.cc:235-244)
  int32_t nfaces = 0, nverts = 0, nnodes = 0, max_depth = 0, max_depth4 = 0, nnodes4 = 0, stack_need4 = 0, top4 = 0;
  std::vector<double> hostVel;   // per-vertex velocities (empty: static mesh)
  void release() { uv.release(); P.release(); vel.release(); nodes.release(); nodes4.release(); nodesq.release(); tri.release(); N.release(); idx.release(); group.release(); }
};

}  // namespace

struct fjgpu_context {
  int device = 0;
  size_t budget_key = 0, budget_cap_env = 0, budget_cached = 0; int budget_tiles = -1;      // plan_frame: batch budget of the last frame (see there)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::string err;
  int sm_count = 148;

  std::map<int, MeshRec> meshes;
  std::vector<fjgpu_instance> inst;
  std::vector<int32_t> group_off, group_ids;
  std::vector<fjgpu_shader> shaders;
  std::vector<fjgpu_light> lights;
  std::vector<fjgpu_texture> textures; std::vector<DevBuf> d_tex_tiles; DevBuf d_textures;
  std::vector<std::vector<double>> dome_dirs;
  std::vector<std::vector<float>> dome_cols;
  fjgpu_camera cam;
  bool have_cam = false, dirty = true;
  // motion blur: per-instance / camera matrices at every entry of the frame's time table (fjgpu.h)
  std::map<int, std::vector<double>> inst_motion;     // instance -> ntimes x (inv[12], fwd[12])
  std::vector<double> cam_motion;                     // ntimes x fwd[12]
  std::map<int, DevBuf> d_inst_motion; DevBuf d_cam_motion; bool cam_motion_dirty = false;
  std::map<int, bool> inst_motion_dirty;              // tables that changed since their last upload

  // device scene
  DevBuf d_meshes, d_inst, d_groups, d_shaders, d_lights;
  std::vector<DevBuf> d_group_nodes, d_group_nodes4, d_group_nodesq, d_group_order, d_group_irec, d_dome;
  fj::DScene sc;
  std::vector<int> mesh_slot_of_id;   // dense slot per mesh id (map order)
  uint64_t tlas_nodes = 0;
  int tlas_depth4 = 0;
  size_t ctl_off = 0;             // offset of the live QueueCtl inside d_ctl
  bool quant_ok = true;        // every tree of the committed scene has a quantised (NodeQ64) copy
  bool all_opaque = true;      // every shader returns Os = 1 (plastic: opacity 1): shadow rays may stop at any hit
  int stack_need = 0;          // worst-case traversal stack of the committed scene (entries)
  const void *top_src = nullptr; int top_avail = 0;     // NodeQ64 array of the largest tree and the length of its breadth-first front (k_extend2<TOP>)
  double shutter[2] = {0., 1.};                         // Renderer::SetSampleTimeRange (fjgpu_shutter_set)
  DevBuf d_timetab; size_t timetab_count = 0; double timetab_range[2] = {0., 0.};
  double build_seconds = 0;
  double device_build_seconds = 0;    // part of build_seconds spent inside fj_device_build (CUDA events)

  // frame resources
  DevBuf d_samples, d_tiles, d_blocks, d_jitter, d_counters, d_frame, d_queue[2], d_hits, d_ctl, d_hist, d_perm;
  float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {0, 0, 0};
  std::vector<cudaEvent_t> evpool;
  size_t jitter_count = 0;
  float *h_blocks = nullptr; size_t h_blocks_bytes = 0;
  DevBuf d_multi_send, d_multi_recv;     // fjgpu_render_frame_multi: this context's tile blocks / the all-gathered blocks of every rank
};

namespace {

int fail(fjgpu_context *c, int code, const std::string &msg) {
  static std::mutex mu;                   // fjgpu_render_frame_multi runs one host thread per context
  std::lock_guard<std::mutex> lock(mu);
  g_last_error = msg;
  if (c) c->err = msg;
  return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, FJGPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

int dev_alloc(fjgpu_context *ctx, DevBuf &b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return 0;
  b.release();
  if (bytes == 0) bytes = 16;
  CK(cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  return 0;
}
int dev_upload(fjgpu_context *ctx, DevBuf &b, const void *src, size_t bytes, bool keep = false) {
  if (int rc = dev_alloc(ctx, b, bytes)) return rc;
  if (keep && bytes) {       // scene arrays go through a retained pinned staging copy
    if (b.h) { cudaFreeHost(b.h); b.h = nullptr; }
    CK(cudaMallocHost(&b.h, bytes));
    memcpy(b.h, src, bytes);
    b.used = bytes;
    src = b.h;
  }
  if (bytes) CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

inline void rows12(const double *m16, double *m12) { memcpy(m12, m16, 12 * sizeof(double)); }

// FP32 box of a set of FP64 points, rounded outward and padded by a few ulps so that culling never rejects
// a box whose content the reference's FP64 tests would reach (fj_bvh.h).
void pad_box(const double lo[3], const double hi[3], fjb::Aabb *b) {
  for (int a = 0; a < 3; a++) {
    const double m = std::max(std::fabs(lo[a]), std::fabs(hi[a]));
    const double pad = 4e-7 * m + 1e-30;
    b->lo[a] = fjb::round_down(lo[a] - pad);
    b->hi[a] = fjb::round_up(hi[a] + pad);
  }
}

// matrix * point with the reference's operation order (MatTransformPoint, src/fj_matrix.cc:209-215)
inline void xpoint(const double *m, const double p[3], double out[3]) {
  for (int r = 0; r < 3; r++) out[r] = m[4 * r] * p[0] + m[4 * r + 1] * p[1] + m[4 * r + 2] * p[2] + m[4 * r + 3];
}

#define FJGPU_TOP_NODES_DEFAULT 0
int env_int(const char *name, int def) { const char *s = getenv(name); return s && *s ? atoi(s) : def; }

// ---- scene commit: instances, TLAS per object group, shader/light tables ------------------------
int commit_scene(fjgpu_context *ctx) {
  if (!ctx->dirty) return 0;
  const auto t0 = std::chrono::steady_clock::now();
  // dense mesh table
  std::map<int, int> slot;
  std::vector<fj::DMesh> dm;
  for (auto &kv : ctx->meshes) { slot[kv.first] = (int)dm.size(); dm.push_back(kv.second.d); }
  if (int rc = dev_upload(ctx, ctx->d_meshes, dm.data(), dm.size() * sizeof(fj::DMesh), true)) return rc;

  const int ninst = (int)ctx->inst.size();
  for (auto it = ctx->d_inst_motion.begin(); it != ctx->d_inst_motion.end();) {
    auto mit = ctx->inst_motion.find(it->first);
    if (it->first >= ninst || mit == ctx->inst_motion.end() || mit->second.empty()) { it->second.release(); it = ctx->d_inst_motion.erase(it); } else ++it;
  }
  std::vector<fj::DInstance> di(ninst);
  std::vector<fjb::Aabb> ibox(ninst);
  for (int i = 0; i < ninst; i++) {
    const fjgpu_instance &s = ctx->inst[i];
    auto it = ctx->meshes.find(s.mesh_id);
    if (it == ctx->meshes.end()) return fail(ctx, FJGPU_ERR_INVALID, "instance refers to an unknown mesh_id");
    fj::DInstance &d = di[i];
    memset(&d, 0, sizeof d);
    rows12(s.inv, d.inv); rows12(s.fwd, d.fwd);
    d.mesh = slot[s.mesh_id];
    const std::vector<double> *motion = nullptr;
    { auto mit = ctx->inst_motion.find(i); if (mit != ctx->inst_motion.end() && !mit->second.empty()) motion = &mit->second; }
    if (motion) {
      DevBuf &mb = ctx->d_inst_motion[i];
      if (!mb.p || ctx->inst_motion_dirty[i]) {
        if (int rc = dev_upload(ctx, mb, motion->data(), motion->size() * sizeof(double), true)) return rc;
        ctx->inst_motion_dirty[i] = false;
      }
      d.motion = (const double *)mb.p;
    }
    for (int g = 0; g < FJGPU_MAX_SHADING_GROUPS; g++) {
      d.shader_of_group[g] = s.shader_of_group[g];
      if (d.shader_of_group[g] >= (int)ctx->shaders.size()) return fail(ctx, FJGPU_ERR_INVALID, "instance refers to an unknown shader slot");
    }
    const int ng = (int)ctx->group_off.size() - 1;
    if (s.reflect_target < 0 || s.reflect_target >= ng || s.refract_target < 0 || s.refract_target >= ng ||
        s.shadow_target < 0 || s.shadow_target >= ng)
      return fail(ctx, FJGPU_ERR_INVALID, "instance target group out of range (call fjgpu_groups_set first)");
    d.reflect_target = s.reflect_target; d.refract_target = s.refract_target; d.shadow_target = s.shadow_target;
    // world bounds: the mesh accelerator's padded bounds through the forward matrix
    // (ObjectInstance::update_bounds, src/fj_object_instance.cc:299-360; MatTransformBounds, src/fj_matrix.cc:225-252)
    const MeshRec &m = it->second;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    if (m.nfaces > 0) {
      // a moving instance is bounded over every entry of its table: rays only ever carry those times, so the union is
      // conservative without the reference's bounding-sphere estimate (merge_sampled_bounds, fj_object_instance.cc:313-360)
      const size_t nt = motion ? motion->size() / 24 : 1;
      for (size_t t = 0; t < nt; t++) {
        double f16[16] = {0};
        if (motion) memcpy(f16, motion->data() + 24 * t + 12, 12 * sizeof(double));
        const double *fwd = motion ? f16 : s.fwd;
        for (int c = 0; c < 8; c++) {
          double p[3], q[3];
          for (int a = 0; a < 3; a++) p[a] = ((c >> a) & 1) ? m.bmax[a] + 1e-4 : m.bmin[a] - 1e-4;
          xpoint(fwd, p, q);
          for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], q[a]); hi[a] = std::max(hi[a], q[a]); }
        }
      }
      pad_box(lo, hi, &ibox[i]);
    } else {
      for (int a = 0; a < 3; a++) { ibox[i].lo[a] = 3e38f; ibox[i].hi[a] = -3e38f; }
    }
  }
  if (int rc = dev_upload(ctx, ctx->d_inst, di.data(), di.size() * sizeof(fj::DInstance), true)) return rc;
  for (int a = 0; a < 3; a++) { ctx->scene_lo[a] = 3e38f; ctx->scene_hi[a] = -3e38f; }
  for (int i = 0; i < ninst; i++) for (int a = 0; a < 3; a++) if (ibox[i].lo[a] <= ibox[i].hi[a]) {
    ctx->scene_lo[a] = std::min(ctx->scene_lo[a], ibox[i].lo[a]); ctx->scene_hi[a] = std::max(ctx->scene_hi[a], ibox[i].hi[a]);
  }

  const int ngroups = (int)ctx->group_off.size() - 1;
  for (auto &b : ctx->d_group_nodes) b.release();
  for (auto &b : ctx->d_group_nodes4) b.release();
  for (auto &b : ctx->d_group_nodesq) b.release();
  for (auto &b : ctx->d_group_order) b.release();
  for (auto &b : ctx->d_group_irec) b.release();
  ctx->d_group_irec.assign(std::max(ngroups, 0), DevBuf());
  ctx->d_group_nodesq.assign(std::max(ngroups, 0), DevBuf());
  ctx->quant_ok = true;
  for (auto &kv : ctx->meshes) ctx->quant_ok = ctx->quant_ok && kv.second.d.nodesq != nullptr;
  // the tree whose top k_extend2<TOP> stages in shared memory: the one with the most nodes that has a breadth-first front
  ctx->top_src = nullptr; ctx->top_avail = 0;
  { int best = 0; for (auto &kv : ctx->meshes) { const MeshRec &m = kv.second; if (m.top4 > 0 && m.d.nodesq && m.nnodes4 > best) { best = m.nnodes4; ctx->top_src = m.d.nodesq; ctx->top_avail = m.top4; } } }
  ctx->d_group_nodes.assign(std::max(ngroups, 0), DevBuf());
  ctx->d_group_nodes4.assign(std::max(ngroups, 0), DevBuf());
  ctx->d_group_order.assign(std::max(ngroups, 0), DevBuf());
  std::vector<fj::DGroup> dg(std::max(ngroups, 0));
  ctx->tlas_nodes = 0;
  int tlas_need = 0;
  for (int g = 0; g < ngroups; g++) {
    const int b = ctx->group_off[g], e = ctx->group_off[g + 1];
    std::vector<fjb::Aabb> boxes; std::vector<int32_t> ids;
    for (int k = b; k < e; k++) {
      const int id = ctx->group_ids[k];
      if (id < 0 || id >= ninst) return fail(ctx, FJGPU_ERR_INVALID, "object group refers to an unknown instance");
      if (ibox[id].lo[0] > ibox[id].hi[0]) continue;      // empty mesh: can never be hit
      boxes.push_back(ibox[id]); ids.push_back(id);
    }
    fjb::BuildResult br;
    fjb::build_bvh(boxes.data(), (int32_t)boxes.size(), 1, 1.f, 0, &br);
    std::vector<int32_t> order(std::max<size_t>(br.order.size(), 1), 0);
    for (size_t k = 0; k < br.order.size(); k++) order[k] = ids[br.order[k]];
    if (int rc = dev_upload(ctx, ctx->d_group_nodes[g], br.nodes.data(), br.nodes.size() * sizeof(fjb::Node64), true)) return rc;
    if (int rc = dev_upload(ctx, ctx->d_group_order[g], order.data(), order.size() * sizeof(int32_t), true)) return rc;
    if (int rc = dev_upload(ctx, ctx->d_group_nodes4[g], br.nodes4.data(), br.nodes4.size() * sizeof(fjb::Node128), true)) return rc;
    dg[g].nodes4 = (const float4 *)ctx->d_group_nodes4[g].p;
    {
      tlas_need = std::max(tlas_need, br.stack_need4);
      std::vector<fjb::NodeQ64> nq(br.nodes4.size());
      float bq = 0;
      if (fjb::quantize_nodes(br.nodes4.data(), br.nodes4.size(), nq.data(), &bq)) {
        if (int rc = dev_upload(ctx, ctx->d_group_nodesq[g], nq.data(), nq.size() * sizeof(fjb::NodeQ64), true)) return rc;
        dg[g].nodesq = (const float4 *)ctx->d_group_nodesq[g].p; dg[g].bmagq = bq;
      } else { dg[g].nodesq = nullptr; dg[g].bmagq = 0; ctx->quant_ok = false; }
    }
    ctx->tlas_depth4 = std::max(ctx->tlas_depth4, br.max_depth4);
    dg[g].nodes = (const float4 *)ctx->d_group_nodes[g].p;
    dg[g].order = (const int32_t *)ctx->d_group_order[g].p;
    {
      std::vector<fj::DInstRec> rec(order.size());
      memset(rec.data(), 0, rec.size() * sizeof(fj::DInstRec));
      for (size_t k = 0; k < br.order.size(); k++) {
        const fj::DInstance &in = di[order[k]];
        const fj::DMesh &me = dm[in.mesh];
        fj::DInstRec &r = rec[k];
        memcpy(r.inv, in.inv, sizeof r.inv);
        r.nodes4 = (const char *)me.nodes4; r.nodesq = (const char *)me.nodesq;
        r.tri64 = me.tri64v ? 2 : (me.tri32 == nullptr ? 1 : 0);
        r.tri = me.tri64v ? (const void *)me.tri64v : (r.tri64 ? (const void *)me.tri64 : (const void *)me.tri32);
        r.bmag = me.bmag; r.bmagq = me.bmagq; r.inst = order[k]; r.motion = in.motion;
      }
      if (int rc = dev_upload(ctx, ctx->d_group_irec[g], rec.data(), rec.size() * sizeof(fj::DInstRec), true)) return rc;
      dg[g].irec = (const fj::DInstRec *)ctx->d_group_irec[g].p;
    }
    dg[g].ninst = (int32_t)ids.size();
    { double b = 0; for (int a = 0; a < 3 && !boxes.empty(); a++) b = std::max(b, std::max(std::fabs((double)br.bounds.lo[a]), std::fabs((double)br.bounds.hi[a]))); dg[g].bmag = fjb::round_up(b); }
    ctx->tlas_nodes += br.nodes.size();
  }
  if (int rc = dev_upload(ctx, ctx->d_groups, dg.data(), dg.size() * sizeof(fj::DGroup), true)) return rc;
  // instance tree + the other instances of a TLAS leaf (<= 7) + the BLAS sentinel + the deepest BLAS
  { int blas_need = 0; for (auto &kv : ctx->meshes) blas_need = std::max(blas_need, kv.second.stack_need4); ctx->stack_need = tlas_need + 7 + 1 + blas_need + 2; }

  std::vector<fj::DShader> ds(ctx->shaders.size());
  for (size_t i = 0; i < ds.size(); i++) {
    const fjgpu_shader &s = ctx->shaders[i]; fj::DShader &d = ds[i];
    d.kind = s.kind; d.do_reflect = s.do_reflect; d.do_color_filter = s.do_color_filter; d.texture = s.texture; d.bump_texture = s.bump_texture; d.bump_amplitude = s.bump_amplitude;
    if (s.bump_texture < 0 || s.bump_texture > (int)ctx->textures.size()) return fail(ctx, FJGPU_ERR_INVALID, "shader refers to an unknown bump texture");
    if (s.texture < 0 || s.texture > (int)ctx->textures.size()) return fail(ctx, FJGPU_ERR_INVALID, "shader refers to an unknown texture");
    memcpy(d.diffuse, s.diffuse, 12); memcpy(d.reflect, s.reflect, 12); memcpy(d.refract, s.refract, 12);
    memcpy(d.emission, s.emission, 12); memcpy(d.transmit, s.transmit, 12);
    d.ior = s.ior; d.opacity = s.opacity;
  }
  if (int rc = dev_upload(ctx, ctx->d_shaders, ds.data(), ds.size() * sizeof(fj::DShader), true)) return rc;
  ctx->all_opaque = true;      // constant / glass / pathtracing shaders and NO_SHADER return Os = 1; plastic returns `opacity`
  for (const fj::DShader &d : ds) if (d.kind == FJGPU_SHADER_PLASTIC && !(d.opacity >= 1.f)) ctx->all_opaque = false;

  for (auto &b : ctx->d_dome) b.release();
  ctx->d_dome.assign(2 * ctx->lights.size(), DevBuf());
  std::vector<fj::DLight> dl(ctx->lights.size());
  for (size_t i = 0; i < dl.size(); i++) {
    const fjgpu_light &s = ctx->lights[i]; fj::DLight &d = dl[i];
    memset(&d, 0, sizeof d);
    d.kind = s.kind; d.sample_count = s.sample_count; d.double_sided = s.double_sided;
    d.dome_count = (int32_t)(ctx->dome_dirs[i].size() / 3);
    memcpy(d.color, s.color, 12); d.intensity = s.intensity;
    memcpy(d.translate, s.translate, 24); rows12(s.fwd, d.fwd);
    if (d.dome_count > 0) {
      if (int rc = dev_upload(ctx, ctx->d_dome[2 * i], ctx->dome_dirs[i].data(), ctx->dome_dirs[i].size() * 8)) return rc;
      if (int rc = dev_upload(ctx, ctx->d_dome[2 * i + 1], ctx->dome_cols[i].data(), ctx->dome_cols[i].size() * 4)) return rc;
      d.dome_dirs = (const double *)ctx->d_dome[2 * i].p; d.dome_colors = (const float *)ctx->d_dome[2 * i + 1].p;
    }
  }
  if (int rc = dev_upload(ctx, ctx->d_lights, dl.data(), dl.size() * sizeof(fj::DLight), true)) return rc;
  CK(cudaStreamSynchronize(ctx->stream));

  fj::DScene &sc = ctx->sc;
  {
    std::vector<fj::DTexture> dt(ctx->textures.size());
    for (size_t i = 0; i < dt.size(); i++) {
      const fjgpu_texture &t = ctx->textures[i];
      dt[i].tiles = (const float *)ctx->d_tex_tiles[i].p; dt[i].width = t.width; dt[i].height = t.height; dt[i].nch = t.nchannels;
      dt[i].tilesize = t.tilesize; dt[i].xnt = t.width / t.tilesize; dt[i].ynt = t.height / t.tilesize;
    }
    if (int rc = dev_upload(ctx, ctx->d_textures, dt.data(), dt.size() * sizeof(fj::DTexture), true)) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    sc.textures = (const fj::DTexture *)ctx->d_textures.p;
  }
  sc.meshes = (const fj::DMesh *)ctx->d_meshes.p; sc.inst = (const fj::DInstance *)ctx->d_inst.p;
  sc.groups = (const fj::DGroup *)ctx->d_groups.p; sc.shaders = (const fj::DShader *)ctx->d_shaders.p;
  sc.lights = (const fj::DLight *)ctx->d_lights.p;
  sc.nmeshes = (int)dm.size(); sc.ninst = ninst; sc.ngroups = ngroups; sc.nshaders = (int)ds.size(); sc.nlights = (int)dl.size(); sc.pad = 0;
  ctx->dirty = false;
  ctx->build_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

// ---- frame set-up ---------------------------------------------------------------------------------
struct FramePlan {
  fj::DFrame fr; fj::DCamera cam;
  uint32_t wstride = 0; int bw = 0, bh = 0;
  int tiles_per_batch = 0;
  size_t budget = 0;                  // bytes the per-batch buffers may take (plan_frame)
  int waves = 1; double peak = 1;     // wavefront: number of extend/shade rounds, worst-case queue records per camera sample in one round
  double factor0 = 1;                 // queue capacity per sample slot the first attempt allocates
};

// Worst-case width of the ray tree per camera sample (rays alive in one wavefront round) and the number of rounds,
// from the lobes the scene's shaders can spawn and the per-type bounce limits (has_reached_bounce_limit,
// src/fj_shading.cc:467-499).
void frontier(const fjgpu_context *ctx, const fjgpu_render_params *p, int *waves, double *peak, double *block_out, bool mega_for_plan = false) {
  bool D = false, R = false, F = false; int b = 0;
  for (const fjgpu_shader &s : ctx->shaders) {
    int n = 0;
    if (s.kind == FJGPU_SHADER_PLASTIC) { if (s.do_reflect) { R = true; n = 1; } }
    else if (s.kind == FJGPU_SHADER_GLASS) { R = true; F = true; n = 2; }          // always one mirror and one refracted child
    else if (s.kind == FJGPU_SHADER_PATHTRACING) {
      if (s.diffuse[0] > 0 || s.diffuse[1] > 0 || s.diffuse[2] > 0) { D = true; n++; }
      if (s.reflect[0] > 0 || s.reflect[1] > 0 || s.reflect[2] > 0) { R = true; n++; }
      if (s.refract[0] > 0 || s.refract[1] > 0 || s.refract[2] > 0) { F = true; n++; }
    }
    b = std::max(b, n);
  }
  const int md = D ? p->max_diffuse_depth : 0, mr = R ? p->max_reflect_depth : 0, mf = F ? p->max_refract_depth : 0;
  *waves = 1 + md + mr + mf;
  *peak = 1;
  // wavefront shadow rays: every plastic hit writes a block of 1 header + one record per light sample into the next queue
  // (fj_kernels.cuh, Shading), and the blocks of the last bounce need one more extend / shade round
  double block = 0;
  {
    bool plastic = false;
    for (const fjgpu_shader &s : ctx->shaders) plastic = plastic || s.kind == FJGPU_SHADER_PLASTIC;
    long samples = 0;
    for (size_t i = 0; i < ctx->lights.size(); i++) {
      const fjgpu_light &l = ctx->lights[i];
      samples += l.kind == FJGPU_LIGHT_POINT ? 1 : (l.kind == FJGPU_LIGHT_DOME ? std::min<long>(l.sample_count, (long)(ctx->dome_dirs[i].size() / 3)) : l.sample_count);
    }
    if (plastic && p->cast_shadow && samples > 0 && !mega_for_plan) { block = 1.0 + (double)samples; *waves += 1; }
  }
  *block_out = block;
  if (b <= 1) { *peak = 1 + block; return; }
  std::vector<double> fact(md + mr + mf + 1, 1.);
  for (size_t i = 1; i < fact.size(); i++) fact[i] = fact[i - 1] * (double)i;
  for (int w = 1; w <= md + mr + mf; w++) {
    double width = 0;
    for (int nd = 0; nd <= md; nd++) for (int nr = 0; nr <= mr; nr++) { const int nf = w - nd - nr; if (nf < 0 || nf > mf) continue; width += fact[w] / (fact[nd] * fact[nr] * fact[nf]); }
    *peak = std::max(*peak, std::min(width, std::pow((double)b, w)));
  }
  *peak *= 1 + block;
}

int plan_frame(fjgpu_context *ctx, const fjgpu_render_params *p, const fjgpu_tile *tiles, int ntiles, FramePlan *pl) {
  if (!p || (!tiles && ntiles > 0) || ntiles < 0) return fail(ctx, FJGPU_ERR_INVALID, "null params/tiles");
  if (p->xres <= 0 || p->yres <= 0 || p->xrate <= 0 || p->yrate <= 0 || !(p->xfwidth > 0) || !(p->yfwidth > 0))
    return fail(ctx, FJGPU_ERR_INVALID, "resolution, pixelsamples and filterwidth must be positive");
  if (!ctx->have_cam) return fail(ctx, FJGPU_ERR_INVALID, "no camera set");
  if (int rc = commit_scene(ctx)) return rc;
  if (p->target_group < 0 || p->target_group >= ctx->sc.ngroups) return fail(ctx, FJGPU_ERR_INVALID, "target_group out of range");
  if (p->max_diffuse_depth < 0 || p->max_reflect_depth < 0 || p->max_refract_depth < 0 ||
      p->max_diffuse_depth + p->max_reflect_depth + p->max_refract_depth > 250)
    return fail(ctx, FJGPU_ERR_INVALID, "max_*_depth out of range");
  fj::DFrame &fr = pl->fr;
  memset(&fr, 0, sizeof fr);
  fr.xres = p->xres; fr.yres = p->yres; fr.xrate = p->xrate; fr.yrate = p->yrate;
  // count_samples_in_margin, src/fj_fixed_grid_sampler.cc:131-136
  fr.mx = (int)std::ceil(((p->xfwidth - 1) * p->xrate) * .5);
  fr.my = (int)std::ceil(((p->yfwidth - 1) * p->yrate) * .5);
  if (fr.mx < 0) fr.mx = 0; if (fr.my < 0) fr.my = 0;
  fr.xfw = p->xfwidth; fr.yfw = p->yfwidth; fr.jitter = p->jitter;
  fr.udelta = 1. / (p->xrate * p->xres); fr.vdelta = 1. / (p->yrate * p->yres);       // :45-46
  fr.max_diffuse = p->max_diffuse_depth; fr.max_reflect = p->max_reflect_depth; fr.max_refract = p->max_refract_depth;
  fr.cast_shadow = p->cast_shadow; fr.target_group = p->target_group; fr.seed = p->seed; fr.flags = p->flags;
  int tw = 1, th = 1;
  for (int i = 0; i < ntiles; i++) {
    const fjgpu_tile &t = tiles[i];
    if (t.xmax <= t.xmin || t.ymax <= t.ymin) return fail(ctx, FJGPU_ERR_INVALID, "empty tile");
    if (t.xmin < 0 || t.ymin < 0 || t.xmax > p->xres || t.ymax > p->yres) return fail(ctx, FJGPU_ERR_INVALID, "tile outside the frame");
    tw = std::max(tw, t.xmax - t.xmin); th = std::max(th, t.ymax - t.ymin);
  }
  pl->bw = tw; pl->bh = th;
  const long nsx = (long)p->xrate * tw + 2 * fr.mx, nsy = (long)p->yrate * th + 2 * fr.my;
  const long slots = ((nsx + 7) / 8) * ((nsy + 3) / 4) * 32;
  if (slots > (1l << 30)) return fail(ctx, FJGPU_ERR_UNSUPPORTED, "tile sample grid too large");
  pl->wstride = (uint32_t)slots;
  fr.max_ns = (int32_t)(nsx * nsy);
  // jitter table: the first 2*max_ns draws of a default-seeded XorShift (src/fj_random.cc:10-43) —
  // every tile restarts the same stream (src/fj_fixed_grid_sampler.cc:41-42)
  const size_t need = 2 * (size_t)fr.max_ns;
  if (ctx->jitter_count < need) {
    std::vector<uint32_t> tab(need);
    uint32_t s[4] = {123456789u, 362436069u, 521288629u, 88675123u};
    for (size_t i = 0; i < need; i++) {
      const uint32_t t = s[0] ^ (s[0] << 11);
      s[0] = s[1]; s[1] = s[2]; s[2] = s[3];
      s[3] = (s[3] ^ (s[3] >> 19)) ^ (t ^ (t >> 8));
      tab[i] = s[3];
    }
    if (int rc = dev_upload(ctx, ctx->d_jitter, tab.data(), need * 4)) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->jitter_count = need;
  }
  fr.jitter_tab = (const uint32_t *)ctx->d_jitter.p;
  // Camera: uv_size_ (src/fj_camera.cc:97-101 via update_uv_size) — aspect = xres / (double) yres
  const double PI = 3.14159265358979323846;
  const double aspect = p->xres / (double)p->yres;
  pl->cam.uvy = 2 * std::tan((ctx->cam.fov / 2.) * PI / 180.);
  pl->cam.uvx = pl->cam.uvy * aspect;
  rows12(ctx->cam.fwd, pl->cam.fwd); pl->cam.znear = ctx->cam.znear; pl->cam.zfar = ctx->cam.zfar;
  // time-sampled transforms: every table must cover the frame's time table (one entry per sample of the largest tile)
  pl->cam.motion = nullptr;
  for (auto &kv : ctx->inst_motion)
    if (!kv.second.empty() && kv.second.size() / 24 < (size_t)fr.max_ns)
      return fail(ctx, FJGPU_ERR_INVALID, "instance motion table is shorter than the frame's time table (fjgpu_time_table)");
  if (!ctx->cam_motion.empty()) {
    if (ctx->cam_motion.size() / 12 < (size_t)fr.max_ns) return fail(ctx, FJGPU_ERR_INVALID, "camera motion table is shorter than the frame's time table (fjgpu_time_table)");
    if (ctx->cam_motion_dirty) {
      if (int rc = dev_upload(ctx, ctx->d_cam_motion, ctx->cam_motion.data(), ctx->cam_motion.size() * sizeof(double))) return rc;
      CK(cudaStreamSynchronize(ctx->stream));
      ctx->cam_motion_dirty = false;
    }
    pl->cam.motion = (const double *)ctx->d_cam_motion.p;
  }
  // meshes with vertex velocity read the time VALUE of a ray's table entry: the frame's table over the shutter, on the device
  {
    bool any_vel = false;
    for (auto &kv : ctx->meshes) any_vel = any_vel || kv.second.d.tri64v != nullptr;
    if (any_vel) {
      const size_t need = (size_t)fr.max_ns;
      if (ctx->timetab_count < need || ctx->timetab_range[0] != ctx->shutter[0] || ctx->timetab_range[1] != ctx->shutter[1]) {
        std::vector<double> tab(need);
        fjgpu_tile whole; whole.id = 0; whole.xmin = 0; whole.ymin = 0; whole.xmax = tw; whole.ymax = th;
        if (fjgpu_time_table(p, &whole, 1, ctx->shutter[0], ctx->shutter[1], tab.data(), (int32_t)need) != (int32_t)need)
          return fail(ctx, FJGPU_ERR_INVALID, "cannot build the frame's time table");
        if (int rc = dev_upload(ctx, ctx->d_timetab, tab.data(), need * sizeof(double))) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->timetab_count = need; ctx->timetab_range[0] = ctx->shutter[0]; ctx->timetab_range[1] = ctx->shutter[1];
      }
      ctx->sc.time_tab = (const double *)ctx->d_timetab.p;
    } else ctx->sc.time_tab = nullptr;
  }
  // batch so the per-batch buffers (accumulators + two ray queues + hit records) stay bounded
  double block = 0;
  frontier(ctx, p, &pl->waves, &pl->peak, &block);
  if (block > 1023) return fail(ctx, FJGPU_ERR_UNSUPPORTED, "more than 1022 light samples per shading point");
  // ray trees that can branch start optimistic — two rays per sample slot, each with its shadow-ray block — and the queue is
  // grown when a batch overflows it (any batch: render_impl); a tree that cannot branch gets its exact worst case at once
  pl->factor0 = std::min(pl->peak, 2.0 * (1.0 + block));
  // budget of the per-batch buffers: FJGPU_SAMPLE_MB (default 16 GiB), never more than 80 % of what the device has free
  // now plus what this context already holds for the purpose
  size_t cap = (size_t)env_int("FJGPU_SAMPLE_MB", 16384) << 20;
  const size_t per_slot = sizeof(fj::Accum) + (size_t)std::ceil(pl->factor0 * (2 * sizeof(fj::RayRec) + sizeof(fj::HitRec)));
  {
    // cudaMemGetInfo goes through the kernel driver (and its global lock, which every nvidia-smi / NVML poll on the box takes
    // too): a frame whose buffers the context already holds reuses the budget of the frame that sized them
    const size_t want_key = (size_t)pl->wstride * per_slot, held = ctx->d_samples.bytes + ctx->d_queue[0].bytes;
    if (ctx->budget_key == want_key && ctx->budget_cap_env == cap && ctx->budget_tiles == ntiles && held > 0) cap = ctx->budget_cached;
    else {
      size_t free_b = 0, total_b = 0;
      const size_t cap_env = cap;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        const size_t avail = (size_t)(0.8 * (double)(free_b + held));
        cap = std::min(cap, std::max<size_t>(avail, (size_t)64 << 20));
      }
      ctx->budget_key = want_key; ctx->budget_cap_env = cap_env; ctx->budget_tiles = ntiles; ctx->budget_cached = cap;
    }
  }
  pl->budget = cap;
  long per = (long)(cap / ((size_t)pl->wstride * per_slot));
  per = std::max(1l, std::min<long>(per, std::max(ntiles, 1)));
  const long nbatches = (std::max(ntiles, 1) + per - 1) / per;             // equal batches: no short last batch with long tails
  pl->tiles_per_batch = (int)((std::max(ntiles, 1) + nbatches - 1) / nbatches);
  return 0;
}

enum { OUT_HOST = 0, OUT_DEVICE_BLOCKS = 1, OUT_RESIDENT = 2, OUT_SAMPLES_ONLY = 3 };

template <int MINB, int SD, bool F2, bool RING>
void launch_extend_ring(fjgpu_context *ctx, const fj::RenderArgs &a, int blocks) {
  const size_t per_cta = (RING ? sizeof(fj::RingShared) : sizeof(fj::ExtShared)) + (size_t)SD * FJ_XT * sizeof(int) + 1024 + 1024;
  static size_t configured[64] = {0};
  size_t &done = configured[ctx->device & 63];
  if (done != per_cta) {
    int pct = (int)std::min(100.0, std::ceil(100.0 * MINB * per_cta / (228.0 * 1024)));
    pct = env_int("FJGPU_CARVEOUT_PCT", pct);
    cudaFuncSetAttribute(fj::k_extend_ring<MINB, SD, F2, RING>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    done = per_cta;
  }
  fj::k_extend_ring<MINB, SD, F2, RING><<<blocks, FJ_XT, 0, ctx->stream>>>(a);
}

template <int MINB, bool QUANT, bool COOP, int SD, bool TOP>
void launch_extend2(fjgpu_context *ctx, const fj::RenderArgs &a, int blocks) {
  // shared-memory carveout: MINB CTAs x (static + dynamic shared memory + 1 KB the driver reserves per CTA), the rest stays L1
  const size_t dyn = TOP ? (size_t)a.top_count * 64 : 0;
  const size_t per_cta = sizeof(fj::ExtShared) + (size_t)SD * FJ_XT * sizeof(int) + (COOP ? 1024 : 0) + 64 + dyn + 1024;
  static size_t configured[64] = {0};      // function attributes are per device: one process may drive several GPUs
  size_t &done = configured[ctx->device & 63];
  if (done != per_cta) {
    int pct = (int)std::min(100.0, std::ceil(100.0 * MINB * per_cta / (228.0 * 1024)));
    pct = env_int("FJGPU_CARVEOUT_PCT", pct);
    if (dyn) cudaFuncSetAttribute(fj::k_extend2<MINB, true, QUANT, COOP, SD, TOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    cudaFuncSetAttribute(fj::k_extend2<MINB, true, QUANT, COOP, SD, TOP>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    done = per_cta;
  }
  fj::k_extend2<MINB, true, QUANT, COOP, SD, TOP><<<blocks, FJ_XT, dyn, ctx->stream>>>(a);
}

// The closest-hit kernel of one wavefront round.  FJGPU_EXTEND=3 (default) is k_extend_ring (fj_extend_ring.cuh: quantised trees,
// cooperative leaves; FJGPU_B1_MIN / FJGPU_B2_MIN gate its leaf-test / instance-entry phases, FJGPU_RING=1 adds the per-warp ring
// of prepared rays), 2 k_extend2 (also the fallback for trees that cannot be quantised), 1 the register-resident first version
// (both kept as cross-checks); FJGPU_EXTEND_MINBLOCKS = resident CTAs per SM, FJGPU_STACK_SMEM = stack entries per lane kept in
// shared memory (8 / 12 / 16), FJGPU_TOP_NODES = nodes of the largest tree staged in shared memory by one bulk copy per CTA
// (k_extend2 only; 0 = off).
void launch_extend(fjgpu_context *ctx, fj::RenderArgs &a, int grid) {
  a.refill = std::min(32, std::max(1, env_int("FJGPU_REFILL", 12)));
  a.phase_a_min = std::min(32, std::max(1, env_int("FJGPU_PHASE_A_MIN", 16)));
  a.park = env_int("FJGPU_PARK", 1);
  const int version = env_int("FJGPU_EXTEND", 3);
  if (version >= 2) {
    const int minb = env_int("FJGPU_EXTEND_MINBLOCKS", 7);
    const int cap = std::max(1, grid / 4 * minb);      // `grid` is 4 CTAs per SM worth of work (or fewer for small probes)
    const bool quant = ctx->quant_ok && env_int("FJGPU_QUANT", 1) != 0;
    const bool coop = quant && env_int("FJGPU_COOP", 1) != 0;       // warp-cooperative exact triangle tests (fj_extend.cuh, phase B1)
    const int sd = env_int("FJGPU_STACK_SMEM", 12);
    a.top_src = nullptr; a.top_count = 0;
    a.shadow_anyhit = ctx->all_opaque && env_int("FJGPU_ANYHIT", 1) != 0 ? 1 : 0;
    if (version >= 3 && coop && ctx->sc.ngroups < (1 << 20)) {      // k_extend_ring (fj_extend_ring.cuh; the target group travels in 20 bits of the lane's state word)
      a.b1_min = std::max(1, env_int("FJGPU_B1_MIN", 20)); a.b2_min = std::max(1, env_int("FJGPU_B2_MIN", 6));
      const bool ring = env_int("FJGPU_RING", 0) != 0;       // per-warp ring of prepared rays, or the direct refill of k_extend2
      a.refill = std::min(32, std::max(1, env_int("FJGPU_REFILL", ring ? 32 : 8)));
      const bool f2 = env_int("FJGPU_FMA2", 1) != 0;         // packed FP32 FMAs (FFMA2) in the node step
      if (!ring) {
        if (minb >= 8) launch_extend_ring<8, 8, true, false>(ctx, a, cap);
        else if (minb == 7 && sd <= 8) launch_extend_ring<7, 8, true, false>(ctx, a, cap);
        else if (minb == 7) launch_extend_ring<7, 12, true, false>(ctx, a, cap);
        else launch_extend_ring<6, 12, true, false>(ctx, a, cap);
      } else if (minb >= 8) launch_extend_ring<8, 8, true, true>(ctx, a, cap);
      else if (minb == 7 && sd >= 12) launch_extend_ring<7, 12, true, true>(ctx, a, cap);
      else if (minb == 7 && f2) launch_extend_ring<7, 8, true, true>(ctx, a, cap);
      else if (minb == 7) launch_extend_ring<7, 8, false, true>(ctx, a, cap);
      else launch_extend_ring<6, 12, true, true>(ctx, a, cap);
      return;
    }
    if (quant && coop && ctx->top_src) {
      const int want = std::min(env_int("FJGPU_TOP_NODES", FJGPU_TOP_NODES_DEFAULT), ctx->top_avail);
      if (want > 0) { a.top_src = (const char *)ctx->top_src; a.top_count = want; }
    }
    if (a.top_count > 0) {                               // staged tree top (TMA bulk copy): 7 or 6 CTAs per SM
      if (minb >= 7) launch_extend2<7, true, true, 12, true>(ctx, a, cap);
      else launch_extend2<6, true, true, 16, true>(ctx, a, cap);
    } else if (coop) {
      if (minb >= 8) launch_extend2<8, true, true, 8, false>(ctx, a, cap);
      else if (minb == 7 && sd >= 16) launch_extend2<7, true, true, 16, false>(ctx, a, cap);
      else if (minb == 7 && sd <= 8) launch_extend2<7, true, true, 8, false>(ctx, a, cap);
      else if (minb == 7) launch_extend2<7, true, true, 12, false>(ctx, a, cap);
      else launch_extend2<6, true, true, 16, false>(ctx, a, cap);
    } else if (quant) {
      if (minb >= 8) launch_extend2<8, true, false, 8, false>(ctx, a, cap);
      else if (minb == 7) launch_extend2<7, true, false, 12, false>(ctx, a, cap);
      else launch_extend2<6, true, false, 16, false>(ctx, a, cap);
    } else {
      if (minb >= 7) launch_extend2<7, false, false, 12, false>(ctx, a, cap);
      else launch_extend2<6, false, false, 16, false>(ctx, a, cap);
    }
    return;
  }
  const int minb = env_int("FJGPU_EXTEND_MINBLOCKS", 5);
  const int g = std::max(1, (grid * minb + 3) / 4);
  if (minb >= 6) fj::k_extend<6><<<g, 128, 0, ctx->stream>>>(a);
  else if (minb == 5) fj::k_extend<5><<<g, 128, 0, ctx->stream>>>(a);
  else fj::k_extend<4><<<grid, 128, 0, ctx->stream>>>(a);
}

cudaEvent_t pool_event(fjgpu_context *ctx, size_t *next) {
  if (*next >= ctx->evpool.size()) { cudaEvent_t e; cudaEventCreate(&e); ctx->evpool.push_back(e); }
  return ctx->evpool[(*next)++];
}

int render_impl(fjgpu_context *ctx, const fjgpu_render_params *p, const fjgpu_tile *tiles, int ntiles, int mode,
                float *rgba_frame, void *d_out_blocks, int out_bw, int out_bh, fjgpu_stats *stats) {
  if (!ctx) return fail(nullptr, FJGPU_ERR_INVALID, "null context");
  CK(cudaSetDevice(ctx->device));
  FramePlan pl;
  if (int rc = plan_frame(ctx, p, tiles, ntiles, &pl)) return rc;
  if (mode == OUT_DEVICE_BLOCKS) {
    if (out_bw < pl.bw || out_bh < pl.bh) return fail(ctx, FJGPU_ERR_INVALID, "tile_w_max/tile_h_max smaller than a tile");
    pl.bw = out_bw; pl.bh = out_bh;
  }
  if (stats) memset(stats, 0, sizeof *stats);
  if (ntiles == 0) return 0;
  const size_t block_floats = (size_t)pl.bw * pl.bh * 4;
  const bool fp64_boxes = (p->flags & FJGPU_FLAG_FP64_BOXES) != 0;
  const bool mega = fp64_boxes || (p->flags & FJGPU_FLAG_MEGAKERNEL) != 0 || env_int("FJGPU_MEGAKERNEL", 0) != 0;
  if (mega && 2 * (p->max_diffuse_depth + p->max_reflect_depth + p->max_refract_depth) + 1 > FJ_PENDING)      // (the cross-check path only)
    return fail(ctx, FJGPU_ERR_UNSUPPORTED, "megakernel: 2*(diffuse+reflect+refract)+1 exceeds the per-path ray stack");

  if (int rc = dev_upload(ctx, ctx->d_tiles, tiles, (size_t)ntiles * sizeof(fjgpu_tile))) return rc;
  if (int rc = dev_alloc(ctx, ctx->d_samples, (size_t)pl.tiles_per_batch * pl.wstride * sizeof(fj::Accum))) return rc;
  if (int rc = dev_alloc(ctx, ctx->d_counters, 2 * sizeof(fj::DCounters) + 64)) return rc;      // live counters, work counter, snapshot at the batch start
  if (int rc = dev_alloc(ctx, ctx->d_ctl, (size_t)4 << 20)) return rc;      // own 2-MB pages; the live QueueCtl sits at ctl_off inside
  // The queue counters get an allocation of their own and sit at its start.  Measured, not yet explained: k_shade, whose
  // every warp takes its output slots with one atomicAdd on QueueCtl::count, runs the north-star frame in 61.5-62.8 ms when
  // the counter lies in the first 2 KB of a 2-MB-aligned block and in 74.6-78.5 ms at any of 13 other offsets tried (4 KB ...
  // 3 MB), while a kernel of nothing but those atomics is equally fast everywhere (profiles/r1_queue_counter_placement.txt).
  // As a 16-byte allocation from the runtime's small-block pool the counter landed on either side from build to build.
  ctx->ctl_off = ((size_t)std::max(0, env_int("FJGPU_CTL_OFFSET", 0)) & ~(size_t)255) % ((size_t)3 << 20);
  char *const ctl_p = (char *)ctx->d_ctl.p + ctx->ctl_off;
  float4 *blocks = nullptr;
  if (mode == OUT_DEVICE_BLOCKS) {
    blocks = (float4 *)d_out_blocks;
    CK(cudaMemsetAsync(blocks, 0, (size_t)ntiles * block_floats * 4, ctx->stream));
  } else if (mode != OUT_SAMPLES_ONLY) {
    if (int rc = dev_alloc(ctx, ctx->d_blocks, (size_t)ntiles * block_floats * 4)) return rc;
    blocks = (float4 *)ctx->d_blocks.p;
  }
  if (mode == OUT_RESIDENT)
    if (int rc = dev_alloc(ctx, ctx->d_frame, (size_t)p->xres * p->yres * sizeof(float4))) return rc;
  if (mode == OUT_HOST && ctx->h_blocks_bytes < (size_t)ntiles * block_floats * 4) {
    if (ctx->h_blocks) cudaFreeHost(ctx->h_blocks);
    ctx->h_blocks = nullptr; ctx->h_blocks_bytes = 0;
    CK(cudaMallocHost((void **)&ctx->h_blocks, (size_t)ntiles * block_floats * 4));
    ctx->h_blocks_bytes = (size_t)ntiles * block_floats * 4;
  }
  CK(cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(fj::DCounters) + 64, ctx->stream));

  const int grid = ctx->sm_count * env_int("FJGPU_BLOCKS_PER_SM", 4);
  bool has_plastic = false;
  for (const fjgpu_shader &sdr : ctx->shaders) has_plastic = has_plastic || sdr.kind == FJGPU_SHADER_PLASTIC;
  uint64_t launches = 0;
  size_t evn = 0;
  std::vector<cudaEvent_t> ev_extend, ev_shade, ev_resolve;     // (start, stop) pairs read after the final sync
  // ray-queue capacity: optimistic when the ray tree can branch, grown on overflow up to the worst case
  double factor = pl.factor0;
  CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  int per_batch = pl.tiles_per_batch;           // shrinks when an overflowed ray queue is grown (the budget stays)
  uint32_t n_batches = 0, n_regrows = 0; int32_t first_regrow = -1;
  for (int b0 = 0, nb = 0; b0 < ntiles; b0 += nb) {
    nb = std::min(per_batch, ntiles - b0);
    n_batches++;
    fj::RenderArgs a;
    memset(&a, 0, sizeof a);
    a.sc = ctx->sc; a.cam = pl.cam; a.fr = pl.fr;
    a.tiles = (const fj::DTile *)ctx->d_tiles.p + b0; a.ntiles = nb; a.wstride = pl.wstride;
    a.accum = (fj::Accum *)ctx->d_samples.p;
    a.counters = (fj::DCounters *)ctx->d_counters.p;
    a.work = (unsigned long long *)((char *)ctx->d_counters.p + sizeof(fj::DCounters));
    void *const counters_snapshot = (char *)ctx->d_counters.p + sizeof(fj::DCounters) + 64;
    if (mega) {
      CK(cudaMemsetAsync(a.work, 0, 8, ctx->stream));
      cudaEvent_t e0 = pool_event(ctx, &evn), e1 = pool_event(ctx, &evn);
      CK(cudaEventRecord(e0, ctx->stream));
      if (fp64_boxes) fj::k_render_samples<double><<<grid, 128, 0, ctx->stream>>>(a);
      else fj::k_render_samples<float><<<grid, 128, 0, ctx->stream>>>(a);
      launches++;
      CK(cudaGetLastError());
      CK(cudaEventRecord(e1, ctx->stream));
      ev_extend.push_back(e0); ev_extend.push_back(e1);
    } else {
      // what the earlier batches counted: restored if this batch has to be rendered again with a larger queue
      if (factor < pl.peak) CK(cudaMemcpyAsync(counters_snapshot, ctx->d_counters.p, sizeof(fj::DCounters), cudaMemcpyDeviceToDevice, ctx->stream));
      for (;;) {      // retried with a larger queue only if a branching ray tree overflowed the optimistic capacity
        const double want = (double)nb * pl.wstride * factor;
        if (want > 4.0e9) return fail(ctx, FJGPU_ERR_UNSUPPORTED, "ray tree too wide for one tile's ray queue");
        // chunked slot reservation (fj_kernels.cuh, QueueSink): every k_shade warp may leave up to FJ_QCHUNK + FJ_QRESERVE
        // reserved slots as fillers, on top of the records the ray tree can produce
        bool moving = !ctx->cam_motion.empty() || ctx->sc.time_tab != nullptr;      // RayRec::key carries the time-table entry then (moving
        for (auto &kv : ctx->inst_motion) moving = moving || !kv.second.empty();      // transforms, moving triangles): no sorting
        const int sort_bits = pl.waves > 1 && !moving ? std::min(7, std::max(0, env_int("FJGPU_SORT_BITS", 0))) : 0;
        // (the RAY_DEAD fillers of a chunked queue are not filed under a sort key: sorting takes one atomic per spawn)
        const bool chunked = !has_plastic && sort_bits == 0 && env_int("FJGPU_QUEUE_CHUNK", 1) != 0;
        // grid of the k_shade variant without plastic shaders: resident CTAs per SM (template parameter) x CTAs per slot
        const int shade_env = env_int("FJGPU_SHADE_MINBLOCKS", 5), shade_per = std::max(1, env_int("FJGPU_SHADE_CTAS", 2));
        a.shade_prefetch = env_int("FJGPU_SHADE_PREFETCH", 1);
        const int shade_smb = shade_env >= 8 ? 8 : (shade_env >= 6 ? 6 : 5);
        const int shade_ctas = ctx->sm_count * shade_smb * shade_per;
        const size_t shade_warps = (size_t)shade_ctas * 4;      // 128 threads per CTA
        const size_t capacity = (size_t)want + (chunked ? shade_warps * (FJ_QCHUNK + FJ_QRESERVE) : 0);
        if (capacity > 4000000000ull) return fail(ctx, FJGPU_ERR_UNSUPPORTED, "ray tree too wide for one tile's ray queue");
        a.chunked = chunked ? 1 : 0;
        // the two ray queues and the hit records live in ONE allocation at fixed relative offsets, so that the streams k_shade
        // reads and writes side by side keep the same relative placement whatever the scene allocated before them
        const size_t pad = (size_t)env_int("FJGPU_ARENA_PAD_KB", 0) << 10;
        const size_t qbytes = (capacity * sizeof(fj::RayRec) + 255) & ~(size_t)255, hbytes = (capacity * sizeof(fj::HitRec) + 255) & ~(size_t)255;
        if (int rc = dev_alloc(ctx, ctx->d_queue[0], 2 * (qbytes + pad) + hbytes)) return rc;
        a.queue[0] = (fj::RayRec *)ctx->d_queue[0].p; a.queue[1] = (fj::RayRec *)((char *)ctx->d_queue[0].p + qbytes + pad);
        a.hits = (fj::HitRec *)((char *)ctx->d_queue[0].p + 2 * (qbytes + pad));
        a.ctl = (fj::QueueCtl *)ctl_p; a.capacity = (uint32_t)capacity; a.cur = 0;
        a.hist = nullptr; a.perm = nullptr; a.sort_bits = sort_bits; a.sort_bins = 8u << (3 * sort_bits);
        if (sort_bits > 0) {
          if (int rc = dev_alloc(ctx, ctx->d_hist, ((size_t)a.sort_bins + 1) * 4)) return rc;
          if (int rc = dev_alloc(ctx, ctx->d_perm, capacity * 4)) return rc;
          a.hist = (unsigned int *)ctx->d_hist.p;
          for (int k = 0; k < 3; k++) {
            const float ext = ctx->scene_hi[k] - ctx->scene_lo[k];
            a.sort_lo[k] = ctx->scene_lo[k]; a.sort_scale[k] = ext > 0 ? (float)(1 << sort_bits) / ext : 0.f;
          }
        }
        CK(cudaMemsetAsync(ctl_p, 0, sizeof(fj::QueueCtl), ctx->stream));
        cudaEvent_t s0 = pool_event(ctx, &evn), s1 = pool_event(ctx, &evn);
        CK(cudaEventRecord(s0, ctx->stream));
        const unsigned long long total = (unsigned long long)nb * pl.wstride;
        fj::k_generate<<<(unsigned)std::min<unsigned long long>((total + 255) / 256, (unsigned long long)ctx->sm_count * 16), 256, 0, ctx->stream>>>(a);
        launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(s1, ctx->stream));
        ev_shade.push_back(s0); ev_shade.push_back(s1);
        for (int w = 0; w < pl.waves; w++) {
          // head of the current queue and the count of the next one start at zero
          if (env_int("FJGPU_CTL_MEMSET", 0)) {     // (the kernels reset the counters themselves: k_extend* the next queue's count, k_shade the head)
            CK(cudaMemsetAsync(ctl_p + offsetof(fj::QueueCtl, head), 0, 4, ctx->stream));
            CK(cudaMemsetAsync(ctl_p + offsetof(fj::QueueCtl, count) + 4 * (a.cur ^ 1), 0, 4, ctx->stream));
          }
          cudaEvent_t e0 = pool_event(ctx, &evn), e1 = pool_event(ctx, &evn), e2 = pool_event(ctx, &evn);
          if (a.hist) CK(cudaMemsetAsync(a.hist, 0, ((size_t)a.sort_bins + 1) * 4, ctx->stream));
          CK(cudaEventRecord(e0, ctx->stream));
          launch_extend(ctx, a, grid);
          CK(cudaGetLastError());
          CK(cudaEventRecord(e1, ctx->stream));
          if (has_plastic) fj::k_shade<float, true><<<ctx->sm_count * 8, 128, 0, ctx->stream>>>(a);
          else {
            if (shade_smb == 8) fj::k_shade<float, false, 8><<<shade_ctas, 128, 0, ctx->stream>>>(a);
            else if (shade_smb == 6) fj::k_shade<float, false, 6><<<shade_ctas, 128, 0, ctx->stream>>>(a);
            else fj::k_shade<float, false, 5><<<shade_ctas, 128, 0, ctx->stream>>>(a);
          }
          CK(cudaGetLastError());
          launches += 2;
          if (a.hist && w + 1 < pl.waves) {       // order the next queue for the next extend
            fj::RenderArgs n = a; n.cur = a.cur ^ 1; n.perm = (unsigned int *)ctx->d_perm.p;
            fj::k_sort_scan<<<1, 1024, 0, ctx->stream>>>(a.hist, a.sort_bins);
            fj::k_sort_scatter<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(n);
            CK(cudaGetLastError());
            launches += 2;
            a.perm = n.perm;
          }
          CK(cudaEventRecord(e2, ctx->stream));
          ev_extend.push_back(e0); ev_extend.push_back(e1); ev_shade.push_back(e1); ev_shade.push_back(e2);
          a.cur ^= 1;
        }
        if (factor >= pl.peak) break;           // cannot overflow
        fj::QueueCtl hctl;
        CK(cudaMemcpyAsync(&hctl, ctl_p, sizeof hctl, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (!hctl.overflow) break;
        factor = std::min(pl.peak, factor * 4.0);
        if (first_regrow < 0) first_regrow = (int32_t)n_batches - 1;
        n_regrows++;
        {
          const size_t per_slot = sizeof(fj::Accum) + (size_t)std::ceil(factor * (2 * sizeof(fj::RayRec) + sizeof(fj::HitRec)));
          per_batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)per_batch, pl.budget / ((size_t)pl.wstride * per_slot)));
          nb = std::min(per_batch, ntiles - b0); a.ntiles = nb;
        }
        // this batch is rendered again from scratch with the larger queue (later batches keep it): discard what the
        // overflowed attempt counted, keep what the batches before it did
        CK(cudaMemcpyAsync(ctx->d_counters.p, counters_snapshot, sizeof(fj::DCounters), cudaMemcpyDeviceToDevice, ctx->stream));
      }
    }
    if (mode != OUT_SAMPLES_ONLY) {
      cudaEvent_t e0 = pool_event(ctx, &evn), e1 = pool_event(ctx, &evn);
      CK(cudaEventRecord(e0, ctx->stream));
      fj::k_resolve_tiles<<<nb * FJ_RESOLVE_SPLIT, 256, 0, ctx->stream>>>(pl.fr, a.tiles, pl.wstride, a.accum, blocks + (size_t)b0 * pl.bw * pl.bh, pl.bw, pl.bh);
      launches++;
      CK(cudaGetLastError());
      CK(cudaEventRecord(e1, ctx->stream));
      ev_resolve.push_back(e0); ev_resolve.push_back(e1);
    }
  }
  if (mode == OUT_RESIDENT) {
    fj::k_blocks_to_frame<<<ntiles, 256, 0, ctx->stream>>>((const fj::DTile *)ctx->d_tiles.p, ntiles, blocks, pl.bw, pl.bh, (float4 *)ctx->d_frame.p, p->xres);
    launches++;
    CK(cudaGetLastError());
  }
  if (mode == OUT_HOST) CK(cudaMemcpyAsync(ctx->h_blocks, blocks, (size_t)ntiles * block_floats * 4, cudaMemcpyDeviceToHost, ctx->stream));
  fj::DCounters hc; memset(&hc, 0, sizeof hc);
  if (stats) CK(cudaMemcpyAsync(&hc, ctx->d_counters.p, sizeof hc, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaEventRecord(ctx->ev[3], ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (mode == OUT_HOST) {
    if (!rgba_frame) return fail(ctx, FJGPU_ERR_INVALID, "null framebuffer");
    for (int i = 0; i < ntiles; i++) {
      const fjgpu_tile &t = tiles[i];
      const int w = t.xmax - t.xmin;
      for (int y = t.ymin; y < t.ymax; y++)
        memcpy(rgba_frame + ((size_t)y * p->xres + t.xmin) * 4, ctx->h_blocks + (size_t)i * block_floats + (size_t)(y - t.ymin) * pl.bw * 4, (size_t)w * 16);
    }
  }
  if (stats) {
    float tot = 0; cudaEventElapsedTime(&tot, ctx->ev[0], ctx->ev[3]);
    stats->rays_camera = hc.rays[0]; stats->rays_shadow = hc.rays[1]; stats->rays_diffuse = hc.rays[2];
    stats->rays_reflect = hc.rays[3]; stats->rays_refract = hc.rays[4]; stats->camera_samples = hc.samples;
    stats->rays_hit = hc.hits; stats->hit_mesh_levels = hc.levels; stats->node_steps = hc.node_steps; stats->tri_tests = hc.tri_tests;
    stats->leaf_phases = hc.leaf_phases; stats->leaf_rounds = hc.leaf_rounds;
    float ms_trace = 0, ms_shade = 0, ms_resolve = 0, t = 0;
    for (size_t i = 0; i + 1 < ev_extend.size(); i += 2) { cudaEventElapsedTime(&t, ev_extend[i], ev_extend[i + 1]); ms_trace += t; }
    for (size_t i = 0; i + 1 < ev_shade.size(); i += 2) { cudaEventElapsedTime(&t, ev_shade[i], ev_shade[i + 1]); ms_shade += t; }
    for (size_t i = 0; i + 1 < ev_resolve.size(); i += 2) { cudaEventElapsedTime(&t, ev_resolve[i], ev_resolve[i + 1]); ms_resolve += t; }
    stats->kernel_launches = launches; stats->trace_launches = ev_extend.size() / 2;
    stats->batches = n_batches; stats->queue_regrows = n_regrows; stats->first_regrow_batch = first_regrow;
    stats->ms_trace = ms_trace; stats->ms_shade = ms_shade; stats->ms_resolve = ms_resolve; stats->ms_total = tot;
  }
  return 0;
}

// ---- single-process multi-GPU: NCCL, bound at first use with dlopen so that libfjgpu itself has no link-time dependency on it
// (a torch process brings its own NCCL; this path is for `fjscene` and other C / C++ hosts that drive several contexts)
struct Nccl {
  void *lib = nullptr;
  int (*CommInitAll)(void **, int, const int *) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::vector<int> devices; std::vector<void *> comms;      // the communicators of the last device list
};
Nccl g_nccl;

int multi_all_gather(fjgpu_context *const *ctxs, int nranks, size_t send_bytes) {
  fjgpu_context *ctx = ctxs[0];
  Nccl &N = g_nccl;
  if (!N.lib) {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) if ((N.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!N.lib) return fail(ctx, FJGPU_ERR_UNSUPPORTED, "fjgpu_render_frame_multi: libnccl.so.2 cannot be loaded");
    N.CommInitAll = (int (*)(void **, int, const int *))dlsym(N.lib, "ncclCommInitAll");
    N.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(N.lib, "ncclAllGather");
    N.GroupStart = (int (*)())dlsym(N.lib, "ncclGroupStart");
    N.GroupEnd = (int (*)())dlsym(N.lib, "ncclGroupEnd");
    N.GetErrorString = (const char *(*)(int))dlsym(N.lib, "ncclGetErrorString");
    if (!N.CommInitAll || !N.AllGather || !N.GroupStart || !N.GroupEnd) { N.lib = nullptr; return fail(ctx, FJGPU_ERR_UNSUPPORTED, "fjgpu_render_frame_multi: NCCL symbols missing"); }
  }
  auto nerr = [&](const char *what, int rc) { return fail(ctx, FJGPU_ERR_CUDA, std::string(what) + ": " + (N.GetErrorString ? N.GetErrorString(rc) : "NCCL error")); };
  std::vector<int> devs(nranks);
  for (int r = 0; r < nranks; r++) devs[r] = ctxs[r]->device;
  if (devs != N.devices) {
    N.comms.assign(nranks, nullptr);
    if (int rc = N.CommInitAll(N.comms.data(), nranks, devs.data())) return nerr("ncclCommInitAll", rc);
    N.devices = devs;
  }
  if (int rc = N.GroupStart()) return nerr("ncclGroupStart", rc);
  for (int r = 0; r < nranks; r++) {
    fjgpu_context *c = ctxs[r];
    cudaSetDevice(c->device);
    if (int rc = N.AllGather(c->d_multi_send.p, c->d_multi_recv.p, send_bytes, /* ncclChar */ 0, N.comms[r], c->stream)) { N.GroupEnd(); return nerr("ncclAllGather", rc); }
  }
  if (int rc = N.GroupEnd()) return nerr("ncclGroupEnd", rc);
  for (int r = 0; r < nranks; r++) {
    cudaSetDevice(ctxs[r]->device);
    cudaError_t e = cudaStreamSynchronize(ctxs[r]->stream);
    if (e != cudaSuccess) return fail(ctx, FJGPU_ERR_CUDA, std::string("all-gather: ") + cudaGetErrorString(e));
  }
  cudaSetDevice(ctx->device);
  return 0;
}

}  // namespace

// ================================================================================== extern "C"
extern "C" {

int fjgpu_api_version(void) { return FJGPU_API_VERSION; }

const char *fjgpu_last_error(const fjgpu_context *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int fjgpu_create(int device_ordinal, fjgpu_context **out_ctx) {
  if (!out_ctx) return fail(nullptr, FJGPU_ERR_INVALID, "null out_ctx");
  *out_ctx = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) return fail(nullptr, FJGPU_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libfjgpu has no CPU fallback)");
  if (device_ordinal < 0 || device_ordinal >= n) return fail(nullptr, FJGPU_ERR_INVALID, "device ordinal out of range");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess || prop.major < 10)
    return fail(nullptr, FJGPU_ERR_NO_DEVICE, "device is not sm_100 class (this library carries sm_100a code only)");
  fjgpu_context *ctx = new fjgpu_context();
  ctx->device = device_ordinal; ctx->sm_count = prop.multiProcessorCount;
  memset(&ctx->sc, 0, sizeof ctx->sc); memset(&ctx->cam, 0, sizeof ctx->cam);
  ctx->group_off.assign(1, 0);
  if (cudaSetDevice(device_ordinal) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx; return fail(nullptr, FJGPU_ERR_CUDA, "cannot create a stream");
  }
  for (int i = 0; i < 4; i++) cudaEventCreate(&ctx->ev[i]);
  cudaDeviceSetLimit(cudaLimitStackSize, 8192);
  *out_ctx = ctx;
  return FJGPU_OK;
}

void fjgpu_destroy(fjgpu_context *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &kv : ctx->meshes) kv.second.release();
  for (auto &b : ctx->d_group_nodes) b.release();
  for (auto &b : ctx->d_group_nodes4) b.release();
  for (auto &b : ctx->d_group_nodesq) b.release();
  for (auto &b : ctx->d_group_order) b.release();
  for (auto &b : ctx->d_group_irec) b.release();
  for (auto &b : ctx->d_dome) b.release();
  for (auto &b : ctx->d_tex_tiles) b.release();
  ctx->d_textures.release();
  for (auto &kv : ctx->d_inst_motion) kv.second.release();
  ctx->d_cam_motion.release();
  DevBuf *all[] = {&ctx->d_meshes, &ctx->d_inst, &ctx->d_groups, &ctx->d_shaders, &ctx->d_lights, &ctx->d_samples,
                   &ctx->d_tiles, &ctx->d_blocks, &ctx->d_jitter, &ctx->d_timetab, &ctx->d_multi_send, &ctx->d_multi_recv, &ctx->d_counters, &ctx->d_frame, &ctx->d_queue[0], &ctx->d_queue[1], &ctx->d_hits, &ctx->d_ctl, &ctx->d_hist, &ctx->d_perm};
  for (cudaEvent_t e : ctx->evpool) cudaEventDestroy(e);
  for (DevBuf *b : all) b->release();
  if (ctx->h_blocks) cudaFreeHost(ctx->h_blocks);
  for (int i = 0; i < 4; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

static int mesh_upload_impl(fjgpu_context *ctx, int32_t mesh_id, const double *P, const double *N, int32_t nverts,
                            const int32_t *idx3, const int32_t *face_group_id, int32_t nfaces, const double *vel);

int fjgpu_mesh_upload(fjgpu_context *ctx, int32_t mesh_id, const double *P, const double *N, int32_t nverts,
                      const int32_t *idx3, const int32_t *face_group_id, int32_t nfaces) {
  return fjgpu_mesh_upload_velocity(ctx, mesh_id, P, N, nverts, idx3, face_group_id, nfaces, nullptr);
}

int fjgpu_shutter_set(fjgpu_context *ctx, double time_start, double time_end) {
  if (!ctx) return fail(nullptr, FJGPU_ERR_INVALID, "null context");
  if (!(time_start <= time_end)) return fail(ctx, FJGPU_ERR_INVALID, "shutter: time_start must not exceed time_end");      // assert of src/fj_renderer.cc:528
  ctx->shutter[0] = time_start; ctx->shutter[1] = time_end;
  return FJGPU_OK;
}

int fjgpu_mesh_upload_velocity(fjgpu_context *ctx, int32_t mesh_id, const double *P, const double *N, int32_t nverts,
                               const int32_t *idx3, const int32_t *face_group_id, int32_t nfaces, const double *vel) {
  if (!ctx) return fail(nullptr, FJGPU_ERR_INVALID, "null context");
  if (nverts < 0 || nfaces < 0 || (nverts > 0 && !P) || (nfaces > 0 && !idx3)) return fail(ctx, FJGPU_ERR_INVALID, "bad mesh arrays");
  if (nfaces >= (1 << 28)) return fail(ctx, FJGPU_ERR_UNSUPPORTED, "more than 2^28 faces in one mesh");
  for (size_t i = 0; i < 3 * (size_t)nfaces; i++)
    if (idx3[i] < 0 || idx3[i] >= nverts) return fail(ctx, FJGPU_ERR_INVALID, "face index out of range");
  const int rc = mesh_upload_impl(ctx, mesh_id, P, N, nverts, idx3, face_group_id, nfaces, vel);
  if (rc != FJGPU_OK) {      // no half-built record stays behind: a later commit would hand freed or null arrays to the kernels
    auto it = ctx->meshes.find(mesh_id);
    if (it != ctx->meshes.end()) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); it->second.release(); ctx->meshes.erase(it); }
    ctx->dirty = true;
  }
  return rc;
}

static int mesh_upload_impl(fjgpu_context *ctx, int32_t mesh_id, const double *P, const double *N, int32_t nverts,
                            const int32_t *idx3, const int32_t *face_group_id, int32_t nfaces, const double *vel) {
  CK(cudaSetDevice(ctx->device));
  const auto t0 = std::chrono::steady_clock::now();
  MeshRec &m = ctx->meshes[mesh_id];
  m.release();
  m.nfaces = nfaces; m.nverts = nverts;
  m.hostP.assign(P, P + 3 * (size_t)nverts);
  m.hostVel.clear(); m.top4 = 0;
  if (vel && nverts > 0) m.hostVel.assign(vel, vel + 3 * (size_t)nverts);
  bool f32ok = true;
  for (size_t i = 0; i < 3 * (size_t)nverts && f32ok; i++) f32ok = ((double)(float)P[i] == P[i]);
  if (env_int("FJGPU_FORCE_TRI64", 0)) f32ok = false;
  // ---- device build (FJGPU_BUILD=device): linear BVH in HBM, fj_build.cu.  The host builder (binned SAH, below) stays the
  // default because its trees are walked faster; the device builder is for time-to-first-pixel on large meshes.
  {
    const char *bm = getenv("FJGPU_BUILD");
    if (bm && std::string(bm) == "device" && nfaces >= env_int("FJGPU_BUILD_DEVICE_MIN", 1024) && !vel) {      // (moving triangles: host builder)
      DevBuf dP;
      if (int rc = dev_upload(ctx, dP, P, (size_t)nverts * 24)) return rc;
      if (int rc = dev_upload(ctx, m.idx, idx3, (size_t)nfaces * 12, true)) { dP.release(); return rc; }
      FjDeviceBuild db; std::string berr;
      const int brc = fj_device_build(ctx->stream, (const double *)dP.p, nverts, (const int32_t *)m.idx.p, nfaces, env_int("FJGPU_MAX_LEAF", 4),
                                      (float)env_int("FJGPU_LEAF_COST_X10", 15) / 10.f, env_int("FJGPU_FORCE_TRI64", 0) != 0, &db, &berr);
      dP.release();
      if (brc == 0 && (db.max_depth + 8 > FJ_STACK || 3 * db.max_depth4 + 16 > FJ_STACK4)) {     // too deep for the traversal stacks: host build
        for (void *q : {db.nodes, db.nodes4, db.nodesq, db.tri}) cudaFree(q);
      } else if (brc == 0) {
        m.nodes.p = db.nodes; m.nodes.bytes = db.nodes_bytes; m.nodes4.p = db.nodes4; m.nodes4.bytes = db.nodes4_bytes;
        m.nodesq.p = db.nodesq; m.nodesq.bytes = db.nodesq_bytes;
        m.tri.p = db.tri; m.tri.bytes = db.tri_bytes;
        for (int a = 0; a < 3; a++) { m.bmin[a] = db.bmin[a]; m.bmax[a] = db.bmax[a]; }
        m.nnodes = db.nnodes; m.max_depth = db.max_depth; m.max_depth4 = db.max_depth4; m.nnodes4 = db.nnodes4; m.stack_need4 = db.stack_need4;
        memset(&m.d, 0, sizeof m.d);
        m.d.nodes = (const float4 *)m.nodes.p; m.d.nodes4 = (const float4 *)m.nodes4.p;
        if (db.quant_ok) { m.d.nodesq = (const float4 *)m.nodesq.p; m.d.bmagq = db.bmagq; }
        if (db.tri64) m.d.tri64 = (const double *)m.tri.p; else m.d.tri32 = (const float4 *)m.tri.p;
        if (N) { if (int rc = dev_upload(ctx, m.N, N, (size_t)nverts * 24, true)) return rc; m.d.N = (const double *)m.N.p; }
        m.d.idx = (const int32_t *)m.idx.p;
        if (face_group_id) { if (int rc = dev_upload(ctx, m.group, face_group_id, (size_t)nfaces * 4, true)) return rc; m.d.group = (const int32_t *)m.group.p; }
        { int l = 0; while ((1ll << l) < (long long)std::max(nfaces, 1)) l++; m.d.log2_tris = l; }
        m.d.bmag = db.bmag;
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->dirty = true;
        ctx->build_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        ctx->device_build_seconds += db.seconds;
        return FJGPU_OK;
      }
      // any failure falls through to the host builder
    }
  }
  std::vector<fjb::Aabb> boxes(nfaces);
  for (int a = 0; a < 3; a++) { m.bmin[a] = 1.7976931348623157e308; m.bmax[a] = -1.7976931348623157e308; }
  for (int f = 0; f < nfaces; f++) {
    double lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
      const double x = P[3 * (size_t)idx3[3 * f] + a], y = P[3 * (size_t)idx3[3 * f + 1] + a], z = P[3 * (size_t)idx3[3 * f + 2] + a];
      lo[a] = std::min(x, std::min(y, z)); hi[a] = std::max(x, std::max(y, z));
      if (vel) {      // Mesh::get_primitive_bounds, src/fj_mesh.cc:420-439: the bounds also hold the vertices at time 1 (`P + velocity`);
                      // a vertex moves on a straight line, so every time in [0, 1] — the shutter the reference bounds for — is inside
        for (int v = 0; v < 3; v++) { const size_t vi = 3 * (size_t)idx3[3 * f + v] + a; const double q = P[vi] + vel[vi]; lo[a] = std::min(lo[a], q); hi[a] = std::max(hi[a], q); }
      }
      m.bmin[a] = std::min(m.bmin[a], lo[a]); m.bmax[a] = std::max(m.bmax[a], hi[a]);
    }
    pad_box(lo, hi, &boxes[f]);
  }
  fjb::BuildResult br;
  fjb::build_bvh(boxes.data(), nfaces, env_int("FJGPU_MAX_LEAF", 4), (float)env_int("FJGPU_LEAF_COST_X10", 15) / 10.f, 0, &br);
  m.nnodes = (int32_t)br.nodes.size(); m.max_depth = br.max_depth;
  if (br.max_depth + 8 > FJ_STACK || 3 * br.max_depth4 + 16 > FJ_STACK4) return fail(ctx, FJGPU_ERR_UNSUPPORTED, "BVH deeper than the traversal stack");
  if (int rc = dev_upload(ctx, m.nodes, br.nodes.data(), br.nodes.size() * sizeof(fjb::Node64), true)) return rc;
  if (int rc = dev_upload(ctx, m.nodes4, br.nodes4.data(), br.nodes4.size() * sizeof(fjb::Node128), true)) return rc;
  m.max_depth4 = br.max_depth4; m.nnodes4 = (int32_t)br.nodes4.size(); m.stack_need4 = br.stack_need4; m.top4 = br.top_count4;
  memset(&m.d, 0, sizeof m.d);
  m.d.nodes = (const float4 *)m.nodes.p;
  m.d.nodes4 = (const float4 *)m.nodes4.p;
  {
    std::vector<fjb::NodeQ64> nq(br.nodes4.size());
    float bq = 0;
    if (fjb::quantize_nodes(br.nodes4.data(), br.nodes4.size(), nq.data(), &bq)) {
      if (int rc = dev_upload(ctx, m.nodesq, nq.data(), nq.size() * sizeof(fjb::NodeQ64), true)) return rc;
      m.d.nodesq = (const float4 *)m.nodesq.p; m.d.bmagq = bq;
    }
  }
  const size_t nt = br.order.size();
  if (vel) {      // moving triangles: v0 v1 v2, face id, velocity0 velocity1 velocity2, pad (20 doubles, fj_device.cuh DMesh::tri64v)
    std::vector<double> tri(std::max<size_t>(nt, 1) * 20, 0.);
    for (size_t k = 0; k < nt; k++) {
      const int f = br.order[k];
      for (int v = 0; v < 3; v++) for (int a = 0; a < 3; a++) {
        tri[20 * k + 3 * v + a] = P[3 * (size_t)idx3[3 * f + v] + a];
        tri[20 * k + 10 + 3 * v + a] = vel[3 * (size_t)idx3[3 * f + v] + a];
      }
      const long long fl = f; memcpy(&tri[20 * k + 9], &fl, 8);
    }
    if (int rc = dev_upload(ctx, m.tri, tri.data(), tri.size() * 8, true)) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    m.d.tri64v = (const double *)m.tri.p;
  } else if (f32ok) {
    std::vector<float> tri(std::max<size_t>(nt, 1) * 12, 0.f);
    for (size_t k = 0; k < nt; k++) {
      const int f = br.order[k];
      for (int v = 0; v < 3; v++) {
        for (int a = 0; a < 3; a++) tri[12 * k + 4 * v + a] = (float)P[3 * (size_t)idx3[3 * f + v] + a];
      }
      memcpy(&tri[12 * k + 3], &f, 4);
    }
    if (int rc = dev_upload(ctx, m.tri, tri.data(), tri.size() * 4, true)) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    m.d.tri32 = (const float4 *)m.tri.p;
  } else {
    std::vector<double> tri(std::max<size_t>(nt, 1) * 10, 0.);
    for (size_t k = 0; k < nt; k++) {
      const int f = br.order[k];
      for (int v = 0; v < 3; v++) for (int a = 0; a < 3; a++) tri[10 * k + 3 * v + a] = P[3 * (size_t)idx3[3 * f + v] + a];
      const long long fl = f; memcpy(&tri[10 * k + 9], &fl, 8);
    }
    if (int rc = dev_upload(ctx, m.tri, tri.data(), tri.size() * 8, true)) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    m.d.tri64 = (const double *)m.tri.p;
  }
  if (N) { if (int rc = dev_upload(ctx, m.N, N, (size_t)nverts * 24, true)) return rc; m.d.N = (const double *)m.N.p; }
  if (int rc = dev_upload(ctx, m.idx, idx3, (size_t)nfaces * 12, true)) return rc;
  m.d.idx = (const int32_t *)m.idx.p;
  if (face_group_id) { if (int rc = dev_upload(ctx, m.group, face_group_id, (size_t)nfaces * 4, true)) return rc; m.d.group = (const int32_t *)m.group.p; }
  m.d.top_count = br.top_count;
  { int l = 0; while ((1ll << l) < (long long)std::max(nfaces, 1)) l++; m.d.log2_tris = l; }
  { double b = 0; for (int a = 0; a < 3; a++) b = std::max(b, std::max(std::fabs((double)br.bounds.lo[a]), std::fabs((double)br.bounds.hi[a]))); m.d.bmag = nfaces > 0 ? fjb::round_up(b) : 0.f; }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->dirty = true;
  ctx->build_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return FJGPU_OK;
}

int fjgpu_instances_set(fjgpu_context *ctx, int32_t n, const fjgpu_instance *inst) {
  if (!ctx || n < 0 || (n > 0 && !inst)) return fail(ctx, FJGPU_ERR_INVALID, "bad instance array");
  if ((size_t)n == ctx->inst.size() && (n == 0 || memcmp(ctx->inst.data(), inst, (size_t)n * sizeof(fjgpu_instance)) == 0)) return FJGPU_OK;   // unchanged
  ctx->inst.assign(inst, inst + n); ctx->dirty = true;
  for (auto it = ctx->inst_motion.begin(); it != ctx->inst_motion.end();) { if (it->first >= n) it = ctx->inst_motion.erase(it); else ++it; }
  return FJGPU_OK;
}

namespace {
// count_samples_in_region of the largest tile, src/fj_fixed_grid_sampler.cc:119-136
long max_tile_samples(const fjgpu_render_params *p, const fjgpu_tile *tiles, int ntiles) {
  int mx = (int)std::ceil(((p->xfwidth - 1) * p->xrate) * .5), my = (int)std::ceil(((p->yfwidth - 1) * p->yrate) * .5);
  if (mx < 0) mx = 0; if (my < 0) my = 0;
  int tw = 1, th = 1;
  for (int i = 0; i < ntiles; i++) { tw = std::max(tw, tiles[i].xmax - tiles[i].xmin); th = std::max(th, tiles[i].ymax - tiles[i].ymin); }
  return ((long)p->xrate * tw + 2 * mx) * ((long)p->yrate * th + 2 * my);
}
}  // namespace

int fjgpu_time_table(const fjgpu_render_params *p, const fjgpu_tile *tiles, int32_t ntiles, double t0, double t1, double *times, int32_t cap) {
  if (!p || (!tiles && ntiles > 0) || ntiles < 0 || p->xrate <= 0 || p->yrate <= 0 || !(p->xfwidth > 0) || !(p->yfwidth > 0) || (cap > 0 && !times)) return -1;
  const long n = max_tile_samples(p, tiles, ntiles);
  if (n > (1l << 30)) return -1;
  uint32_t s[4] = {123456789u, 362436069u, 521288629u, 88675123u};       // XorShift, src/fj_random.cc:10-43
  for (long i = 0; i < n && i < cap; i++) {
    const uint32_t t = s[0] ^ (s[0] << 11);
    s[0] = s[1]; s[1] = s[2]; s[2] = s[3];
    s[3] = (s[3] ^ (s[3] >> 19)) ^ (t ^ (t >> 8));
    const double rnd = (double)s[3] / 4294967295.0;                      // NextFloat01
    // Fit(rnd, 0, 1, t0, t1), src/fj_numeric.h:84-93
    times[i] = rnd <= 0 ? t0 : (rnd >= 1 ? t1 : t0 + (t1 - t0) * ((rnd - 0) / (1. - 0)));
  }
  return (int32_t)n;
}

int fjgpu_instance_motion_set(fjgpu_context *ctx, int32_t instance, int32_t ntimes, const double *fwd16, const double *inv16) {
  if (!ctx || instance < 0 || instance >= (int32_t)ctx->inst.size() || ntimes < 0 || (ntimes > 0 && (!fwd16 || !inv16)))
    return fail(ctx, FJGPU_ERR_INVALID, "bad instance motion table (call fjgpu_instances_set first)");
  if (ntimes == 0 && ctx->inst_motion.find(instance) == ctx->inst_motion.end()) return FJGPU_OK;      // was static already
  std::vector<double> tab((size_t)ntimes * 24);
  for (int32_t t = 0; t < ntimes; t++) { rows12(inv16 + 16 * (size_t)t, &tab[24 * (size_t)t]); rows12(fwd16 + 16 * (size_t)t, &tab[24 * (size_t)t + 12]); }
  std::vector<double> &cur = ctx->inst_motion[instance];
  if (cur.size() == tab.size() && !tab.empty() && memcmp(cur.data(), tab.data(), tab.size() * sizeof(double)) == 0) return FJGPU_OK;   // unchanged
  cur.swap(tab);
  ctx->inst_motion_dirty[instance] = true; ctx->dirty = true;
  return FJGPU_OK;
}

int fjgpu_camera_motion_set(fjgpu_context *ctx, int32_t ntimes, const double *fwd16) {
  if (!ctx || ntimes < 0 || (ntimes > 0 && !fwd16)) return fail(ctx, FJGPU_ERR_INVALID, "bad camera motion table");
  std::vector<double> tab((size_t)ntimes * 12);
  for (int32_t t = 0; t < ntimes; t++) rows12(fwd16 + 16 * (size_t)t, &tab[12 * (size_t)t]);
  if (tab == ctx->cam_motion) return FJGPU_OK;
  ctx->cam_motion.swap(tab);
  ctx->cam_motion_dirty = true;
  return FJGPU_OK;
}

int fjgpu_groups_set(fjgpu_context *ctx, int32_t ngroups, const int32_t *group_offsets, const int32_t *instance_ids) {
  if (!ctx || ngroups < 0 || (ngroups > 0 && !group_offsets)) return fail(ctx, FJGPU_ERR_INVALID, "bad group arrays");
  if (ngroups > 0 && group_offsets[0] == 0 && (size_t)ngroups + 1 == ctx->group_off.size() &&
      memcmp(ctx->group_off.data(), group_offsets, ((size_t)ngroups + 1) * 4) == 0 && (size_t)group_offsets[ngroups] == ctx->group_ids.size() &&
      (ctx->group_ids.empty() || (instance_ids && memcmp(ctx->group_ids.data(), instance_ids, ctx->group_ids.size() * 4) == 0))) return FJGPU_OK;   // unchanged
  ctx->group_off.assign(1, 0); ctx->group_ids.clear();
  if (ngroups > 0) {
    if (group_offsets[0] != 0) return fail(ctx, FJGPU_ERR_INVALID, "group_offsets[0] must be 0");
    for (int g = 0; g < ngroups; g++) if (group_offsets[g + 1] < group_offsets[g]) return fail(ctx, FJGPU_ERR_INVALID, "group_offsets must be non-decreasing");
    ctx->group_off.assign(group_offsets, group_offsets + ngroups + 1);
    if (group_offsets[ngroups] > 0 && !instance_ids) return fail(ctx, FJGPU_ERR_INVALID, "null instance_ids");
    ctx->group_ids.assign(instance_ids, instance_ids + group_offsets[ngroups]);
  }
  ctx->dirty = true;
  return FJGPU_OK;
}

int fjgpu_shaders_set(fjgpu_context *ctx, int32_t n, const fjgpu_shader *shaders) {
  if (!ctx || n < 0 || (n > 0 && !shaders)) return fail(ctx, FJGPU_ERR_INVALID, "bad shader array");
  for (int i = 0; i < n; i++) if (shaders[i].kind < FJGPU_SHADER_NONE || shaders[i].kind > FJGPU_SHADER_GLASS)
    return fail(ctx, FJGPU_ERR_UNSUPPORTED, "shader kind has no device implementation");
  if ((size_t)n == ctx->shaders.size() && (n == 0 || memcmp(ctx->shaders.data(), shaders, (size_t)n * sizeof(fjgpu_shader)) == 0)) return FJGPU_OK;   // unchanged
  ctx->shaders.assign(shaders, shaders + n); ctx->dirty = true;
  return FJGPU_OK;
}

int fjgpu_lights_set(fjgpu_context *ctx, int32_t n, const fjgpu_light *lights) {
  if (!ctx || n < 0 || (n > 0 && !lights)) return fail(ctx, FJGPU_ERR_INVALID, "bad light array");
  if ((size_t)n == ctx->lights.size()) {        // unchanged (the dome tables compared by content, the caller's pointers ignored)?
    bool same = true;
    for (int i = 0; i < n && same; i++) {
      fjgpu_light a = lights[i], b = ctx->lights[i];
      const size_t nd = a.kind == FJGPU_LIGHT_DOME && a.dome_sample_count > 0 ? (size_t)a.dome_sample_count : 0;
      same = a.dome_sample_count == b.dome_sample_count && ctx->dome_dirs[i].size() == 3 * nd && (nd == 0 || (a.dome_dirs && a.dome_colors &&
             memcmp(ctx->dome_dirs[i].data(), a.dome_dirs, 24 * nd) == 0 && memcmp(ctx->dome_cols[i].data(), a.dome_colors, 12 * nd) == 0));
      a.dome_dirs = b.dome_dirs = nullptr; a.dome_colors = b.dome_colors = nullptr;
      same = same && memcmp(&a, &b, sizeof a) == 0;
    }
    if (same) return FJGPU_OK;
  }
  ctx->lights.assign(lights, lights + n);
  ctx->dome_dirs.assign(n, std::vector<double>()); ctx->dome_cols.assign(n, std::vector<float>());
  for (int i = 0; i < n; i++) {
    fjgpu_light &l = ctx->lights[i];
    if (l.kind < FJGPU_LIGHT_POINT || l.kind > FJGPU_LIGHT_DOME) return fail(ctx, FJGPU_ERR_UNSUPPORTED, "unknown light kind");
    if (l.kind == FJGPU_LIGHT_DOME && l.dome_sample_count > 0) {
      if (!l.dome_dirs || !l.dome_colors) return fail(ctx, FJGPU_ERR_INVALID, "dome light without sample tables");
      ctx->dome_dirs[i].assign(l.dome_dirs, l.dome_dirs + 3 * (size_t)l.dome_sample_count);
      ctx->dome_cols[i].assign(l.dome_colors, l.dome_colors + 3 * (size_t)l.dome_sample_count);
    }
    l.dome_dirs = nullptr; l.dome_colors = nullptr;    // the caller's buffers are not kept
  }
  ctx->dirty = true;
  return FJGPU_OK;
}

int fjgpu_textures_set(fjgpu_context *ctx, int32_t n, const fjgpu_texture *textures) {
  if (!ctx || n < 0 || (n > 0 && !textures)) return fail(ctx, FJGPU_ERR_INVALID, "bad texture array");
  CK(cudaSetDevice(ctx->device));
  for (int i = 0; i < n; i++) {
    const fjgpu_texture &t = textures[i];
    if (t.width <= 0 || t.height <= 0 || t.tilesize <= 0 || !t.tiles || (t.nchannels != 1 && t.nchannels != 3 && t.nchannels != 4))
      return fail(ctx, FJGPU_ERR_INVALID, "bad texture description");
    if (t.tilesize < 64 || t.width / t.tilesize < 1 || t.height / t.tilesize < 1)
      return fail(ctx, FJGPU_ERR_UNSUPPORTED, "texture tiles must be at least 64 texels wide (TextureCache::LookupTexture assumes 64) and the image at least one tile");
  }
  for (auto &b : ctx->d_tex_tiles) b.release();
  ctx->d_tex_tiles.assign(n, DevBuf());
  ctx->textures.assign(textures, textures + n);
  for (int i = 0; i < n; i++) {
    const fjgpu_texture &t = textures[i];
    const size_t floats = (size_t)(t.width / t.tilesize) * (t.height / t.tilesize) * t.tilesize * t.tilesize * t.nchannels;
    if (int rc = dev_upload(ctx, ctx->d_tex_tiles[i], t.tiles, floats * sizeof(float), true)) return rc;
    ctx->textures[i].tiles = nullptr;      // the caller's buffer is not kept
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->dirty = true;
  return FJGPU_OK;
}

int fjgpu_mesh_set_uv(fjgpu_context *ctx, int32_t mesh_id, const float *uv2, int32_t nverts) {
  if (!ctx) return fail(nullptr, FJGPU_ERR_INVALID, "null context");
  auto it = ctx->meshes.find(mesh_id);
  if (it == ctx->meshes.end()) return fail(ctx, FJGPU_ERR_INVALID, "fjgpu_mesh_set_uv: unknown mesh_id");
  MeshRec &m = it->second;
  CK(cudaSetDevice(ctx->device));
  if (!uv2) { m.uv.release(); m.P.release(); m.vel.release(); m.d.uv = nullptr; m.d.P = nullptr; m.d.vel = nullptr; ctx->dirty = true; return FJGPU_OK; }
  if (nverts != m.nverts) return fail(ctx, FJGPU_ERR_INVALID, "fjgpu_mesh_set_uv: vertex count differs from the uploaded mesh");
  if (int rc = dev_upload(ctx, m.uv, uv2, (size_t)nverts * 8, true)) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  if (int rc = dev_upload(ctx, m.P, m.hostP.data(), m.hostP.size() * 8, true)) return rc;       // dPdu / dPdv of bump maps need the vertices by index
  CK(cudaStreamSynchronize(ctx->stream));
  m.d.uv = (const float *)m.uv.p; m.d.P = (const double *)m.P.p;
  if (!m.hostVel.empty()) {
    if (int rc = dev_upload(ctx, m.vel, m.hostVel.data(), m.hostVel.size() * 8, true)) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    m.d.vel = (const double *)m.vel.p;
  }
  ctx->dirty = true;
  return FJGPU_OK;
}

int fjgpu_camera_set(fjgpu_context *ctx, const fjgpu_camera *cam) {
  if (!ctx || !cam) return fail(ctx, FJGPU_ERR_INVALID, "null camera");
  ctx->cam = *cam; ctx->have_cam = true;
  return FJGPU_OK;
}

int fjgpu_render_tiles(fjgpu_context *ctx, const fjgpu_render_params *params, const fjgpu_tile *tiles, int32_t ntiles,
                       float *rgba_frame, fjgpu_stats *stats) {
  if (!rgba_frame && ntiles > 0) return fail(ctx, FJGPU_ERR_INVALID, "null framebuffer");
  return render_impl(ctx, params, tiles, ntiles, OUT_HOST, rgba_frame, nullptr, 0, 0, stats);
}

int fjgpu_render_tiles_device(fjgpu_context *ctx, const fjgpu_render_params *params, const fjgpu_tile *tiles, int32_t ntiles,
                              int32_t tile_w_max, int32_t tile_h_max, void *d_tile_blocks, fjgpu_stats *stats) {
  if (!d_tile_blocks && ntiles > 0) return fail(ctx, FJGPU_ERR_INVALID, "null device buffer");
  return render_impl(ctx, params, tiles, ntiles, OUT_DEVICE_BLOCKS, nullptr, d_tile_blocks, tile_w_max, tile_h_max, stats);
}

int fjgpu_render_tiles_resident(fjgpu_context *ctx, const fjgpu_render_params *params, const fjgpu_tile *tiles, int32_t ntiles,
                                fjgpu_stats *stats) {
  return render_impl(ctx, params, tiles, ntiles, OUT_RESIDENT, nullptr, nullptr, 0, 0, stats);
}

int fjgpu_assemble_frame(fjgpu_context *ctx, const void *d_gathered_blocks, int32_t nranks, int32_t tile_w_max, int32_t tile_h_max,
                         const fjgpu_tile *tiles, int32_t ntiles, int32_t xres, int32_t yres, float *rgba_frame) {
  if (!ctx) return fail(nullptr, FJGPU_ERR_INVALID, "null context");
  if (!d_gathered_blocks || nranks < 1 || tile_w_max < 1 || tile_h_max < 1 || (!tiles && ntiles > 0) || ntiles < 0 || xres < 1 || yres < 1 || !rgba_frame)
    return fail(ctx, FJGPU_ERR_INVALID, "fjgpu_assemble_frame: bad arguments");
  for (int i = 0; i < ntiles; i++) {
    const fjgpu_tile &t = tiles[i];
    if (t.xmin < 0 || t.ymin < 0 || t.xmax > xres || t.ymax > yres || t.xmax - t.xmin > tile_w_max || t.ymax - t.ymin > tile_h_max || t.xmax <= t.xmin || t.ymax <= t.ymin)
      return fail(ctx, FJGPU_ERR_INVALID, "fjgpu_assemble_frame: tile outside the frame or larger than a block");
  }
  CK(cudaSetDevice(ctx->device));
  const size_t fbytes = (size_t)xres * yres * sizeof(float4);
  if (int rc = dev_alloc(ctx, ctx->d_frame, fbytes)) return rc;
  if (int rc = dev_upload(ctx, ctx->d_tiles, tiles, (size_t)ntiles * sizeof(fjgpu_tile))) return rc;
  CK(cudaMemsetAsync(ctx->d_frame.p, 0, fbytes, ctx->stream));      // pixels outside the tiles (render_region) stay zero
  const int per = (ntiles + nranks - 1) / nranks;
  if (ntiles > 0)
    fj::k_gathered_to_frame<<<ntiles, 256, 0, ctx->stream>>>((const fj::DTile *)ctx->d_tiles.p, ntiles, nranks, per, (const float4 *)d_gathered_blocks,
                                                             tile_w_max, tile_h_max, (float4 *)ctx->d_frame.p, xres);
  CK(cudaGetLastError());
  // ONE device -> host copy of the frame: straight into the caller's buffer when it is pinned, through the context's pinned
  // staging buffer otherwise
  cudaPointerAttributes at; memset(&at, 0, sizeof at);
  const bool pinned = cudaPointerGetAttributes(&at, rgba_frame) == cudaSuccess && at.type == cudaMemoryTypeHost;
  cudaGetLastError();
  if (pinned) {
    CK(cudaMemcpyAsync(rgba_frame, ctx->d_frame.p, fbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  } else {
    if (ctx->h_blocks_bytes < fbytes) {
      if (ctx->h_blocks) cudaFreeHost(ctx->h_blocks);
      ctx->h_blocks = nullptr; ctx->h_blocks_bytes = 0;
      CK(cudaMallocHost((void **)&ctx->h_blocks, fbytes));
      ctx->h_blocks_bytes = fbytes;
    }
    CK(cudaMemcpyAsync(ctx->h_blocks, ctx->d_frame.p, fbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(rgba_frame, ctx->h_blocks, fbytes);
  }
  return FJGPU_OK;
}

int fjgpu_render_frame_multi(fjgpu_context *const *ctxs, int32_t nranks, const fjgpu_render_params *params, const fjgpu_tile *tiles, int32_t ntiles,
                             float *rgba_frame, fjgpu_stats *stats) {
  if (!ctxs || nranks < 1 || !ctxs[0]) return fail(nullptr, FJGPU_ERR_INVALID, "fjgpu_render_frame_multi: no contexts");
  fjgpu_context *ctx = ctxs[0];
  if (!params || (!tiles && ntiles > 0) || ntiles < 0 || !rgba_frame) return fail(ctx, FJGPU_ERR_INVALID, "fjgpu_render_frame_multi: bad arguments");
  for (int r = 0; r < nranks; r++) {
    if (!ctxs[r]) return fail(ctx, FJGPU_ERR_INVALID, "fjgpu_render_frame_multi: null context");
    for (int q = 0; q < r; q++) if (ctxs[q]->device == ctxs[r]->device) return fail(ctx, FJGPU_ERR_INVALID, "fjgpu_render_frame_multi: one context per device");
  }
  if (stats) memset(stats, 0, sizeof(fjgpu_stats) * (size_t)nranks);
  if (nranks == 1) return fjgpu_render_tiles(ctx, params, tiles, ntiles, rgba_frame, stats);
  int bw = 1, bh = 1;
  for (int i = 0; i < ntiles; i++) { bw = std::max(bw, tiles[i].xmax - tiles[i].xmin); bh = std::max(bh, tiles[i].ymax - tiles[i].ymin); }
  const int per = (ntiles + nranks - 1) / nranks;
  const size_t block_bytes = (size_t)bw * bh * sizeof(float4), send_bytes = (size_t)per * block_bytes;
  // tile i -> rank i % nranks (the rule of sharding.py and libfjscene): interleaved, cheap static balance
  std::vector<std::vector<fjgpu_tile>> mine(nranks);
  for (int i = 0; i < ntiles; i++) mine[i % nranks].push_back(tiles[i]);
  for (int r = 0; r < nranks; r++) {
    fjgpu_context *c = ctxs[r];
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(ctx, FJGPU_ERR_CUDA, "cudaSetDevice");
    if (int rc = dev_alloc(c, c->d_multi_send, send_bytes)) return fail(ctx, rc, c->err);
    if (int rc = dev_alloc(c, c->d_multi_recv, send_bytes * nranks)) return fail(ctx, rc, c->err);
  }
  // one host thread per context renders its tiles into its send buffer (no data-path communication while tracing)
  std::vector<int> rcs(nranks, 0);
  {
    std::vector<std::thread> th;
    for (int r = 0; r < nranks; r++)
      th.emplace_back([&, r]() {
        fjgpu_context *c = ctxs[r];
        cudaSetDevice(c->device);
        cudaMemsetAsync(c->d_multi_send.p, 0, send_bytes, c->stream);
        rcs[r] = render_impl(c, params, mine[r].data(), (int)mine[r].size(), OUT_DEVICE_BLOCKS, nullptr, c->d_multi_send.p, bw, bh, stats ? stats + r : nullptr);
      });
    for (auto &t : th) t.join();
  }
  for (int r = 0; r < nranks; r++) if (rcs[r]) return fail(ctx, rcs[r], "rank " + std::to_string(r) + ": " + ctxs[r]->err);
  // ONE all-gather of the packed tile blocks ends the frame (NCCL over NVLink; loaded at first use)
  if (int rc = multi_all_gather(ctxs, nranks, send_bytes)) return rc;
  return fjgpu_assemble_frame(ctx, ctx->d_multi_recv.p, nranks, bw, bh, tiles, ntiles, params->xres, params->yres, rgba_frame);
}

int fjgpu_trace_closest(fjgpu_context *ctx, int32_t group, int32_t n, const double *orig3, const double *dir3,
                        const double *tmin, const double *tmax, int32_t flags,
                        double *out_t, double *out_u, double *out_v, int32_t *out_prim, int32_t *out_inst) {
  if (!ctx) return fail(nullptr, FJGPU_ERR_INVALID, "null context");
  if (n < 0 || (n > 0 && (!orig3 || !dir3 || !tmin || !tmax || !out_t || !out_u || !out_v || !out_prim || !out_inst)))
    return fail(ctx, FJGPU_ERR_INVALID, "null ray arrays");
  CK(cudaSetDevice(ctx->device));
  if (int rc = commit_scene(ctx)) return rc;
  if (group < 0 || group >= ctx->sc.ngroups) return fail(ctx, FJGPU_ERR_INVALID, "group out of range");
  if (n == 0) return FJGPU_OK;
  DevBuf o, d, t0, t1, rt, ru, rv, rp, ri;
  int rc = 0;
  const size_t N = (size_t)n;
  if ((rc = dev_upload(ctx, o, orig3, N * 24)) || (rc = dev_upload(ctx, d, dir3, N * 24)) || (rc = dev_upload(ctx, t0, tmin, N * 8)) ||
      (rc = dev_upload(ctx, t1, tmax, N * 8)) || (rc = dev_alloc(ctx, rt, N * 8)) || (rc = dev_alloc(ctx, ru, N * 8)) ||
      (rc = dev_alloc(ctx, rv, N * 8)) || (rc = dev_alloc(ctx, rp, N * 4)) || (rc = dev_alloc(ctx, ri, N * 4))) {
    DevBuf *all[] = {&o, &d, &t0, &t1, &rt, &ru, &rv, &rp, &ri}; for (DevBuf *b : all) b->release();
    return rc;
  }
  const int grid = (n + 127) / 128;
  DevBuf q, hits, ctl;
  if (flags & FJGPU_FLAG_FP64_BOXES)
    fj::k_trace_closest<double><<<grid, 128, 0, ctx->stream>>>(ctx->sc, group, n, (const double *)o.p, (const double *)d.p, (const double *)t0.p, (const double *)t1.p,
                                                               (double *)rt.p, (double *)ru.p, (double *)rv.p, (int32_t *)rp.p, (int32_t *)ri.p);
  else if (flags & FJGPU_FLAG_MEGAKERNEL)
    fj::k_trace_closest<float><<<grid, 128, 0, ctx->stream>>>(ctx->sc, group, n, (const double *)o.p, (const double *)d.p, (const double *)t0.p, (const double *)t1.p,
                                                              (double *)rt.p, (double *)ru.p, (double *)rv.p, (int32_t *)rp.p, (int32_t *)ri.p);
  else {      // the wavefront's closest-hit kernel: rays -> queue records -> k_extend -> hit records
    if ((rc = dev_alloc(ctx, q, N * sizeof(fj::RayRec))) || (rc = dev_alloc(ctx, hits, N * sizeof(fj::HitRec))) || (rc = dev_alloc(ctx, ctl, sizeof(fj::QueueCtl)))) {
      DevBuf *all[] = {&o, &d, &t0, &t1, &rt, &ru, &rv, &rp, &ri, &q, &hits, &ctl}; for (DevBuf *b : all) b->release();
      return rc;
    }
    fj::QueueCtl h; memset(&h, 0, sizeof h); h.count[0] = (unsigned)n;
    cudaMemcpyAsync(ctl.p, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream);
    fj::k_probe_pack<<<grid, 128, 0, ctx->stream>>>(group, n, (const double *)o.p, (const double *)d.p, (const double *)t0.p, (const double *)t1.p, (fj::RayRec *)q.p);
    fj::RenderArgs a; memset(&a, 0, sizeof a);
    a.sc = ctx->sc; a.queue[0] = (fj::RayRec *)q.p; a.hits = (fj::HitRec *)hits.p; a.ctl = (fj::QueueCtl *)ctl.p; a.capacity = (uint32_t)n; a.cur = 0;
    launch_extend(ctx, a, std::min(grid, ctx->sm_count * 4));
    fj::k_probe_unpack<<<grid, 128, 0, ctx->stream>>>(n, (const fj::HitRec *)hits.p, (double *)rt.p, (double *)ru.p, (double *)rv.p, (int32_t *)rp.p, (int32_t *)ri.p);
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_t, rt.p, N * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_u, ru.p, N * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_v, rv.p, N * 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_prim, rp.p, N * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_inst, ri.p, N * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  DevBuf *all[] = {&o, &d, &t0, &t1, &rt, &ru, &rv, &rp, &ri, &q, &hits, &ctl}; for (DevBuf *b : all) b->release();
  if (e != cudaSuccess) return fail(ctx, FJGPU_ERR_CUDA, std::string("trace_closest: ") + cudaGetErrorString(e));
  return FJGPU_OK;
}

int fjgpu_render_tile_samples(fjgpu_context *ctx, const fjgpu_render_params *params, const fjgpu_tile *tile, int32_t max_samples,
                              double *out_uv, float *out_rgba, int32_t *out_nsamples) {
  if (!ctx) return fail(nullptr, FJGPU_ERR_INVALID, "null context");
  if (!tile || !out_nsamples) return fail(ctx, FJGPU_ERR_INVALID, "null tile");
  int rc = render_impl(ctx, params, tile, 1, OUT_SAMPLES_ONLY, nullptr, nullptr, 0, 0, nullptr);
  if (rc) return rc;
  FramePlan pl;
  if ((rc = plan_frame(ctx, params, tile, 1, &pl))) return rc;
  const int n = pl.fr.max_ns;
  *out_nsamples = n;
  if (max_samples < n || !out_uv || !out_rgba) return FJGPU_OK;    // size query
  DevBuf duv, drgba;
  if ((rc = dev_alloc(ctx, duv, (size_t)n * 16)) || (rc = dev_alloc(ctx, drgba, (size_t)n * 16))) { duv.release(); drgba.release(); return rc; }
  fj::DTile t; t.id = tile->id; t.xmin = tile->xmin; t.ymin = tile->ymin; t.xmax = tile->xmax; t.ymax = tile->ymax;
  fj::k_dump_tile_samples<<<std::min((n + 255) / 256, 1024), 256, 0, ctx->stream>>>(pl.fr, t, (const fj::Accum *)ctx->d_samples.p, (double *)duv.p, (float4 *)drgba.p);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_uv, duv.p, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_rgba, drgba.p, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  duv.release(); drgba.release();
  if (e != cudaSuccess) return fail(ctx, FJGPU_ERR_CUDA, std::string("render_tile_samples: ") + cudaGetErrorString(e));
  return FJGPU_OK;
}

int fjgpu_scene_info_get(fjgpu_context *ctx, fjgpu_scene_info *info) {
  if (!ctx || !info) return fail(ctx, FJGPU_ERR_INVALID, "null argument");
  memset(info, 0, sizeof *info);
  for (auto &kv : ctx->meshes) {
    const MeshRec &m = kv.second;
    info->hbm_bytes += m.nodes.bytes + m.nodes4.bytes + m.nodesq.bytes + m.tri.bytes + m.N.bytes + m.idx.bytes + m.group.bytes;
    info->blas_nodes += m.nnodes; info->blas_tris += m.nfaces;
    info->blas_max_depth = std::max<uint32_t>(info->blas_max_depth, (uint32_t)m.max_depth);
  }
  info->hbm_bytes += ctx->d_inst.bytes + ctx->d_shaders.bytes + ctx->d_lights.bytes;
  for (auto &b : ctx->d_group_nodes) info->hbm_bytes += b.bytes;
  info->tlas_nodes = ctx->tlas_nodes; info->instances = ctx->inst.size();
  info->build_seconds = ctx->build_seconds;
  info->device_build_seconds = ctx->device_build_seconds;
  return FJGPU_OK;
}

int fjgpu_scene_resend(fjgpu_context *ctx, uint64_t *bytes_sent) {
  if (!ctx) return fail(nullptr, FJGPU_ERR_INVALID, "null context");
  CK(cudaSetDevice(ctx->device));
  if (int rc = commit_scene(ctx)) return rc;
  std::vector<DevBuf *> all = {&ctx->d_meshes, &ctx->d_inst, &ctx->d_groups, &ctx->d_shaders, &ctx->d_lights};
  for (auto &kv : ctx->meshes) { MeshRec &m = kv.second; for (DevBuf *b : {&m.nodes, &m.nodes4, &m.nodesq, &m.tri, &m.N, &m.idx, &m.group, &m.uv, &m.P, &m.vel}) all.push_back(b); }
  for (auto &b : ctx->d_tex_tiles) all.push_back(&b);
  all.push_back(&ctx->d_textures);
  for (auto &b : ctx->d_group_nodes) all.push_back(&b);
  for (auto &b : ctx->d_group_nodes4) all.push_back(&b);
  for (auto &b : ctx->d_group_nodesq) all.push_back(&b);
  for (auto &b : ctx->d_group_order) all.push_back(&b);
  for (auto &b : ctx->d_group_irec) all.push_back(&b);
  for (auto &kv : ctx->d_inst_motion) all.push_back(&kv.second);
  uint64_t total = 0;
  for (DevBuf *b : all) {
    if (!b->h || !b->used) continue;
    CK(cudaMemcpyAsync(b->p, b->h, b->used, cudaMemcpyHostToDevice, ctx->stream));
    total += b->used;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  if (bytes_sent) *bytes_sent = total;
  return FJGPU_OK;
}

}  // extern "C"
