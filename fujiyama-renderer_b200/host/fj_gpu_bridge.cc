// fj_gpu_bridge.cc — Integration A: the UNMODIFIED reference host (libscene: Scene, Si* interface, plugin loader, .scn parser,
// framebuffer, callbacks) rendering through libfjgpu (include/fjgpu.h).
//
// This file is the only thing added to the reference.  `make -f oracle/Makefile.ref bridge` compiles the reference's own
// src/*.cc as they lie under /root/reference (nothing is patched or copied), compiles this file against the reference's headers,
// and links both into fujiyama-renderer_b200/host/_refgpu/lib/libscene.so with four `-Wl,--wrap=` redirections — the linker
// sends the following calls between the reference's object files through this file:
//
//   Renderer::RenderScene()           (called by SiRenderScene, src/fj_scene_interface.cc:247-278)   remembers which Renderer is rendering
//   MtRunParallelLoop(render_tile…)   (the hot loop, src/fj_renderer.cc:786)                          THE DISPATCH: the frame's tiles go to
//                                     fjgpu_render_tiles when FJ_DEVICE is set and the scene has a device path; otherwise — or
//                                     when anything is outside that path — the reference's CPU workers run exactly as before
//   Scene::NewShader(Plugin *)        (called by SiNewShader, src/fj_scene_interface.cc:559-582)      shader instance -> plugin (name, property list)
//   Property::SetValue(self, value)   (called by set_property, src/fj_scene_interface.cc:1236-1275)   records (shader, property, value): the plugin
//                                     classes keep their parameters in members private to the DSO (SURVEY.md 8b "parameter capture")
//
// Everything else — dlopen + Initialize of the shader / procedure plugins (src/fj_plugin.cc:28-70), the Si* calls, PLY loading,
// ComputeNormals, lights' Preprocess, frame callbacks, the .fb writer — is the reference's own code, unchanged.  The
// unmodified bin/scene and the unmodified ConstantShader.so / PlasticShader.so / PathtracingShader.so / GlassShader.so /
// StanfordPlyProcedure.so / VelocityGeneratorProcedure.so run on this libscene.so (tests/test_bridge_gpu.py).
//
// The CPU fallback described above lives HERE, in the reference's host code; libfjgpu itself has none.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <stdint.h>
#include <string>
#include <typeinfo>
#include <vector>

// the scene description lives in private members of the reference's classes (ObjectInstance::acc_ / transform_samples_,
// Accelerator::primset_, ObjectGroup::surface_set_, Light::…, Camera::…, Texture::filename_); the reference's headers stay as
// they are and this translation unit alone reads through the access specifiers
#define private public
#define protected public
#include "fj_accelerator.h"
#include "fj_camera.h"
#include "fj_dome_light.h"
#include "fj_framebuffer.h"
#include "fj_light.h"
#include "fj_mesh.h"
#include "fj_mipmap.h"
#include "fj_multi_thread.h"
#include "fj_object_group.h"
#include "fj_object_instance.h"
#include "fj_object_set.h"
#include "fj_plugin.h"
#include "fj_point_light.h"
#include "fj_property.h"
#include "fj_rectangle_light.h"
#include "fj_renderer.h"
#include "fj_scene.h"
#include "fj_shader.h"
#include "fj_sphere_light.h"
#include "fj_texture.h"
#include "fj_tiler.h"
#include "fj_transform.h"
#include "fj_volume_accelerator.h"
#undef private
#undef protected

#include "fjgpu.h"

namespace {

using namespace fj;

struct ShaderRecord {
  std::string plugin_name;
  const Property *property_list = nullptr;
  std::map<std::string, PropertyValue> values;        // what Si{SetProperty,AssignTexture} handed to the plugin's setters
};
std::map<const void *, ShaderRecord> g_shaders;        // by Shader instance
Renderer *g_renderer = nullptr;                        // the Renderer inside RenderScene()

struct Bridge {
  fjgpu_context *ctx = nullptr;
  int device = -1;
  std::map<const Mesh *, uint64_t> mesh_key;           // content key of what the device holds for mesh id = position in mesh_order
  std::vector<const Mesh *> mesh_order;
  std::string texture_key;
  std::string motion_key;
} g_bridge;

int why(const char *msg) {       // the reason the CPU workers run instead
  if (getenv("FJ_DEVICE_VERBOSE")) fprintf(stderr, "# fjgpu bridge: %s; rendering on the CPU workers\n", msg);
  return 1;
}

uint64_t fnv(const void *p, size_t n, uint64_t h = 1469598103934665603ull) {
  const uint64_t *w = (const uint64_t *)p;
  for (size_t i = 0; i < n / 8; i++) { h ^= w[i]; h *= 1099511628211ull; }
  const unsigned char *b = (const unsigned char *)p + (n & ~(size_t)7);
  for (size_t i = 0; i < (n & 7); i++) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

// Value of a recorded property or the plugin's default (Property::GetDefaultValue, seeded like PropSetAllDefaultValues does
// in the plugin's constructor).
bool prop(const ShaderRecord &r, const char *name, double out[4]) {
  auto it = r.values.find(name);
  if (it != r.values.end()) { for (int k = 0; k < 4; k++) out[k] = it->second.vector[k]; return true; }
  for (const Property *p = r.property_list; p && p->IsValid(); p++)
    if (strcmp(p->GetName(), name) == 0) { for (int k = 0; k < 4; k++) out[k] = p->GetDefaultValue()[k]; return true; }
  return false;
}
const Texture *prop_texture(const ShaderRecord &r, const char *name) {
  auto it = r.values.find(name);
  return it != r.values.end() && it->second.type == PROP_TEXTURE ? it->second.texture : nullptr;
}
inline float clamp0(double v) { return (float)(v > 0 ? v : 0); }

// The device description of a shader instance: the values AFTER the clamping the plugin's setters apply
// (constant_shader.cc:96-107, plastic_shader.cc:183-300, pathtracing_shader.cc:305-420, glass_shader.cc:135-213).
bool flatten_shader(const ShaderRecord &r, const std::map<const Texture *, int> &tex, fjgpu_shader *o) {
  memset(o, 0, sizeof *o);
  double v[4];
  auto tex_slot = [&](const char *name) { const Texture *t = prop_texture(r, name); auto it = tex.find(t); return t && it != tex.end() ? it->second + 1 : 0; };
  if (r.plugin_name == "ConstantShader") {
    o->kind = FJGPU_SHADER_CONSTANT;
    if (!prop(r, "diffuse", v)) return false;
    for (int k = 0; k < 3; k++) o->diffuse[k] = clamp0(v[k]);
    o->texture = tex_slot("texture");
  } else if (r.plugin_name == "PlasticShader") {
    o->kind = FJGPU_SHADER_PLASTIC;
    if (!prop(r, "diffuse", v)) return false;
    for (int k = 0; k < 3; k++) o->diffuse[k] = clamp0(v[k]);
    if (!prop(r, "reflect", v)) return false;
    for (int k = 0; k < 3; k++) o->reflect[k] = clamp0(v[k]);
    o->do_reflect = (o->reflect[0] > 0 || o->reflect[1] > 0 || o->reflect[2] > 0) ? 1 : 0;
    if (!prop(r, "ior", v)) return false;
    o->ior = (float)std::max(.001, (double)(float)v[0]);
    if (!prop(r, "opacity", v)) return false;
    { const float op = (float)v[0]; o->opacity = op < 0 ? 0 : (op > 1 ? 1 : op); }
    if (!prop(r, "bump_amplitude", v)) return false;
    o->bump_amplitude = (float)v[0];
    o->texture = tex_slot("diffuse_map"); o->bump_texture = tex_slot("bump_map");
  } else if (r.plugin_name == "GlassShader") {
    o->kind = FJGPU_SHADER_GLASS;
    if (!prop(r, "filter_color", v)) return false;
    for (int k = 0; k < 3; k++) o->transmit[k] = (float)std::max(.001, v[k]);
    o->do_color_filter = (o->transmit[0] == 1 && o->transmit[1] == 1 && o->transmit[2] == 1) ? 0 : 1;
    if (!prop(r, "ior", v)) return false;
    { const float ior = (float)v[0]; o->ior = ior > 0 ? ior : 0; }
    o->opacity = 1;
  } else if (r.plugin_name == "PathtracingShader") {
    o->kind = FJGPU_SHADER_PATHTRACING;
    const char *names[5] = {"emission", "diffuse", "reflect", "refract", "transmit"};
    float *dst[5] = {o->emission, o->diffuse, o->reflect, o->refract, o->transmit};
    for (int n = 0; n < 5; n++) {
      if (!prop(r, names[n], v)) return false;
      for (int k = 0; k < 3; k++) dst[n][k] = n == 4 ? (float)std::max(.001, v[k]) : clamp0(v[k]);
    }
    o->do_color_filter = (o->transmit[0] == 1 && o->transmit[1] == 1 && o->transmit[2] == 1) ? 0 : 1;
    if (!prop(r, "ior", v)) return false;
    o->ior = (float)std::max(.001, (double)(float)v[0]);
    o->opacity = 1;
    o->texture = tex_slot("diffuse_map"); o->bump_texture = tex_slot("bump_map");
    if (!prop(r, "bump_amplitude", v)) return false;
    o->bump_amplitude = (float)v[0];
  } else return false;
  return true;
}

// A `.mip` file's tiles in file order through the reference's own reader (MipInput, src/fj_mipmap.cc:124-180).
bool read_mip(const std::string &filename, fjgpu_texture *t, std::vector<float> *tiles) {
  MipInput mip;
  if (mip.Open(filename)) return false;
  if (mip.ReadHeader()) return false;
  const int w = mip.GetWidth(), h = mip.GetHeight(), nc = mip.GetChannelCount(), ts = mip.GetTileSize();
  if (w <= 0 || h <= 0 || ts <= 0) return false;
  const int xnt = w / ts, ynt = h / ts;
  tiles->resize((size_t)xnt * ynt * ts * ts * nc);
  for (int y = 0; y < ynt; y++) for (int x = 0; x < xnt; x++)
    if (mip.ReadTile(x, y, tiles->data() + (size_t)(y * xnt + x) * ts * ts * nc)) return false;
  t->width = w; t->height = h; t->nchannels = nc; t->tilesize = ts; t->tiles = tiles->data();
  return true;
}

bool moving(const TransformSampleList &l) { return l.translate.sample_count > 1 || l.rotate.sample_count > 1 || l.scale.sample_count > 1; }

// Renders the frame of `r` on the device.  0 = done (the framebuffer holds the frame); anything else = not rendered, nothing
// was written, the caller runs the CPU workers.
int device_render(Renderer *r) {
  const char *dev = getenv("FJ_DEVICE");
  if (!dev || !*dev) return 1;                                     // (silent: the device path was not asked for)
  if (!r || !r->camera_ || !r->framebuffer_ || !r->target_objects_) return why("renderer without camera / framebuffer / objects");
  if (r->sampler_type_ != RENDERER_FIXED_GRID_SAMPLER) return why("adaptive sampler");
  const ObjectGroup *all = r->target_objects_;
  if (all->volume_set_.GetObjectCount() > 0) return why("volumes in the scene");

  // ---- object groups reachable from the frame: group 0 = the renderer's target objects, then every reflect / refract /
  // shadow target (SlReflectContext / SlRefractContext / SlShadowContext, src/fj_shading.cc:242-279)
  std::vector<const ObjectGroup *> groups{all};
  std::map<const ObjectGroup *, int> group_index{{all, 0}};
  std::vector<const ObjectInstance *> inst;
  std::map<const ObjectInstance *, int> inst_index;
  for (size_t g = 0; g < groups.size(); g++) {
    const ObjectSet &set = groups[g]->surface_set_;
    if (groups[g]->volume_set_.GetObjectCount() > 0) return why("volumes in a target group");
    for (Index k = 0; k < set.GetObjectCount(); k++) {
      const ObjectInstance *o = set.GetObject(k);
      if (!inst_index.count(o)) { inst_index[o] = (int)inst.size(); inst.push_back(o); }
      for (const ObjectGroup *t : {o->GetReflectTarget(), o->GetRefractTarget(), o->GetShadowTarget()})
        if (t && !group_index.count(t)) { group_index[t] = (int)groups.size(); groups.push_back(t); }
    }
  }

  // ---- meshes, shaders, textures
  std::vector<const Mesh *> meshes; std::map<const Mesh *, int> mesh_index;
  std::vector<const void *> shaders; std::map<const void *, int> shader_index;
  for (const ObjectInstance *o : inst) {
    if (!o->IsSurface() || !o->acc_ || !o->acc_->primset_) return why("an object instance without a surface");
    const Mesh *m = dynamic_cast<const Mesh *>(o->acc_->primset_);
    if (!m) return why("a primitive set that is not a Mesh (curves / point clouds)");
    if (!mesh_index.count(m)) { mesh_index[m] = (int)meshes.size(); meshes.push_back(m); }
    if (o->shader_list_.size() > FJGPU_MAX_SHADING_GROUPS) return why("more shading groups than the device table holds");
    for (const Shader *s : o->shader_list_) {
      if (!s) continue;
      if (!g_shaders.count(s)) return why("a shader instance of an unknown plugin");
      if (!shader_index.count(s)) { shader_index[s] = (int)shaders.size(); shaders.push_back(s); }
    }
  }
  std::vector<const Texture *> textures; std::map<const Texture *, int> tex_index;
  for (const void *s : shaders)
    for (auto &kv : g_shaders[s].values)
      if (kv.second.type == PROP_TEXTURE && kv.second.texture && !tex_index.count(kv.second.texture)) {
        tex_index[kv.second.texture] = (int)textures.size(); textures.push_back(kv.second.texture);
      }
  std::vector<fjgpu_shader> fshaders(shaders.size());
  for (size_t i = 0; i < shaders.size(); i++)
    if (!flatten_shader(g_shaders[shaders[i]], tex_index, &fshaders[i])) return why("a shader plugin without a device kind");

  // ---- the context (one per process, re-created when FJ_DEVICE changes)
  Bridge &B = g_bridge;
  const int ordinal = atoi(dev);
  if (B.ctx && B.device != ordinal) { fjgpu_destroy(B.ctx); B = Bridge(); }
  if (!B.ctx) {
    if (fjgpu_create(ordinal, &B.ctx) != FJGPU_OK) { fprintf(stderr, "# fjgpu bridge: %s\n", fjgpu_last_error(nullptr)); B.ctx = nullptr; return why("no device"); }
    B.device = ordinal;
  }
  fjgpu_context *ctx = B.ctx;
  auto failed = [&](const char *what) { fprintf(stderr, "# fjgpu bridge: %s: %s\n", what, fjgpu_last_error(ctx)); return why(what); };

  // ---- meshes: uploaded when new or changed (content key over positions, indices, velocities, texture coordinates)
  if (meshes != B.mesh_order) { B.mesh_key.clear(); B.mesh_order = meshes; }
  for (size_t mi = 0; mi < meshes.size(); mi++) {
    const Mesh *m = meshes[mi];
    const int nv = m->GetPointCount(), nf = m->GetFaceCount();
    static_assert(sizeof(Vector) == 24 && sizeof(Index3) == 12 && sizeof(TexCoord) == 8, "reference vector layouts");
    uint64_t key = fnv(&nv, 4); key = fnv(&nf, 4, key);
    if (m->HasPointPosition()) key = fnv(m->P_.data(), m->P_.size() * sizeof(Vector), key);
    if (m->HasPointNormal()) key = fnv(m->N_.data(), m->N_.size() * sizeof(Vector), key);
    if (m->HasFaceIndices()) key = fnv(m->indices_.data(), m->indices_.size() * sizeof(Index3), key);
    if (m->HasPointVelocity()) key = fnv(m->velocity_.data(), m->velocity_.size() * sizeof(Vector), key);
    if (m->HasPointTexture()) key = fnv(m->uv_.data(), m->uv_.size() * sizeof(TexCoord), key);
    if (m->HasFaceGroupID()) key = fnv(m->face_group_id_.data(), m->face_group_id_.size() * sizeof(int), key);
    auto it = B.mesh_key.find(m);
    if (it != B.mesh_key.end() && it->second == key) continue;
    if (!m->HasPointPosition() || !m->HasFaceIndices() || (int)m->P_.size() < nv || (int)m->indices_.size() < nf) return why("a mesh without positions / faces");
    const double *P = reinterpret_cast<const double *>(m->P_.data());
    const double *N = m->HasPointNormal() && (int)m->N_.size() >= nv ? reinterpret_cast<const double *>(m->N_.data()) : nullptr;
    const int32_t *idx = reinterpret_cast<const int32_t *>(m->indices_.data());
    const int32_t *grp = m->HasFaceGroupID() && (int)m->face_group_id_.size() >= nf ? m->face_group_id_.data() : nullptr;
    const double *vel = m->HasPointVelocity() && (int)m->velocity_.size() >= nv ? reinterpret_cast<const double *>(m->velocity_.data()) : nullptr;
    if (fjgpu_mesh_upload_velocity(ctx, (int32_t)mi, P, N, nv, idx, grp, nf, vel) != FJGPU_OK) return failed("fjgpu_mesh_upload");
    if (m->HasPointTexture() && (int)m->uv_.size() >= nv) {
      if (fjgpu_mesh_set_uv(ctx, (int32_t)mi, reinterpret_cast<const float *>(m->uv_.data()), nv) != FJGPU_OK) return failed("fjgpu_mesh_set_uv");
    }
    B.mesh_key[m] = key;
  }

  // ---- textures: the .mip files named at SiNewTexture, read tile by tile with the reference's MipInput
  {
    std::string key;
    for (const Texture *t : textures) key += t->filename_ + "\n";
    if (key != B.texture_key) {
      std::vector<fjgpu_texture> ft(textures.size());
      std::vector<std::vector<float>> tiles(textures.size());
      for (size_t i = 0; i < textures.size(); i++)
        if (!read_mip(textures[i]->filename_, &ft[i], &tiles[i])) return why("a texture file that cannot be read");
      if (fjgpu_textures_set(ctx, (int32_t)ft.size(), ft.data()) != FJGPU_OK) return failed("fjgpu_textures_set");
      B.texture_key = key;
    }
  }

  // ---- frame parameters and tiles (the reference's own Tiler: what execute_rendering hands to its workers)
  fjgpu_render_params p; memset(&p, 0, sizeof p);
  p.xres = r->resolution_[0]; p.yres = r->resolution_[1];
  p.xrate = r->pixelsamples_[0]; p.yrate = r->pixelsamples_[1];
  p.xfwidth = r->filterwidth_[0]; p.yfwidth = r->filterwidth_[1];
  p.jitter = r->jitter_;
  p.max_diffuse_depth = r->max_diffuse_depth_; p.max_reflect_depth = r->max_reflect_depth_; p.max_refract_depth = r->max_refract_depth_;
  p.cast_shadow = r->cast_shadow_; p.target_group = 0;
  { const char *seed = getenv("FJ_SEED"); p.seed = seed ? (uint32_t)strtoul(seed, nullptr, 10) : 1u; }
  Tiler tiler;
  tiler.Divide(p.xres, p.yres, r->tilesize_[0], r->tilesize_[1]);
  tiler.GenerateTiles(r->frame_region_);
  std::vector<fjgpu_tile> tiles(tiler.GetTileCount());
  for (size_t i = 0; i < tiles.size(); i++) {
    const Tile *t = tiler.GetTile((int)i);
    tiles[i].id = t->id; tiles[i].xmin = t->xmin; tiles[i].ymin = t->ymin; tiles[i].xmax = t->xmax; tiles[i].ymax = t->ymax;
  }
  if (tiles.empty()) return why("no tiles");

  // ---- instances, groups, lights, camera: transforms at time 0 through the reference's own XfmLerpTransformSample
  std::vector<fjgpu_instance> finst(inst.size());
  for (size_t i = 0; i < inst.size(); i++) {
    const ObjectInstance *o = inst[i]; fjgpu_instance &d = finst[i];
    memset(&d, 0, sizeof d);
    d.mesh_id = mesh_index[dynamic_cast<const Mesh *>(o->acc_->primset_)];
    for (int g = 0; g < FJGPU_MAX_SHADING_GROUPS; g++)
      d.shader_of_group[g] = g < (int)o->shader_list_.size() && o->shader_list_[g] ? shader_index[o->shader_list_[g]] : -1;
    d.reflect_target = o->GetReflectTarget() ? group_index[o->GetReflectTarget()] : 0;
    d.refract_target = o->GetRefractTarget() ? group_index[o->GetRefractTarget()] : 0;
    d.shadow_target = o->GetShadowTarget() ? group_index[o->GetShadowTarget()] : 0;
    Transform t; XfmLerpTransformSample(&o->transform_samples_, 0., &t);
    static_assert(sizeof(t.matrix.e) == 16 * sizeof(double), "Matrix layout");
    memcpy(d.fwd, t.matrix.e, sizeof d.fwd); memcpy(d.inv, t.inverse.e, sizeof d.inv);
  }
  std::vector<int32_t> goff{0}, gids;
  for (const ObjectGroup *g : groups) {
    const ObjectSet &set = g->surface_set_;
    for (Index k = 0; k < set.GetObjectCount(); k++) gids.push_back(inst_index[set.GetObject(k)]);
    goff.push_back((int32_t)gids.size());
  }
  std::vector<fjgpu_light> flights(r->nlights_);
  std::vector<std::vector<double>> dome_dirs(r->nlights_); std::vector<std::vector<float>> dome_cols(r->nlights_);
  for (int i = 0; i < r->nlights_; i++) {
    const Light *l = r->target_lights_[i]; fjgpu_light &d = flights[i];
    memset(&d, 0, sizeof d);
    if (dynamic_cast<const PointLight *>(l)) d.kind = FJGPU_LIGHT_POINT;
    else if (dynamic_cast<const RectangleLight *>(l)) d.kind = FJGPU_LIGHT_GRID;
    else if (dynamic_cast<const SphereLight *>(l)) d.kind = FJGPU_LIGHT_SPHERE;
    else if (dynamic_cast<const DomeLight *>(l)) d.kind = FJGPU_LIGHT_DOME;
    else return why("a light type without a device kind");
    d.sample_count = l->sample_count_; d.double_sided = l->double_sided_ ? 1 : 0;
    d.color[0] = l->color_.r; d.color[1] = l->color_.g; d.color[2] = l->color_.b; d.intensity = l->intensity_;
    Transform t; XfmLerpTransformSample(&l->transform_samples_, 0., &t);       // lights sample time 0 (fj_point_light.cc:27-29)
    d.translate[0] = t.translate.x; d.translate[1] = t.translate.y; d.translate[2] = t.translate.z;
    memcpy(d.fwd, t.matrix.e, sizeof d.fwd);
    if (const DomeLight *dl = dynamic_cast<const DomeLight *>(l)) {            // dome_samples_ as Light::Preprocess left them
      for (const DomeSample &s : dl->dome_samples_) {
        dome_dirs[i].insert(dome_dirs[i].end(), {s.dir.x, s.dir.y, s.dir.z});
        dome_cols[i].insert(dome_cols[i].end(), {s.color.r, s.color.g, s.color.b});
      }
      d.dome_sample_count = (int32_t)dl->dome_samples_.size();
      d.dome_dirs = dome_dirs[i].data(); d.dome_colors = dome_cols[i].data();
    }
  }
  fjgpu_camera cam; memset(&cam, 0, sizeof cam);
  { Transform t; XfmLerpTransformSample(&r->camera_->transform_samples_, 0., &t); memcpy(cam.fwd, t.matrix.e, sizeof cam.fwd); }
  cam.fov = r->camera_->fov_; cam.znear = r->camera_->znear_; cam.zfar = r->camera_->zfar_;

  int rc = fjgpu_shaders_set(ctx, (int32_t)fshaders.size(), fshaders.data());
  if (!rc) rc = fjgpu_shutter_set(ctx, r->sample_time_start_, r->sample_time_end_);
  if (!rc) rc = fjgpu_groups_set(ctx, (int32_t)goff.size() - 1, goff.data(), gids.data());
  if (!rc) rc = fjgpu_instances_set(ctx, (int32_t)finst.size(), finst.data());
  if (!rc) rc = fjgpu_lights_set(ctx, (int32_t)flights.size(), flights.data());
  if (!rc) rc = fjgpu_camera_set(ctx, &cam);
  if (rc) return failed("scene description");

  // ---- time-sampled transforms: the reference's XfmLerpTransformSample (glibc sin / cos) once per entry of the frame's time
  // table instead of once per ray per candidate instance (include/fjgpu.h "motion blur")
  {
    std::vector<int> movers;
    for (size_t i = 0; i < inst.size(); i++) if (moving(inst[i]->transform_samples_)) movers.push_back((int)i);
    const bool cam_moves = moving(r->camera_->transform_samples_);
    std::ostringstream key;
    std::vector<double> times;
    if (!movers.empty() || cam_moves) {
      const int n = fjgpu_time_table(&p, tiles.data(), (int32_t)tiles.size(), r->sample_time_start_, r->sample_time_end_, nullptr, 0);
      if (n <= 0) return why("fjgpu_time_table");
      times.resize(n);
      fjgpu_time_table(&p, tiles.data(), (int32_t)tiles.size(), r->sample_time_start_, r->sample_time_end_, times.data(), n);
      key << n << ' ' << r->sample_time_start_ << ' ' << r->sample_time_end_ << ' ';
    }
    auto keys_of = [&](const TransformSampleList &l) { key << fnv(&l, sizeof l) << ' '; };
    for (int i : movers) { key << i << ':'; keys_of(inst[i]->transform_samples_); }
    if (cam_moves) { key << "cam:"; keys_of(r->camera_->transform_samples_); }
    if (key.str() != B.motion_key) {
      std::vector<double> fwd(times.size() * 16), inv(times.size() * 16);
      for (size_t i = 0; i < inst.size(); i++) {
        if (!moving(inst[i]->transform_samples_)) { if (fjgpu_instance_motion_set(ctx, (int32_t)i, 0, nullptr, nullptr)) return failed("fjgpu_instance_motion_set"); continue; }
        for (size_t k = 0; k < times.size(); k++) {
          Transform t; XfmLerpTransformSample(&inst[i]->transform_samples_, times[k], &t);
          memcpy(&fwd[16 * k], t.matrix.e, 128); memcpy(&inv[16 * k], t.inverse.e, 128);
        }
        if (fjgpu_instance_motion_set(ctx, (int32_t)i, (int32_t)times.size(), fwd.data(), inv.data())) return failed("fjgpu_instance_motion_set");
      }
      if (cam_moves) {
        for (size_t k = 0; k < times.size(); k++) { Transform t; XfmLerpTransformSample(&r->camera_->transform_samples_, times[k], &t); memcpy(&fwd[16 * k], t.matrix.e, 128); }
        if (fjgpu_camera_motion_set(ctx, (int32_t)times.size(), fwd.data())) return failed("fjgpu_camera_motion_set");
      } else if (fjgpu_camera_motion_set(ctx, 0, nullptr)) return failed("fjgpu_camera_motion_set");
      B.motion_key = key.str();
    }
  }

  // ---- the frame: FrameBuffer::buf_ is the row-major RGBA float frame fjgpu_render_tiles writes (src/fj_framebuffer.cc:103-133)
  FrameBuffer *fb = r->framebuffer_;
  if (fb->GetWidth() != p.xres || fb->GetHeight() != p.yres || fb->GetChannelCount() != 4) return why("framebuffer is not RGBA at the frame's resolution");
  fjgpu_stats st;
  if (fjgpu_render_tiles(ctx, &p, tiles.data(), (int32_t)tiles.size(), fb->GetWritable(0, 0, 0), &st) != FJGPU_OK) return failed("fjgpu_render_tiles");
  if (getenv("FJ_DEVICE_VERBOSE"))
    fprintf(stderr, "# fjgpu bridge: device %d rendered %zu tiles, %llu rays, %.1f ms\n", ordinal, tiles.size(),
            (unsigned long long)(st.rays_camera + st.rays_shadow + st.rays_diffuse + st.rays_reflect + st.rays_refract), st.ms_total);
  // tile callbacks after the fact (per-sample callbacks cannot be honoured on a device)
  for (size_t i = 0; i < tiles.size(); i++) {
    TileInfo info;
    info.worker_id = 0; info.region_id = tiles[i].id; info.total_region_count = (int)tiles.size(); info.frame_id = r->frame_id_;
    info.tile_region.min[0] = tiles[i].xmin; info.tile_region.min[1] = tiles[i].ymin; info.tile_region.max[0] = tiles[i].xmax; info.tile_region.max[1] = tiles[i].ymax;
    info.framebuffer = fb;
    CbReportTileStart(&r->tile_report_, &info);
    CbReportTileDone(&r->tile_report_, &info);
  }
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ the four redirections
// (`-Wl,--wrap=<mangled name>`: references to <name> from other object files resolve to __wrap_<name>, and __real_<name> is
// the reference's own definition.)  A non-static member function takes `this` as its first argument in the Itanium C++ ABI.
extern "C" {

int __real__ZN2fj8Renderer11RenderSceneEv(fj::Renderer *self);
int __wrap__ZN2fj8Renderer11RenderSceneEv(fj::Renderer *self) {
  g_renderer = self;
  const int rc = __real__ZN2fj8Renderer11RenderSceneEv(self);
  g_renderer = nullptr;
  return rc;
}

fj::LoopStatus __real__ZN2fj17MtRunParallelLoopEPvPFNS_10LoopStatusES0_RKNS_13ThreadContextEEiRKSt6vectorIiSaIiEE(
    void *data, fj::TaskFunction task_fn, int thread_count, const std::vector<int> &iteration_que);
fj::LoopStatus __wrap__ZN2fj17MtRunParallelLoopEPvPFNS_10LoopStatusES0_RKNS_13ThreadContextEEiRKSt6vectorIiSaIiEE(
    void *data, fj::TaskFunction task_fn, int thread_count, const std::vector<int> &iteration_que) {
  // src/fj_renderer.cc:786 — the only MtRunParallelLoop inside Renderer::RenderScene is the tile loop
  if (g_renderer && device_render(g_renderer) == 0) return fj::LoopStatus::Continue;
  return __real__ZN2fj17MtRunParallelLoopEPvPFNS_10LoopStatusES0_RKNS_13ThreadContextEEiRKSt6vectorIiSaIiEE(data, task_fn, thread_count, iteration_que);
}

fj::Shader *__real__ZN2fj5Scene9NewShaderEPNS_6PluginE(fj::Scene *self, fj::Plugin *plugin);
fj::Shader *__wrap__ZN2fj5Scene9NewShaderEPNS_6PluginE(fj::Scene *self, fj::Plugin *plugin) {
  fj::Shader *s = __real__ZN2fj5Scene9NewShaderEPNS_6PluginE(self, plugin);
  if (s && plugin) {
    ShaderRecord &r = g_shaders[s];
    r = ShaderRecord();
    r.plugin_name = plugin->GetName() ? plugin->GetName() : "";
    r.property_list = plugin->GetPropertyList();
  }
  return s;
}

int __real__ZNK2fj8Property8SetValueEPvRKNS_13PropertyValueE(const fj::Property *prop, void *self, const fj::PropertyValue &value);
int __wrap__ZNK2fj8Property8SetValueEPvRKNS_13PropertyValueE(const fj::Property *prop, void *self, const fj::PropertyValue &value) {
  const int rc = __real__ZNK2fj8Property8SetValueEPvRKNS_13PropertyValueE(prop, self, value);
  if (rc == 0 && self && prop && prop->GetName()) {
    auto it = g_shaders.find(self);
    if (it != g_shaders.end()) { fj::PropertyValue v = value; v.string = nullptr; it->second.values[prop->GetName()] = v; }
  }
  return rc;
}

}  // extern "C"
