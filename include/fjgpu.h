/*
 * fjgpu.h — the drop-in C-ABI of the B200-native Fujiyama hot path (libfjgpu.so).
 *
 * This is the ONLY boundary between the host renderer (the reference's libscene,
 * or this repo's host mirror `libfjscene`) and the sm_100a CUDA implementation of
 *
 *      Renderer::execute_rendering  (src/fj_renderer.cc:747-791)
 *        -> render_tile / integrate_samples / reconstruct_image (:1061-1121, :939-995)
 *        -> SlTrace (src/fj_shading.cc:140) -> closest hit -> Shader::Evaluate -> SlTrace...
 *
 * Plain C: opaque context handle, plain pointers and sizes, int status returns
 * (0 = FJGPU_OK, negative = error; message via fjgpu_last_error), no C++ or torch
 * types, no callbacks into the host while a launch is in flight, the caller owns
 * every host buffer (they may be freed as soon as the call returns).
 * One context per GPU; a context is thread-compatible (one host thread at a time).
 *
 * Every entry point names the reference interface it replaces (file:line under
 * /root/reference of tsubo164/Fujiyama-Renderer @ a451548).  INTEGRATION.md shows the
 * patch a reference maintainer adds to Renderer::execute_rendering to call these.
 *
 * Numeric contract: geometry is FP64 exactly as the reference (`using Real = double`,
 * src/fj_types.h:13) wherever a decision is taken (ray/triangle test, instance
 * transforms, hit attributes, sample positions, filter weights); colours are FP32
 * (src/fj_color.h:78).  Bounding-volume culling may run in FP32 but is conservative.
 */
#ifndef FJGPU_H
#define FJGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FJGPU_API_VERSION 2

enum {
  FJGPU_OK = 0,
  FJGPU_ERR_INVALID = -1,   /* bad argument / inconsistent scene description      */
  FJGPU_ERR_CUDA = -2,      /* a CUDA runtime call failed (see fjgpu_last_error)  */
  FJGPU_ERR_NO_DEVICE = -3, /* no usable sm_100 device: there is NO CPU fallback  */
  FJGPU_ERR_UNSUPPORTED = -4/* feature outside the device path (caller must use the CPU renderer) */
};

typedef struct fjgpu_context fjgpu_context;

/* ---- context -------------------------------------------------------------------------- */

/* Creates a context on CUDA device `device_ordinal`.  Fails with FJGPU_ERR_NO_DEVICE when
 * no GPU is present — the product path never falls back to a CPU implementation. */
int  fjgpu_create(int device_ordinal, fjgpu_context **out_ctx);
void fjgpu_destroy(fjgpu_context *ctx);
const char *fjgpu_last_error(const fjgpu_context *ctx);   /* ctx may be NULL: last global error */
int  fjgpu_api_version(void);

/* ---- geometry: replaces Mesh storage + GridAccelerator::build ---------------------------
 * src/fj_mesh.h:200-216 (P_, N_, indices_, face_group_id_), src/fj_scene_interface.cc:667-690
 * (SiNewMesh creates a GridAccelerator), src/fj_grid_accelerator.cc:69-160 (build).
 * P, N: nverts*3 doubles (x,y,z AoS as std::vector<Vector>); N may be NULL (shading normal = 0).
 * idx3: nfaces*3 vertex indices; face_group_id: nfaces ints or NULL (all 0).
 * Builds the bottom-level BVH on the host and uploads it.  Re-uploading a mesh_id replaces it. */
int fjgpu_mesh_upload(fjgpu_context *ctx, int32_t mesh_id,
                      const double *P, const double *N, int32_t nverts,
                      const int32_t *idx3, const int32_t *face_group_id, int32_t nfaces);

/* ---- instances: replaces ObjectInstance (src/fj_object_instance.cc:213-243) -------------
 * fwd/inv are the row-major 4x4 matrix and inverse the reference rebuilds per ray in
 * XfmLerpTransformSample (src/fj_transform.cc:306-322); the host computes them once with
 * the reference's own arithmetic (make_transform_matrix + MatInverse) for single-sample
 * (static) transforms.  Time-sampled transforms (motion blur): fjgpu_instance_motion_set below. */
#define FJGPU_MAX_SHADING_GROUPS 8
typedef struct fjgpu_instance {
  int32_t mesh_id;
  int32_t shader_of_group[FJGPU_MAX_SHADING_GROUPS]; /* shader slot per shading_group_id; -1 = unset
                                                        (ObjectInstance::GetShader falls back to slot 0,
                                                        src/fj_object_instance.cc:177-191) */
  int32_t reflect_target;   /* object-group index for diffuse/reflect rays (SlDiffuseContext/SlReflectContext) */
  int32_t refract_target;   /* SlRefractContext, src/fj_shading.cc:254-264 */
  int32_t shadow_target;    /* SlShadowContext,  src/fj_shading.cc:266-279 */
  int32_t _pad;
  double  fwd[16];
  double  inv[16];
} fjgpu_instance;

int fjgpu_instances_set(fjgpu_context *ctx, int32_t n, const fjgpu_instance *inst);

/* Object groups (src/fj_object_group.cc:21-53): group g holds
 * instance_ids[group_offsets[g] .. group_offsets[g+1]).  Builds one top-level BVH per group
 * (the reference's BVHAccelerator over ObjectSet, src/fj_bvh_accelerator.cc:79-107). */
int fjgpu_groups_set(fjgpu_context *ctx, int32_t ngroups,
                     const int32_t *group_offsets, const int32_t *instance_ids);

/* ---- shaders: device re-implementations keyed on plugin_name ----------------------------
 * shaders/constant_shader/constant_shader.cc:72-94, shaders/plastic_shader/plastic_shader.cc:101-179,
 * shaders/pathtracing_shader/pathtracing_shader.cc:125-257, shaders/glass_shader/glass_shader.cc:88-133.  Values are
 * AFTER the clamping the plugin's property setters apply (Max(0,.), ior>=.001, opacity in [0,1], transmit>=.001; glass:
 * ior = Max(0, ior), filter_color >= .001 carried in `transmit`, do_color_filter = filter_color != (1,1,1)). */
enum { FJGPU_SHADER_NONE = 0,      /* no shader: NO_SHADER_COLOR (.5,1,0), Os 1 (src/fj_shading.cc:26,555-560) */
       FJGPU_SHADER_CONSTANT = 1, FJGPU_SHADER_PLASTIC = 2, FJGPU_SHADER_PATHTRACING = 3, FJGPU_SHADER_GLASS = 4 };
typedef struct fjgpu_shader {
  int32_t kind;
  int32_t do_reflect;       /* plastic: any(reflect > 0)  (plastic_shader.cc:231-250)          */
  int32_t do_color_filter;  /* pathtracing: transmit != (1,1,1) (pathtracing_shader.cc:358-378); glass: filter_color != (1,1,1) */
  int32_t texture;          /* constant: `texture`, plastic: `diffuse_map` — 1 + index into fjgpu_textures_set, 0 = none */
  float diffuse[3];
  float reflect[3];
  float refract[3];
  float emission[3];
  float transmit[3];
  float ior;
  float opacity;
  int32_t bump_texture;     /* plastic / pathtracing `bump_map`: 1 + texture index, 0 = none (SlBumpMapping, src/fj_shading.cc:418-465) */
  float bump_amplitude;     /* `bump_amplitude` (plastic_shader.cc:293-300, pathtracing_shader.cc:457-465) */
} fjgpu_shader;

int fjgpu_shaders_set(fjgpu_context *ctx, int32_t n, const fjgpu_shader *shaders);

/* ---- lights: src/fj_point_light.cc:21-39, src/fj_rectangle_light.cc:21-61 ("GridLight"),
 *              src/fj_sphere_light.cc, src/fj_dome_light.cc:25-97 --------------------------
 * Every instance sees every light (create_implicit_groups, src/fj_scene_interface.cc:1100-1104). */
enum { FJGPU_LIGHT_POINT = 0, FJGPU_LIGHT_GRID = 1, FJGPU_LIGHT_SPHERE = 2, FJGPU_LIGHT_DOME = 3 };
typedef struct fjgpu_light {
  int32_t kind;
  int32_t sample_count;     /* Light::sample_count_ (default 16); point light always 1 sample */
  int32_t double_sided;
  int32_t dome_sample_count;
  float   color[3];
  float   intensity;
  double  translate[3];     /* PointLight::get_samples uses the raw translate (fj_point_light.cc:33) */
  double  fwd[16];          /* light transform matrix (grid/sphere lights) */
  const double *dome_dirs;   /* dome_sample_count*3: DomeLight::dome_samples_ directions (host Preprocess) */
  const float  *dome_colors; /* dome_sample_count*3 */
} fjgpu_light;

int fjgpu_lights_set(fjgpu_context *ctx, int32_t n, const fjgpu_light *lights);

/* ---- camera: Camera::GetRay, src/fj_camera.cc:79-110 ----------------------------------- */
typedef struct fjgpu_camera {
  double fwd[16];           /* camera transform matrix (a moving camera adds fjgpu_camera_motion_set) */
  double fov;               /* degrees, default 30 */
  double znear, zfar;       /* ray [tmin,tmax], defaults .01 / 1000 */
} fjgpu_camera;

int fjgpu_camera_set(fjgpu_context *ctx, const fjgpu_camera *cam);


/* ---- frame: Renderer properties (src/internal/fj_property_list_include.cc:451-474) ----- */
enum { FJGPU_RNG_COUNTER = 0 };   /* Philox-4x32-10 keyed (seed, tile id, sample, path node): rank/thread independent */
typedef struct fjgpu_render_params {
  int32_t xres, yres;               /* resolution        (default 320x240) */
  int32_t xrate, yrate;             /* pixelsamples      (default 3x3)     */
  double  xfwidth, yfwidth;         /* filterwidth       (default 2x2), Gaussian (fj_filter.cc:49-58) */
  double  jitter;                   /* sample_jitter     (default 1)       */
  int32_t max_diffuse_depth;        /* default 3 */
  int32_t max_reflect_depth;        /* default 3 */
  int32_t max_refract_depth;        /* default 3 */
  int32_t cast_shadow;              /* default 1 */
  int32_t target_group;             /* object group the camera rays trace (all_objects) */
  uint32_t seed;                    /* stream seed of the stochastic shaders/lights */
  int32_t flags;                    /* FJGPU_FLAG_* */
  int32_t _pad;
} fjgpu_render_params;

enum { FJGPU_FLAG_FP64_BOXES = 1,   /* cull BVH boxes with FP64 slab arithmetic instead of the default conservative
                                       FP32 (same hits; a cross-check for the parity tests; implies the megakernel) */
       FJGPU_FLAG_MEGAKERNEL = 2    /* one camera sample per lane with a private ray stack instead of the wavefront
                                       (generate / extend / shade rounds over ray queues): the independent second
                                       implementation the wavefront is checked against */ };

typedef struct fjgpu_tile {         /* Tile of src/fj_tiler.h; [xmin,xmax) x [ymin,ymax) pixels */
  int32_t id;                       /* global tile id in the frame's tile list (keys the RNG) */
  int32_t xmin, ymin, xmax, ymax;
} fjgpu_tile;

/* ---- motion blur: time-sampled transforms ---------------------------
 * The reference draws ONE time per camera sample and every ray of that sample's tree inherits it (TraceContext::time,
 * src/fj_renderer.cc:1073-1074, src/fj_shading.cc:226-264).  The k-th sample of EVERY tile (k = y * nsamples_x + x) takes
 * the k-th draw of a freshly seeded XorShift mapped into sample_time_range (src/fj_fixed_grid_sampler.cc:42,72-77), so a
 * frame only ever sees a finite table of times.  fjgpu_time_table returns that table; the caller evaluates its own
 * XfmLerpTransformSample (src/fj_transform.cc:306-322 -- glibc sin/cos, which a device cannot return bit for bit) once per
 * table entry and hands the matrices over; rays carry the table index.  An instance / camera without a table is static
 * (fjgpu_instances_set / fjgpu_camera_set matrices).  Lights sample time 0 in the reference (fj_point_light.cc:27-29).
 *
 * fjgpu_time_table: writes min(cap, count) times and returns count = the largest per-tile sample count of `tiles`
 * (< 0 on invalid arguments); call with cap = 0 to size the buffer.
 * fjgpu_instance_motion_set: fwd16 / inv16 = ntimes x 16 doubles (Transform::matrix / inverse at each table time);
 * ntimes = 0 makes the instance static again.  A table belongs to its instance index: it survives fjgpu_instances_set
 * as long as the index exists, and handing over an unchanged table (or instance array) costs one comparison, no upload.
 * Rendering fails with FJGPU_ERR_INVALID if a table is shorter than the frame's time table. */
int fjgpu_time_table(const fjgpu_render_params *params, const fjgpu_tile *tiles, int32_t ntiles,
                     double time_start, double time_end, double *times, int32_t cap);
int fjgpu_instance_motion_set(fjgpu_context *ctx, int32_t instance, int32_t ntimes, const double *fwd16, const double *inv16);
int fjgpu_camera_motion_set(fjgpu_context *ctx, int32_t ntimes, const double *fwd16);

/* ---- motion blur, second half: per-vertex velocity ----------------------------------------
 * Mesh::velocity_ (src/fj_mesh.h:210; written by VelocityGeneratorProcedure): Mesh::ray_intersect tests the triangle where
 * the ray's time puts it, `P0 += time * velocity0` (src/fj_mesh.cc:252-259), the primitive and mesh bounds also hold the
 * vertices at time 1 (src/fj_mesh.cc:420-439) and the shading normal stays the mesh's (:273-276).
 * fjgpu_mesh_upload_velocity = fjgpu_mesh_upload + `velocity` (nverts*3 doubles; NULL = static mesh): the triangle packets
 * carry the velocities, the BVH bounds both ends (host builder), and rays read the time VALUE of their table entry.
 * fjgpu_shutter_set = Renderer::SetSampleTimeRange (src/fj_renderer.cc:526-532; default 0, 1): the range the frame's time
 * table (fjgpu_time_table) is drawn over.  The reference's bounds — and therefore these — cover times in [0, 1]. */
int fjgpu_mesh_upload_velocity(fjgpu_context *ctx, int32_t mesh_id,
                               const double *P, const double *N, int32_t nverts,
                               const int32_t *idx3, const int32_t *face_group_id, int32_t nfaces,
                               const double *velocity);
int fjgpu_shutter_set(fjgpu_context *ctx, double time_start, double time_end);

typedef struct fjgpu_stats {
  uint64_t rays_camera, rays_shadow, rays_diffuse, rays_reflect, rays_refract;
  uint64_t camera_samples;          /* incl. filter-margin samples */
  uint64_t rays_hit;                /* rays (any type) that found a surface */
  uint64_t hit_mesh_levels;         /* sum over those rays of ceil(log2(triangles of the mesh hit)): the root-to-leaf
                                       path lengths of the algorithmic-bytes model */
  uint64_t node_steps, tri_tests;   /* k_extend: 4-wide BVH node visits and exact FP64 triangle tests (all rays) */
  uint64_t kernel_launches;         /* launches of this library's kernels in the call */
  uint64_t trace_launches;          /* launches of the closest-hit kernel (k_extend; k_render_samples in megakernel mode) */
  float    ms_trace;                /* device time of the closest-hit kernel launches (CUDA events on the launching stream) */
  float    ms_resolve;              /* device time of the pixel-filter kernels */
  float    ms_total;                /* device time of the whole call incl. copies */
  float    ms_shade;                /* device time of the generate + shade kernels */
  uint32_t batches;                 /* tile batches the frame was rendered in (per-batch buffers are bounded: FJGPU_SAMPLE_MB) */
  uint32_t queue_regrows;           /* batches rendered again because a branching ray tree overflowed the optimistic ray queue */
  int32_t  first_regrow_batch;      /* index of the first such batch, -1 if none */
  uint32_t _pad;
  uint64_t leaf_phases, leaf_rounds; /* k_extend2: warp-level leaf phases and the rounds of 32 (ray, triangle) pairs they took */
} fjgpu_stats;

/* Renders `ntiles` tiles (sampler -> camera rays -> trace/shade -> Gaussian resolve), i.e. the
 * body of MtRunParallelLoop(render_tile) (src/fj_renderer.cc:786, :1098-1121), and writes the
 * RGBA32F pixels of those tiles into the caller's HOST framebuffer `rgba_frame`
 * (xres*yres*4 floats, row-major, y = 0 on top: src/fj_framebuffer.cc:130-133).
 * Pixels outside the given tiles are left untouched.  `stats` may be NULL. */
int fjgpu_render_tiles(fjgpu_context *ctx, const fjgpu_render_params *params,
                       const fjgpu_tile *tiles, int32_t ntiles,
                       float *rgba_frame, fjgpu_stats *stats);

/* Same, but leaves the result in DEVICE memory as packed tile blocks for the multi-GPU
 * gather: block i = tile i, tile_w_max*tile_h_max*4 floats (row-major inside the tile, unused
 * texels zero).  `d_tile_blocks` is a device pointer owned by the caller (e.g. a torch tensor
 * that is the NCCL all-gather send buffer).  The call returns after the work is complete. */
int fjgpu_render_tiles_device(fjgpu_context *ctx, const fjgpu_render_params *params,
                              const fjgpu_tile *tiles, int32_t ntiles,
                              int32_t tile_w_max, int32_t tile_h_max,
                              void *d_tile_blocks, fjgpu_stats *stats);

/* Device-resident timing leg for bench.py: renders the tiles like fjgpu_render_tiles but
 * keeps the frame on the device (no D2H).  Fills stats with CUDA-event times. */
int fjgpu_render_tiles_resident(fjgpu_context *ctx, const fjgpu_render_params *params,
                                const fjgpu_tile *tiles, int32_t ntiles, fjgpu_stats *stats);

/* ---- multi-GPU: tiles are the independent units of the path (src/fj_renderer.cc:1098-1121; the reference deals them to
 * its worker threads through MtRunParallelLoop, src/fj_multi_thread.cc:86-100).  Tile i of the frame's tile list goes to rank
 * i % nranks, every rank renders into packed blocks in its own HBM with the scene replicated, ONE all-gather of the blocks
 * ends the frame, rank 0 un-permutes on the device and copies the frame to the host once.
 *
 * fjgpu_assemble_frame: the last step for hosts that ran the all-gather themselves (bench.py: one process per GPU,
 * torch.distributed over NCCL).  d_gathered_blocks = nranks x ceil(ntiles / nranks) blocks of tile_w_max x tile_h_max x 4
 * floats in rank order on ctx's device, `tiles` = the whole frame's tile list; writes xres*yres*4 floats to rgba_frame
 * (pixels outside the tiles zero).  The copy goes straight into rgba_frame when it is pinned host memory.
 *
 * fjgpu_render_frame_multi: the whole frame from ONE process driving one context per GPU (the same scene uploaded to each):
 * one host thread per context renders its tiles, ncclAllGather over the contexts' streams (libnccl.so.2 is loaded at first
 * use: FJGPU_ERR_UNSUPPORTED without it), rank 0 assembles.  stats (may be NULL) = nranks entries.  The image is identical
 * for every nranks (the counter RNG is keyed by tile id). */
int fjgpu_assemble_frame(fjgpu_context *ctx, const void *d_gathered_blocks, int32_t nranks,
                         int32_t tile_w_max, int32_t tile_h_max, const fjgpu_tile *tiles, int32_t ntiles,
                         int32_t xres, int32_t yres, float *rgba_frame);
int fjgpu_render_frame_multi(fjgpu_context *const *ctxs, int32_t nranks, const fjgpu_render_params *params,
                             const fjgpu_tile *tiles, int32_t ntiles, float *rgba_frame, fjgpu_stats *stats);

/* ---- probes used by the parity tests (each mirrors one reference function) ------------- */

/* Closest hit of n rays in object group `group` = Accelerator::Intersect of the group's
 * surface accelerator (src/fj_shading.cc:527-541): out_t (REAL_MAX on miss), barycentrics u,v,
 * primitive (face) index in the mesh and instance index (-1 on miss). */
int fjgpu_trace_closest(fjgpu_context *ctx, int32_t group, int32_t n,
                        const double *orig3, const double *dir3,
                        const double *tmin, const double *tmax, int32_t flags,
                        double *out_t, double *out_u, double *out_v,
                        int32_t *out_prim, int32_t *out_inst);

/* Per-sample radiance of one tile before the pixel filter (Sample::data, src/fj_pixel_sample.h):
 * out_uv = nsamples*2 doubles (screen uv), out_rgba = nsamples*4 floats, row-major over the
 * tile's (rate*w+2m) x (rate*h+2m) sample grid (src/fj_fixed_grid_sampler.cc:33-84). */
int fjgpu_render_tile_samples(fjgpu_context *ctx, const fjgpu_render_params *params,
                              const fjgpu_tile *tile, int32_t max_samples,
                              double *out_uv, float *out_rgba, int32_t *out_nsamples);

/* Scene statistics (BVH nodes, bytes resident in HBM) for DESIGN.md / bench config. */
typedef struct fjgpu_scene_info {
  uint64_t hbm_bytes;       /* device bytes held by the scene (nodes + triangles + attributes) */
  uint64_t blas_nodes, blas_tris, tlas_nodes, instances;
  uint32_t blas_max_depth, _pad;
  double   build_seconds;   /* host time spent in BVH construction since context creation */
  double   device_build_seconds;   /* part of it spent in the device builder (FJGPU_BUILD=device; CUDA events), 0 for host builds */
} fjgpu_scene_info;
int fjgpu_scene_info_get(fjgpu_context *ctx, fjgpu_scene_info *info);

/* ---- textures ------------------------------------------------------------
 * Replaces Texture / TextureCache (src/fj_texture.cc:51-78) over a `.mip` file (src/fj_mipmap.cc:124-180: "MIPM", version,
 * width, height, nchannels, tilesize, then (width/tilesize) x (height/tilesize) row-major tiles of tilesize x tilesize x
 * nchannels floats).  `tiles` is the file's tile data in file order; the lookup is the reference's nearest-texel tile
 * arithmetic (64 texels per tile side).  nchannels 1, 3 or 4 (FrameBuffer::GetColor, src/fj_framebuffer.cc:84-101). */
typedef struct fjgpu_texture {
  int32_t width, height, nchannels, tilesize;
  const float *tiles;
} fjgpu_texture;
int fjgpu_textures_set(fjgpu_context *ctx, int32_t n, const fjgpu_texture *textures);
/* Per-vertex texture coordinates of an uploaded mesh (Mesh::AddPointTexture, src/fj_mesh.h; interpolated as
 * Mesh::ray_intersect does, src/fj_mesh.cc:280-291).  uv2 = nverts x 2 floats; NULL removes them. */
int fjgpu_mesh_set_uv(fjgpu_context *ctx, int32_t mesh_id, const float *uv2, int32_t nverts);

/* Re-sends the retained pinned host copy of every scene array (BVH nodes, triangle packets, normals, indices,
 * instance / group / shader / light tables) host -> device and reports the bytes copied: what a host that
 * edited or re-loaded the scene pays before a frame (the reference re-walks its host scene in prepare_render,
 * src/fj_scene_interface.cc:1204-1222).  bench.py's end-to-end leg calls it every step so that each step carries
 * its inputs over PCIe. */
int fjgpu_scene_resend(fjgpu_context *ctx, uint64_t *bytes_sent);

#ifdef __cplusplus
}
#endif
#endif /* FJGPU_H */
