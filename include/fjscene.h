/*
 * fjscene.h — host mirror of the reference's scene interface for the device hot path (libfjscene.so).
 *
 * The functions in `namespace fj` carry the SAME names, argument meaning, ID encoding and error
 * behaviour as src/fj_scene_interface.h:64-127 of tsubo164/Fujiyama-Renderer @ a451548, so a host
 * program written against the reference's API (scenes/cube.cc) or a `.scn` command file written for
 * bin/scene (tools/scene_parser) runs unchanged — SiRenderScene() then executes
 * Renderer::execute_rendering (src/fj_renderer.cc:747-791) through the C-ABI of fjgpu.h instead of the
 * CPU worker threads.  What differs, by design:
 *   - SiOpenPlugin() does not dlopen: the device shaders are keyed on the plugin NAME
 *     (ConstantShader, PlasticShader, PathtracingShader, GlassShader; StanfordPlyProcedure as the mesh loader,
 *     VelocityGeneratorProcedure, whose per-vertex velocities reach the device as moving triangles).  The reference's OWN
 *     host with its dlopen'ed plugin DSOs runs on libfjgpu through host/fj_gpu_bridge.cc.
 *     Any other plugin yields SI_BADID / SI_ERR_PLUGIN_NOT_FOUND — there is no CPU fallback here.
 *   - entity kinds outside the path (Volume, Curve, PointCloud, Turbulence) yield SI_BADID / SI_ERR_FAILNEW; the
 *     adaptive sampler makes SiRenderScene() return SI_FAIL.  Time-sampled transforms of instances and cameras
 *     (SiSetSampleProperty3: motion blur) are rendered; lights are evaluated at time 0 as in the reference.
 */
#ifndef FJSCENE_H
#define FJSCENE_H

#include <stdint.h>
#include "fjgpu.h"

#ifdef __cplusplus
namespace fj {

typedef long int ID;
typedef int Status;
enum { SI_BADID = -1 };
enum { SI_FAIL = -1, SI_SUCCESS = 0 };

enum SiErrorNo {                    /* src/fj_scene_interface.h:17-32 */
  SI_ERR_NONE = 0, SI_ERR_NO_MEMORY, SI_ERR_BADTYPE, SI_ERR_FAILLOAD, SI_ERR_FAILNEW,
  SI_ERR_PLUGIN_NOT_FOUND, SI_ERR_INIT_PLUGIN_FUNC_NOT_EXIST, SI_ERR_INIT_PLUGIN_FUNC_FAIL,
  SI_ERR_BAD_PLUGIN_INFO, SI_ERR_CLOSE_PLUGIN_FAIL, SI_ERR_UNDEFINED
};
enum SiTransformOrder {             /* :34-49 */
  SI_ORDER_SRT = 0, SI_ORDER_STR, SI_ORDER_RST, SI_ORDER_RTS, SI_ORDER_TRS, SI_ORDER_TSR,
  SI_ORDER_XYZ, SI_ORDER_XZY, SI_ORDER_YXZ, SI_ORDER_YZX, SI_ORDER_ZXY, SI_ORDER_ZYX
};
enum SiLightType { SI_POINT_LIGHT = 0, SI_GRID_LIGHT, SI_SPHERE_LIGHT, SI_DOME_LIGHT };
enum SiSamplerType { SI_FIXED_GRID_SAMPLER = 0, SI_ADAPTIVE_GRID_SAMPLER = 1 };

/* callbacks: src/fj_callback.h:14-74 (Rectangle = int min[2], max[2]) */
struct Rectangle { int min[2]; int max[2]; };
class FrameBuffer;
struct FrameInfo { int32_t frame_id; int worker_count; int tile_count; int xres; int yres; Rectangle frame_region; const FrameBuffer *framebuffer; };
struct TileInfo { int32_t frame_id; int worker_id; int region_id; int total_region_count; Rectangle tile_region; const FrameBuffer *framebuffer; };
enum { CALLBACK_CONTINUE = 0, CALLBACK_INTERRUPT = -1 };
typedef int Interrupt;
typedef Interrupt (*FrameStartCallback)(void *data, const FrameInfo *info);
typedef Interrupt (*FrameAbortCallback)(void *data, const FrameInfo *info);
typedef Interrupt (*FrameDoneCallback)(void *data, const FrameInfo *info);
typedef Interrupt (*TileStartCallback)(void *data, const TileInfo *info);
typedef Interrupt (*TileDoneCallback)(void *data, const TileInfo *info);
typedef Interrupt (*SampleDoneCallback)(void *data);

int SiGetErrorNo(void);
ID SiOpenPlugin(const char *filename);
Status SiOpenScene(void);
Status SiCloseScene(void);
Status SiRenderScene(ID renderer);
Status SiSaveFrameBuffer(ID framebuffer, const char *filename);
Status SiRunProcedure(ID procedure);
Status SiAddObjectToGroup(ID group, ID object);

ID SiNewObjectInstance(ID primset);
ID SiNewFrameBuffer(const char *arg);
ID SiNewObjectGroup(void);
ID SiNewPointCloud(void);
ID SiNewTurbulence(void);
ID SiNewProcedure(ID plugin);
ID SiNewRenderer(void);
ID SiNewTexture(const char *filename);
ID SiNewCamera(const char *arg);
ID SiNewShader(ID plugin);
ID SiNewVolume(void);
ID SiNewCurve(void);
ID SiNewLight(int light_type);
ID SiNewMesh(void);

Status SiAssignFrameBuffer(ID renderer, ID framebuffer);
Status SiAssignObjectGroup(ID id, const char *name, ID group);
Status SiAssignPointCloud(ID id, const char *name, ID pointcloud);
Status SiAssignTurbulence(ID id, const char *name, ID turbulence);
Status SiAssignTexture(ID id, const char *name, ID texture);
Status SiAssignVolume(ID id, const char *name, ID volume);
Status SiAssignCamera(ID renderer, ID camera);
Status SiAssignShader(ID object, const char *shading_group, ID shader);
Status SiAssignCurve(ID id, const char *name, ID curve);
Status SiAssignMesh(ID id, const char *name, ID mesh);

Status SiSetProperty1(ID id, const char *name, double v0);
Status SiSetProperty2(ID id, const char *name, double v0, double v1);
Status SiSetProperty3(ID id, const char *name, double v0, double v1, double v2);
Status SiSetProperty4(ID id, const char *name, double v0, double v1, double v2, double v3);
Status SiSetStringProperty(ID id, const char *name, const char *string);
Status SiSetSampleProperty3(ID id, const char *name, double v0, double v1, double v2, double time);

Status SiSetFrameReportCallback(ID id, void *data, FrameStartCallback frame_start, FrameAbortCallback frame_abort,
                                FrameDoneCallback frame_done);
Status SiSetTileReportCallback(ID id, void *data, TileStartCallback tile_start, SampleDoneCallback sample_done,
                               TileDoneCallback tile_done);

}  /* namespace fj */
extern "C" {
#endif

/* ---- plain-C entry points (bindings: ctypes / cgo / JNI) --------------------------------------- */

/* `.scn` command interpreter — the grammar of tools/scene_parser (parser.cc:45-97, command.cc:502-543):
 * one command per line, `#` comments, names bound by New* / OpenPlugin commands.  A fresh interpreter opens a
 * scene (SiOpenScene) and closes it on free.  Returns 0, or -1 after printing "error: <file>:<line>: ...". */
typedef struct fjscene_parser fjscene_parser;
fjscene_parser *fjscene_parser_new(void);
void fjscene_parser_free(fjscene_parser *p);
int fjscene_parse_line(fjscene_parser *p, const char *line);
int fjscene_parse_text(fjscene_parser *p, const char *text);
int fjscene_parse_file(fjscene_parser *p, const char *path);
long fjscene_lookup(fjscene_parser *p, const char *name);           /* ID bound to a name, -1 if none */
void fjscene_set_echo(fjscene_parser *p, int echo);                 /* echo commands like parser.cc:274-285 (default on) */

/* Direct mesh fill (what a Procedure plugin does through Mesh::Set*, src/fj_mesh.h): P = nverts*3 doubles,
 * idx3 = nfaces*3.  Runs Mesh::ComputeNormals + ComputeBounds like ply2mesh.cc:164-165. */
int fjscene_mesh_set(long mesh_id, const double *P, int32_t nverts, const int32_t *idx3, int32_t nfaces);

/* Frame access and device statistics of the last SiRenderScene. */
const float *fjscene_framebuffer(long framebuffer_id, int32_t *width, int32_t *height, int32_t *channels);
int fjscene_last_stats(fjgpu_stats *stats, fjgpu_scene_info *info, double *upload_seconds);

/* Device selection for SiRenderScene: tiles with (index % world_size) == rank are rendered on `device_ordinal`;
 * the others are left untouched in the framebuffer (multi-GPU sharding).  Default 0, 0, 1. */
void fjscene_set_device(int device_ordinal, int rank, int world_size);
/* Render mode of SiRenderScene: 0 = host framebuffer (default), 1 = keep the frame on the device (bench leg). */
void fjscene_set_resident(int resident);
/* 1 = every SiRenderScene first re-sends the whole scene host -> device (fjgpu_scene_resend); bytes of the last one. */
void fjscene_set_resend(int resend);
uint64_t fjscene_last_resend_bytes(void);
/* GPUs ONE process renders a frame on (default: the environment variable FJ_GPU_COUNT, else 1): contexts on `count`
 * consecutive devices from fjscene_set_device's ordinal on, the scene uploaded to each, tiles dealt round-robin, one NCCL
 * all-gather of the tile blocks, rank 0 assembles the host framebuffer (fjgpu_render_frame_multi).  `fjscene file.scn` with
 * FJ_GPU_COUNT=8 renders on the 8 GPUs of a box without Python. */
void fjscene_set_gpu_count(int count);
/* One process per GPU (torchrun): after the host's all-gather of the ranks' tile blocks (fjscene_set_device_blocks), rank 0
 * un-permutes them on the device into the last frame's layout and copies the frame to `rgba_frame` (xres*yres*4 floats; pinned
 * memory is written directly) — fjgpu_assemble_frame with the frame's own tile list.  0 or -1. */
int fjscene_assemble_gathered(const void *d_gathered_blocks, int nranks, int tile_w_max, int tile_h_max, float *rgba_frame);
/* Non-NULL: SiRenderScene leaves this rank's tiles as packed blocks (tile_w_max*tile_h_max*4 floats per tile, in
 * the rank's tile order) in the caller's DEVICE buffer — the send buffer of the multi-GPU all-gather
 * (fjgpu_render_tiles_device).  NULL restores the host-framebuffer mode. */
void fjscene_set_device_blocks(void *d_tile_blocks, int tile_w_max, int tile_h_max);
/* The flat scene description handed to libfjgpu by the last prepare_render (parity tests compare it with the
 * oracle's): row-major 4x4 forward/inverse matrices of instance `index`; returns 0 or -1. */
int fjscene_instance_matrices(int32_t index, double *fwd16, double *inv16);
int fjscene_mesh_normals(long mesh_id, double *N_out, int32_t nverts);
/* Per-vertex velocities VelocityGeneratorProcedure wrote on a mesh (3 doubles per vertex); -1 if it has none.  The
 * procedure is mirrored bit for bit; SiRenderScene hands the velocities to fjgpu_mesh_upload_velocity. */
int fjscene_mesh_velocity(long mesh_id, double *vel_out, int32_t nverts);
const char *fjscene_last_message(void);
/* The flat scene description SiRenderScene would hand to libfjgpu for a renderer (instances, lights, shaders, camera,
 * frame parameters, tiles), built without touching a device: host-logic tests compare it with an independent
 * flattening.  Pointers inside the returned structs stay valid until the next fjscene_flatten.  0 or -1. */
int fjscene_flatten(long renderer_id, int32_t *ninst, int32_t *nlights, int32_t *nshaders, int32_t *ntiles);
int fjscene_flat_instance(int32_t i, fjgpu_instance *out);
int fjscene_flat_light(int32_t i, fjgpu_light *out);
int fjscene_flat_shader(int32_t i, fjgpu_shader *out);
int fjscene_flat_tile(int32_t i, fjgpu_tile *out);
int fjscene_flat_frame(fjgpu_camera *cam, fjgpu_render_params *params);
/* XfmLerpTransformSample (src/fj_transform.cc:306-322) of an ObjectInstance / Camera / Light entry at `time`: the
 * matrices SiRenderScene tabulates per entry of the frame's time table for fjgpu_instance_motion_set /
 * fjgpu_camera_motion_set (motion blur); returns 0 or -1. */
int fjscene_lerp_transform(long id, double time, double *fwd16, double *inv16);
/* make_transform_matrix + MatInverse with the reference's arithmetic (src/fj_transform.cc:335-391,
 * src/fj_matrix.cc:119-207): the matrices SiRenderScene hands to fjgpu_instances_set. */
void fjscene_make_transform(int transform_order, int rotate_order, const double *T, const double *R, const double *S,
                            double *fwd16, double *inv16);

#ifdef __cplusplus
}
#endif
#endif /* FJSCENE_H */
