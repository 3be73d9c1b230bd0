// fj_build.h — bottom-level BVH construction ON THE DEVICE.
//
// Replaces, for large meshes, the host builder of fj_bvh.cc on the path that the reference runs as
// "Building Accelerators" (build_accelerators, src/fj_scene_interface.cc:1161-1202 → GridAccelerator::build,
// src/fj_grid_accelerator.cc:69-160: ≈ 1.3 s per million triangles on one core).  Any conservative BVH returns the
// reference's closest hit (fj_bvh.h), so the tree may be built differently: here a linear BVH — 63-bit Morton codes of
// the padded triangle boxes' centres, one radix sort, the binary radix tree of Karras (HPG 2012), bottom-up boxes, then
// the same products as the host builder: binary Node64, 4-wide Node128 (largest-area child opened first),
// 8-bit quantised NodeQ64, triangle packets in leaf order, worst-case stack depths.  Everything stays in HBM.
#ifndef FJ_BUILD_H
#define FJ_BUILD_H

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

struct FjDeviceBuild {
  void *nodes = nullptr, *nodes4 = nullptr, *nodesq = nullptr, *tri = nullptr;   // cudaMalloc'ed, owned by the caller
  size_t nodes_bytes = 0, nodes4_bytes = 0, nodesq_bytes = 0, tri_bytes = 0;
  int32_t nnodes = 0, nnodes4 = 0, max_depth = 0, max_depth4 = 0, stack_need4 = 0;
  float bmag = 0, bmagq = 0;
  int quant_ok = 0, tri64 = 0;
  double bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};     // exact FP64 bounds of the referenced vertices (Mesh::ComputeBounds)
  double seconds = 0;                                    // device time of the whole build (CUDA events)
};

// dP: nverts x 3 doubles, didx: nfaces x 3 int32 (already validated), both in device memory.  nfaces >= 2.
// Returns 0, or -1 with *err set (the caller falls back to the host builder).
int fj_device_build(cudaStream_t stream, const double *dP, int32_t nverts, const int32_t *didx, int32_t nfaces, int max_leaf, float leaf_cost,
                    bool force_tri64, FjDeviceBuild *out, std::string *err);

#endif
