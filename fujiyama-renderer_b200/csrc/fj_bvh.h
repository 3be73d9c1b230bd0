// fj_bvh.h — host-side BVH construction for the device tracer (product code).
//
// Replaces, on the device path, both acceleration structures of the reference:
//   GridAccelerator (per mesh, src/fj_grid_accelerator.cc:69-160) -> bottom-level BVH2 over triangles
//   BVHAccelerator  (per object group, src/fj_bvh_accelerator.cc:79-107) -> top-level BVH2 over instances
// Both reference structures implement the same contract — the closest hit with tmin <= t <= tmax
// (src/fj_primitive_set.cc:10-26) — so any conservative BVH returns the same hit.
//
// Node layout (64 B, one 128-B line holds two nodes; fetched as 4 x 16-B loads by one lane):
//   f[0..3]  = child0.lo.x, child0.hi.x, child0.lo.y, child0.hi.y
//   f[4..7]  = child1.lo.x, child1.hi.x, child1.lo.y, child1.hi.y
//   f[8..11] = child0.lo.z, child0.hi.z, child1.lo.z, child1.hi.z
//   c[0], c[1] = children: >= 0 inner node index, < 0 leaf: ~c = (first << 3) | (count - 1)
// Boxes are FP32, rounded outward and padded so that FP32/FP64 slab culling never rejects a box
// whose triangle the reference's FP64 Moller-Trumbore test would hit.
#ifndef FJ_BVH_H
#define FJ_BVH_H

#include <cstdint>
#include <vector>

namespace fjb {

struct Node64 {
  float f[12];
  int32_t c[2];
  int32_t pad[2];
};
static_assert(sizeof(Node64) == 64, "node must be 64 bytes");

// 4-wide node of the wavefront's closest-hit kernel (128 B = one cache line, fetched as 8 x 16-B loads by one lane):
// child boxes SoA (lo.x[4] hi.x[4] lo.y[4] hi.y[4] lo.z[4] hi.z[4]), then the four child references in the
// encoding above.  Unused slots hold a box that is never hit (lo = +3e38, hi = -3e38).  Built by collapsing the
// binary tree (the child with the largest surface area is opened first).
struct Node128 {
  float lox[4], hix[4], loy[4], hiy[4], loz[4], hiz[4];
  int32_t c[4];
  int32_t pad[4];
};
static_assert(sizeof(Node128) == 128, "wide node must be 128 bytes");

struct Aabb {
  float lo[3], hi[3];
};

struct BuildResult {
  std::vector<Node64> nodes;     // nodes[0] is the root; the first `top_count` nodes are the top levels in BFS order
  std::vector<int32_t> order;    // primitive indices in leaf order (leaf `first` indexes this array)
  int32_t top_count = 0;
  int32_t max_depth = 0;
  Aabb bounds;
  std::vector<Node128> nodes4;   // the same tree collapsed to 4-wide nodes, DFS preorder, nodes4[0] is the root
  int32_t max_depth4 = 0;
};

// Builds a binned-SAH BVH2 over `n` primitive boxes.  max_leaf <= 8.  `top_levels` BFS levels are laid
// out contiguously at the front of `nodes` (they are staged in shared memory by the kernels).
void build_bvh(const Aabb *prims, int32_t n, int max_leaf, float leaf_cost, int top_levels, BuildResult *out);

// Outward-rounded float bounds of a double value.
float round_down(double v);
float round_up(double v);

}  // namespace fjb
#endif
