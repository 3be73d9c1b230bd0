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
  int32_t stack_need4 = 0;       // worst-case traversal-stack entries of a nearest-first walk of nodes4 (see stack_need())
  int32_t top_count4 = 0;        // nodes4[0 .. top_count4) are the top of the tree in breadth-first order (any prefix of them is
                                 // a breadth-first prefix too): what k_extend2 stages in shared memory with one bulk copy
};
#define FJB_TOP4_MAX 341         // levels 0..4 of a full 4-wide tree

// The 4-wide tree with child boxes quantised to 8 bits per plane inside the node's own box (64 B per node, two 32-B
// loads): the closest-hit kernel is bound by the L1 data pipe, which moves 16 B per lane per pass when every lane reads
// its own node, so bytes per node step are what it pays for.
//   w[0..2]  p = lo corner of the union of the valid child boxes (float)
//   w[3]     (bits of sx) & 0xffff0000 | (bits of sy) >> 16      scale per axis, a power of two (exact in 16 bits)
//   w[4..7]  qlo.x[4] qhi.x[4] qlo.y[4] qhi.y[4]                  one byte per child: plane = p + q * s
//   w[8..9]  qlo.z[4] qhi.z[4]
//   w[10..13] child references (as Node128::c)
//   w[14]    (bits of sz) & 0xffff0000
// qlo is rounded down and qhi up, so a decoded box contains the FP32 box it came from.  Unused slots hold the inverted
// box qlo = 255, qhi = 0, which the sign-selected slab test never enters.
struct NodeQ64 { uint32_t w[16]; };
static_assert(sizeof(NodeQ64) == 64, "quantised node must be 64 bytes");
// Returns false when a node cannot be represented (extent beyond the exponent range kept for the FP32 error bound).
// *bmag receives the largest |coordinate| of any decoded plane.
bool quantize_nodes(const Node128 *in, size_t n, NodeQ64 *out, float *bmag);

// Builds a binned-SAH BVH2 over `n` primitive boxes.  max_leaf <= 8.  `top_levels` BFS levels are laid
// out contiguously at the front of `nodes` (they are staged in shared memory by the kernels).
void build_bvh(const Aabb *prims, int32_t n, int max_leaf, float leaf_cost, int top_levels, BuildResult *out);

// Outward-rounded float bounds of a double value.
float round_down(double v);
float round_up(double v);

}  // namespace fjb
#endif
