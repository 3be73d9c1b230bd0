// fj_bvh.cc — binned-SAH BVH2 builder (host, multi-threaded).  See fj_bvh.h for the layout.
#include "fj_bvh.h"
#include "fj_quant.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>

namespace fjb {

float round_down(double v) {
  float f = (float)v;
  if ((double)f > v) f = std::nextafterf(f, -FLT_MAX);
  return f;
}
float round_up(double v) {
  float f = (float)v;
  if ((double)f < v) f = std::nextafterf(f, FLT_MAX);
  return f;
}

namespace {

const int NBINS = 16;

struct BNode {
  Aabb box;
  int32_t left, right;   // -1 for leaves
  int32_t first, count;
};

inline void box_empty(Aabb &b) { for (int a = 0; a < 3; a++) { b.lo[a] = FLT_MAX; b.hi[a] = -FLT_MAX; } }
inline void box_add(Aabb &b, const Aabb &o) {
  for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], o.lo[a]); b.hi[a] = std::max(b.hi[a], o.hi[a]); }
}
inline float box_area(const Aabb &b) {
  const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  if (dx < 0 || dy < 0 || dz < 0) return 0.f;
  return 2.f * (dx * dy + dy * dz + dz * dx);
}

struct Builder {
  const Aabb *prims;
  std::vector<float> cx[3];
  std::vector<int32_t> idx;
  std::vector<BNode> nodes;
  std::atomic<int32_t> next_node{0};
  std::atomic<int> threads_left{0};
  int max_leaf;
  float leaf_cost;

  int32_t alloc() { return next_node.fetch_add(1); }

  int32_t build(int32_t b, int32_t e) {
    const int32_t me = alloc();
    Aabb box, cbox;
    box_empty(box); box_empty(cbox);
    for (int32_t i = b; i < e; i++) {
      const int32_t p = idx[i];
      box_add(box, prims[p]);
      for (int a = 0; a < 3; a++) { cbox.lo[a] = std::min(cbox.lo[a], cx[a][p]); cbox.hi[a] = std::max(cbox.hi[a], cx[a][p]); }
    }
    BNode nd; nd.box = box; nd.left = nd.right = -1; nd.first = b; nd.count = e - b;
    const int32_t n = e - b;
    if (n <= 1) { nodes[me] = nd; return me; }

    // binned SAH over the three axes
    float best_cost = FLT_MAX; int best_axis = -1, best_bin = -1;
    for (int a = 0; a < 3; a++) {
      const float ext = cbox.hi[a] - cbox.lo[a];
      if (!(ext > 0)) continue;
      const float scale = NBINS * (1.f - 1e-6f) / ext;
      Aabb bb[NBINS]; int32_t bc[NBINS];
      for (int k = 0; k < NBINS; k++) { box_empty(bb[k]); bc[k] = 0; }
      for (int32_t i = b; i < e; i++) {
        const int32_t p = idx[i];
        int k = (int)((cx[a][p] - cbox.lo[a]) * scale);
        k = k < 0 ? 0 : (k >= NBINS ? NBINS - 1 : k);
        box_add(bb[k], prims[p]); bc[k]++;
      }
      float ra[NBINS]; int32_t rc[NBINS];
      Aabb acc; box_empty(acc); int32_t cnt = 0;
      for (int k = NBINS - 1; k > 0; k--) { box_add(acc, bb[k]); cnt += bc[k]; ra[k] = box_area(acc); rc[k] = cnt; }
      box_empty(acc); cnt = 0;
      for (int k = 0; k < NBINS - 1; k++) {
        box_add(acc, bb[k]); cnt += bc[k];
        if (cnt == 0 || rc[k + 1] == 0) continue;
        const float cost = box_area(acc) * cnt + ra[k + 1] * rc[k + 1];
        if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = k; }
      }
    }
    const float parent_area = std::max(box_area(box), 1e-30f);
    if (n <= max_leaf) {
      // leaf if cheaper than splitting: node cost 1 (two boxes), triangle cost leaf_cost
      const float split_cost = 1.f + leaf_cost * best_cost / parent_area;
      if (best_axis < 0 || leaf_cost * n <= split_cost) { nodes[me] = nd; return me; }
    }
    int32_t mid;
    if (best_axis >= 0) {
      const int a = best_axis;
      const float lo = cbox.lo[a], scale = NBINS * (1.f - 1e-6f) / (cbox.hi[a] - cbox.lo[a]);
      const int bin = best_bin;
      const std::vector<float> &c = cx[a];
      int32_t *first = idx.data() + b, *last = idx.data() + e;
      int32_t *m = std::partition(first, last, [&](int32_t p) {
        int k = (int)((c[p] - lo) * scale); k = k < 0 ? 0 : (k >= NBINS ? NBINS - 1 : k); return k <= bin; });
      mid = (int32_t)(m - idx.data());
    } else mid = b;
    if (mid == b || mid == e) {   // all centroids coincide: split the range in half
      int a = 0; float ext = -1;
      for (int k = 0; k < 3; k++) if (box.hi[k] - box.lo[k] > ext) { ext = box.hi[k] - box.lo[k]; a = k; }
      mid = b + n / 2;
      const std::vector<float> &c = cx[a];
      std::nth_element(idx.begin() + b, idx.begin() + mid, idx.begin() + e, [&](int32_t p, int32_t q) { return c[p] < c[q]; });
    }
    int32_t l, r;
    if (n > 32768 && threads_left.fetch_sub(1) > 0) {
      std::thread t([&]() { l = build(b, mid); });
      r = build(mid, e);
      t.join();
      threads_left.fetch_add(1);
    } else {
      if (n > 32768) threads_left.fetch_add(1);
      l = build(b, mid);
      r = build(mid, e);
    }
    nd.left = l; nd.right = r;
    nodes[me] = nd;
    return me;
  }
};

inline int32_t leaf_ref(int32_t first, int32_t count) { return ~((first << 3) | (count - 1)); }

}  // namespace

void build_bvh(const Aabb *prims, int32_t n, int max_leaf, float leaf_cost, int top_levels, BuildResult *out) {
  out->nodes.clear(); out->order.clear(); out->top_count = 0; out->max_depth = 0;
  out->nodes4.clear(); out->max_depth4 = 0; out->top_count4 = 0; out->stack_need4 = 0;
  Node128 empty4; memset(&empty4, 0, sizeof empty4);
  // unused slots: a valid box (lo == hi) at +3e38 that no ray interval reaches — the min/max slab test of k_extend2 needs lo <= hi
  for (int k = 0; k < 4; k++) { empty4.lox[k] = empty4.loy[k] = empty4.loz[k] = 3e38f; empty4.hix[k] = empty4.hiy[k] = empty4.hiz[k] = 3e38f; empty4.c[k] = ~0; }
  box_empty(out->bounds);
  Node64 empty_node; memset(&empty_node, 0, sizeof empty_node);
  for (int k = 0; k < 12; k++) empty_node.f[k] = (k % 2 == 0) ? 3e38f : -3e38f;   // lo = +big, hi = -big: never hit
  empty_node.c[0] = empty_node.c[1] = leaf_ref(0, 1);
  if (n <= 0) {   // an empty set: one node with two never-hit children
    out->nodes.push_back(empty_node); out->top_count = 1;
    out->nodes4.push_back(empty4); out->max_depth4 = 1;
    return;
  }
  Builder B;
  B.prims = prims; B.max_leaf = std::min(std::max(max_leaf, 1), 8); B.leaf_cost = leaf_cost;
  for (int a = 0; a < 3; a++) B.cx[a].resize(n);
  B.idx.resize(n);
  for (int32_t i = 0; i < n; i++) {
    B.idx[i] = i;
    for (int a = 0; a < 3; a++) B.cx[a][i] = .5f * prims[i].lo[a] + .5f * prims[i].hi[a];
  }
  B.nodes.resize((size_t)2 * n + 2);
  B.threads_left = (int)std::max(1u, std::thread::hardware_concurrency()) - 1;
  const int32_t root = B.build(0, n);
  out->bounds = B.nodes[root].box;
  out->order = B.idx;

  auto set_child = [&](Node64 &dst, int which, const BNode &c, int32_t ref) {
    if (which == 0) { dst.f[0] = c.box.lo[0]; dst.f[1] = c.box.hi[0]; dst.f[2] = c.box.lo[1]; dst.f[3] = c.box.hi[1]; dst.f[8] = c.box.lo[2]; dst.f[9] = c.box.hi[2]; }
    else            { dst.f[4] = c.box.lo[0]; dst.f[5] = c.box.hi[0]; dst.f[6] = c.box.lo[1]; dst.f[7] = c.box.hi[1]; dst.f[10] = c.box.lo[2]; dst.f[11] = c.box.hi[2]; }
    dst.c[which] = ref;
  };

  if (B.nodes[root].left < 0) {   // the whole set is one leaf
    Node64 nd = empty_node;
    set_child(nd, 0, B.nodes[root], leaf_ref(0, B.nodes[root].count));
    out->nodes.push_back(nd); out->top_count = 1; out->max_depth = 1;
    Node128 w = empty4;
    const BNode &r = B.nodes[root];
    w.lox[0] = r.box.lo[0]; w.hix[0] = r.box.hi[0]; w.loy[0] = r.box.lo[1]; w.hiy[0] = r.box.hi[1]; w.loz[0] = r.box.lo[2]; w.hiz[0] = r.box.hi[2];
    w.c[0] = leaf_ref(0, r.count);
    out->nodes4.push_back(w); out->max_depth4 = 1;
    return;
  }

  // numbering: BFS over the first `top_levels` levels, then each remaining subtree in DFS preorder
  const int32_t total = B.next_node.load();
  std::vector<int32_t> newid(total, -1);
  std::vector<int32_t> emit;            // build-node ids of inner nodes in final order
  std::vector<std::pair<int32_t, int>> frontier{{root, 0}}, rest;
  size_t head = 0;
  while (head < frontier.size()) {
    const int32_t id = frontier[head].first; const int d = frontier[head].second; head++;
    if (d >= top_levels) { rest.push_back({id, d}); continue; }
    newid[id] = (int32_t)emit.size(); emit.push_back(id);
    const BNode &nd = B.nodes[id];
    if (B.nodes[nd.left].left >= 0) frontier.push_back({nd.left, d + 1});
    if (B.nodes[nd.right].left >= 0) frontier.push_back({nd.right, d + 1});
  }
  out->top_count = (int32_t)emit.size();
  int max_depth = 0;
  std::vector<std::pair<int32_t, int>> stack;
  for (size_t k = 0; k < rest.size(); k++) {
    stack.push_back(rest[k]);
    while (!stack.empty()) {
      const int32_t id = stack.back().first; const int d = stack.back().second; stack.pop_back();
      newid[id] = (int32_t)emit.size(); emit.push_back(id);
      max_depth = std::max(max_depth, d + 1);
      const BNode &nd = B.nodes[id];
      if (B.nodes[nd.right].left >= 0) stack.push_back({nd.right, d + 1});
      if (B.nodes[nd.left].left >= 0) stack.push_back({nd.left, d + 1});
    }
  }
  max_depth = std::max(max_depth, std::min(top_levels, 64));
  out->max_depth = max_depth + 1;
  out->nodes.resize(emit.size());
  for (size_t k = 0; k < emit.size(); k++) {
    const BNode &nd = B.nodes[emit[k]];
    Node64 o; memset(&o, 0, sizeof o);
    const BNode &l = B.nodes[nd.left], &r = B.nodes[nd.right];
    set_child(o, 0, l, l.left >= 0 ? newid[nd.left] : leaf_ref(l.first, l.count));
    set_child(o, 1, r, r.left >= 0 ? newid[nd.right] : leaf_ref(r.first, r.count));
    out->nodes[k] = o;
  }

  // ---- collapse to 4-wide nodes: open the inner child with the largest surface area until four slots are used
  struct Item { int32_t bnode; int32_t wide; int depth; };
  std::vector<Item> todo{{root, 0, 1}};
  out->nodes4.push_back(empty4);
  while (!todo.empty()) {
    const Item it = todo.back(); todo.pop_back();
    out->max_depth4 = std::max(out->max_depth4, it.depth);
    int32_t slots[4]; int ns = 0;
    slots[ns++] = B.nodes[it.bnode].left; slots[ns++] = B.nodes[it.bnode].right;
    while (ns < 4) {
      int best = -1; float best_area = -1.f;
      for (int k = 0; k < ns; k++) {
        const BNode &c = B.nodes[slots[k]];
        if (c.left < 0) continue;
        const float ar = box_area(c.box);
        if (ar > best_area) { best_area = ar; best = k; }
      }
      if (best < 0) break;
      const BNode c = B.nodes[slots[best]];
      slots[best] = c.left; slots[ns++] = c.right;
    }
    // slot order = tie-break of the walk (children whose entry distances are equal are visited in slot order — a ray that
    // starts inside several child boxes): smaller boxes first, they hold the nearer geometry more often (an object inside an
    // enclosing environment mesh is walked before the environment and cuts its traversal short)
    std::sort(slots, slots + ns, [&](int32_t x, int32_t y) { return box_area(B.nodes[x].box) < box_area(B.nodes[y].box); });
    Node128 w = empty4;
    for (int k = 0; k < ns; k++) {
      const BNode &c = B.nodes[slots[k]];
      w.lox[k] = c.box.lo[0]; w.hix[k] = c.box.hi[0]; w.loy[k] = c.box.lo[1]; w.hiy[k] = c.box.hi[1]; w.loz[k] = c.box.lo[2]; w.hiz[k] = c.box.hi[2];
      if (c.left < 0) w.c[k] = leaf_ref(c.first, c.count);
      else {
        w.c[k] = (int32_t)out->nodes4.size();
        out->nodes4.push_back(empty4);
        todo.push_back({slots[k], w.c[k], it.depth + 1});
      }
    }
    out->nodes4[it.wide] = w;
  }
  // Worst-case stack depth of the wavefront's walk (nearest hit child first, the other hit children pushed): entering a
  // child leaves at most (valid children - 1) siblings on the stack at every level.  Children have larger indices than
  // their parent (DFS preorder), so one backward sweep is enough.
  {
    const size_t n4 = out->nodes4.size();
    std::vector<int32_t> need(n4, 0);
    for (size_t i = n4; i-- > 0;) {
      const Node128 &w = out->nodes4[i];
      int valid = 0, deepest = 0;
      for (int k = 0; k < 4; k++) {
        if (w.lox[k] > w.hix[k] || w.lox[k] >= 3e38f) continue;
        valid++;
        if (w.c[k] >= 0) deepest = std::max(deepest, need[w.c[k]]);
      }
      need[i] = std::max(valid - 1, 0) + deepest;
    }
    out->stack_need4 = n4 ? need[0] : 0;
  }
  // Breadth-first front: the first min(n4, FJB_TOP4_MAX) nodes in level order move to the front of the array (the others
  // keep their relative order), so that a prefix of the array is the top of the tree.
  {
    const size_t n4 = out->nodes4.size();
    std::vector<int32_t> bfs; bfs.reserve(FJB_TOP4_MAX);
    if (n4) bfs.push_back(0);
    for (size_t h = 0; h < bfs.size() && bfs.size() < (size_t)FJB_TOP4_MAX; h++)
      for (int k = 0; k < 4 && bfs.size() < (size_t)FJB_TOP4_MAX; k++) { const int32_t c = out->nodes4[bfs[h]].c[k]; if (c >= 0) bfs.push_back(c); }
    std::vector<int32_t> newid(n4, -1);
    for (size_t k = 0; k < bfs.size(); k++) newid[bfs[k]] = (int32_t)k;
    int32_t next = (int32_t)bfs.size();
    for (size_t i = 0; i < n4; i++) if (newid[i] < 0) newid[i] = next++;
    std::vector<Node128> moved(n4);
    for (size_t i = 0; i < n4; i++) {
      Node128 w = out->nodes4[i];
      for (int k = 0; k < 4; k++) if (w.c[k] >= 0) w.c[k] = newid[w.c[k]];
      moved[newid[i]] = w;
    }
    out->nodes4.swap(moved);
    out->top_count4 = (int32_t)bfs.size();
  }
}

bool quantize_nodes(const Node128 *in, size_t n, NodeQ64 *out, float *bmag) {
  double mag = 0;
  for (size_t i = 0; i < n; i++)
    if (!quantize_one(in[i], out[i], &mag)) return false;
  *bmag = round_up(mag * (1.0 + 1e-6));
  return true;
}

}  // namespace fjb
