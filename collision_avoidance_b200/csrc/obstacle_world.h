// obstacle_world.h -- host-side obstacle preprocessing for the batched ORCA simulator.
//
// Replaces RVOSimulator::addObstacle + processObstacles (reference call sites
// collision_avoidence_env.py:118-123,145 ; ALAN_true.py:196-210,476), SURVEY.md A.1/A.3:
// polygons -> linked vertex ring (point, unit direction, convexity) -> BSP over the edges,
// splitting straddling edges and appending the new vertices.  It runs once per world on the
// host; the flat tables it produces are what the kernels traverse (orca_core.cuh).
#pragma once

#include <cmath>
#include <cstdint>
#include <utility>
#include <algorithm>
#include <cmath>
#include <vector>

namespace orca_host {

struct ObstacleTables {
  // per vertex (an edge is identified by its first vertex)
  std::vector<float> px, py, ux, uy;
  std::vector<int32_t> next, prev;
  std::vector<uint8_t> convex;
  // BSP nodes; node 0 is the root when non-empty
  std::vector<int32_t> node_vertex, node_left, node_right;
  int depth = 0;

  int num_vertices() const { return static_cast<int>(px.size()); }
  int num_nodes() const { return static_cast<int>(node_vertex.size()); }
};

namespace detail {

constexpr float kEps = 0.00001f;

inline float det2(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }
// leftOf(a, b, c) = det(a - c, b - a)
inline float left_of(float ax, float ay, float bx, float by, float cx, float cy) {
  return det2(ax - cx, ay - cy, bx - ax, by - ay);
}

struct Builder {
  ObstacleTables& T;
  explicit Builder(ObstacleTables& t) : T(t) {}

  using Rank = std::pair<size_t, size_t>;
  static Rank rank(size_t l, size_t r) { return {l > r ? l : r, l > r ? r : l}; }

  void classify(int ei, int ej, float* a, float* b) const {
    const int i2 = T.next[ei], j2 = T.next[ej];
    *a = left_of(T.px[ei], T.py[ei], T.px[i2], T.py[i2], T.px[ej], T.py[ej]);
    *b = left_of(T.px[ei], T.py[ei], T.px[i2], T.py[i2], T.px[j2], T.py[j2]);
  }

  int build(const std::vector<int>& edges, int level) {
    if (edges.empty()) return -1;
    if (level + 1 > T.depth) T.depth = level + 1;
    const size_t n = edges.size();
    size_t best = 0;
    Rank best_rank = rank(n, n);
    size_t best_l = n, best_r = n;
    for (size_t i = 0; i < n; ++i) {
      size_t nl = 0, nr = 0;
      for (size_t j = 0; j < n; ++j) {
        if (i == j) continue;
        float a, b;
        classify(edges[i], edges[j], &a, &b);
        if (a >= -kEps && b >= -kEps) {
          ++nl;
        } else if (a <= kEps && b <= kEps) {
          ++nr;
        } else {
          ++nl;
          ++nr;
        }
        if (rank(nl, nr) >= best_rank) break;
      }
      if (rank(nl, nr) < best_rank) {
        best_rank = rank(nl, nr);
        best_l = nl;
        best_r = nr;
        best = i;
      }
    }
    std::vector<int> lefts, rights;
    lefts.reserve(best_l);
    rights.reserve(best_r);
    const int ei = edges[best];
    for (size_t j = 0; j < n; ++j) {
      if (j == best) continue;
      const int ej = edges[j];
      float a, b;
      classify(ei, ej, &a, &b);
      if (a >= -kEps && b >= -kEps) {
        lefts.push_back(ej);
      } else if (a <= kEps && b <= kEps) {
        rights.push_back(ej);
      } else {
        // split edge ej where the supporting line of ei crosses it
        const int i2 = T.next[ei], j2 = T.next[ej];
        const float ix = T.px[i2] - T.px[ei], iy = T.py[i2] - T.py[ei];
        const float t = det2(ix, iy, T.px[ej] - T.px[ei], T.py[ej] - T.py[ei]) /
                        det2(ix, iy, T.px[ej] - T.px[j2], T.py[ej] - T.py[j2]);
        const float sx = T.px[ej] + t * (T.px[j2] - T.px[ej]);
        const float sy = T.py[ej] + t * (T.py[j2] - T.py[ej]);
        const int nid = T.num_vertices();
        T.px.push_back(sx);
        T.py.push_back(sy);
        T.ux.push_back(T.ux[ej]);
        T.uy.push_back(T.uy[ej]);
        T.next.push_back(j2);
        T.prev.push_back(ej);
        T.convex.push_back(1);
        T.next[ej] = nid;
        T.prev[j2] = nid;
        if (a > 0.f) {
          lefts.push_back(ej);
          rights.push_back(nid);
        } else {
          rights.push_back(ej);
          lefts.push_back(nid);
        }
      }
    }
    const int me = T.num_nodes();
    T.node_vertex.push_back(ei);
    T.node_left.push_back(-1);
    T.node_right.push_back(-1);
    const int l = build(lefts, level + 1);
    const int r = build(rights, level + 1);
    T.node_left[me] = l;
    T.node_right[me] = r;
    return me;
  }
};

}  // namespace detail

// Appends one polygon (>= 2 vertices) as a closed ring; returns the id of its first vertex or -1.
inline int add_polygon(ObstacleTables& T, const float* xy, int n) {
  if (n < 2) return -1;
  const int first = T.num_vertices();
  for (int i = 0; i < n; ++i) {
    const int nx = (i == n - 1) ? 0 : i + 1;
    const int pv = (i == 0) ? n - 1 : i - 1;
    const float dx = xy[2 * nx] - xy[2 * i], dy = xy[2 * nx + 1] - xy[2 * i + 1];
    const float inv = 1.0f / std::sqrt(dx * dx + dy * dy);
    T.px.push_back(xy[2 * i]);
    T.py.push_back(xy[2 * i + 1]);
    T.ux.push_back(dx * inv);
    T.uy.push_back(dy * inv);
    T.next.push_back(first + nx);
    T.prev.push_back(first + pv);
    bool cvx = true;
    if (n != 2) {
      cvx = detail::left_of(xy[2 * pv], xy[2 * pv + 1], xy[2 * i], xy[2 * i + 1], xy[2 * nx], xy[2 * nx + 1]) >= 0.f;
    }
    T.convex.push_back(cvx ? 1 : 0);
  }
  return first;
}

// processObstacles: build the BSP over every edge added so far.
inline void process(ObstacleTables& T) {
  T.node_vertex.clear();
  T.node_left.clear();
  T.node_right.clear();
  T.depth = 0;
  std::vector<int> all(static_cast<size_t>(T.num_vertices()));
  for (size_t i = 0; i < all.size(); ++i) all[i] = static_cast<int>(i);
  detail::Builder(T).build(all, 0);
}

// Obstacle-free map of a processed world: a kCullGrid x kCullGrid grid over the bounding box of all
// obstacle vertices; bit (cx, cy) is set when some edge MAY come within `range` of a point of the
// cell (distance from the cell's centre to the edge <= range + the cell's half diagonal, plus 5 % of a
// cell for the rounding of the cell lookup).  A clear bit proves that an agent in that cell has no
// obstacle neighbor, so the kernel skips the BSP walk (RVO2's result -- an empty list -- unchanged).
constexpr int kCullGrid = 32;
struct CullMap {
  uint32_t rows[kCullGrid];  // rows[cy] bit cx
  float x0, y0, inv_cx, inv_cy;
};
inline CullMap build_cull_map(const ObstacleTables& T, float range) {
  CullMap M;
  for (int i = 0; i < kCullGrid; ++i) M.rows[i] = 0xffffffffu;
  M.x0 = M.y0 = 0.f;
  M.inv_cx = M.inv_cy = 0.f;  // inv = 0: every point maps outside -> always walk
  const int V = T.num_vertices();
  if (V == 0) return M;
  double bx0 = T.px[0], bx1 = T.px[0], by0 = T.py[0], by1 = T.py[0];
  for (int v = 1; v < V; ++v) {
    bx0 = std::min<double>(bx0, T.px[v]);
    bx1 = std::max<double>(bx1, T.px[v]);
    by0 = std::min<double>(by0, T.py[v]);
    by1 = std::max<double>(by1, T.py[v]);
  }
  const double cw = (bx1 - bx0) / kCullGrid, ch = (by1 - by0) / kCullGrid;
  if (!(cw > 0.0) || !(ch > 0.0)) return M;
  const double reach = (double)range + 0.5 * std::sqrt(cw * cw + ch * ch) + 0.05 * std::max(cw, ch);
  for (int cy = 0; cy < kCullGrid; ++cy) {
    uint32_t bits = 0u;
    for (int cx = 0; cx < kCullGrid; ++cx) {
      const double qx = bx0 + (cx + 0.5) * cw, qy = by0 + (cy + 0.5) * ch;
      bool near = false;
      for (int v = 0; v < V && !near; ++v) {
        const int w = T.next[v];
        const double ax = T.px[v], ay = T.py[v], ex = T.px[w] - ax, ey = T.py[w] - ay;
        const double len2 = ex * ex + ey * ey;
        double t = len2 > 0.0 ? ((qx - ax) * ex + (qy - ay) * ey) / len2 : 0.0;
        t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
        const double dx = qx - (ax + t * ex), dy = qy - (ay + t * ey);
        near = dx * dx + dy * dy <= reach * reach;
      }
      bits |= near ? (1u << cx) : 0u;
    }
    M.rows[cy] = bits;
  }
  M.x0 = (float)bx0;
  M.y0 = (float)by0;
  M.inv_cx = (float)(1.0 / cw);
  M.inv_cy = (float)(1.0 / ch);
  return M;
}

}  // namespace orca_host
