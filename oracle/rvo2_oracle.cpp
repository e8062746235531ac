// rvo2_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// A from-scratch scalar float32 restatement of the RVO2 Library semantics that
// the reference reaches through `import rvo2` (Python-RVO2):
//   /root/reference/collision_avoidance/envs/collision_avoidence_env.py:16,62-68,126-148,385,448
//   /root/reference/collision_avoidance/ALAN/ALAN_true.py:6,22-28,461-479,601,632
// The RVO2 sources themselves are a third-party dependency that is NOT vendored
// in /root/reference and is not installed in the build container (no version is
// pinned anywhere in the reference: README.md:11,39, setup.py:5).  This file
// therefore restates the published RVO2 v2.0.x algorithm (van den Berg et al.,
// "Reciprocal n-body collision avoidance") and is anchored
// on analytic known-answers + invariants (tests/test_oracle_*.py).
//
// PARITY UNPINNED vs. upstream rvo2 (module unavailable); pinned vs. analytic
// known-answer vectors and invariants.  See DESIGN.md "Oracle".
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product path never does.
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math  (no FMA contraction, so every
// float op rounds exactly once, which is what a stock x86-64 build of RVO2 does).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <memory>
#include <thread>
#include <utility>
#include <vector>

namespace {

constexpr float kEps = 0.00001f;  // RVO_EPSILON
constexpr int kMaxLeaf = 10;      // kd-tree MAX_LEAF_SIZE

struct V2 {
  float x = 0.f, y = 0.f;
  V2() = default;
  V2(float x_, float y_) : x(x_), y(y_) {}
};
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 operator-(V2 a) { return {-a.x, -a.y}; }
inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
inline V2 operator*(float s, V2 a) { return {s * a.x, s * a.y}; }
inline V2 operator*(V2 a, float s) { return {a.x * s, a.y * s}; }
// RVO2's vector/scalar division multiplies by the reciprocal (Appendix A helpers).
inline V2 operator/(V2 a, float s) {
  const float inv = 1.0f / s;
  return {a.x * inv, a.y * inv};
}
inline float det(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
inline float abs_sq(V2 a) { return dot(a, a); }
inline float vabs(V2 a) { return std::sqrt(dot(a, a)); }
inline V2 unit(V2 a) { return a / vabs(a); }
inline float sqr(float s) { return s * s; }
inline float left_of(V2 a, V2 b, V2 c) { return det(a - c, b - a); }
inline float dist_sq_point_segment(V2 a, V2 b, V2 c) {
  const float r = dot(c - a, b - a) / abs_sq(b - a);
  if (r < 0.f) return abs_sq(c - a);
  if (r > 1.f) return abs_sq(c - b);
  return abs_sq(c - (a + r * (b - a)));
}

struct Line {
  V2 point, dir;
};

struct ObstVertex {
  V2 point, unit_dir;
  int next = -1, prev = -1;
  bool convex = false;
};

struct AgentRec {
  V2 pos, vel, pref, new_vel;
  float nd = 0, th = 0, tho = 0, radius = 0, vmax = 0;
  int k = 0;
  std::vector<std::pair<float, int>> agent_nbrs;  // (distSq, agent id) ascending
  std::vector<std::pair<float, int>> obst_nbrs;   // (distSq, vertex id) ascending
  std::vector<Line> lines;
  int n_obst_lines = 0;
};

struct KdNode {
  int begin, end, left, right;
  float max_x, max_y, min_x, min_y;
};

struct BspNode {
  int vertex = -1;
  int left = -1, right = -1;
};

// ---- linear programs (Appendix A.6) ---------------------------------------
bool lp1(const std::vector<Line>& L, size_t i, float radius, V2 opt, bool dir_opt, V2& result) {
  const float dp = dot(L[i].point, L[i].dir);
  const float disc = sqr(dp) + sqr(radius) - abs_sq(L[i].point);
  if (disc < 0.f) return false;
  const float sq = std::sqrt(disc);
  float t_lo = -dp - sq;
  float t_hi = -dp + sq;
  for (size_t j = 0; j < i; ++j) {
    const float den = det(L[i].dir, L[j].dir);
    const float num = det(L[j].dir, L[i].point - L[j].point);
    if (std::fabs(den) <= kEps) {
      if (num < 0.f) return false;
      continue;
    }
    const float t = num / den;
    if (den >= 0.f)
      t_hi = std::min(t_hi, t);
    else
      t_lo = std::max(t_lo, t);
    if (t_lo > t_hi) return false;
  }
  if (dir_opt) {
    if (dot(opt, L[i].dir) > 0.f)
      result = L[i].point + t_hi * L[i].dir;
    else
      result = L[i].point + t_lo * L[i].dir;
  } else {
    const float t = dot(L[i].dir, opt - L[i].point);
    if (t < t_lo)
      result = L[i].point + t_lo * L[i].dir;
    else if (t > t_hi)
      result = L[i].point + t_hi * L[i].dir;
    else
      result = L[i].point + t * L[i].dir;
  }
  return true;
}

size_t lp2(const std::vector<Line>& L, float radius, V2 opt, bool dir_opt, V2& result) {
  if (dir_opt)
    result = opt * radius;
  else if (abs_sq(opt) > sqr(radius))
    result = unit(opt) * radius;
  else
    result = opt;
  for (size_t i = 0; i < L.size(); ++i) {
    if (det(L[i].dir, L[i].point - result) > 0.f) {
      const V2 keep = result;
      if (!lp1(L, i, radius, opt, dir_opt, result)) {
        result = keep;
        return i;
      }
    }
  }
  return L.size();
}

void lp3(const std::vector<Line>& L, size_t n_obst, size_t begin, float radius, V2& result) {
  float distance = 0.f;
  for (size_t i = begin; i < L.size(); ++i) {
    if (det(L[i].dir, L[i].point - result) > distance) {
      std::vector<Line> proj(L.begin(), L.begin() + static_cast<std::ptrdiff_t>(n_obst));
      for (size_t j = n_obst; j < i; ++j) {
        Line ln;
        const float d = det(L[i].dir, L[j].dir);
        if (std::fabs(d) <= kEps) {
          if (dot(L[i].dir, L[j].dir) > 0.f) continue;
          ln.point = 0.5f * (L[i].point + L[j].point);
        } else {
          ln.point = L[i].point + (det(L[j].dir, L[i].point - L[j].point) / d) * L[i].dir;
        }
        ln.dir = unit(L[j].dir - L[i].dir);
        proj.push_back(ln);
      }
      const V2 keep = result;
      if (lp2(proj, radius, V2(-L[i].dir.y, L[i].dir.x), true, result) < proj.size()) result = keep;
      distance = det(L[i].dir, L[i].point - result);
    }
  }
}

// ---- simulator ---------------------------------------------------------------
struct Sim {
  float dt, d_nd, d_th, d_tho, d_radius, d_vmax;
  int d_k;
  V2 d_vel;
  float global_time = 0.f;
  std::vector<AgentRec> agents;
  std::vector<ObstVertex> verts;
  // agent kd-tree (persistent permutation, Appendix A.2)
  std::vector<int> perm;
  std::vector<KdNode> kd;
  // obstacle BSP (Appendix A.3)
  std::vector<BspNode> bsp;
  int bsp_root = -1;

  int add_agent(V2 p, float nd, int k, float th, float tho, float r, float vmax, V2 v) {
    AgentRec a;
    a.pos = p;
    a.vel = v;
    a.nd = nd;
    a.k = k;
    a.th = th;
    a.tho = tho;
    a.radius = r;
    a.vmax = vmax;
    agents.push_back(std::move(a));
    return static_cast<int>(agents.size()) - 1;
  }

  int add_obstacle(const std::vector<V2>& pts) {
    const int n = static_cast<int>(pts.size());
    if (n < 2) return -1;
    const int first = static_cast<int>(verts.size());
    for (int i = 0; i < n; ++i) {
      ObstVertex v;
      v.point = pts[i];
      const int self = first + i;
      if (i != 0) {
        v.prev = self - 1;
        verts[self - 1].next = self;
      }
      if (i == n - 1) {
        v.next = first;
        // first vertex's prev is fixed up after push (single vertex list guarded by n>=2)
      }
      const V2 nxt = pts[i == n - 1 ? 0 : i + 1];
      v.unit_dir = unit(nxt - pts[i]);
      if (n == 2)
        v.convex = true;
      else
        v.convex = left_of(pts[i == 0 ? n - 1 : i - 1], pts[i], nxt) >= 0.f;
      verts.push_back(v);
      if (i == n - 1) verts[first].prev = self;
    }
    return first;
  }

  // --- obstacle BSP build; may split edges and append vertices (A.3)
  int build_bsp(const std::vector<int>& edges) {
    if (edges.empty()) return -1;
    const size_t n = edges.size();
    size_t best = 0, min_l = n, min_r = n;
    for (size_t i = 0; i < n; ++i) {
      size_t nl = 0, nr = 0;
      const V2 i1 = verts[edges[i]].point;
      const V2 i2 = verts[verts[edges[i]].next].point;
      for (size_t j = 0; j < n; ++j) {
        if (i == j) continue;
        const V2 j1 = verts[edges[j]].point;
        const V2 j2 = verts[verts[edges[j]].next].point;
        const float a = left_of(i1, i2, j1);
        const float b = left_of(i1, i2, j2);
        if (a >= -kEps && b >= -kEps)
          ++nl;
        else if (a <= kEps && b <= kEps)
          ++nr;
        else {
          ++nl;
          ++nr;
        }
        if (std::make_pair(std::max(nl, nr), std::min(nl, nr)) >=
            std::make_pair(std::max(min_l, min_r), std::min(min_l, min_r)))
          break;
      }
      if (std::make_pair(std::max(nl, nr), std::min(nl, nr)) <
          std::make_pair(std::max(min_l, min_r), std::min(min_l, min_r))) {
        min_l = nl;
        min_r = nr;
        best = i;
      }
    }
    std::vector<int> lefts, rights;
    lefts.reserve(min_l);
    rights.reserve(min_r);
    const int e_i = edges[best];
    for (size_t j = 0; j < n; ++j) {
      if (j == best) continue;
      const int e_j = edges[j];
      const V2 i1 = verts[e_i].point;
      const V2 i2 = verts[verts[e_i].next].point;
      const int j2_id = verts[e_j].next;
      const V2 j1 = verts[e_j].point;
      const V2 j2 = verts[j2_id].point;
      const float a = left_of(i1, i2, j1);
      const float b = left_of(i1, i2, j2);
      if (a >= -kEps && b >= -kEps) {
        lefts.push_back(e_j);
      } else if (a <= kEps && b <= kEps) {
        rights.push_back(e_j);
      } else {
        const float t = det(i2 - i1, j1 - i1) / det(i2 - i1, j1 - j2);
        ObstVertex nv;
        nv.point = j1 + t * (j2 - j1);
        nv.prev = e_j;
        nv.next = j2_id;
        nv.convex = true;
        nv.unit_dir = verts[e_j].unit_dir;
        const int nid = static_cast<int>(verts.size());
        verts.push_back(nv);
        verts[e_j].next = nid;
        verts[j2_id].prev = nid;
        if (a > 0.f) {
          lefts.push_back(e_j);
          rights.push_back(nid);
        } else {
          rights.push_back(e_j);
          lefts.push_back(nid);
        }
      }
    }
    const int me = static_cast<int>(bsp.size());
    bsp.emplace_back();
    bsp[me].vertex = e_i;
    const int l = build_bsp(lefts);
    const int r = build_bsp(rights);
    bsp[me].left = l;
    bsp[me].right = r;
    return me;
  }

  void process_obstacles() {
    bsp.clear();
    std::vector<int> all(verts.size());
    for (size_t i = 0; i < verts.size(); ++i) all[i] = static_cast<int>(i);
    bsp_root = build_bsp(all);
  }

  // --- agent kd-tree
  void build_kd_rec(int begin, int end, int node) {
    KdNode& nd = kd[node];
    nd.begin = begin;
    nd.end = end;
    nd.min_x = nd.max_x = agents[perm[begin]].pos.x;
    nd.min_y = nd.max_y = agents[perm[begin]].pos.y;
    for (int i = begin + 1; i < end; ++i) {
      const V2 p = agents[perm[i]].pos;
      nd.max_x = std::max(nd.max_x, p.x);
      nd.min_x = std::min(nd.min_x, p.x);
      nd.max_y = std::max(nd.max_y, p.y);
      nd.min_y = std::min(nd.min_y, p.y);
    }
    if (end - begin > kMaxLeaf) {
      const bool vertical = (nd.max_x - nd.min_x > nd.max_y - nd.min_y);
      const float split = vertical ? 0.5f * (nd.max_x + nd.min_x) : 0.5f * (nd.max_y + nd.min_y);
      int l = begin, r = end;
      while (l < r) {
        while (l < r && (vertical ? agents[perm[l]].pos.x : agents[perm[l]].pos.y) < split) ++l;
        while (r > l && (vertical ? agents[perm[r - 1]].pos.x : agents[perm[r - 1]].pos.y) >= split) --r;
        if (l < r) {
          std::swap(perm[l], perm[r - 1]);
          ++l;
          --r;
        }
      }
      if (l == begin) {
        ++l;
        ++r;
      }
      const int left_node = node + 1;
      const int right_node = node + 2 * (l - begin);
      kd[node].left = left_node;
      kd[node].right = right_node;
      build_kd_rec(begin, l, left_node);
      build_kd_rec(l, end, right_node);
    }
  }

  void build_kd() {
    const int n = static_cast<int>(agents.size());
    if (static_cast<int>(perm.size()) < n) {
      for (int i = static_cast<int>(perm.size()); i < n; ++i) perm.push_back(i);
      kd.assign(static_cast<size_t>(2 * n > 1 ? 2 * n - 1 : 1), KdNode{});
    }
    if (n > 0) build_kd_rec(0, n, 0);
  }

  void insert_agent_nbr(AgentRec& a, int self, int other, float& range_sq) {
    if (self == other) return;
    const float d = abs_sq(a.pos - agents[other].pos);
    if (d < range_sq) {
      auto& nb = a.agent_nbrs;
      if (static_cast<int>(nb.size()) < a.k) nb.emplace_back(d, other);
      size_t i = nb.size() - 1;
      while (i != 0 && d < nb[i - 1].first) {
        nb[i] = nb[i - 1];
        --i;
      }
      nb[i] = std::make_pair(d, other);
      if (static_cast<int>(nb.size()) == a.k) range_sq = nb.back().first;
    }
  }

  void query_kd(AgentRec& a, int self, float& range_sq, int node) {
    const KdNode& nd = kd[node];
    if (nd.end - nd.begin <= kMaxLeaf) {
      for (int i = nd.begin; i < nd.end; ++i) insert_agent_nbr(a, self, perm[i], range_sq);
      return;
    }
    const KdNode& L = kd[nd.left];
    const KdNode& R = kd[nd.right];
    const float dl = sqr(std::max(0.f, L.min_x - a.pos.x)) + sqr(std::max(0.f, a.pos.x - L.max_x)) +
                     sqr(std::max(0.f, L.min_y - a.pos.y)) + sqr(std::max(0.f, a.pos.y - L.max_y));
    const float dr = sqr(std::max(0.f, R.min_x - a.pos.x)) + sqr(std::max(0.f, a.pos.x - R.max_x)) +
                     sqr(std::max(0.f, R.min_y - a.pos.y)) + sqr(std::max(0.f, a.pos.y - R.max_y));
    if (dl < dr) {
      if (dl < range_sq) {
        query_kd(a, self, range_sq, nd.left);
        if (dr < range_sq) query_kd(a, self, range_sq, nd.right);
      }
    } else {
      if (dr < range_sq) {
        query_kd(a, self, range_sq, nd.right);
        if (dl < range_sq) query_kd(a, self, range_sq, nd.left);
      }
    }
  }

  void insert_obst_nbr(AgentRec& a, int vid, float range_sq) {
    const float d = dist_sq_point_segment(verts[vid].point, verts[verts[vid].next].point, a.pos);
    if (d < range_sq) {
      auto& nb = a.obst_nbrs;
      nb.emplace_back(d, vid);
      size_t i = nb.size() - 1;
      while (i != 0 && d < nb[i - 1].first) {
        nb[i] = nb[i - 1];
        --i;
      }
      nb[i] = std::make_pair(d, vid);
    }
  }

  void query_bsp(AgentRec& a, float range_sq, int node) {
    if (node < 0) return;
    const int v1 = bsp[node].vertex;
    const V2 p1 = verts[v1].point;
    const V2 p2 = verts[verts[v1].next].point;
    const float side = left_of(p1, p2, a.pos);
    query_bsp(a, range_sq, side >= 0.f ? bsp[node].left : bsp[node].right);
    const float d_line = sqr(side) / abs_sq(p2 - p1);
    if (d_line < range_sq) {
      if (side < 0.f) insert_obst_nbr(a, v1, range_sq);
      query_bsp(a, range_sq, side >= 0.f ? bsp[node].right : bsp[node].left);
    }
  }

  void compute_neighbors(int self) {
    AgentRec& a = agents[self];
    a.obst_nbrs.clear();
    float range_sq = sqr(a.tho * a.vmax + a.radius);
    query_bsp(a, range_sq, bsp_root);
    a.agent_nbrs.clear();
    if (a.k > 0) {
      range_sq = sqr(a.nd);
      query_kd(a, self, range_sq, 0);
    }
  }

  // Appendix A.5
  void compute_new_velocity(int self) {
    AgentRec& a = agents[self];
    auto& L = a.lines;
    L.clear();
    const float inv_tho = 1.0f / a.tho;
    for (const auto& on : a.obst_nbrs) {
      int o1 = on.second;
      int o2 = verts[o1].next;
      const V2 rp1 = verts[o1].point - a.pos;
      const V2 rp2 = verts[o2].point - a.pos;
      bool covered = false;
      for (const Line& ln : L) {
        if (det(inv_tho * rp1 - ln.point, ln.dir) - inv_tho * a.radius >= -kEps &&
            det(inv_tho * rp2 - ln.point, ln.dir) - inv_tho * a.radius >= -kEps) {
          covered = true;
          break;
        }
      }
      if (covered) continue;
      const float d1 = abs_sq(rp1), d2 = abs_sq(rp2), r_sq = sqr(a.radius);
      const V2 ov = verts[o2].point - verts[o1].point;
      const float s = dot(-rp1, ov) / abs_sq(ov);
      const float d_line = abs_sq(-rp1 - s * ov);
      Line ln;
      if (s < 0.f && d1 <= r_sq) {
        if (verts[o1].convex) {
          ln.point = V2(0.f, 0.f);
          ln.dir = unit(V2(-rp1.y, rp1.x));
          L.push_back(ln);
        }
        continue;
      } else if (s > 1.f && d2 <= r_sq) {
        if (verts[o2].convex && det(rp2, verts[o2].unit_dir) >= 0.f) {
          ln.point = V2(0.f, 0.f);
          ln.dir = unit(V2(-rp2.y, rp2.x));
          L.push_back(ln);
        }
        continue;
      } else if (s >= 0.f && s < 1.f && d_line <= r_sq) {
        ln.point = V2(0.f, 0.f);
        ln.dir = -verts[o1].unit_dir;
        L.push_back(ln);
        continue;
      }
      V2 left_leg, right_leg;
      if (s < 0.f && d_line <= r_sq) {
        if (!verts[o1].convex) continue;
        o2 = o1;
        const float leg1 = std::sqrt(d1 - r_sq);
        left_leg = V2(rp1.x * leg1 - rp1.y * a.radius, rp1.x * a.radius + rp1.y * leg1) / d1;
        right_leg = V2(rp1.x * leg1 + rp1.y * a.radius, -rp1.x * a.radius + rp1.y * leg1) / d1;
      } else if (s > 1.f && d_line <= r_sq) {
        if (!verts[o2].convex) continue;
        o1 = o2;
        const float leg2 = std::sqrt(d2 - r_sq);
        left_leg = V2(rp2.x * leg2 - rp2.y * a.radius, rp2.x * a.radius + rp2.y * leg2) / d2;
        right_leg = V2(rp2.x * leg2 + rp2.y * a.radius, -rp2.x * a.radius + rp2.y * leg2) / d2;
      } else {
        if (verts[o1].convex) {
          const float leg1 = std::sqrt(d1 - r_sq);
          left_leg = V2(rp1.x * leg1 - rp1.y * a.radius, rp1.x * a.radius + rp1.y * leg1) / d1;
        } else {
          left_leg = -verts[o1].unit_dir;
        }
        if (verts[o2].convex) {
          const float leg2 = std::sqrt(d2 - r_sq);
          right_leg = V2(rp2.x * leg2 + rp2.y * a.radius, -rp2.x * a.radius + rp2.y * leg2) / d2;
        } else {
          right_leg = verts[o1].unit_dir;
        }
      }
      const int left_nbr = verts[o1].prev;
      bool left_foreign = false, right_foreign = false;
      if (verts[o1].convex && det(left_leg, -verts[left_nbr].unit_dir) >= 0.f) {
        left_leg = -verts[left_nbr].unit_dir;
        left_foreign = true;
      }
      if (verts[o2].convex && det(right_leg, verts[o2].unit_dir) <= 0.f) {
        right_leg = verts[o2].unit_dir;
        right_foreign = true;
      }
      const V2 lc = inv_tho * (verts[o1].point - a.pos);
      const V2 rc = inv_tho * (verts[o2].point - a.pos);
      const V2 cv = rc - lc;
      const float t = (o1 == o2) ? 0.5f : dot(a.vel - lc, cv) / abs_sq(cv);
      const float t_left = dot(a.vel - lc, left_leg);
      const float t_right = dot(a.vel - rc, right_leg);
      if ((t < 0.f && t_left < 0.f) || (o1 == o2 && t_left < 0.f && t_right < 0.f)) {
        const V2 uw = unit(a.vel - lc);
        ln.dir = V2(uw.y, -uw.x);
        ln.point = lc + a.radius * inv_tho * uw;
        L.push_back(ln);
        continue;
      } else if (t > 1.f && t_right < 0.f) {
        const V2 uw = unit(a.vel - rc);
        ln.dir = V2(uw.y, -uw.x);
        ln.point = rc + a.radius * inv_tho * uw;
        L.push_back(ln);
        continue;
      }
      const float inf = std::numeric_limits<float>::infinity();
      const float dc = (t < 0.f || t > 1.f || o1 == o2) ? inf : abs_sq(a.vel - (lc + t * cv));
      const float dl = (t_left < 0.f) ? inf : abs_sq(a.vel - (lc + t_left * left_leg));
      const float dr = (t_right < 0.f) ? inf : abs_sq(a.vel - (rc + t_right * right_leg));
      if (dc <= dl && dc <= dr) {
        ln.dir = -verts[o1].unit_dir;
        ln.point = lc + a.radius * inv_tho * V2(-ln.dir.y, ln.dir.x);
        L.push_back(ln);
        continue;
      } else if (dl <= dr) {
        if (left_foreign) continue;
        ln.dir = left_leg;
        ln.point = lc + a.radius * inv_tho * V2(-ln.dir.y, ln.dir.x);
        L.push_back(ln);
        continue;
      } else {
        if (right_foreign) continue;
        ln.dir = -right_leg;
        ln.point = rc + a.radius * inv_tho * V2(-ln.dir.y, ln.dir.x);
        L.push_back(ln);
        continue;
      }
    }
    const size_t n_obst = L.size();
    a.n_obst_lines = static_cast<int>(n_obst);
    const float inv_th = 1.0f / a.th;
    for (const auto& an : a.agent_nbrs) {
      const AgentRec& o = agents[an.second];
      const V2 rp = o.pos - a.pos;
      const V2 rv = a.vel - o.vel;
      const float d_sq = abs_sq(rp);
      const float cr = a.radius + o.radius;
      const float cr_sq = sqr(cr);
      Line ln;
      V2 u;
      if (d_sq > cr_sq) {
        const V2 w = rv - inv_th * rp;
        const float w_sq = abs_sq(w);
        const float dp1 = dot(w, rp);
        if (dp1 < 0.f && sqr(dp1) > cr_sq * w_sq) {
          const float wl = std::sqrt(w_sq);
          const V2 uw = w / wl;
          ln.dir = V2(uw.y, -uw.x);
          u = (cr * inv_th - wl) * uw;
        } else {
          const float leg = std::sqrt(d_sq - cr_sq);
          if (det(rp, w) > 0.f)
            ln.dir = V2(rp.x * leg - rp.y * cr, rp.x * cr + rp.y * leg) / d_sq;
          else
            ln.dir = -(V2(rp.x * leg + rp.y * cr, -rp.x * cr + rp.y * leg) / d_sq);
          const float dp2 = dot(rv, ln.dir);
          u = dp2 * ln.dir - rv;
        }
      } else {
        const float inv_dt = 1.0f / dt;
        const V2 w = rv - inv_dt * rp;
        const float wl = vabs(w);
        const V2 uw = w / wl;
        ln.dir = V2(uw.y, -uw.x);
        u = (cr * inv_dt - wl) * uw;
      }
      ln.point = a.vel + 0.5f * u;
      L.push_back(ln);
    }
    const size_t fail = lp2(L, a.vmax, a.pref, false, a.new_vel);
    if (fail < L.size()) lp3(L, n_obst, fail, a.vmax, a.new_vel);
  }

  void do_step() {
    build_kd();
    const int n = static_cast<int>(agents.size());
    for (int i = 0; i < n; ++i) {
      compute_neighbors(i);
      compute_new_velocity(i);
    }
    for (int i = 0; i < n; ++i) {
      AgentRec& a = agents[i];
      a.vel = a.new_vel;
      a.pos = a.pos + a.vel * dt;
    }
    global_time += dt;
  }
};

inline Sim* S(void* h) { return static_cast<Sim*>(h); }

}  // namespace

extern "C" {

void* rvo_create(float dt, float nd, int k, float th, float tho, float radius, float vmax, float vx, float vy) {
  Sim* s = new Sim();
  s->dt = dt;
  s->d_nd = nd;
  s->d_k = k;
  s->d_th = th;
  s->d_tho = tho;
  s->d_radius = radius;
  s->d_vmax = vmax;
  s->d_vel = V2(vx, vy);
  return s;
}
void rvo_destroy(void* h) { delete S(h); }

int rvo_add_agent(void* h, float x, float y, float nd, int k, float th, float tho, float radius, float vmax, float vx,
                  float vy) {
  return S(h)->add_agent(V2(x, y), nd, k, th, tho, radius, vmax, V2(vx, vy));
}
int rvo_add_agent_default(void* h, float x, float y) {
  Sim* s = S(h);
  return s->add_agent(V2(x, y), s->d_nd, s->d_k, s->d_th, s->d_tho, s->d_radius, s->d_vmax, s->d_vel);
}
int rvo_add_obstacle(void* h, const float* xy, int n) {
  std::vector<V2> pts(static_cast<size_t>(n > 0 ? n : 0));
  for (int i = 0; i < n; ++i) pts[i] = V2(xy[2 * i], xy[2 * i + 1]);
  return S(h)->add_obstacle(pts);
}
void rvo_process_obstacles(void* h) { S(h)->process_obstacles(); }
void rvo_do_step(void* h) { S(h)->do_step(); }
int rvo_num_agents(void* h) { return static_cast<int>(S(h)->agents.size()); }
int rvo_num_obstacle_vertices(void* h) { return static_cast<int>(S(h)->verts.size()); }
float rvo_global_time(void* h) { return S(h)->global_time; }

void rvo_set_agent_pref_velocity(void* h, int i, float x, float y) { S(h)->agents[i].pref = V2(x, y); }
void rvo_set_agent_position(void* h, int i, float x, float y) { S(h)->agents[i].pos = V2(x, y); }
void rvo_set_agent_velocity(void* h, int i, float x, float y) { S(h)->agents[i].vel = V2(x, y); }
void rvo_get_agent_position(void* h, int i, float* out) {
  out[0] = S(h)->agents[i].pos.x;
  out[1] = S(h)->agents[i].pos.y;
}
void rvo_get_agent_velocity(void* h, int i, float* out) {
  out[0] = S(h)->agents[i].vel.x;
  out[1] = S(h)->agents[i].vel.y;
}
void rvo_get_agent_pref_velocity(void* h, int i, float* out) {
  out[0] = S(h)->agents[i].pref.x;
  out[1] = S(h)->agents[i].pref.y;
}
int rvo_get_agent_num_agent_neighbors(void* h, int i) { return static_cast<int>(S(h)->agents[i].agent_nbrs.size()); }
int rvo_get_agent_agent_neighbor(void* h, int i, int j) { return S(h)->agents[i].agent_nbrs[j].second; }
float rvo_get_agent_agent_neighbor_distsq(void* h, int i, int j) { return S(h)->agents[i].agent_nbrs[j].first; }
int rvo_get_agent_num_obstacle_neighbors(void* h, int i) { return static_cast<int>(S(h)->agents[i].obst_nbrs.size()); }
int rvo_get_agent_obstacle_neighbor(void* h, int i, int j) { return S(h)->agents[i].obst_nbrs[j].second; }
float rvo_get_agent_obstacle_neighbor_distsq(void* h, int i, int j) { return S(h)->agents[i].obst_nbrs[j].first; }
int rvo_get_next_obstacle_vertex_no(void* h, int v) { return S(h)->verts[v].next; }
int rvo_get_prev_obstacle_vertex_no(void* h, int v) { return S(h)->verts[v].prev; }
void rvo_get_obstacle_vertex(void* h, int v, float* out) {
  out[0] = S(h)->verts[v].point.x;
  out[1] = S(h)->verts[v].point.y;
}
int rvo_get_obstacle_vertex_convex(void* h, int v) { return S(h)->verts[v].convex ? 1 : 0; }
void rvo_get_obstacle_vertex_unit_dir(void* h, int v, float* out) {
  out[0] = S(h)->verts[v].unit_dir.x;
  out[1] = S(h)->verts[v].unit_dir.y;
}
int rvo_get_agent_num_orca_lines(void* h, int i) { return static_cast<int>(S(h)->agents[i].lines.size()); }
int rvo_get_agent_num_obst_orca_lines(void* h, int i) { return S(h)->agents[i].n_obst_lines; }
void rvo_get_agent_orca_line(void* h, int i, int j, float* out) {
  const Line& l = S(h)->agents[i].lines[j];
  out[0] = l.point.x;
  out[1] = l.point.y;
  out[2] = l.dir.x;
  out[3] = l.dir.y;
}

// ---- bulk accessors (test convenience; same state, arrays of [n][2]) ----------
// addAgent(pos) with the simulator's default parameters for n agents; vel = initial velocities
int rvo_add_agents(void* h, const float* pos, const float* vel, int n) {
  Sim* s = S(h);
  int last = -1;
  for (int i = 0; i < n; ++i)
    last = s->add_agent(V2(pos[2 * i], pos[2 * i + 1]), s->d_nd, s->d_k, s->d_th, s->d_tho, s->d_radius, s->d_vmax,
                        V2(vel[2 * i], vel[2 * i + 1]));
  return last;
}
void rvo_get_positions(void* h, float* out) {
  Sim* s = S(h);
  for (size_t i = 0; i < s->agents.size(); ++i) {
    out[2 * i] = s->agents[i].pos.x;
    out[2 * i + 1] = s->agents[i].pos.y;
  }
}
void rvo_get_velocities(void* h, float* out) {
  Sim* s = S(h);
  for (size_t i = 0; i < s->agents.size(); ++i) {
    out[2 * i] = s->agents[i].vel.x;
    out[2 * i + 1] = s->agents[i].vel.y;
  }
}
void rvo_set_positions(void* h, const float* in) {
  Sim* s = S(h);
  for (size_t i = 0; i < s->agents.size(); ++i) s->agents[i].pos = V2(in[2 * i], in[2 * i + 1]);
}
void rvo_set_velocities(void* h, const float* in) {
  Sim* s = S(h);
  for (size_t i = 0; i < s->agents.size(); ++i) s->agents[i].vel = V2(in[2 * i], in[2 * i + 1]);
}
void rvo_set_pref_velocities(void* h, const float* in) {
  Sim* s = S(h);
  for (size_t i = 0; i < s->agents.size(); ++i) s->agents[i].pref = V2(in[2 * i], in[2 * i + 1]);
}

// Standalone LP entry for known-answer tests: lines = [n][4] (point.xy, dir.xy).
// Returns the index of the failing line from LP2 (n on success) after running the
// same LP2 -> LP3 sequence as computeNewVelocity.
int rvo_solve_lp(const float* lines, int n, int n_obst, float radius, float pref_x, float pref_y, float* out) {
  std::vector<Line> L(static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) {
    L[i].point = V2(lines[4 * i], lines[4 * i + 1]);
    L[i].dir = V2(lines[4 * i + 2], lines[4 * i + 3]);
  }
  V2 res;
  const size_t fail = lp2(L, radius, V2(pref_x, pref_y), false, res);
  if (fail < L.size()) lp3(L, static_cast<size_t>(n_obst), fail, radius, res);
  out[0] = res.x;
  out[1] = res.y;
  return static_cast<int>(fail);
}

// ---- CPU baseline driver (bench.py cpu_baseline / --impl reference only) -------
// Steps `n` independent simulators `steps` times with the reference's orca_step
// policy (doStep, then goal-directed preferred velocity in float64 as the shell
// does: ALAN_true.py:483-495,631-633), envs split over `threads` host threads.
// goals: [n][agents][2] float64.
void rvo_batch_orca_steps(void** sims, int n, const double* goals, int agents_per_sim, int steps, int threads) {
  std::atomic<int> next{0};
  auto work = [&]() {
    for (;;) {
      const int e = next.fetch_add(1);
      if (e >= n) break;
      Sim* s = S(sims[e]);
      const double* g = goals + static_cast<size_t>(e) * agents_per_sim * 2;
      for (int t = 0; t < steps; ++t) {
        s->do_step();
        for (int i = 0; i < agents_per_sim; ++i) {
          const double ang = std::atan2(g[2 * i + 1] - static_cast<double>(s->agents[i].pos.y),
                                        g[2 * i] - static_cast<double>(s->agents[i].pos.x));
          s->agents[i].pref = V2(static_cast<float>(std::cos(ang)), static_cast<float>(std::sin(ang)));
        }
      }
    }
  };
  if (threads <= 1) {
    work();
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) pool.emplace_back(work);
  for (auto& th : pool) th.join();
}

}  // extern "C"
