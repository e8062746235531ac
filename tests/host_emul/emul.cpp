// emul.cpp -- TEST-ONLY host emulation of the fused step kernel.
//
// Compiles the library's own per-agent device code (csrc/orca_core.cuh, agent_step_body in
// csrc/orca_step_small.cuh) with g++ and runs it serially over a batch.  It lets the CPU test
// suite check the kernel LOGIC against the oracle without a GPU.  It is not a fallback: the
// package never loads it, and the GPU parity tests run the real CUDA path.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

static inline void sincosf_(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
#define sincosf sincosf_
#include "../../collision_avoidance_b200/csrc/obstacle_world.h"
#include "../../collision_avoidance_b200/csrc/orca_step_small.cuh"
#include "../../collision_avoidance_b200/csrc/orca_obs.cuh"
#include "../../collision_avoidance_b200/csrc/orca_grid.cuh"

namespace {
template <int K, bool KFULL>
void run_k(const orca::StepArgs& a, int policy) {
  const int E = a.E, N = a.N;
  std::vector<float2> spos((size_t)N), svel((size_t)N);
  std::vector<float4> lines((size_t)(K + ORCA_MAX_OBST_LINES));
  for (int e = 0; e < E; ++e) {
    for (int i = 0; i < N; ++i) {
      spos[(size_t)i] = a.pos[(size_t)e * N + i];
      svel[(size_t)i] = a.vel[(size_t)e * N + i];
    }
    const int estep = a.env_step ? a.env_step[e] : 0;
    // host twin of build_tile_grid (orca_step_small.cuh): same cells, ids inside a cell in
    // DESCENDING id order here -- any order is legal, and this one differs from the id scan
    std::vector<unsigned short> cell_start(orca::kTileCells + 1, 0);
    std::vector<unsigned char> sorted((size_t)N);
    std::vector<int> cell_of((size_t)N);
    const bool tile_grid = a.tile_grid_inv_cell > 0.f;
    if (tile_grid) {
      float ox = INFINITY, oy = INFINITY;
      for (int i = 0; i < N; ++i) { ox = std::fmin(ox, spos[(size_t)i].x); oy = std::fmin(oy, spos[(size_t)i].y); }
      std::vector<int> cnt(orca::kTileCells, 0);
      for (int i = 0; i < N; ++i) {
        int cx = (int)floorf((spos[(size_t)i].x - ox) * a.tile_grid_inv_cell);
        int cy = (int)floorf((spos[(size_t)i].y - oy) * a.tile_grid_inv_cell);
        cx = cx < 0 ? 0 : (cx >= orca::kTileGrid ? orca::kTileGrid - 1 : cx);
        cy = cy < 0 ? 0 : (cy >= orca::kTileGrid ? orca::kTileGrid - 1 : cy);
        cell_of[(size_t)i] = cy * orca::kTileGrid + cx;
        cnt[cell_of[(size_t)i]]++;
      }
      for (int c = 0; c < orca::kTileCells; ++c) cell_start[c + 1] = (unsigned short)(cell_start[c] + cnt[c]);
      std::vector<int> fill(orca::kTileCells, 0);
      for (int i = N - 1; i >= 0; --i) sorted[cell_start[cell_of[(size_t)i]] + fill[cell_of[(size_t)i]]++] = (unsigned char)i;
    }
    for (int i = 0; i < N; ++i) {
      orca::Lines L;
      L.base = lines.data();
      L.stride = 1;
      const int g = e * N + i;
      orca::TileSource src;
      src.env_pos = spos.data();
      src.env_vel = svel.data();
      src.n = N;
      src.self = i;
      src.full_range_sq = a.nd_sq;
      if (tile_grid) {
        src.cell_start = cell_start.data();
        src.sorted = sorted.data();
        src.cx = cell_of[(size_t)i] % orca::kTileGrid;
        src.cy = cell_of[(size_t)i] / orca::kTileGrid;
      }
      switch (policy) {
        case 0: orca::agent_step_body<K, KFULL, 0>(a, e, i, g, spos[(size_t)i], svel[(size_t)i], estep, src, L, 0xffffffffu); break;
        case 1: orca::agent_step_body<K, KFULL, 1>(a, e, i, g, spos[(size_t)i], svel[(size_t)i], estep, src, L, 0xffffffffu); break;
        case 2: orca::agent_step_body<K, KFULL, 2>(a, e, i, g, spos[(size_t)i], svel[(size_t)i], estep, src, L, 0xffffffffu); break;
        default: orca::agent_step_body<K, KFULL, 3>(a, e, i, g, spos[(size_t)i], svel[(size_t)i], estep, src, L, 0xffffffffu); break;
      }
    }
  }
}
// Host twin of the uniform-grid pipeline (G1-G6 of orca_grid.cuh): bounds, cell keys, counting
// sort (serial, so slots inside a cell follow agent order -- any order is legal), then the same
// agent_step_body with a GridSource.
template <int K, bool KFULL>
void run_grid_k(const orca::StepArgs& a0, int policy) {
  orca::StepArgs a = a0;
  a.grid_path = 1;
  const int E = a.E, N = a.N, T = E * N;
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  for (int i = 0; i < T; ++i) {
    mnx = std::fmin(mnx, a.pos[i].x); mny = std::fmin(mny, a.pos[i].y);
    mxx = std::fmax(mxx, a.pos[i].x); mxy = std::fmax(mxy, a.pos[i].y);
  }
  orca::GridParams gp;
  const float cell = orca::grid_cell_size(std::sqrt(a.nd_sq));
  gp.origin_x = mnx; gp.origin_y = mny; gp.inv_cell = 1.0f / cell;
  gp.W = (int)(std::floor((mxx - mnx) / cell) + 1.f);
  gp.H = (int)(std::floor((mxy - mny) / cell) + 1.f);
  gp.ncells = E * gp.W * gp.H;
  std::vector<int> key((size_t)T), start((size_t)gp.ncells + 1, 0), sidx((size_t)T);
  for (int i = 0; i < T; ++i) {
    const int cx = orca::GridSource::cell_coord(a.pos[i].x, gp.origin_x, gp.inv_cell, gp.W);
    const int cy = orca::GridSource::cell_coord(a.pos[i].y, gp.origin_y, gp.inv_cell, gp.H);
    key[(size_t)i] = (i / N) * gp.W * gp.H + cy * gp.W + cx;
    start[(size_t)key[(size_t)i] + 1]++;
  }
  for (int c = 0; c < gp.ncells; ++c) start[(size_t)c + 1] += start[(size_t)c];
  std::vector<int> fill(start.begin(), start.end() - 1);
  // reversed agent order inside each cell on purpose: the result must not depend on it
  for (int i = T - 1; i >= 0; --i) sidx[(size_t)fill[(size_t)key[(size_t)i]]++] = i;
  std::vector<float2> spos((size_t)T), svel((size_t)T);
  std::vector<float4> spv((size_t)T);
  for (int j = 0; j < T; ++j) {
    spos[(size_t)j] = a.pos[sidx[(size_t)j]];
    svel[(size_t)j] = a.vel[sidx[(size_t)j]];
    spv[(size_t)j].x = spos[(size_t)j].x; spv[(size_t)j].y = spos[(size_t)j].y;
    spv[(size_t)j].z = svel[(size_t)j].x; spv[(size_t)j].w = svel[(size_t)j].y;
  }
  std::vector<int> estep0((size_t)E, 0);
  if (a.env_step) for (int e = 0; e < E; ++e) estep0[(size_t)e] = a.env_step[e];
  std::vector<float4> lines((size_t)(K + ORCA_MAX_OBST_LINES));
  for (int j = 0; j < T; ++j) {
    const int g = sidx[(size_t)j], env = g / N, la = g - env * N;
    orca::GridSource src;
    src.spv = spv.data(); src.orig = sidx.data(); src.cell_start = start.data(); src.full_range_sq = a.nd_sq;
    src.gp = gp; src.env = env; src.env_n0 = env * N; src.self = j;
    orca::Lines L; L.base = lines.data(); L.stride = 1;
    const int es = estep0[(size_t)env];
    switch (policy) {
      case 0: orca::agent_step_body<K, KFULL, 0>(a, env, la, g, spos[(size_t)j], svel[(size_t)j], es, src, L, 0xffffffffu); break;
      case 1: orca::agent_step_body<K, KFULL, 1>(a, env, la, g, spos[(size_t)j], svel[(size_t)j], es, src, L, 0xffffffffu); break;
      case 2: orca::agent_step_body<K, KFULL, 2>(a, env, la, g, spos[(size_t)j], svel[(size_t)j], es, src, L, 0xffffffffu); break;
      default: orca::agent_step_body<K, KFULL, 3>(a, env, la, g, spos[(size_t)j], svel[(size_t)j], es, src, L, 0xffffffffu); break;
    }
  }
  if (a.env_step && !a.neighbors_only) for (int e = 0; e < E; ++e) a.env_step[e] += 1;
}
}  // namespace

extern "C" {

// Mirrors orca_set_obstacles + the table packing of orca_api.cu for ONE shared world.
// Returns number of vertices; fills pd[v*4], link[v*4], bsp[v*4] (caller allocates max_v rows).
int emul_build_world(const float* xy, const int* poly_sizes, int num_polys, int max_v, float* pd, int* link, int* bsp,
                     float* seg, int* depth, float obst_range, unsigned* cull_rows, float* cull_geo) {
  orca_host::ObstacleTables T;
  size_t off = 0;
  for (int p = 0; p < num_polys; ++p) {
    if (orca_host::add_polygon(T, xy + 2 * off, poly_sizes[p]) < 0) return -1;
    off += (size_t)poly_sizes[p];
  }
  orca_host::process(T);
  const int nv = T.num_vertices();
  if (nv > max_v) return -2;
  for (int v = 0; v < nv; ++v) {
    pd[4 * v] = T.px[v]; pd[4 * v + 1] = T.py[v]; pd[4 * v + 2] = T.ux[v]; pd[4 * v + 3] = T.uy[v];
    link[4 * v] = T.next[v]; link[4 * v + 1] = T.prev[v]; link[4 * v + 2] = T.convex[v]; link[4 * v + 3] = 0;
    bsp[4 * v] = T.node_vertex[v]; bsp[4 * v + 1] = T.node_left[v]; bsp[4 * v + 2] = T.node_right[v]; bsp[4 * v + 3] = 0;
    const int e1 = T.node_vertex[v], e2 = T.next[e1];
    seg[4 * v] = T.px[e1]; seg[4 * v + 1] = T.py[e1]; seg[4 * v + 2] = T.px[e2]; seg[4 * v + 3] = T.py[e2];
  }
  *depth = T.depth;
  const orca_host::CullMap M = orca_host::build_cull_map(T, obst_range);
  for (int i = 0; i < orca_host::kCullGrid; ++i) cull_rows[i] = M.rows[i];
  cull_geo[0] = M.x0; cull_geo[1] = M.y0; cull_geo[2] = M.inv_cx; cull_geo[3] = M.inv_cy;
  return nv;
}

// args: a fully populated orca::StepArgs with HOST pointers.
int emul_step(const orca::StepArgs* a, int policy) {
  if (a->k == 5) run_k<5, true>(*a, policy);
  else if (a->k == 10) run_k<10, true>(*a, policy);
  else if (a->k <= 16) run_k<16, false>(*a, policy);
  else return -1;
  return 0;
}

// Laser observation for every (agent, ray); tables filled like orca_observe does.
int emul_observe(orca::ObsArgs* a, float neighbor_dist, float radius) {
  const double two_pi = 6.283185307179586476925286766559;
  for (int i = 0; i < a->R; ++i) {
    const double th = i * (two_pi / a->R);
    a->ray_end[i].x = (float)((double)neighbor_dist * std::cos(th));
    a->ray_end[i].y = (float)(-(double)neighbor_dist * std::sin(th));
  }
  for (int i = 0; i < a->C; ++i) {
    const double th = i * (two_pi / a->C);
    a->poly[i].x = (float)((double)radius * std::cos(th));
    a->poly[i].y = (float)(-(double)radius * std::sin(th));
  }
  for (int g = 0; g < a->E * a->N; ++g)
    for (int r = 0; r < a->R; ++r) a->obs[(size_t)g * a->R + r] = orca::observe_ray(*a, g, r);
  return 0;
}
int emul_obsargs_size() { return (int)sizeof(orca::ObsArgs); }

int emul_step_grid(const orca::StepArgs* a, int policy) {
  if (a->k == 5) run_grid_k<5, true>(*a, policy);
  else if (a->k == 10) run_grid_k<10, true>(*a, policy);
  else if (a->k <= 16) run_grid_k<16, false>(*a, policy);
  else return -1;
  return 0;
}

int emul_stepargs_size() { return (int)sizeof(orca::StepArgs); }

float emul_philox_uniform(unsigned long long seed, unsigned c0, unsigned c1) { return orca::philox_uniform(seed, c0, c1); }

}  // extern "C"
