// orca_obs.cuh -- laser-scan observation kernel (K-obs).
//
// Replaces Collision_Avoidance_Env._get_obs + utils.comp_laser + utils.line_intersection
// (collision_avoidence_env.py:231-318,321-350 ; utils.py:5-113): every agent casts `R` rays of
// length neighborDist in the frame whose x axis points at its goal; each ray reports the nearest
// hit among the octagon approximations of its agent neighbors and its obstacle-neighbor edges,
// plus the velocity of what it hit -> 4 floats per ray (hit.x, hit.y, vel.x, vel.y).
// Neighbor lists are the ones the last step produced from PRE-update positions, combined with
// POST-update positions/velocities; the caller passes exactly those buffers.
//
// The reference rotates every segment into the agent's frame and tests every ray against every
// segment.  Here the RAY is rotated into the world frame instead (one rotation per ray; the hit
// parameter t is frame-independent and the hit point is t * ray_end in the agent's frame), and the
// work is split so that few lanes idle while another tests polygons:
//   phase 0  the block stages what the rays of an agent share, once per agent instead of once per
//            ray: frame, neighbor offsets (neighbor - agent), obstacle-edge endpoints.  Every
//            global load of the kernel is issued here, one or two per thread, all in flight at
//            once -- the ray loops below only touch shared memory;
//   phase A  one thread per (agent, ray): exact test against the agent's obstacle-neighbor edges
//            (a handful), and a cull of its agent neighbors -- which polygons can this ray touch
//            at all (bounding circle)? -> a 16-bit mask, typically 0-2 bits;
//   phase B  the block turns the set bits of its 256 rays into a dense queue of (ray, neighbor)
//            pairs and its warps take 32 pairs at a time off it (a ray that owns its polygons
//            makes its whole warp wait for the ray with the most of them): find the edges whose
//            endpoints straddle the ray line (bit mask over the C edges), exact
//            line_intersection only for those.  Hits go back to the owning ray through a 64-bit
//            atomicMin on (distance bits, item, edge) in shared memory;
//   phase C  the ray's thread recomputes the winning edge's hit and writes its float4 of the row
//            (coalesced).
// Ties keep the first segment in the reference's order (neighbors in list order, edges in ring
// order, then obstacle edges): the order index is the low word of the key.
//
// Parity: float32 against the float64 shell within 2e-4 (tests/test_gpu_parity_shell.py).
#pragma once

#include "orca_core.cuh"

#define ORCA_MAX_LASER 32
#define ORCA_MAX_CIRCLE_APPROX 16

namespace orca {

struct ObsArgs {
  int E, N, k;
  int R;  // rays per agent
  int C;  // segments per neighbor polygon
  const float2* pos;
  const float2* vel;
  const float2* goal;
  const int* nbr_idx;   // [E*N][k]
  const int* nbr_cnt;   // [E*N]
  const int* onbr_idx;  // [E*N][ORCA_MAX_OBST_NEIGHBORS]
  const int* onbr_cnt;  // [E*N]
  const float4* vert_pd;
  const int4* vert_link;
  int vert_stride;
  float4* obs;                          // [E*N][R] (hit.x, hit.y, vel.x, vel.y)
  float2 ray_end[ORCA_MAX_LASER];       // (nd cos t, -nd sin t)          env:321-332
  float2 poly[ORCA_MAX_CIRCLE_APPROX];  // (r cos t, -r sin t) ring points  env:335-350
};

constexpr unsigned long long kObsNoHit = ~0ull;
constexpr float kObsSideEps = 1e-3f;  // slack of the "endpoints on one side of the ray" pre-tests (exact test follows)

// utils.line_intersection for a ray from the origin to `e` against segment (p2, p3): same tests
// in the same order as the reference; on a hit returns true and the ray parameter t in [0, 1].
ORCA_HD bool ray_hit_t(float2 e, float2 p2, float2 p3, float* t_out) {
  const float bx = p3.x - p2.x, by = p3.y - p2.y;
  const float denom = e.x * by - bx * e.y;
  if (denom == 0.f) return false;
  const bool pos = denom > 0.f;
  const float cx = -p2.x, cy = -p2.y;
  const float s_num = e.x * cy - e.y * cx;
  if ((s_num < 0.f) == pos) return false;
  const float t_num = bx * cy - by * cx;
  if ((t_num < 0.f) == pos) return false;
  if (((s_num > denom) == pos) || ((t_num > denom) == pos)) return false;
  *t_out = t_num / denom;
  return true;
}

// What the rays of one agent share (staged in shared memory by the kernel, in local arrays by the
// host twin).  (c, s): rotation into the agent's frame, (x, y) -> (c x - s y, s x + c y) with
// c = pref.x, s = -pref.y (rotation by -atan2(pref), utils.py:48-51).
struct AgentScan {
  float c, s;
  int cnt, ocnt;      // agent neighbors, obstacle neighbors
  const float2* rel;  // [cnt]  neighbor position - agent position
  const int* nbr;     // [cnt]  global index of the neighbor (its velocity is read for the winner only)
  const float4* edge; // [ocnt] obstacle edge endpoints - agent position: (a.x, a.y, b.x, b.y)
};

// stage slot q of agent g: neighbor offset / obstacle edge
ORCA_HD void stage_neighbor(const ObsArgs& a, int g, int q, float2* rel, int* nbr) {
  const int j = (g / a.N) * a.N + a.nbr_idx[(size_t)g * a.k + q];
  *rel = sub(a.pos[j], a.pos[g]);
  *nbr = j;
}
ORCA_HD float4 stage_edge(const ObsArgs& a, int g, int q) {
  const size_t voff = (size_t)(g / a.N) * a.vert_stride;
  const int v1 = a.onbr_idx[(size_t)g * ORCA_MAX_OBST_NEIGHBORS + q];
  const float4 A = ORCA_LDG(&a.vert_pd[voff + v1]);
  const int v2i = ORCA_LDG(&a.vert_link[voff + v1]).x;
  const float4 B = ORCA_LDG(&a.vert_pd[voff + v2i]);
  const float2 p = a.pos[g];
  float4 o;
  o.x = A.x - p.x;
  o.y = A.y - p.y;
  o.z = B.x - p.x;
  o.w = B.y - p.y;
  return o;
}
ORCA_HD float2 agent_frame(const ObsArgs& a, int g) {
  const float2 pref = goal_direction(a.pos[g], a.goal[g]);
  return v2(pref.x, -pref.y);
}
// ray end in the world frame (relative to the agent): inverse rotation (x, y) -> (c x + s y, -s x + c y)
ORCA_HD float2 ray_world(float c, float s, float2 e) { return v2(c * e.x + s * e.y, c * e.y - s * e.x); }

ORCA_HD unsigned long long hit_key(float2 e, float t, int item, int edge) {
  const float hx = t * e.x, hy = t * e.y;
  const float d = sqrtf(hx * hx + hy * hy);
  return ((unsigned long long)(unsigned)float_to_bits(d) << 32) | (unsigned)((item << 5) | edge);
}

// Phase A, agent neighbors: bit q = neighbor q's polygon may be hit by the ray (e: agent frame,
// ew: world frame).
ORCA_HD unsigned ray_cull(const ObsArgs& a, const AgentScan& A, float2 e, float2 ew) {
  unsigned mask = 0u;
  const float ray_len = sqrtf(e.x * e.x + e.y * e.y);
  const float inv_len = 1.0f / ray_len;
  const float2 u = v2(ew.x * inv_len, ew.y * inv_len);
  // a neighbor polygon is inscribed in the circle of radius |poly[0]| around the neighbor; the bound
  // is padded far beyond float32 rounding so no hit is ever dropped
  const float reach = sqrtf(a.poly[0].x * a.poly[0].x + a.poly[0].y * a.poly[0].y) * 1.001f + 1e-4f;
  for (int q = 0; q < A.cnt; ++q) {
    const float2 rel = A.rel[q];
    const float along = rel.x * u.x + rel.y * u.y;
    const float perp_sq = (rel.x * rel.x + rel.y * rel.y) - along * along;
    if (along >= -reach && along <= ray_len + reach && perp_sq <= reach * reach) mask |= 1u << q;
  }
  return mask;
}

// Phase A, obstacle neighbors: nearest hit of the ray on the agent's obstacle-neighbor edges, as a
// key with item = 16 + q (so that at equal distance any agent neighbor wins, as in the reference's
// segment order), or kObsNoHit.
ORCA_HD unsigned long long obstacle_hits(const AgentScan& A, float2 e, float2 ew) {
  unsigned long long best = kObsNoHit;
  for (int q = 0; q < A.ocnt; ++q) {
    const float4 ed = A.edge[q];
    float t;
    if (ray_hit_t(ew, v2(ed.x, ed.y), v2(ed.z, ed.w), &t)) {
      const unsigned long long key = hit_key(e, t, 16 + q, 0);
      best = key < best ? key : best;
    }
  }
  return best;
}

// Phase B: nearest hit of the ray on the polygon around `rel` (agent neighbor `item` < 16) as a key
// (distance bits << 32 | item << 5 | edge), or kObsNoHit.  `enabled` = false: no work, the lane
// only keeps its warp company.
ORCA_HD unsigned long long pair_test(const ObsArgs& a, float2 rel, float2 e, float2 ew, int item, bool enabled) {
  unsigned long long best = kObsNoHit;
  unsigned edges = 0u;  // bit m - 1: the endpoints of edge (m - 1, m mod C) are not strictly on one side of the ray line
  if (enabled) {
    const float f_first = det(ew, add(a.poly[0], rel));
    float f_prev = f_first;
    for (int m = 1; m <= a.C; ++m) {
      const float f_cur = (m < a.C) ? det(ew, add(a.poly[m], rel)) : f_first;
      const bool one_side = (f_prev > kObsSideEps && f_cur > kObsSideEps) || (f_prev < -kObsSideEps && f_cur < -kObsSideEps);
      edges |= one_side ? 0u : (1u << (m - 1));
      f_prev = f_cur;
    }
  }
  // lanes walk their own candidate edges (2 of C, typically) in lock step
  while (edges != 0u) {
#if defined(__CUDA_ARCH__)
    const int m = __ffs(edges);
#else
    int m = 1;
    while (!((edges >> (m - 1)) & 1u)) ++m;
#endif
    edges &= edges - 1u;
    float t;
    if (ray_hit_t(ew, add(a.poly[m - 1], rel), add(a.poly[m < a.C ? m : 0], rel), &t)) {
      const unsigned long long key = hit_key(e, t, item, m);
      best = key < best ? key : best;
    }
  }
  return best;
}

// Phase C: the observation row entry of the ray given the winning key.
ORCA_HD float4 ray_result(const ObsArgs& a, const AgentScan& A, float2 e, float2 ew, unsigned long long key) {
  float4 out;
  out.x = out.y = out.z = out.w = 0.f;
  if (key == kObsNoHit) return out;
  const int item = (int)((key >> 5) & 31ull), m = (int)(key & 31ull);
  float t = 0.f;
  float2 vel_r = v2(0.f, 0.f);
  if (item < 16) {
    const float2 rel = A.rel[item];
    ray_hit_t(ew, add(a.poly[m - 1], rel), add(a.poly[m < a.C ? m : 0], rel), &t);
    const float2 nv = a.vel[A.nbr[item]];
    vel_r = v2(A.c * nv.x - A.s * nv.y, A.s * nv.x + A.c * nv.y);
  } else {
    const float4 ed = A.edge[item - 16];
    ray_hit_t(ew, v2(ed.x, ed.y), v2(ed.z, ed.w), &t);
  }
  out.x = t * e.x;
  out.y = t * e.y;
  out.z = vel_r.x;
  out.w = vel_r.y;
  return out;
}

// All phases for one ray, serially (host twin of the kernel; same functions, same order).
ORCA_HD float4 observe_ray(const ObsArgs& a, int g, int ray) {
  float2 rel[16];
  int nbr[16];
  float4 edge[ORCA_MAX_OBST_NEIGHBORS];
  AgentScan A;
  const float2 cs = agent_frame(a, g);
  A.c = cs.x;
  A.s = cs.y;
  A.cnt = a.nbr_cnt[g];
  A.ocnt = a.onbr_cnt[g];
  for (int q = 0; q < A.cnt; ++q) stage_neighbor(a, g, q, &rel[q], &nbr[q]);
  for (int q = 0; q < A.ocnt; ++q) edge[q] = stage_edge(a, g, q);
  A.rel = rel;
  A.nbr = nbr;
  A.edge = edge;
  const float2 e = a.ray_end[ray];
  const float2 ew = ray_world(A.c, A.s, e);
  const unsigned mask = ray_cull(a, A, e, ew);
  unsigned long long best = obstacle_hits(A, e, ew);
  for (int item = 0; item < 16; ++item) {
    if ((mask >> item) & 1u) {
      const unsigned long long key = pair_test(a, rel[item], e, ew, item, true);
      best = key < best ? key : best;
    }
  }
  return ray_result(a, A, e, ew, best);
}

#if defined(__CUDACC__)

#ifndef ORCA_OBS_THREADS
#define ORCA_OBS_THREADS 256
#endif
#ifndef ORCA_OBS_MIN_BLOCKS
#define ORCA_OBS_MIN_BLOCKS 8  // 32 registers: the kernel waits on barriers and shared memory, occupancy is what hides it (551 -> 444 us)
#endif
constexpr int kObsThreads = ORCA_OBS_THREADS;
constexpr int kObsWarps = kObsThreads / 32;
constexpr int kObsAgents = 32;  // agents per block at most (8 rays each; 16 agents for the usual 16 rays)

// agents per block for R rays per agent
inline int obs_agents_per_block(int R) { return (kObsThreads / R) < kObsAgents ? (kObsThreads / R) : kObsAgents; }

__global__ void __launch_bounds__(kObsThreads, ORCA_OBS_MIN_BLOCKS) observe_kernel(const ObsArgs a, const int agents_per_block) {
  __shared__ unsigned long long s_best[kObsThreads];
  __shared__ float2 s_ew[kObsThreads];                        // world-frame ray ends of the block's rays
  __shared__ unsigned short s_queue[kObsThreads * 16];        // (ray of the block << 4) | neighbor item
  __shared__ float2 s_rel[kObsAgents][16];
  __shared__ int s_nbr[kObsAgents][16];
  __shared__ float4 s_edge[kObsAgents][ORCA_MAX_OBST_NEIGHBORS];
  __shared__ float2 s_frame[kObsAgents];
  __shared__ int s_cnt[kObsAgents];                           // cnt | ocnt << 8
  __shared__ int s_warp_total[kObsWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total_agents = a.E * a.N;
  const int g0 = blockIdx.x * agents_per_block;
  const int n_agents = (total_agents - g0) < agents_per_block ? (total_agents - g0) : agents_per_block;

  // ---- phase 0: stage what the rays of an agent share ----
  if (tid < n_agents) {
    const int g = g0 + tid;
    s_frame[tid] = agent_frame(a, g);
    s_cnt[tid] = a.nbr_cnt[g] | (a.onbr_cnt[g] << 8);
  }
  for (int i = tid; i < n_agents * 16; i += kObsThreads) {
    const int al = i >> 4, q = i & 15, g = g0 + al;
    if (q < a.nbr_cnt[g]) stage_neighbor(a, g, q, &s_rel[al][q], &s_nbr[al][q]);
    if (q < a.onbr_cnt[g]) s_edge[al][q] = stage_edge(a, g, q);
  }
  __syncthreads();

  // ---- phase A: obstacle edges (exact), neighbor cull ----
  const int al = tid / a.R;
  const int ray = tid - al * a.R;
  const bool valid = al < n_agents;
  AgentScan A;
  A.c = A.s = 0.f;
  A.cnt = A.ocnt = 0;
  A.rel = s_rel[0];
  A.nbr = s_nbr[0];
  A.edge = s_edge[0];
  float2 e = v2(0.f, 0.f), ew = v2(0.f, 0.f);
  unsigned mask = 0u;
  unsigned long long best = kObsNoHit;
  if (valid) {
    const float2 cs = s_frame[al];
    const int cc = s_cnt[al];
    A.c = cs.x;
    A.s = cs.y;
    A.cnt = cc & 255;
    A.ocnt = cc >> 8;
    A.rel = s_rel[al];
    A.nbr = s_nbr[al];
    A.edge = s_edge[al];
    e = a.ray_end[ray];
    ew = ray_world(A.c, A.s, e);
    best = obstacle_hits(A, e, ew);
    mask = ray_cull(a, A, e, ew);
  }
  s_best[tid] = best;
  s_ew[tid] = ew;

  // ---- the block's (ray, neighbor) pairs as a dense queue ----
  const int mine = __popc(mask);
  int incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  if (lane == 31) s_warp_total[warp] = incl;
  __syncthreads();
  int base = 0, pairs = 0;
#pragma unroll
  for (int w = 0; w < kObsWarps; ++w) {
    const int c = s_warp_total[w];
    base += (w < warp) ? c : 0;
    pairs += c;
  }
  {
    int slot = base + incl - mine;
    unsigned m = mask;
    while (m != 0u) {
      const int item = __ffs(m) - 1;
      m &= m - 1u;
      s_queue[slot++] = (unsigned short)((tid << 4) | item);
    }
  }
  __syncthreads();

  // ---- phase B: pair tests, 32 pairs per warp until the queue is empty ----
  for (int first = warp << 5; first < pairs; first += kObsThreads) {  // warp-uniform
    const int i = first + lane;
    const bool on = i < pairs;
    const int entry = on ? (int)s_queue[i] : 0;
    const int src = entry >> 4, item = entry & 15;
    const int sal = src / a.R;
    const unsigned long long key = pair_test(a, s_rel[sal][item], a.ray_end[src - sal * a.R], s_ew[src], item, on);
    if (key != kObsNoHit) atomicMin(&s_best[src], key);
  }
  __syncthreads();

  // ---- phase C: winners ----
  if (valid) a.obs[(size_t)(g0 + al) * a.R + ray] = ray_result(a, A, e, ew, s_best[tid]);
}
#endif

}  // namespace orca
