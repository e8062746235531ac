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
//            atomicMin on (ray-parameter bits, item, edge) in shared memory;
//   phase C  the ray's thread turns the winning key (it holds t) into its float4 of the row
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

// Key of a hit: (bits of the ray parameter t, item, edge).  All hits of a key's ray lie on that ray,
// so their distance t * |e| orders like t (t in [0, 1]: non-negative floats order like their bit
// patterns); keeping t itself, not the distance, lets the row be written from the key alone.
ORCA_HD unsigned long long hit_key(float t, int item, int edge) {
  return ((unsigned long long)(unsigned)float_to_bits(fabsf(t)) << 32) | (unsigned)((item << 5) | edge);
}
ORCA_HD float hit_key_t(unsigned long long key) { return bits_to_float((int)(unsigned)(key >> 32)); }

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
      const unsigned long long key = hit_key(t, 16 + q, 0);
      best = key < best ? key : best;
    }
  }
  return best;
}

// Phase B: nearest hit of the ray on the polygon around `rel` (agent neighbor `item` < 16) as a key
// (ray-parameter bits << 32 | item << 5 | edge), or kObsNoHit.  `enabled` = false: no work, the lane
// only keeps its warp company.
//
// `drop_exit` (kernel only; the host twin tests every candidate edge): the caller guarantees that
// the ray's origin lies outside the polygon's circumcircle.  A line crosses a convex polygon through
// two edges; from outside, the ray ENTERS through the one whose denominator in ray_hit_t has the
// sign opposite to the ring's orientation (`ccw`) and leaves through the other, later.  An edge
// whose two endpoints are both further than `clear_eps` from the ray line (in units of det(ew, .))
// and whose sign says "exit" is skipped: its hit, if any, lies at least 1.15 * clear_eps / |ew|^2
// (>= 1e-4 of the ray) behind the entry hit, which is tested, so the minimum key cannot change.
// Edges with an endpoint near the line are always tested, like in the twin.
struct PairFilter {
  bool drop_exit, ccw;
  float clear_eps;
};
ORCA_HD PairFilter no_pair_filter() {
  PairFilter f;
  f.drop_exit = false;
  f.ccw = false;
  f.clear_eps = 0.f;
  return f;
}
ORCA_HD unsigned long long pair_test(const float2* poly, int C, float2 rel, float2 e, float2 ew, int item, bool enabled,
                                     PairFilter flt) {
  unsigned long long best = kObsNoHit;
  unsigned edges = 0u;  // bit m - 1: the endpoints of edge (m - 1, m mod C) are not strictly on one side of the ray line
  if (enabled) {
    const float f_first = det(ew, add(poly[0], rel));
    float f_prev = f_first;
    for (int m = 1; m <= C; ++m) {
      const float f_cur = (m < C) ? det(ew, add(poly[m], rel)) : f_first;
      const bool one_side = fminf(f_prev, f_cur) > kObsSideEps || fmaxf(f_prev, f_cur) < -kObsSideEps;
      const bool clear_exit = flt.drop_exit && fminf(fabsf(f_prev), fabsf(f_cur)) > flt.clear_eps && ((f_cur > 0.f) == flt.ccw);
      edges |= (one_side || clear_exit) ? 0u : (1u << (m - 1));
      f_prev = f_cur;
    }
  }
  // lanes walk their own candidate edges (1 or 2 of C, typically) in lock step
  while (edges != 0u) {
#if defined(__CUDA_ARCH__)
    const int m = __ffs(edges);
#else
    int m = 1;
    while (!((edges >> (m - 1)) & 1u)) ++m;
#endif
    edges &= edges - 1u;
    float t;
    if (ray_hit_t(ew, add(poly[m - 1], rel), add(poly[m < C ? m : 0], rel), &t)) {
      const unsigned long long key = hit_key(t, item, m);
      best = key < best ? key : best;
    }
  }
  return best;
}

// Phase C: the observation row entry of a ray from its winning key: hit point = t * ray_end in the
// agent's frame; a polygon hit also reports the neighbor's velocity nv rotated into that frame.
ORCA_HD float4 ray_result_polygon(float t, float2 nv, float c, float s, float2 e) {
  float4 out;
  out.x = t * e.x;
  out.y = t * e.y;
  out.z = c * nv.x - s * nv.y;
  out.w = s * nv.x + c * nv.y;
  return out;
}
ORCA_HD float4 ray_result_edge(float t, float2 e) {
  float4 out;
  out.x = t * e.x;
  out.y = t * e.y;
  out.z = 0.f;
  out.w = 0.f;
  return out;
}
ORCA_HD float4 ray_result(const ObsArgs& a, const AgentScan& A, float2 e, unsigned long long key) {
  float4 out;
  out.x = out.y = out.z = out.w = 0.f;
  if (key == kObsNoHit) return out;
  const int item = (int)((key >> 5) & 31ull);
  if (item < 16) return ray_result_polygon(hit_key_t(key), a.vel[A.nbr[item]], A.c, A.s, e);
  return ray_result_edge(hit_key_t(key), e);
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
      const unsigned long long key = pair_test(a.poly, a.C, rel[item], e, ew, item, true, no_pair_filter());
      best = key < best ? key : best;
    }
  }
  return ray_result(a, A, e, best);
}

#if defined(__CUDACC__)

// ---- the kernel: warps own whole agents, nothing crosses a block barrier -------------------------
//
// A warp takes a chunk of `A` consecutive agents.
//   staging  lane = agent: counts, position, frame towards the goal (one square root and one
//            division per AGENT); then lane = (agent, neighbor slot): one gather of the neighbor's
//            position each, all of the chunk's gathers in flight together; then lane = (agent,
//            obstacle-edge slot).  Everything the ray loops need is in shared memory afterwards.
//   passes   of `AP` agents, lane = (agent, ray slot), `L` lanes per agent.  When the ray table is
//            antipodal (even ray count: ray i + R/2 = -ray i, checked by the host) a ray slot owns
//            BOTH rays of a line through the agent: one cross / dot product per neighbor culls for
//            the two of them, R = 16 rays need 8 lanes and a pass covers 4 agents.  The cull marks
//            which neighbor polygons a ray can touch at all (bounding circle; a superset of the
//            hits, so the exact tests decide and the result is independent of it); obstacle-
//            neighbor edges are tested exactly right away.
//   queue    the marked (agent, ray, neighbor) pairs -- 5 to 10 per agent -- go to a warp-private
//            queue; whenever 32 are pending the warp runs the exact polygon test on them with every
//            lane live (pair_test, the function the host twin calls; exit edges skipped, see there),
//            and a hit lowers the ray's 64-bit key (ray-parameter bits, item, edge) with a shared-memory
//            atomicMin, which also encodes the reference's first-segment-wins order.
//   rows     after the last pass the rest of the queue is drained and the rows are written (float4
//            per ray, 128 contiguous bytes per agent and store).
// Only __syncwarp() orders the phases.
#ifndef ORCA_OBS_BLOCKS_PER_SM
#define ORCA_OBS_BLOCKS_PER_SM 5  // measured per 1 M agents: 4 blocks 291 us, 5 blocks 274 us, 6 blocks (40 registers, spills) 285 us
#endif
constexpr int kObsWarps = 8;
constexpr int kObsThreads = kObsWarps * 32;
constexpr int kObsWarpBytes = ((227 * 1024) / ORCA_OBS_BLOCKS_PER_SM - 1024 - 640) / kObsWarps;  // shared memory a warp may use
constexpr int kObsQueue = 64;      // < 32 pending before a push round, at most 32 pushed per round
constexpr int kObsEdgeSlots = 4;   // obstacle edges staged per agent; further ones are read from global memory

struct ObsPlan {
  int HR;       // ray slots per agent (R / 2 when paired, else R)
  int L, logL;  // lanes per agent in the ray role (power of two >= HR)
  int AP;       // agents per pass = 32 / L
  int KP, logKP;  // neighbor slots per agent (power of two >= k)
  int A;        // agents per warp chunk (multiple of AP, at most 32)
  int warp_bytes;
  unsigned long long n_magic;  // ceil(2^64 / N): g / N = umul64hi(g, n_magic) for N > 1
  long long last_chunk;        // index and size of the last chunk (the only one that may be partial)
  int last_n;
  int ccw;                     // orientation of the polygon ring
  float clear_eps;             // see pair_test
  float far_sq;                // (longest ray, padded)^2: an obstacle edge further than this from the agent meets no ray
};

inline int obs_log2_ceil(int x) {
  int l = 0;
  while ((1 << l) < x) ++l;
  return l;
}
inline bool obs_table_is_antipodal(const ObsArgs& a) {
  if (a.R < 2 || (a.R & 1)) return false;
  for (int i = 0; i < a.R / 2; ++i) {
    const float2 p = a.ray_end[i], q = a.ray_end[i + a.R / 2];
    const float tol = 1e-6f * (fabsf(p.x) + fabsf(p.y));
    if (fabsf(p.x + q.x) > tol || fabsf(p.y + q.y) > tol) return false;
  }
  return true;
}
// bytes per agent of a warp's shared memory: keys, neighbor offsets + indices, edges, frame, position, counts
inline int obs_bytes_per_agent(int R, int KP) { return R * 8 + KP * 12 + kObsEdgeSlots * 16 + 8 + 8 + 4; }
inline ObsPlan obs_plan(const ObsArgs& a, bool paired) {
  ObsPlan p;
  p.HR = paired ? a.R / 2 : a.R;
  p.logL = obs_log2_ceil(p.HR);
  if (p.logL < 2) p.logL = 2;  // at most 8 agents per pass
  p.L = 1 << p.logL;
  p.AP = 32 / p.L;
  p.logKP = obs_log2_ceil(a.k);
  p.KP = 1 << p.logKP;
  // chunk size: what fits a warp's share of shared memory (five blocks of 8 warps per SM: 5.5 KB), at most 32 agents (lane = agent
  // while staging; 5 bits of a queue entry), in whole passes, even (16-byte alignment of the edge array)
  p.A = (kObsWarpBytes - kObsQueue * 2) / obs_bytes_per_agent(a.R, p.KP);
  if (p.A > 32) p.A = 32;
  p.A -= p.A % p.AP;
  if (p.A < p.AP) p.A = p.AP;
  if (p.A & 1) p.A += 1;
  p.warp_bytes = (kObsQueue * 2 + p.A * obs_bytes_per_agent(a.R, p.KP) + 15) & ~15;
  p.n_magic = a.N > 1 ? (~0ull / (unsigned long long)a.N + 1ull) : 0ull;
  const long long total = (long long)a.E * a.N;
  p.last_chunk = (total - 1) / p.A;
  p.last_n = (int)(total - p.last_chunk * p.A);
  // ring orientation and the "clearly off the line" threshold of the exit-edge filter
  const float2 b0 = sub(a.poly[1], a.poly[0]), b1 = sub(a.poly[2], a.poly[1]);
  p.ccw = det(b0, b1) > 0.f ? 1 : 0;
  float len_sq = 0.f;
  for (int i = 0; i < a.R; ++i) len_sq = fmaxf(len_sq, a.ray_end[i].x * a.ray_end[i].x + a.ray_end[i].y * a.ray_end[i].y);
  p.clear_eps = fmaxf(kObsSideEps, 1e-4f * len_sq);
  const float far = sqrtf(len_sq) * 1.001f + 1e-3f;
  p.far_sq = far * far;
  return p;
}

__device__ __forceinline__ int obs_env_of(int g, int N, unsigned long long magic) {
  return N > 1 ? (int)__umul64hi((unsigned long long)(unsigned)g, magic) : g;
}

// obstacle edge q of agent g relative to the agent (same arithmetic as stage_edge)
__device__ __forceinline__ float4 obs_edge(const ObsArgs& a, size_t voff, int g, int q, float2 p) {
  const int v1 = a.onbr_idx[(size_t)g * ORCA_MAX_OBST_NEIGHBORS + q];
  const float4 A = ORCA_LDG(&a.vert_pd[voff + v1]);
  const int v2i = ORCA_LDG(&a.vert_link[voff + v1]).x;
  const float4 B = ORCA_LDG(&a.vert_pd[voff + v2i]);
  float4 o;
  o.x = A.x - p.x;
  o.y = A.y - p.y;
  o.z = B.x - p.x;
  o.w = B.y - p.y;
  return o;
}

struct ObsWarpMem {
  unsigned long long* best;  // [A][R]
  float4* edge;              // [A][kObsEdgeSlots]
  float2* rel;               // [A][KP]   neighbor position - agent position
  float2* frame;             // [A]       (c, s)
  float2* pos;               // [A]
  int* nbr;                  // [A][KP]   global index of the neighbor
  int* cnt;                  // [A]       agent neighbors | obstacle neighbors << 8 | out-of-reach bits of the staged edges << 16
  unsigned short* queue;     // [kObsQueue]  (agent of the chunk << 9) | (ray << 4) | item
};

// exact tests of up to 32 queued pairs, one per lane
__device__ __forceinline__ void obs_drain(const ObsArgs& a, const ObsPlan& p, const ObsWarpMem& M, const float4* s_ray,
                                          const float2* s_poly, float reach_sq, int first, int count, int lane) {
  const bool on = lane < count;
  const int entry = on ? (int)M.queue[first + lane] : 0;
  const int ac = entry >> 9, ray = (entry >> 4) & 31, item = entry & 15;
  ORCA_DCHECK(first >= 0 && first + count <= kObsQueue && ac < p.A && ray < a.R && item < p.KP);
  const float2 cs = M.frame[ac];
  const float4 rt = s_ray[ray];
  const float2 e = v2(rt.x, rt.y);
  const float2 ew = ray_world(cs.x, cs.y, e);
  const float2 rel = M.rel[(ac << p.logKP) + item];
  PairFilter flt;
  flt.drop_exit = abs_sq(rel) > reach_sq;  // the agent is outside the neighbor polygon's circumcircle
  flt.ccw = p.ccw != 0;
  flt.clear_eps = p.clear_eps;
  const unsigned long long key = pair_test(s_poly, a.C, rel, e, ew, item, on, flt);
  if (key != kObsNoHit) atomicMin(&M.best[ac * a.R + ray], key);
}

template <bool PAIRED>
__global__ void __launch_bounds__(kObsThreads, ORCA_OBS_BLOCKS_PER_SM) observe_kernel(const ObsArgs a, const ObsPlan p) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ float4 s_ray[ORCA_MAX_LASER];  // (e.x, e.y, |e|, 1 / |e|)
  __shared__ float2 s_poly[ORCA_MAX_CIRCLE_APPROX];
  constexpr unsigned kFull = 0xffffffffu;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < a.R) {
    const float2 e = a.ray_end[threadIdx.x];
    const float len = sqrtf(e.x * e.x + e.y * e.y);
    s_ray[threadIdx.x] = make_float4(e.x, e.y, len, 1.0f / len);
  }
  if (threadIdx.x < a.C) s_poly[threadIdx.x] = a.poly[threadIdx.x];
  __syncthreads();

  [[maybe_unused]] const int total = a.E * a.N;  // ORCA_DCHECK only
  const long long chunk = (long long)blockIdx.x * kObsWarps + warp;
  if (chunk > p.last_chunk) return;
  const int g_first = (int)(chunk * p.A);
  // The size of the (only) partial chunk comes from the host: ptxas 12.9.86 turns
  // min(total - g_first, A) -- in either spelling, also through an asm sub -- into ONE VIADDMNMX
  // with the uniform `total` as an operand and no negate bit, i.e. min(total + g_first, A): the last
  // chunk ran past the end of the batch (tools/probes/viaddmnmx_probe.cu reproduces it).
  const int n_chunk = chunk == p.last_chunk ? p.last_n : p.A;

  ObsWarpMem M;
  {
    unsigned char* base = s_dyn + (size_t)warp * p.warp_bytes;
    M.best = reinterpret_cast<unsigned long long*>(base);
    M.edge = reinterpret_cast<float4*>(M.best + p.A * a.R);
    M.rel = reinterpret_cast<float2*>(M.edge + p.A * kObsEdgeSlots);
    M.frame = M.rel + p.A * p.KP;
    M.pos = M.frame + p.A;
    M.nbr = reinterpret_cast<int*>(M.pos + p.A);
    M.cnt = M.nbr + p.A * p.KP;
    M.queue = reinterpret_cast<unsigned short*>(M.cnt + p.A);
  }
  // a neighbor polygon is inscribed in the circle of radius |poly[0]| around the neighbor; the bound
  // is padded far beyond float32 rounding (of the products below and of the antipodal pairing)
  const float reach = sqrtf(s_poly[0].x * s_poly[0].x + s_poly[0].y * s_poly[0].y) * 1.001f + 1e-4f;
  const float reach_sq = reach * reach;

  // ---- staging, lane = agent
  int my_base = 0;  // first agent of this lane's agent's world
  if (lane < n_chunk) {
    const int g = g_first + lane;
    ORCA_DCHECK(g >= 0 && g < total);
    const float2 P = a.pos[g];
    const float2 pref = goal_direction(P, a.goal[g]);
    const int cnt = a.nbr_cnt[g], ocnt = a.onbr_cnt[g];
    ORCA_DCHECK(cnt >= 0 && cnt <= a.k && ocnt >= 0 && ocnt <= ORCA_MAX_OBST_NEIGHBORS);
    M.frame[lane] = v2(pref.x, -pref.y);
    M.pos[lane] = P;
    M.cnt[lane] = cnt | (ocnt << 8);
    my_base = obs_env_of(g, a.N, p.n_magic) * a.N;
  }
  for (int i = lane; i < n_chunk * a.R; i += 32) M.best[i] = kObsNoHit;
  __syncwarp();
  // ---- staging, lane = (agent, neighbor slot): index first, then the gather
  const int n_slots = n_chunk << p.logKP;
#pragma unroll 2
  for (int i0 = 0; i0 < n_slots; i0 += 32) {  // warp-uniform trip count: every lane takes part in the shuffle
    const int i = i0 + lane;
    const int ac = (i < n_slots ? i : 0) >> p.logKP, q = i & (p.KP - 1);
    const int base = __shfl_sync(kFull, my_base, ac);
    if (i < n_slots) {
      int j = -1;
      if (q < (M.cnt[ac] & 255)) j = base + a.nbr_idx[(size_t)(g_first + ac) * a.k + q];
      M.nbr[i] = j;
    }
  }
  __syncwarp();
#pragma unroll 4
  for (int i = lane; i < n_slots; i += 32) {
    const int j = M.nbr[i];
    if (j >= 0) {
      ORCA_DCHECK(j < total);
      M.rel[i] = sub(a.pos[j], M.pos[i >> p.logKP]);
    }
  }
  // ---- staging, lane = (agent, obstacle-edge slot)
  for (int i = lane; i < n_chunk * kObsEdgeSlots; i += 32) {
    const int ac = i / kObsEdgeSlots, q = i % kObsEdgeSlots;
    if (q < ((M.cnt[ac] >> 8) & 255)) {
      const int g = g_first + ac;
      const size_t voff = a.vert_stride ? (size_t)obs_env_of(g, a.N, p.n_magic) * a.vert_stride : 0;
      const float4 ed = obs_edge(a, voff, g, q, M.pos[ac]);
      M.edge[i] = ed;
      // The step kernel lists the edges within timeHorizonObst * maxSpeed + radius of the agent; the rays are
      // usually much shorter (gym world: 4.4 against 1.5).  An edge whose nearest point is further away than the
      // longest ray (padded) cannot meet any ray -- every exact test would reject it -- so the ray loop skips it.
      if (dist_sq_point_segment(v2(ed.x, ed.y), v2(ed.z, ed.w), v2(0.f, 0.f)) > p.far_sq) atomicOr(&M.cnt[ac], 0x10000 << q);
    }
  }
  __syncwarp();

  int pending = 0;  // warp-uniform
  const int al = lane >> p.logL, lr = lane & (p.L - 1);
  const bool ray_lane = lr < p.HR;
  const int r0 = lr, r1 = lr + p.HR;
  float4 rt0 = make_float4(0.f, 0.f, 1.f, 1.f);
  if (ray_lane) rt0 = s_ray[r0];

  for (int p0 = 0; p0 < n_chunk; p0 += p.AP) {
    const int ac = p0 + al;
    const bool valid = ac < n_chunk && ray_lane;
    unsigned mask = 0u;  // bit q: neighbor q may be hit by ray r0; bit 16 + q: by ray r1
    if (valid) {
      const float2 cs = M.frame[ac];
      const int cc = M.cnt[ac];
      const int cnt = cc & 255, ocnt = (cc >> 8) & 255;
      const unsigned far_edges = (unsigned)cc >> 16;  // staged edges (q < kObsEdgeSlots) out of every ray's reach
      const float2 e0 = v2(rt0.x, rt0.y);
      const float2 ew0 = ray_world(cs.x, cs.y, e0);
      const float2 u = v2(ew0.x * rt0.w, ew0.y * rt0.w);
      const float2* rel = M.rel + (ac << p.logKP);
      for (int q = 0; q < cnt; ++q) {
        const float2 d = rel[q];
        const float along = d.x * u.x + d.y * u.y;
        const float across = d.x * u.y - d.y * u.x;
        const bool near_line = fabsf(across) <= reach;
        if (near_line && along >= -reach) mask |= 1u << q;
        if (PAIRED && near_line && along <= reach) mask |= 0x10000u << q;
      }
      if (ocnt > 0) {
        // exact, in list order (key item 16 + q: at equal distance any agent neighbor wins)
        const int g = g_first + ac;
        const float2 e1 = PAIRED ? v2(s_ray[r1].x, s_ray[r1].y) : e0;
        const float2 ew1 = ray_world(cs.x, cs.y, e1);
        unsigned long long b0 = kObsNoHit, b1 = kObsNoHit;
        for (int q = 0; q < ocnt; ++q) {
          if ((far_edges >> q) & 1u) continue;
          float4 ed;
          if (q < kObsEdgeSlots) {
            ed = M.edge[ac * kObsEdgeSlots + q];
          } else {
            const size_t voff = a.vert_stride ? (size_t)obs_env_of(g, a.N, p.n_magic) * a.vert_stride : 0;
            ed = obs_edge(a, voff, g, q, M.pos[ac]);
          }
          float t;
          if (ray_hit_t(ew0, v2(ed.x, ed.y), v2(ed.z, ed.w), &t)) {
            const unsigned long long key = hit_key(t, 16 + q, 0);
            b0 = key < b0 ? key : b0;
          }
          if (PAIRED && ray_hit_t(ew1, v2(ed.x, ed.y), v2(ed.z, ed.w), &t)) {
            const unsigned long long key = hit_key(t, 16 + q, 0);
            b1 = key < b1 ? key : b1;
          }
        }
        if (b0 != kObsNoHit) M.best[ac * a.R + r0] = b0;
        if (PAIRED && b1 != kObsNoHit) M.best[ac * a.R + r1] = b1;
      }
    }
    // ---- marked pairs -> queue, one per lane and round; 32 pending -> exact tests
    while (__any_sync(kFull, mask != 0u)) {
      const bool has = mask != 0u;
      const unsigned votes = __ballot_sync(kFull, has);
      if (has) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1u;
        const int slot = pending + __popc(votes & ((1u << lane) - 1u));
        ORCA_DCHECK(slot >= 0 && slot < kObsQueue && ac < p.A);
        M.queue[slot] = (unsigned short)((ac << 9) | ((bit < 16 ? r0 : r1) << 4) | (bit & 15));
      }
      pending += __popc(votes);
      __syncwarp();
      if (pending >= 32) {
        pending -= 32;
        obs_drain(a, p, M, s_ray, s_poly, reach_sq, pending, 32, lane);
        __syncwarp();
      }
    }
  }
  __syncwarp();
  if (pending > 0) obs_drain(a, p, M, s_ray, s_poly, reach_sq, 0, pending, lane);
  __syncwarp();

  // ---- rows
  for (int p0 = 0; p0 < n_chunk; p0 += p.AP) {
    const int ac = p0 + al;
    if (ac < n_chunk && ray_lane) {
      const int g = g_first + ac;
      ORCA_DCHECK(g >= 0 && g < total && r0 < a.R && (!PAIRED || r1 < a.R));
      const float2 cs = M.frame[ac];
#pragma unroll
      for (int h = 0; h < (PAIRED ? 2 : 1); ++h) {
        const int ray = h == 0 ? r0 : r1;
        const unsigned long long key = M.best[ac * a.R + ray];
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        if (key != kObsNoHit) {
          const float4 rt = s_ray[ray];
          const float2 e = v2(rt.x, rt.y);
          const int item = (int)((key >> 5) & 31ull);
          if (item < 16) {
            out = ray_result_polygon(hit_key_t(key), a.vel[M.nbr[(ac << p.logKP) + item]], cs.x, cs.y, e);
          } else {
            out = ray_result_edge(hit_key_t(key), e);
          }
        }
        a.obs[(size_t)g * a.R + ray] = out;
      }
    }
  }
}
#endif

}  // namespace orca
