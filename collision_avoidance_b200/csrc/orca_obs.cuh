// orca_obs.cuh -- laser-scan observation kernel (K-obs).
//
// Replaces Collision_Avoidance_Env._get_obs + utils.comp_laser + utils.line_intersection
// (collision_avoidence_env.py:231-318,321-350 ; utils.py:5-113): every agent casts `R` rays of
// length neighborDist in the frame whose x axis points at its goal; each ray reports the nearest
// hit among the octagon approximations of its agent neighbors and its obstacle-neighbor edges,
// plus the velocity of what it hit -> 4 floats per ray (hit.x, hit.y, vel.x, vel.y).
//
// One thread per (agent, ray): the 16 rays of an agent sit in 16 consecutive lanes, so an
// agent's 256-byte observation row is written as one coalesced float4 store per lane and the
// neighbor loads are broadcast within the half-warp.  Neighbor lists are the ones the last
// step produced from PRE-update positions, combined with POST-update positions/velocities
// (SURVEY Q3); the caller passes exactly those buffers.
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

// utils.line_intersection for a ray from the origin to `e` against segment (p2, p3); returns
// true on a hit and the hit point.  Same tests in the same order as the reference.
ORCA_HD bool ray_hit(float2 e, float2 p2, float2 p3, float2* hit) {
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
  const float t = t_num / denom;
  *hit = v2(t * e.x, t * e.y);
  return true;
}

ORCA_HD float4 observe_ray(const ObsArgs& a, int g, int ray) {
  const int env = g / a.N;
  const float2 p = a.pos[g];
  const float2 pref = goal_direction(p, a.goal[g]);
  // rotation by -atan2(pref): (x, y) -> (c x - s y, s x + c y) with c = pref.x, s = -pref.y
  const float c = pref.x, s = -pref.y;
  const float2 e = a.ray_end[ray];
  float best = INFINITY;
  float2 best_hit = v2(0.f, 0.f), best_vel = v2(0.f, 0.f);

  // A neighbor polygon is inscribed in the circle of radius |poly[0]| around the neighbor, so a
  // ray that stays farther than that from the centre cannot hit any of its segments.  Each lane
  // (= ray) first collects the neighbors it can hit at all (typically 1-2 of 5) in a bit mask and
  // then walks only those: lanes of a warp work on DIFFERENT neighbors in the same iteration, so
  // the culling survives SIMT (a plain `continue` would not: some ray always hits).  The bound is
  // padded far beyond float32 rounding of the vertex positions, so no hit is ever dropped.
  const int cnt = a.nbr_cnt[g];
  unsigned todo = 0u;
  {
    const float ray_len = sqrtf(e.x * e.x + e.y * e.y);
    const float inv_len = 1.0f / ray_len;
    const float2 u = v2(e.x * inv_len, e.y * inv_len);
    const float reach = sqrtf(a.poly[0].x * a.poly[0].x + a.poly[0].y * a.poly[0].y) * 1.001f + 1e-4f;
    for (int q = 0; q < cnt; ++q) {
      const int j = env * a.N + a.nbr_idx[(size_t)g * a.k + q];
      const float2 rel = sub(a.pos[j], p);
      const float2 ctr = v2(c * rel.x - s * rel.y, s * rel.x + c * rel.y);
      const float along = ctr.x * u.x + ctr.y * u.y;
      const float perp_sq = (ctr.x * ctr.x + ctr.y * ctr.y) - along * along;
      if (along >= -reach && along <= ray_len + reach && perp_sq <= reach * reach) todo |= 1u << q;
    }
  }
  // neighbors are visited in list order (lowest bit first) and ties keep the first minimum, so the
  // winner is the same as when every neighbor is tested
  while (todo != 0u) {
#if defined(__CUDA_ARCH__)
    const int q = __ffs(todo) - 1;
#else
    int q = 0;
    while (!((todo >> q) & 1u)) ++q;
#endif
    todo &= todo - 1u;
    const int j = env * a.N + a.nbr_idx[(size_t)g * a.k + q];
    const float2 rel = sub(a.pos[j], p);
    const float2 nv = a.vel[j];
    const float2 nv_r = v2(c * nv.x - s * nv.y, s * nv.x + c * nv.y);
    float2 prev = add(a.poly[0], rel);
    float2 prev_r = v2(c * prev.x - s * prev.y, s * prev.x + c * prev.y);
    const float2 first_r = prev_r;
    for (int m = 1; m <= a.C; ++m) {
      float2 cur_r;
      if (m < a.C) {
        const float2 cur = add(a.poly[m], rel);
        cur_r = v2(c * cur.x - s * cur.y, s * cur.x + c * cur.y);
      } else {
        cur_r = first_r;
      }
      float2 h;
      if (ray_hit(e, prev_r, cur_r, &h)) {
        const float d = sqrtf(h.x * h.x + h.y * h.y);
        if (d < best) {
          best = d;
          best_hit = h;
          best_vel = nv_r;
        }
      }
      prev_r = cur_r;
    }
  }
  const int ocnt = a.onbr_cnt[g];
  const size_t voff = (size_t)env * a.vert_stride;
  for (int q = 0; q < ocnt; ++q) {
    const int v1 = a.onbr_idx[(size_t)g * ORCA_MAX_OBST_NEIGHBORS + q];
    const float4 A = ORCA_LDG(&a.vert_pd[voff + v1]);
    const int v2i = ORCA_LDG(&a.vert_link[voff + v1]).x;
    const float4 B = ORCA_LDG(&a.vert_pd[voff + v2i]);
    const float2 pa = sub(v2(A.x, A.y), p), pb = sub(v2(B.x, B.y), p);
    const float2 pa_r = v2(c * pa.x - s * pa.y, s * pa.x + c * pa.y);
    const float2 pb_r = v2(c * pb.x - s * pb.y, s * pb.x + c * pb.y);
    float2 h;
    if (ray_hit(e, pa_r, pb_r, &h)) {
      const float d = sqrtf(h.x * h.x + h.y * h.y);
      if (d < best) {
        best = d;
        best_hit = h;
        best_vel = v2(0.f, 0.f);
      }
    }
  }
  float4 out;
  out.x = best_hit.x;
  out.y = best_hit.y;
  out.z = best_vel.x;
  out.w = best_vel.y;
  return out;
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(256) observe_kernel(const ObsArgs a) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)a.E * a.N * a.R;
  if (t >= total) return;
  const int g = (int)(t / a.R);
  const int ray = (int)(t - (long long)g * a.R);
  a.obs[t] = observe_ray(a, g, ray);
}
#endif

}  // namespace orca
