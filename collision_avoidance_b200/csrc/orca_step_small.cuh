// orca_step_small.cuh -- fused ORCA environment step for worlds that fit one thread block
// (agents_per_env <= 256): the shape of BASELINE configs 1-4.
//
// One thread per agent, one or more whole envs per block.  The block stages the pre-step
// positions/velocities of its envs in shared memory (one coalesced float2 load each), so the
// brute-force neighbor scan and every neighbor fetch after it are shared-memory reads and HBM
// sees exactly the algorithmic traffic: pos + vel + goal in, pos + vel out (40 B / agent-step,
// SURVEY.md 8d).  Because a whole env lives in one block and all reads of other agents go
// through the staged copy, the update is done in place (RVO2's two-phase doStep).
//
// Replaces, per step and for every env of the batch at once:
//   RVO2 doStep (kd-tree build + computeNeighbors + computeNewVelocity + update)
//       <- collision_avoidence_env.py:385,448 ; ALAN_true.py:601,632
//   comp_pref_vel / update_pref_vel            <- env :151-162 ; ALAN :483-495
//   action rotation, reward                    <- env :371-400 ; ALAN :588-618
//   done test / goal switch                    <- env :352-365 ; ALAN :547-566
//   ALAN softmax bandit select + update        <- ALAN :576-586, 620-628
#pragma once

#include "orca_core.cuh"

namespace orca {

enum : int { POLICY_EXTERNAL = 0, POLICY_GOAL = 1, POLICY_RL = 2, POLICY_ALAN = 3 };
enum : int { DONE_NONE = 0, DONE_GOAL_RADIUS = 1, DONE_X_BELOW = 2 };
enum : int {
  STAT_AGENT_STEPS = 0,
  STAT_FINISHED = 1,
  STAT_COLLISIONS = 2,
  STAT_LP3_CALLS = 3,
  STAT_OVERFLOW = 4,
  STAT_SUM_ARRIVAL = 5,
  STAT_SUM_ARRIVAL2 = 6,
  STAT_SUM_REWARD = 7
};

struct StepArgs {
  // shape
  int E, N, envs_per_block;
  // simulator parameters (uniform over the batch)
  int k;
  float dt, inv_dt, nd_sq, inv_th, inv_tho, radius, vmax, obst_range_sq;
  // state
  float2* pos;
  float2* vel;
  const float2* pref;
  float2* goal;
  const float2* goal2;
  // RL policy
  const float* action_theta;
  float rl_scale, done_x;
  // ALAN
  float* alan_w;
  const float2* alan_actions;
  uint8_t* alan_action_out;
  const float* alan_uniform_in;
  int A, alan_window;
  float alan_gamma, alan_inv_temp;
  unsigned long long seed;
  // outputs
  float* reward;
  uint8_t* done;
  float* arrival;
  int* env_step;
  int* env_done_cnt;
  int done_mode;
  int* nbr_idx;
  float* nbr_dsq;
  int* nbr_cnt;
  int* onbr_idx;
  int* onbr_cnt;
  unsigned long long* stats;
  // obstacles
  const float4* vert_pd;
  const int4* vert_link;
  const int4* bsp;
  const float4* bsp_seg;
  const int* env_nodes;  // per-env node count, or NULL when the world is shared
  int shared_nodes;      // node count of the shared world
  int vert_stride;       // per-env table stride (0 when shared)
  // 1: stop after the neighbor search (orca_neighbors parity hook)
  int neighbors_only;
  // 1: uniform-grid path; an env spans several blocks, so its step counter is bumped by a
  // separate kernel instead of by the env's first agent
  int grid_path;
};

struct LocalLines {  // LP3 scratch in local memory (rarely touched; lives in L1)
  float4* base;
  ORCA_HD float4 get(int i) const { return base[i]; }
  ORCA_HD void set(int i, float2 point, float2 dir) const {
    float4 v;
    v.x = point.x;
    v.y = point.y;
    v.z = dir.x;
    v.w = dir.y;
    base[i] = v;
  }
};

// Where an agent's neighbor candidates come from.  TileSource: the pre-step snapshot of the
// agent's own env in shared memory, candidates = every other agent of the env in id order.
struct TileSource {
  const float2* env_pos;
  const float2* env_vel;
  int n;
  int self;
  template <class NK>
  ORCA_HD void gather(NK& nk, float2 p) const {
    for (int j = 0; j < n; ++j) {
      if (j == self) continue;
      const float2 q = env_pos[j];
      nk.offer(abs_sq(sub(p, q)), j);
    }
  }
  ORCA_HD float2 pos(int q) const { return env_pos[q]; }
  ORCA_HD float2 vel(int q) const { return env_vel[q]; }
  ORCA_HD int local_id(int q) const { return q; }
};

// atomics: device atomics in the kernels, plain adds in the single-threaded host emulation
ORCA_HD void stat_add_u64(unsigned long long* stats, int slot, unsigned long long v) {
  if (stats != nullptr && v != 0ull) {
#if defined(__CUDA_ARCH__)
    atomicAdd(&stats[slot], v);
#else
    stats[slot] += v;
#endif
  }
}
ORCA_HD void stat_add_f64(unsigned long long* stats, int slot, double v) {
  if (stats != nullptr) {
#if defined(__CUDA_ARCH__)
    atomicAdd(reinterpret_cast<double*>(&stats[slot]), v);
#else
    *reinterpret_cast<double*>(&stats[slot]) += v;
#endif
  }
}
ORCA_HD void counter_inc(int* c) {
#if defined(__CUDA_ARCH__)
  atomicAdd(c, 1);
#else
  *c += 1;
#endif
}

// Everything one agent does in one fused step.  `src` yields the PRE-step state of the other
// agents (shared-memory tile or uniform-grid cells), `L` is the agent's private line storage,
// (p, v) its own pre-step state, `estep` the env's step counter before this step.
template <int K, bool KFULL, int POLICY, class Src>
ORCA_HD void agent_step_body(const StepArgs& a, const int env, const int la, const int g, float2 p, float2 v,
                             const int estep, const Src& src, const Lines L, const unsigned warp_mask) {

  // ---------------- preferred velocity (policy) ----------------
  float2 gdir = v2(0.f, 0.f);
  float2 pref;
  int act = 0;
  float w_act[POLICY == POLICY_ALAN ? ORCA_MAX_ACTIONS : 1];
  if (POLICY == POLICY_EXTERNAL) {
    pref = a.pref[g];
  } else {
    gdir = goal_direction(p, a.goal[g]);
    pref = gdir;
    if (POLICY == POLICY_RL) {
      float sn, cs;
      sincosf(a.action_theta[g], &sn, &cs);
      pref = rotate(gdir, v2(cs, sn));
    }
    if (POLICY == POLICY_ALAN) {
      // softmax over the last rewards; inverse-CDF draw like np.random.choice(p=ps)
      const float* wrow = a.alan_w + (size_t)g * a.A;
      float total = 0.f;
#pragma unroll
      for (int i = 0; i < ORCA_MAX_ACTIONS; ++i) {
        if (i < a.A) {
          w_act[i] = expf(wrow[i] * a.alan_inv_temp);
          total += w_act[i];
        }
      }
      const float u = (a.alan_uniform_in != nullptr)
                          ? a.alan_uniform_in[g]
                          : philox_uniform(a.seed, (uint32_t)g, (uint32_t)estep);
      const float target = u * total;
      float run = 0.f;
      act = a.A - 1;
      bool found = false;
#pragma unroll
      for (int i = 0; i < ORCA_MAX_ACTIONS; ++i) {
        if (i < a.A) {
          run += w_act[i];
          if (!found && run > target) {
            act = i;
            found = true;
          }
        }
      }
      pref = rotate(gdir, a.alan_actions[act]);
      if (a.alan_action_out != nullptr) a.alan_action_out[g] = (uint8_t)act;
    }
  }

  // ---------------- neighbors ----------------
  ObstacleWorld W;
  {
    const size_t voff = (size_t)env * a.vert_stride;
    W.vert_pd = a.vert_pd + voff;
    W.vert_link = a.vert_link + voff;
    W.bsp = a.bsp + voff;
    W.bsp_seg = a.bsp_seg + voff;
    W.n_nodes = (a.env_nodes != nullptr) ? a.env_nodes[env] : a.shared_nodes;
  }
  bool overflow = false;
  float od[ORCA_MAX_OBST_NEIGHBORS];
  int oid[ORCA_MAX_OBST_NEIGHBORS];
  int ocnt = 0;
  if (W.n_nodes > 0) obstacle_neighbors(W, p, a.obst_range_sq, od, oid, &ocnt, &overflow);

  NearestK<K, KFULL> nk;
  nk.init(a.k, a.nd_sq);
  src.gather(nk, p);

  if (a.nbr_idx != nullptr) {
    int c = 0;
#pragma unroll
    for (int s = 0; s < K; ++s) {
      if (s < a.k) {
        a.nbr_idx[(size_t)g * a.k + s] = (nk.id[s] >= 0) ? src.local_id(nk.id[s]) : -1;
        if (a.nbr_dsq != nullptr) a.nbr_dsq[(size_t)g * a.k + s] = (nk.id[s] >= 0) ? nk.d[s] : 0.f;
        c += (nk.id[s] >= 0) ? 1 : 0;
      }
    }
    a.nbr_cnt[g] = c;
  }
  if (a.onbr_idx != nullptr) {
    for (int s = 0; s < ORCA_MAX_OBST_NEIGHBORS; ++s) a.onbr_idx[(size_t)g * ORCA_MAX_OBST_NEIGHBORS + s] = (s < ocnt) ? oid[s] : -1;
    a.onbr_cnt[g] = ocnt;
  }
  if (a.neighbors_only) return;

  // ---------------- ORCA lines ----------------
  int n_obst = 0;
  if (ocnt > 0) n_obst = obstacle_lines(W, p, v, a.radius, a.inv_tho, od, oid, ocnt, L, &overflow);
  int n = n_obst;
  unsigned collisions = 0;
  {
    const float cr = a.radius + a.radius;
#pragma unroll
    for (int s = 0; s < K; ++s) {
      const int j = nk.id[s];
      if (j >= 0) {
        bool hit;
        const float4 ln = agent_line(p, v, src.pos(j), src.vel(j), cr, a.inv_th, a.inv_dt, &hit);
        L.base[n * L.stride] = ln;
        ++n;
        collisions += hit ? 1u : 0u;
      }
    }
  }

  // ---------------- linear programs ----------------
  float2 nv;
  const int fail = lp2(warp_mask, true, L, n, a.vmax, pref, false, nv);
  {
    // every lane takes part (warp-synchronous loops); only lanes whose LP2 failed do work
    float4 proj[K + ORCA_MAX_OBST_LINES];
    LocalLines P;
    P.base = proj;
    lp3(warp_mask, fail < n, L, n, n_obst, fail, a.vmax, P, nv);
    if (fail < n) stat_add_u64(a.stats, STAT_LP3_CALLS, 1ull);
  }

  // ---------------- integrate (Agent::update) ----------------
  v = nv;
  p = add(p, mul(a.dt, v));  // position += velocity * timeStep
  a.pos[g] = p;
  a.vel[g] = v;

  // ---------------- reward / bandit update ----------------
  if (POLICY == POLICY_RL || POLICY == POLICY_ALAN) {
    const float r_goal = dot(v, gdir);
    const float r_polite = dot(v, pref);
    float R;
    if (POLICY == POLICY_RL)
      R = a.rl_scale * r_goal + (1.f - a.rl_scale) * r_polite;
    else
      R = a.alan_gamma * r_goal + (1.f - a.alan_gamma) * r_polite;
    if (a.reward != nullptr) a.reward[g] = R;
    if (POLICY == POLICY_ALAN) {
      float* wrow = a.alan_w + (size_t)g * a.A;
      // every per-action timer advances in lock step, so the 2 s window expires for all of
      // them on the same step: zero the table, then store the fresh reward
      if (a.alan_window > 0 && ((estep + 1) % a.alan_window) == 0) {
        for (int i = 0; i < a.A; ++i) wrow[i] = 0.f;
      }
      wrow[act] = R;
    }
  }

  // ---------------- done test ----------------
  if (a.done_mode != DONE_NONE) {
    if (a.done[g] == 0) {
      bool hit;
      if (a.done_mode == DONE_GOAL_RADIUS) {
        const float2 d = sub(p, a.goal[g]);
        hit = sqrtf(abs_sq(d)) < 2.f * a.radius;
      } else {
        hit = p.x < a.done_x;
      }
      if (hit) {
        a.done[g] = 1;
        const float t_arr = (float)(estep + 1) * a.dt;
        if (a.arrival != nullptr) a.arrival[g] = t_arr;
        if (a.goal2 != nullptr) a.goal[g] = a.goal2[g];
        if (a.env_done_cnt != nullptr) counter_inc(&a.env_done_cnt[env]);
        stat_add_u64(a.stats, STAT_FINISHED, 1ull);
        stat_add_f64(a.stats, STAT_SUM_ARRIVAL, (double)t_arr);
        stat_add_f64(a.stats, STAT_SUM_ARRIVAL2, (double)t_arr * (double)t_arr);
      }
    }
  }
  if (a.env_step != nullptr && la == 0 && !a.grid_path) a.env_step[env] = estep + 1;
  stat_add_u64(a.stats, STAT_COLLISIONS, (unsigned long long)collisions);
  stat_add_u64(a.stats, STAT_OVERFLOW, overflow ? 1ull : 0ull);
}

#if defined(__CUDACC__)

// register budget of the step kernel: 65536 / (threads * min blocks) registers per thread
#ifndef ORCA_STEP_MAX_THREADS
#define ORCA_STEP_MAX_THREADS 256
#endif
#ifndef ORCA_STEP_MIN_BLOCKS
#define ORCA_STEP_MIN_BLOCKS 3
#endif

template <int K, bool KFULL, int POLICY>
__global__ void __launch_bounds__(ORCA_STEP_MAX_THREADS, ORCA_STEP_MIN_BLOCKS) step_small_kernel(const StepArgs a) {
  extern __shared__ float4 smem4[];
  const int tpb = blockDim.x;
  float2* s_pos = reinterpret_cast<float2*>(smem4);
  float2* s_vel = s_pos + tpb;
  float4* s_lines = smem4 + tpb;  // after 2 * tpb float2 = tpb float4

  const int tid = threadIdx.x;
  const int N = a.N;
  const int le = tid / N;
  const int la = tid - le * N;
  const int env = blockIdx.x * a.envs_per_block + le;
  const bool valid = (le < a.envs_per_block) && (env < a.E);
  const int g = env * N + la;

  float2 p = v2(0.f, 0.f), v = v2(0.f, 0.f);
  int estep = 0;
  if (valid) {
    p = a.pos[g];
    v = a.vel[g];
    s_pos[tid] = p;
    s_vel[tid] = v;
    if (a.env_step != nullptr) estep = a.env_step[env];
  }
  __syncthreads();
  const unsigned warp_mask = __ballot_sync(0xffffffffu, valid);  // lanes that run the step
  if (!valid) return;
  Lines L;
  L.base = s_lines + tid;
  L.stride = tpb;
  TileSource src;
  src.env_pos = s_pos + le * N;
  src.env_vel = s_vel + le * N;
  src.n = N;
  src.self = la;
  agent_step_body<K, KFULL, POLICY>(a, env, la, g, p, v, estep, src, L, warp_mask);
}

#endif  // __CUDACC__

}  // namespace orca
