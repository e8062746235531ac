// orca_policy.cuh -- shared policy network of the RL shell, evaluated for every agent of the batch.
//
// The reference trains one weight-shared model for all agents (run_rllib.py:35-52, CustomModel1):
//     fc1: 64 -> 64, ReLU ;  fc2: 64 -> 64, ReLU ;  fc_out: 64 -> num_outputs, linear
// on the 64-float laser observation of Collision_Avoidance_Env (collision_avoidence_env.py:231-277,
// 16 rays x (hit.x, hit.y, vel.x, vel.y)).  In the reference that forward pass runs in TensorFlow
// inside RLlib, once per env per step; here one kernel evaluates it for all E*N agents straight
// from the observation buffer the observe kernel wrote, so the RL loop
//     obs -> policy -> action -> orca_env_step -> obs
// never leaves the device.
//
// Arithmetic: float32 with fused multiply-adds (explicit fmaf: the library is built with
// -fmad=false for the ORCA code, see build.py), sums over k in ascending order.  Parity target:
// a float32 torch evaluation of the same layers, |delta| <= 1e-4 (tests/test_gpu_policy.py).
//
// Kernel shape: a block of 256 threads owns a tile of 128 agents.  The observation tile and both
// weight matrices sit in shared memory; each thread accumulates an 8 agents x 4 outputs register
// tile, k unrolled by 4 (12 LDS.128 per 128 FMA).  The hidden activations go back to shared memory
// (row-major, stride 68 floats so that the two agent rows a warp touches land in different banks);
// the 64 -> num_outputs head is one dot product per (agent, output).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace orca {

constexpr int kMlpIn = 64;       // observation width: 16 rays x 4
constexpr int kMlpHidden = 64;   // fc1 / fc2 width
constexpr int kMlpMaxOut = 8;    // num_outputs upper bound (PPO on Box(1): mean + log-std = 2)
constexpr int kMlpTile = 128;    // agents per block
constexpr int kMlpThreads = 256;
constexpr int kMlpStride = 68;   // floats per staged row (64 + 4 pad: 16-byte aligned, bank shift 4)

struct MlpArgs {
  const float* obs;  // [rows][64]
  long long rows;
  const float* w1;   // [64 in][64 out]
  const float* b1;   // [64]
  const float* w2;   // [64][64]
  const float* b2;   // [64]
  const float* w3;   // [64][n_out]
  const float* b3;   // [n_out]
  int n_out;
  float* out;        // [rows][n_out]
};

inline size_t mlp_smem_bytes() {
  // two activation tiles + w1 + w2 + w3 + biases
  return sizeof(float) * (2 * (size_t)kMlpTile * kMlpStride + 2 * kMlpIn * kMlpHidden + kMlpHidden * kMlpMaxOut +
                          2 * kMlpHidden + kMlpMaxOut);
}

#if defined(__CUDACC__)

// One hidden layer for the block's tile: dst[a][n] = relu(b[n] + sum_k src[a][k] * w[k][n]).
// Thread (ty = tid / 16, tx = tid % 16) owns agents ty + 16 r (r < 8) and outputs 4 tx .. 4 tx + 3.
__device__ __forceinline__ void mlp_hidden_layer(const float* __restrict__ src, const float* __restrict__ w,
                                                 const float* __restrict__ bias, float* __restrict__ dst) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][4];
  {
    const float4 b = *reinterpret_cast<const float4*>(bias + 4 * tx);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      acc[r][0] = b.x;
      acc[r][1] = b.y;
      acc[r][2] = b.z;
      acc[r][3] = b.w;
    }
  }
#pragma unroll 2
  for (int k = 0; k < kMlpIn; k += 4) {
    float4 wv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wv[i] = *reinterpret_cast<const float4*>(w + (k + i) * kMlpHidden + 4 * tx);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(src + (ty + 16 * r) * kMlpStride + k);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[r][0] = fmaf(av[i], wv[i].x, acc[r][0]);
        acc[r][1] = fmaf(av[i], wv[i].y, acc[r][1]);
        acc[r][2] = fmaf(av[i], wv[i].z, acc[r][2]);
        acc[r][3] = fmaf(av[i], wv[i].w, acc[r][3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    float4 o;
    o.x = fmaxf(acc[r][0], 0.f);
    o.y = fmaxf(acc[r][1], 0.f);
    o.z = fmaxf(acc[r][2], 0.f);
    o.w = fmaxf(acc[r][3], 0.f);
    *reinterpret_cast<float4*>(dst + (ty + 16 * r) * kMlpStride + 4 * tx) = o;
  }
}

__global__ void __launch_bounds__(kMlpThreads, 2) policy_mlp_kernel(const MlpArgs a) {
  extern __shared__ float4 mlp_smem4[];
  float* xa = reinterpret_cast<float*>(mlp_smem4);   // observation tile, later fc2 activations
  float* xb = xa + kMlpTile * kMlpStride;            // fc1 activations
  float* w1 = xb + kMlpTile * kMlpStride;
  float* w2 = w1 + kMlpIn * kMlpHidden;
  float* w3 = w2 + kMlpHidden * kMlpHidden;          // [64][n_out]
  float* b1 = w3 + kMlpHidden * kMlpMaxOut;
  float* b2 = b1 + kMlpHidden;
  float* b3 = b2 + kMlpHidden;
  const int tid = threadIdx.x;

  // weights: 2 x 16 KB, the same for every block (L2-resident after the first wave)
  for (int i = tid; i < kMlpIn * kMlpHidden / 4; i += kMlpThreads) {
    reinterpret_cast<float4*>(w1)[i] = __ldg(reinterpret_cast<const float4*>(a.w1) + i);
    reinterpret_cast<float4*>(w2)[i] = __ldg(reinterpret_cast<const float4*>(a.w2) + i);
  }
  for (int i = tid; i < kMlpHidden * a.n_out; i += kMlpThreads) w3[i] = __ldg(a.w3 + i);
  if (tid < kMlpHidden) {
    b1[tid] = __ldg(a.b1 + tid);
    b2[tid] = __ldg(a.b2 + tid);
  }
  if (tid < a.n_out) b3[tid] = __ldg(a.b3 + tid);

  // persistent over tiles: grid = min(#tiles, 2 blocks per SM), weights staged once per block
  const long long tiles = (a.rows + kMlpTile - 1) / kMlpTile;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long row0 = tile * kMlpTile;
    // observation rows are 256 B each, the tile is one contiguous 32 KB run: coalesced float4 loads
    for (int i = tid; i < kMlpTile * (kMlpIn / 4); i += kMlpThreads) {
      const int r = i >> 4, c = i & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < a.rows) v = __ldcs(reinterpret_cast<const float4*>(a.obs + (row0 + r) * kMlpIn) + c);  // read once
      *reinterpret_cast<float4*>(xa + r * kMlpStride + 4 * c) = v;
    }
    __syncthreads();  // tile (and, first time round, the weights) staged
    mlp_hidden_layer(xa, w1, b1, xb);
    __syncthreads();
    mlp_hidden_layer(xb, w2, b2, xa);
    __syncthreads();
    // head: thread -> (agent, output); 128 * n_out dot products of length 64
    for (int i = tid; i < kMlpTile * a.n_out; i += kMlpThreads) {
      const int r = i / a.n_out, o = i - r * a.n_out;
      float acc = b3[o];
      const float* h = xa + r * kMlpStride;
#pragma unroll 8
      for (int k = 0; k < kMlpHidden; ++k) acc = fmaf(h[k], w3[k * a.n_out + o], acc);
      if (row0 + r < a.rows) a.out[(row0 + r) * a.n_out + o] = acc;
    }
    __syncthreads();  // xa is overwritten by the next tile
  }
}

#endif  // __CUDACC__

}  // namespace orca
