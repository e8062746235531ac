// Stand-alone timing harness of policy_mlp_tc_kernel (dev tool): builds in seconds, so kernel
// variants (-DORCA_TC_*) can be compared in one GPU call.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -I collision_avoidance_b200/csrc \
//        tools/probes/policy_tc_probe.cu -o policy_tc_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "orca_policy_tc.cuh"

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) {                                                   \
      std::printf("%s: %s\n", #x, cudaGetErrorString(e_));                     \
      return 1;                                                                \
    }                                                                          \
  } while (0)

int main(int argc, char** argv) {
  const long long rows = argc > 1 ? std::atoll(argv[1]) : 1000000;
  const int n_out = 2;
  std::vector<float> h((size_t)rows * 64);
  unsigned s = 12345u;
  for (auto& x : h) {
    s = s * 1664525u + 1013904223u;
    x = ((s >> 8) & 0xffff) / 65536.0f - 0.5f;
  }
  std::vector<float> w(64 * 64 * 2 + 64 * n_out + 64 * 2 + n_out);
  for (auto& x : w) {
    s = s * 1664525u + 1013904223u;
    x = (((s >> 8) & 0xffff) / 65536.0f - 0.5f) * 0.25f;
  }
  float *d_obs, *d_w, *d_out;
  unsigned char* d_flush;
  CK(cudaMalloc(&d_obs, h.size() * 4));
  CK(cudaMalloc(&d_w, w.size() * 4));
  CK(cudaMalloc(&d_out, (size_t)rows * n_out * 4));
  CK(cudaMalloc(&d_flush, 256u << 20));
  CK(cudaMemcpy(d_obs, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  orca::MlpArgs a{};
  a.obs = d_obs;
  a.w1 = d_w;
  a.w2 = d_w + 4096;
  a.w3 = d_w + 8192;
  a.b1 = d_w + 8192 + 64 * n_out;
  a.b2 = a.b1 + 64;
  a.b3 = a.b2 + 64;
  a.out = d_out;
  a.rows = rows;
  a.n_out = n_out;
  CK(cudaFuncSetAttribute(orca::policy_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)orca::mlp_tc_smem_bytes()));
  const long long tiles = (rows + orca::kTcTile - 1) / orca::kTcTile;
  int blocks = (int)((tiles + orca::kTcGroups - 1) / orca::kTcGroups);
  if (blocks > 148) blocks = 148;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float total = 0.f;
  const int iters = 10;
  for (int it = 0; it < iters + 3; ++it) {
    CK(cudaMemset(d_flush, it, 256u << 20));
    cudaEventRecord(e0);
    orca::policy_mlp_tc_kernel<<<blocks, orca::kTcThreads, orca::mlp_tc_smem_bytes()>>>(a);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it >= 3) total += ms;
  }
  std::vector<float> o(8);
  CK(cudaMemcpy(o.data(), d_out, 32, cudaMemcpyDeviceToHost));
  std::printf("%s us %.1f  out0 %.9f %.9f\n", argc > 2 ? argv[2] : "", total / iters * 1e3, o[0], o[1]);
  return 0;
}
