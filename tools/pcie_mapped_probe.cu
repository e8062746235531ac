// Dev tool: bandwidth of GPU-initiated PCIe traffic on mapped (pinned) host memory -- what the
// direct route of orca_step_host is made of.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
//   reads  : every thread loads one float2 from the mapped host buffer (goals)
//   writes : every thread stores two float2 into mapped host buffers (new pos / vel)
// against the copy engine over the same buffers.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); \
      std::exit(1);                                                                \
    }                                                                              \
  } while (0)

__global__ void rd(const float2* __restrict__ h, float2* __restrict__ d, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = h[i];
}
__global__ void wr(const float2* __restrict__ d, float2* __restrict__ h1, float2* __restrict__ h2, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float2 v = d[i];
    h1[i] = v;
    h2[i] = v;
  }
}
__global__ void rw(const float2* __restrict__ hin, float2* __restrict__ h1, float2* __restrict__ h2, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float2 v = hin[i];
    h1[i] = v;
    h2[i] = v;
  }
}

template <class F>
static float timed(F f, int reps = 20) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main() {
  const int n = 1 << 20;  // agents
  const size_t bytes = (size_t)n * sizeof(float2);
  float2 *h_in, *h_o1, *h_o2, *d_a, *d_b;
  CK(cudaHostAlloc(&h_in, bytes, cudaHostAllocMapped));
  CK(cudaHostAlloc(&h_o1, bytes, cudaHostAllocMapped));
  CK(cudaHostAlloc(&h_o2, bytes, cudaHostAllocMapped));
  CK(cudaMalloc(&d_a, bytes));
  CK(cudaMalloc(&d_b, bytes));
  float2 *m_in, *m_o1, *m_o2;
  CK(cudaHostGetDevicePointer(&m_in, h_in, 0));
  CK(cudaHostGetDevicePointer(&m_o1, h_o1, 0));
  CK(cudaHostGetDevicePointer(&m_o2, h_o2, 0));
  const int tpb = 256, blocks = n / tpb;
  float t;
  t = timed([&] { rd<<<blocks, tpb>>>(m_in, d_a, n); });
  std::printf("kernel reads 8 MiB from mapped host : %.3f ms  %.1f GB/s\n", t, bytes / t / 1e6);
  t = timed([&] { wr<<<blocks, tpb>>>(d_a, m_o1, m_o2, n); });
  std::printf("kernel writes 16 MiB to mapped host : %.3f ms  %.1f GB/s\n", t, 2 * bytes / t / 1e6);
  t = timed([&] { rw<<<blocks, tpb>>>(m_in, m_o1, m_o2, n); });
  std::printf("kernel reads 8 + writes 16 MiB      : %.3f ms\n", t);
  t = timed([&] { CK(cudaMemcpyAsync(d_a, h_in, bytes, cudaMemcpyHostToDevice)); });
  std::printf("copy engine H2D 8 MiB               : %.3f ms  %.1f GB/s\n", t, bytes / t / 1e6);
  t = timed([&] {
    CK(cudaMemcpyAsync(h_o1, d_a, bytes, cudaMemcpyDeviceToHost));
    CK(cudaMemcpyAsync(h_o2, d_b, bytes, cudaMemcpyDeviceToHost));
  });
  std::printf("copy engine D2H 16 MiB              : %.3f ms  %.1f GB/s\n", t, 2 * bytes / t / 1e6);
  cudaStream_t s1;
  CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  cudaEvent_t ev, ev2;
  CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&ev2, cudaEventDisableTiming));
  t = timed([&] {  // copy-engine upload running under a kernel that writes to the host
    CK(cudaEventRecord(ev, 0));
    CK(cudaStreamWaitEvent(s1, ev, 0));
    CK(cudaMemcpyAsync(d_b, h_in, bytes, cudaMemcpyHostToDevice, s1));
    wr<<<blocks, tpb>>>(d_a, m_o1, m_o2, n);
    CK(cudaEventRecord(ev2, s1));
    CK(cudaStreamWaitEvent(0, ev2, 0));
  });
  std::printf("CE H2D 8 MiB || kernel writes 16 MiB : %.3f ms\n", t);
  return 0;
}
