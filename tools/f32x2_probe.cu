// Dev tool: issue cost of the packed FP32 instructions of sm_100a (FADD2 / FMUL2) against their scalar
// forms.  Each thread runs ILP independent dependency chains; time per (lane, component) operation.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; } } while (0)

template <int ILP>
__global__ void scalar_k(float* out, float a, float b, int iters) {
  float x[2 * ILP];
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) x[i] = a + i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) x[i] = x[i] * b;
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) x[i] = x[i] * a;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void packed_k(float* out, float a, float b, int iters) {
  unsigned long long x[ILP], aa, bb;
  float2 t;
  t.x = a; t.y = a; aa = *reinterpret_cast<unsigned long long*>(&t);
  t.x = b; t.y = b; bb = *reinterpret_cast<unsigned long long*>(&t);
#pragma unroll
  for (int i = 0; i < ILP; ++i) { t.x = a + 2 * i + threadIdx.x; t.y = a + 2 * i + 1 + threadIdx.x; x[i] = *reinterpret_cast<unsigned long long*>(&t); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(bb));
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(aa));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { t = *reinterpret_cast<float2*>(&x[i]); s += t.x + t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timed(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  float* out; CK(cudaMalloc(&out, 148 * 8 * 1024 * 4));
  const int iters = 20000;
  for (int warps = 1; warps <= 8; warps *= 2) {  // warps per scheduler: blocks of 128 threads x `warps` per SM
    const int blocks = 148 * warps, tpb = 128;
    const double ops = (double)blocks * tpb * iters * 2.0;  // per component-pair chain step: two multiplications
    float t1 = timed([&] { scalar_k<1><<<blocks, tpb>>>(out, 1.001f, 0.999f, iters); });
    float t2 = timed([&] { packed_k<1><<<blocks, tpb>>>(out, 1.001f, 0.999f, iters); });
    float t3 = timed([&] { scalar_k<4><<<blocks, tpb>>>(out, 1.001f, 0.999f, iters); });
    float t4 = timed([&] { packed_k<4><<<blocks, tpb>>>(out, 1.001f, 0.999f, iters); });
    printf("warps/scheduler %d: ILP1 scalar %.3f ms packed %.3f ms | ILP4 scalar %.3f ms packed %.3f ms  (2-component ops: ILP1 %.0f M, ILP4 %.0f M)\n",
           warps, t1, t2, t3, t4, ops / 1e6, ops * 4 / 1e6);
  }
  return 0;
}
