// Probe: does ptxas 12.9.86 (sm_100a) miscompile min(total - g_first, A) when `total` is warp-uniform?
// nvcc -arch=sm_100a -o viaddmnmx_probe viaddmnmx_probe.cu && ./viaddmnmx_probe
#include <cstdio>
struct P { int E, N, A; };
__global__ void v0(int* o, const P p) {
  const int total = p.E * p.N;
  const long long chunk = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long gl = chunk * p.A;
  if (gl >= total) return;
  const int g_first = (int)gl;
  const int n = (total - g_first) < p.A ? (total - g_first) : p.A;
  if ((threadIdx.x & 31) == 0) o[chunk] = n + 1000 * (g_first & 1);
}
__global__ void v1(int* o, const P p) {
  const int total = p.E * p.N;
  const long long chunk = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long gl = chunk * p.A;
  if (gl >= total) return;
  const int g_first = (int)gl;
  int left;
  asm volatile("sub.s32 %0, %1, %2;" : "=r"(left) : "r"(total), "r"(g_first));
  const int n = left < p.A ? left : p.A;
  if ((threadIdx.x & 31) == 0) o[chunk] = n + 1000 * (g_first & 1);
}
__global__ void v2(int* o, const P p, long long last_chunk, int last_n) {
  const int total = p.E * p.N;
  const long long chunk = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long gl = chunk * p.A;
  if (gl >= total) return;
  const int g_first = (int)gl;
  const int n = chunk == last_chunk ? last_n : p.A;
  if ((threadIdx.x & 31) == 0) o[chunk] = n + 1000 * (g_first & 1);
}
int main() {
  P p{320, 10, 28};
  const int total = p.E * p.N, chunks = (total + p.A - 1) / p.A, blocks = (chunks + 7) / 8;
  int* d; cudaMalloc(&d, sizeof(int) * blocks * 8);
  int* h = new int[blocks * 8];
  for (int v = 0; v < 3; ++v) {
    cudaMemset(d, 0, sizeof(int) * blocks * 8);
    if (v == 0) v0<<<blocks, 256>>>(d, p); else if (v == 1) v1<<<blocks, 256>>>(d, p); else v2<<<blocks, 256>>>(d, p, chunks - 1, total - (chunks - 1) * p.A);
    cudaMemcpy(h, d, sizeof(int) * blocks * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < chunks; ++c) {
      const int g = c * p.A, want = (total - g < p.A ? total - g : p.A) + 1000 * (g & 1);
      if (h[c] != want) { if (!bad) printf("variant %d: chunk %d got %d want %d\n", v, c, h[c], want); ++bad; }
    }
    printf("variant %d: %d of %d chunks wrong\n", v, bad, chunks);
  }
  return 0;
}
