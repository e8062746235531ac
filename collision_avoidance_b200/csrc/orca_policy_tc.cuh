// orca_policy_tc.cuh -- the shared policy network (orca_policy.cuh) on the 5th-generation tensor
// cores: tcgen05.mma with TMEM accumulators, hand-written for sm_100a.
//
// This is the one dense contraction next to the ORCA path: for every agent
//     h1 = relu(obs . W1 + b1) ; h2 = relu(h1 . W2 + b2) ; out = h2 . W3 + b3        (64-64-64-out)
// i.e. two [rows x 64] x [64 x 64] GEMMs whose A operand is produced on chip.  On the FP32 pipes
// (policy_mlp_kernel) it is FMA-bound at ~0.41 ms per million rows; here the products run as
// tcgen05.mma kind::tf32 and the kernel is bound by reading the observations once.
//
// Accuracy: a single TF32 product drops 13 mantissa bits of each operand (~1e-3 relative), too
// coarse for a parity target of 1e-4 on O(1) outputs.  Every product is therefore split
//     a = a_hi + a_lo,  w = w_hi + w_lo      (x_hi = x with the low 13 mantissa bits cleared)
//     a . w  ~=  a_hi . w_hi + a_hi . w_lo + a_lo . w_hi                 ("3xTF32")
// accumulated in FP32 in TMEM: error ~2^-21 relative per product, the same order as FP32 itself.
// Tensor-core time triples and still hides behind the HBM read.
//
// One CTA = FOUR independent groups of 128 threads; a group = 128 agents (one TMEM lane each),
// persistent over row tiles.  The groups share the weight tiles and nothing else (own A tile, own
// accumulator columns, own mbarrier, named barriers): while one waits for its MMAs the others run
// their epilogues.  Per group:
//   - weights (B operands): split once per CTA into K-major SWIZZLE_NONE core-matrix layout in
//     shared memory, 4 x 16 KB;
//   - observation rows: a warp stages ITS 32 rows (8 KB of contiguous global memory) with coalesced
//     16-byte cp.async into a padded buffer and thread r reads row r back with conflict-free
//     LDS.128.  (Per-thread row loads -- 32 lines per warp-level load -- kept the L1 data pipe at
//     63-68 % of its wavefront peak: that, not HBM and not the tensor pipe, bound the first version.)
//     The next tile's copies are issued as soon as this tile's rows are in registers;
//   - the A operand lives in TENSOR MEMORY (tcgen05.mma with A from TMEM): row r of the tile is
//     TMEM lane r, which is exactly what thread r owns -- it splits its own row and writes it with
//     tcgen05.st, no shared-memory staging, no proxy fence.  With N = 64 an MMA that fetches A from
//     shared memory is bound by that fetch (4 KB of A + 2 KB of B per 128x64x8 MMA, measured 53 ns
//     each); from TMEM only the 2 KB of B cross the shared-memory port.  A group has ONE A tile
//     (64 columns): a_lo goes in first (8 MMAs with w_hi), then a_hi (16 MMAs with w_lo and w_hi) --
//     and a_hi is the row as it is: kind::tf32 ignores the low 13 mantissa bits of its operands;
//   - an elected lane of a converged warp issues the MMAs (elect.sync: issued under `lane == 0`
//     ptxas wraps every tcgen05.mma in a loop over the warp's operand values) and commits to an
//     mbarrier; every thread then pulls ITS row of the accumulator with tcgen05.ld, adds the bias,
//     applies ReLU, splits again and stores the row back as the A operand of the second GEMM; the
//     64 -> out head is a handful of FMAs on the registers that tcgen05.ld delivered.
// Tensor memory per group: A 64 + accumulator 64 columns (4 x 128 = all 512); 128 registers per
// thread; shared memory ~205 KB (64 KB weights + 16 x 8.5 KB staging).  -DORCA_TC_GROUPS=2 builds
// the round-1 arrangement (a_hi and a_lo resident, 192 columns per group) for A/B runs.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "orca_policy.cuh"

namespace orca {

#ifndef ORCA_TC_GROUPS
#define ORCA_TC_GROUPS 4  // dev A/B: 2 = the round-1 pipeline (a_hi and a_lo resident, next row prefetched into registers)
#endif
constexpr int kTcTile = 128;     // rows per CTA tile = MMA M = TMEM lanes
constexpr int kTcGroups = ORCA_TC_GROUPS;  // independent 128-thread pipelines per CTA
constexpr bool kTcTwoPhase = kTcGroups > 2;  // one A tile per group, a_lo and a_hi take turns in it
constexpr int kTcThreads = 128 * kTcGroups;
constexpr int kTcN = 64;         // MMA N = hidden width
constexpr int kTcK = 64;         // reduction length of both GEMMs
// per group: a_hi | a_lo | accumulator, 64 columns each; two-phase: A | accumulator
constexpr int kTcAccCols = (kTcTwoPhase ? 2 : 3) * kTcN;
constexpr int kTcDCol = kTcAccCols - kTcN;   // first accumulator column of a group
constexpr int kTcTmemCols = 512;             // power of two >= kTcGroups * kTcAccCols
static_assert(kTcGroups * kTcAccCols <= kTcTmemCols, "tensor memory has 512 columns");

// byte sizes / strides of the K-major SWIZZLE_NONE operand tiles: a core matrix is 8 rows x 16 B
// (4 tf32), stored as 128 contiguous bytes; core matrices of consecutive 8-row groups follow each
// other (SBO = 128 B), the next 16-byte chunk along K starts after all row groups (LBO).
constexpr uint32_t kTcSbo = 128;
constexpr uint32_t kTcLboB = (kTcN / 8) * 128;     // 1024
constexpr uint32_t kTcBytesB = kTcN * kTcK * 4;     // 16 KB

// Observation rows are staged per WARP (32 rows = 8 KB of contiguous global memory) with coalesced
// 16-byte cp.async into rows of 68 floats: thread r then reads row r with 16 conflict-free LDS.128
// (68 words per row: eight lanes cover the 32 banks).  One buffer per warp: a tile's rows are in
// registers before the next tile's copies are issued.
constexpr int kTcStagePitch = kTcK + 4;                       // floats per staged row
constexpr int kTcStageWarpBytes = 32 * kTcStagePitch * 4;     // 8,704
inline size_t mlp_tc_smem_bytes() {
  return 4 * (size_t)kTcBytesB + (size_t)(kTcThreads / 32) * kTcStageWarpBytes +
         sizeof(float) * (kMlpHidden * kMlpMaxOut + 2 * kMlpHidden + kMlpMaxOut) + 64 /* mbarrier + tmem base */ +
         1024 /* alignment slack */;
}

#if defined(__CUDACC__)

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (SM100 UMMA, version 1), SWIZZLE_NONE, K-major
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address, bits [0, 14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;    // leading-dimension byte offset, bits [16, 30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;    // stride-dimension byte offset, bits [32, 46)
  d |= (uint64_t)1 << 46;                               // descriptor version 1 (Blackwell)
  return d;                                             // base offset 0, layout type 0 = SWIZZLE_NONE
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = 64, dense, no negate
__device__ __forceinline__ uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcTile >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (lane = row, column = k), B from shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t zero = 0u;  // disable-output-lane mask: none
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(zero)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void group_sync(int group) { asm volatile("bar.sync %0, 128;\n" ::"r"(group + 1) : "memory"); }
// 16 bytes global -> shared without passing through registers; src_bytes = 0 fills with zeros
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
// one lane of the (converged) warp: the form the compiler recognises for single-thread tcgen05 issue
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0u;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// this thread's row (64 floats in v) -> its TMEM lane of the a_hi (columns [0, 64)) and a_lo
// (columns [64, 128)) operands of the group
__device__ __forceinline__ void store_row_split(uint32_t lane_base, const float* v) {
#pragma unroll
  for (int q = 0; q < kTcK / 16; ++q) {
    float hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      hi[i] = tf32_hi(v[16 * q + i]);
      lo[i] = v[16 * q + i] - hi[i];
    }
    tmem_st16(lane_base + (uint32_t)(16 * q), hi);
    tmem_st16(lane_base + (uint32_t)(kTcN + 16 * q), lo);
  }
  tmem_st_wait();
}

// two-phase pipeline: the a_lo (part 0) or a_hi (part 1) half of this thread's row -> its TMEM lane of
// the group's single A tile (columns [0, 64))
template <int PART>
__device__ __forceinline__ void store_row_part(uint32_t lane_base, const float* v) {
#pragma unroll
  for (int q = 0; q < kTcK / 16; ++q) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float hi = tf32_hi(v[16 * q + i]);
      // kind::tf32 reads the upper 19 bits of its 32-bit operands: the low 13 mantissa bits are ignored, so
      // the row itself IS the a_hi operand (bit-identical results to storing the truncated values, measured:
      // 109 -> 97 us; tests/test_gpu_policy.py holds the FP32-level accuracy a rounding tensor core would break)
      x[i] = PART ? v[16 * q + i] : v[16 * q + i] - hi;
    }
    tmem_st16(lane_base + (uint32_t)(16 * q), x);
  }
  tmem_st_wait();
}
// phase 0: D = a_lo . w_hi (the small terms first, 8 MMAs); phase 1: D += a_hi . w_lo + a_hi . w_hi (16 MMAs)
template <int PART>
__device__ __forceinline__ void issue_gemm_part(uint32_t tmem_group, uint32_t b_hi, uint32_t b_lo, uint32_t idesc) {
#if defined(ORCA_TC_DEV_NO_MMA)  // dev A/B (wrong results): no tensor work, the commit arrives at once
  return;
#endif
  const uint32_t d = tmem_group + (uint32_t)kTcDCol;
#pragma unroll
  for (int ks = 0; ks < kTcK / 8; ++ks) {
    const uint32_t at = tmem_group + (uint32_t)(8 * ks);
    const uint64_t dbh = make_desc(b_hi + (uint32_t)ks * 2u * kTcLboB, kTcLboB, kTcSbo);
    if (PART == 0) {
      mma_tf32_ts(d, at, dbh, idesc, 1u);  // the accumulator starts from the bias row (run_layer)
    } else {
      const uint64_t dbl = make_desc(b_lo + (uint32_t)ks * 2u * kTcLboB, kTcLboB, kTcSbo);
      mma_tf32_ts(d, at, dbl, idesc, 1u);
      mma_tf32_ts(d, at, dbh, idesc, 1u);
    }
  }
}

// D[tmem] = A . B over K = 64 with the 3xTF32 split: 8 k-steps x 3 MMAs (issued by ONE thread).
// tmem_group: first column of the group (a_hi | a_lo | D), lane 0.
__device__ __forceinline__ void issue_gemm(uint32_t tmem_group, uint32_t b_hi, uint32_t b_lo, uint32_t idesc) {
  const uint32_t d = tmem_group + (uint32_t)kTcDCol;
#pragma unroll
  for (int ks = 0; ks < kTcK / 8; ++ks) {  // one MMA consumes K = 8 tf32: 8 TMEM columns of A, two 16-byte chunks of B
    const uint32_t ah = tmem_group + (uint32_t)(8 * ks), al = ah + (uint32_t)kTcN;
    const uint64_t dbh = make_desc(b_hi + (uint32_t)ks * 2u * kTcLboB, kTcLboB, kTcSbo);
    const uint64_t dbl = make_desc(b_lo + (uint32_t)ks * 2u * kTcLboB, kTcLboB, kTcSbo);
    mma_tf32_ts(d, al, dbh, idesc, ks > 0 ? 1u : 0u);  // small terms first
    mma_tf32_ts(d, ah, dbl, idesc, 1u);
    mma_tf32_ts(d, ah, dbh, idesc, 1u);
  }
}

}  // namespace tc

__global__ void __launch_bounds__(kTcThreads, 1) policy_mlp_tc_kernel(const MlpArgs a) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, lane = tid & 31;
  // warp index through a shuffle: the compiler then knows that everything derived from it (tensor-memory
  // addresses, descriptors, barrier addresses) is warp-uniform and issues the MMAs from uniform registers
  // directly instead of wrapping each one in a loop over the distinct operand values of the warp
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int group = warp >> 2, gtid = tid & 127;  // pipeline of this thread, row of the tile / TMEM lane
  // the issuing thread of group g is lane 0 of the group's warp g: the four issuers sit on four different schedulers
  const bool issuer_warp = (warp & 3) == (group & 3);
  uint8_t* w1_hi = base;
  uint8_t* w1_lo = w1_hi + kTcBytesB;
  uint8_t* w2_hi = w1_lo + kTcBytesB;
  uint8_t* w2_lo = w2_hi + kTcBytesB;
  uint8_t* stage = w2_lo + kTcBytesB + (size_t)warp * kTcStageWarpBytes;  // this warp's 32 staged rows
  float* w3 = reinterpret_cast<float*>(w2_lo + kTcBytesB + (size_t)(kTcThreads / 32) * kTcStageWarpBytes);  // [n_out][64]
  float* b1 = w3 + kMlpHidden * kMlpMaxOut;
  float* b2 = b1 + kMlpHidden;
  float* b3 = b2 + kMlpHidden;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b3 + kMlpMaxOut);  // one per group
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kTcGroups);
  uint64_t* bar = bars + group;

  const long long tiles = (a.rows + kTcTile - 1) / kTcTile;
  const long long first_tile = (long long)blockIdx.x * kTcGroups + group, tile_step = (long long)gridDim.x * kTcGroups;
  // coalesced copy of this warp's 32 rows of `tile` (512 chunks of 16 bytes, 16 per lane) into its buffer
  const uint32_t s_stage = tc::smem_u32(stage);
  const uint32_t s_stage_lane = s_stage + (uint32_t)((lane >> 4) * kTcStagePitch * 4 + (lane & 15) * 16);
  auto stage_rows = [&](long long tile) {
    if (tile < tiles) {
      const long long row0 = tile * kTcTile + (gtid & ~31);
      const float* src = a.obs + row0 * kMlpIn;
#if !defined(ORCA_TC_DEV_NO_LOAD)  // dev A/B (wrong results): no observation traffic
      if (row0 + 32 <= a.rows) {  // warp-uniform: all 32 rows exist -- constant offsets from one address pair
        const float* src_lane = src + lane * 4;
#pragma unroll
        for (int i = 0; i < 16; ++i) tc::cp_async16(s_stage_lane + (uint32_t)(i * 2 * kTcStagePitch * 4), src_lane + i * 128, 16u);
      } else
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = i * 32 + lane, r = c >> 4, q = c & 15;
        const bool have = row0 + r < a.rows;
        tc::cp_async16(s_stage + (uint32_t)(r * kTcStagePitch * 4 + q * 16), have ? src + c * 4 : a.obs, have ? 16u : 0u);
      }
#endif
    }
    tc::cp_async_commit();
  };
  stage_rows(first_tile);  // in flight during the set-up

  // ---- one-time setup: weights (split, K-major core-matrix layout), barrier, TMEM ----
  for (int i = tid; i < kTcK * kTcN; i += kTcThreads) {
    const int k = i >> 6, n = i & 63;  // global [k][n] (n contiguous): coalesced reads
    const uint32_t off = (uint32_t)(k >> 2) * kTcLboB + (uint32_t)(n >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 3) * 4u;
    const float x1 = __ldg(a.w1 + i), x2 = __ldg(a.w2 + i);
    const float h1 = tc::tf32_hi(x1), h2 = tc::tf32_hi(x2);
    *reinterpret_cast<float*>(w1_hi + off) = h1;
    *reinterpret_cast<float*>(w1_lo + off) = x1 - h1;
    *reinterpret_cast<float*>(w2_hi + off) = h2;
    *reinterpret_cast<float*>(w2_lo + off) = x2 - h2;
  }
  // head weights transposed to [o][n]: the head reads them as float4 over n
  for (int i = tid; i < kMlpHidden * a.n_out; i += kTcThreads) w3[(i % a.n_out) * kMlpHidden + i / a.n_out] = __ldg(a.w3 + i);
  if (tid < kMlpHidden) {
    b1[tid] = __ldg(a.b1 + tid);
    b2[tid] = __ldg(a.b2 + tid);
  }
  if (tid < a.n_out) b3[tid] = __ldg(a.b3 + tid);
  if (tid == 0) {
    for (int gq = 0; gq < kTcGroups; ++gq) tc::mbar_init(tc::smem_u32(bars + gq), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {  // one warp allocates the accumulator columns and passes the permit on
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc::smem_u32(tmem_slot)), "n"(kTcTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc::fence_async_smem();  // the weight tiles were written through the generic proxy, the MMA reads them through the async proxy
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_alloc = *tmem_slot;
  const uint32_t tmem_base = tmem_alloc + (uint32_t)(group * kTcAccCols);  // this group's accumulator columns
  const uint32_t idesc = tc::make_idesc();
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's 32 TMEM lanes
  const uint32_t s_w1_hi = tc::smem_u32(w1_hi), s_w1_lo = tc::smem_u32(w1_lo);
  const uint32_t s_w2_hi = tc::smem_u32(w2_hi), s_w2_lo = tc::smem_u32(w2_lo);
  const uint32_t s_bar = tc::smem_u32(bar);
  uint32_t parity = 0;
  float v[kTcK];  // this thread's row: observation, then h1, then h2

  // one layer: v (this thread's row of A) -> v (its row of A . W), 3xTF32
  auto run_layer = [&](uint32_t s_hi, uint32_t s_lo, const float* bias) {
    if constexpr (kTcTwoPhase) {
      // the bias enters through the accumulator: every thread writes the bias row into its lane of D
      // (free since the previous layer was read out) and all MMAs accumulate -- 16 LDS.128 + 4 tcgen05.st
      // instead of 64 additions after the read-out
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float bq[16];
#pragma unroll
        for (int n = 0; n < 16; n += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(bias + 16 * q + n);  // broadcast
          bq[n + 0] = bb.x;
          bq[n + 1] = bb.y;
          bq[n + 2] = bb.z;
          bq[n + 3] = bb.w;
        }
        tc::tmem_st16(lane_base + (uint32_t)(kTcDCol + 16 * q), bq);
      }
      // a_lo -> the A tile, 8 MMAs, wait (the MMAs have read the tile), a_hi -> the same tile, 16 MMAs
      tc::store_row_part<0>(lane_base, v);
      tc::fence_before_sync();
      tc::group_sync(group);
      if (issuer_warp && tc::elect_one()) {
        tc::fence_after_sync();
        tc::issue_gemm_part<0>(tmem_base, s_hi, s_lo, idesc);
        tc::mma_commit(s_bar);
      }
      tc::mbar_wait(s_bar, parity);
      parity ^= 1u;
      tc::fence_after_sync();
      tc::store_row_part<1>(lane_base, v);
      tc::fence_before_sync();
      tc::group_sync(group);
      if (issuer_warp && tc::elect_one()) {
        tc::fence_after_sync();
        tc::issue_gemm_part<1>(tmem_base, s_hi, s_lo, idesc);
        tc::mma_commit(s_bar);
      }
    } else {
      tc::store_row_split(lane_base, v);
      tc::fence_before_sync();
      tc::group_sync(group);
      if (issuer_warp && tc::elect_one()) {
        tc::fence_after_sync();
        tc::issue_gemm(tmem_base, s_hi, s_lo, idesc);
        tc::mma_commit(s_bar);
      }
    }
    tc::mbar_wait(s_bar, parity);
    parity ^= 1u;
    tc::fence_after_sync();
#pragma unroll
    for (int q = 0; q < 4; ++q) tc::tmem_ld16(lane_base + (uint32_t)(kTcDCol + 16 * q), v + 16 * q);
    tc::tmem_ld_wait();
  };
  auto bias_relu = [&](const float* b) {
    if constexpr (kTcTwoPhase) {  // the bias is already in the accumulator
#pragma unroll
      for (int n = 0; n < kTcN; ++n) v[n] = fmaxf(v[n], 0.f);
      return;
    }
#pragma unroll
    for (int n = 0; n < kTcN; n += 4) {
      const float4 bb = *reinterpret_cast<const float4*>(b + n);  // broadcast
      v[n + 0] = fmaxf(v[n + 0] + bb.x, 0.f);
      v[n + 1] = fmaxf(v[n + 1] + bb.y, 0.f);
      v[n + 2] = fmaxf(v[n + 2] + bb.z, 0.f);
      v[n + 3] = fmaxf(v[n + 3] + bb.w, 0.f);
    }
  };

  for (long long tile = first_tile; tile < tiles; tile += tile_step) {
    const long long row = tile * kTcTile + gtid;
    // this thread's row out of the warp's staging buffer, then the next tile's copies into it
    tc::cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int c = 0; c < kTcK / 4; ++c) {
      const float4 x = *reinterpret_cast<const float4*>(stage + (size_t)lane * (kTcStagePitch * 4) + c * 16);
      v[4 * c + 0] = x.x;
      v[4 * c + 1] = x.y;
      v[4 * c + 2] = x.z;
      v[4 * c + 3] = x.w;
    }
    __syncwarp();
    stage_rows(tile + tile_step);

    run_layer(s_w1_hi, s_w1_lo, b1);
    bias_relu(b1);
    run_layer(s_w2_hi, s_w2_lo, b2);
    bias_relu(b2);

    // ---- head: 64 -> n_out on the registers ----
    if (row < a.rows) {
      for (int o = 0; o < a.n_out; ++o) {
        float acc = b3[o];
#pragma unroll
        for (int n = 0; n < kTcN; n += 4) {
          const float4 w = *reinterpret_cast<const float4*>(w3 + o * kMlpHidden + n);  // broadcast
          acc = fmaf(v[n + 0], w.x, acc);
          acc = fmaf(v[n + 1], w.y, acc);
          acc = fmaf(v[n + 2], w.z, acc);
          acc = fmaf(v[n + 3], w.w, acc);
        }
        a.out[row * a.n_out + o] = acc;
      }
    }
  }
  tc::cp_async_wait_all();

  // ---- teardown ----
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_alloc), "n"(kTcTmemCols) : "memory");
  }
}

#endif  // __CUDACC__

}  // namespace orca
