// l2_tc.cu -- K2: exact 2-NN under the L2 norm on the 5th-gen tensor cores (tcgen05 + TMEM + TMA)
// for integer-valued 128-d descriptors (SIFT, FeatureDetector.cpp:20-24: 128 floats in [0,255]).
//
// Replaces knnMatch at Mapper/libMapper/FeatureMatcher.cpp:48-49.  Parity: cv::BFMatcher(NORM_L2),
// bit-exact.  Why exact: operands are fp16 (integers <= 510 are exact), every product and every
// partial sum is an integer of magnitude < 2^24, so the fp32 accumulator in TMEM holds the exact
// integer  nb - 2 a.b  whatever the accumulation order.  d^2 = na + (nb - 2 a.b).
//
// The distance matrix is the dense contraction it is:
//     acc[q][t] = sum_k (-2 a_qk) * b_tk   +   1*lo(nb_t) + 2048*mid(nb_t) + 2048*(2048*hi(nb_t))
// i.e. the train-row norm rides in one extra K=16 MMA step (three fp16-exact pieces), so the
// epilogue only has to select minima -- no per-element add.
//
// CTA = 12 warps, persistent, one per SM:
//   warp 0      TMA producer: A (256 query rows, resident per work item) and a 4-stage ring of
//               B tiles (128 train rows), SWIZZLE_128B for the two 64-wide K atoms and
//               SWIZZLE_32B for the 16-wide norm block
//   warp 1      MMA issuer: one thread, tcgen05.mma.kind::f16 M=128 N=128 K=16, 9 steps per
//               (m-tile, B tile); four 128-column fp32 accumulators in TMEM (2 m-tiles x 2 stages)
//   warp 2      TMEM allocator (512 columns)
//   warps 4-11  epilogue: tcgen05.ld 32x32b.x32 -> registers, min-tree over 32 columns, rare
//               slow path inserts into the per-row running top-2 (strict <, ascending columns
//               => lowest index wins ties); overlaps the next tile's MMAs via the 2-deep
//               accumulator ring
// Work item = (pair, 256-row query super-tile); items are dealt round-robin to the CTAs so that
// concurrently running CTAs share the same train image in L2.
//
// Algorithmic work: 2 * nq * nt * 128 FLOP per pair.
#include "common.cuh"
#include "kernels.h"

namespace pm {

static constexpr int TC_BM = 128;             // rows per m-tile (UMMA M)
static constexpr int TC_MT = 2;               // m-tiles per CTA
static constexpr int TC_ROWS = TC_BM * TC_MT; // query rows per work item
static constexpr int TC_BN = 128;             // train rows per B tile (UMMA N)
static constexpr int TC_STAGES = 3;   // 3 x 36 KB: leaves ~40 KB of shared memory per SM for the tail kernels to co-reside
static constexpr int TC_ATOM_BYTES = TC_BM * 128;                 // 128 rows x 128 B  (SW128 atom column)
static constexpr int TC_EXT_BYTES = TC_BM * 32;                   // 128 rows x 32 B   (SW32 block)
static constexpr int TC_TILE_BYTES = 2 * TC_ATOM_BYTES + TC_EXT_BYTES;   // 36 KB
static constexpr int TC_SMEM_A = TC_MT * TC_TILE_BYTES;           // 72 KB
static constexpr int TC_SMEM_B = TC_STAGES * TC_TILE_BYTES;       // 144 KB
static constexpr int TC_XCHG_BYTES = 8 * 32 * 16;               // top-2 hand-over between column-split warps
static constexpr int TC_SMEM_BYTES = TC_SMEM_A + TC_SMEM_B + 1024 /*align slack*/ + 256 /*barriers*/ + TC_XCHG_BYTES;
// EPI 0..2: 8 epilogue warps (one per m-tile x lane quarter); EPI 3: 16 epilogue warps, the two
// warps of a (m-tile, quarter) split the 128 columns of a tile in halves.
__host__ __device__ constexpr int tc_threads(int epi) { return epi >= 3 ? 640 : 384; }
// EPI 5: like EPI 3 (16 epilogue warps) but the epilogue only tracks values + winning chunk; exact
// indices are recovered by l2_fixup.cu.
static constexpr uint32_t TC_TMEM_COLS = 512;

// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, N=128, M=128.
static constexpr uint32_t TC_IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((TC_BN >> 3) << 17) |
                                     ((TC_BM >> 4) << 24);

// Shared-memory matrix descriptor, K-major, swizzled.  addr/SBO in bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);           // start address
  d |= static_cast<uint64_t>(1) << 16;                            // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;          // stride between 8-row groups
  d |= static_cast<uint64_t>(1) << 46;                            // descriptor version (sm_100)
  d |= static_cast<uint64_t>(layout) << 61;                       // 2 = SWIZZLE_128B, 6 = SWIZZLE_32B
  return d;
}

__device__ __forceinline__ void wait_bounded(uint64_t* bar, uint32_t parity) {
  // A protocol bug must fault, not hang the GPU -- but only a lost arrival may trap: the bound is 20 s of %globaltimer
  // (read every 4096 polls), orders of magnitude above what a profiler, a sanitizer or co-resident kernels add to a wait.
  unsigned long long t0 = 0;
  for (uint32_t spin = 1; !mbar_try_wait(bar, parity); ++spin) {
    if ((spin & 0xFFFu) == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 20ull * 1000000000ull) __trap();
    }
  }
}

struct Top2 {
  float m1, m2;
  int i1, i2;
};

__device__ __forceinline__ void top2_push(Top2& s, float v, int col) {
  if (v < s.m2) {
    if (v < s.m1) { s.m2 = s.m1; s.i2 = s.i1; s.m1 = v; s.i1 = col; }
    else { s.m2 = v; s.i2 = col; }
  }
}

// One group of 4 columns whose minimum g is known to be below s.m2.  Common case: exactly one
// element of the group enters the top-2 -- locate the first element equal to g (lowest column on
// ties), insert it with two min/max and two selects, and check with the group's second-smallest
// value that nothing else qualifies.  Otherwise (rare) redo the group element by element.
__device__ __forceinline__ void group4_insert(Top2& s, float v0, float v1, float v2, float v3, float g,
                                              int col) {
  int pos = 3;
  pos = v2 == g ? 2 : pos;
  pos = v1 == g ? 1 : pos;
  pos = v0 == g ? 0 : pos;
  const float second = fminf(fmaxf(fminf(v0, v1), fminf(v2, v3)),
                             fminf(fmaxf(v0, v1), fmaxf(v2, v3)));
  const float nm2 = fmaxf(g, s.m1);                 // g < s.m2  =>  new second = max(g, old first)
  if (second < nm2) {                               // another element also enters: generic path
    top2_push(s, v0, col); top2_push(s, v1, col + 1); top2_push(s, v2, col + 2); top2_push(s, v3, col + 3);
  } else {
    const bool first = g < s.m1;
    const int c = col + pos;
    s.i2 = first ? s.i1 : c;
    s.i1 = first ? c : s.i1;
    s.m2 = nm2;
    s.m1 = fminf(g, s.m1);
  }
}

// Rare generic path of a group (two or more elements enter the top-2); kept out of line so the hot
// code stays small in the instruction cache.
__device__ __noinline__ void group4_generic(Top2* sp, float v0, float v1, float v2, float v3, int col) {
  Top2 s = *sp;
  top2_push(s, v0, col); top2_push(s, v1, col + 1); top2_push(s, v2, col + 2); top2_push(s, v3, col + 3);
  *sp = s;
}

// Processes 32 consecutive columns held in registers (r must resolve to registers: call sites are
// fully unrolled with compile-time offsets).
//   MODE 0: nested tests, element-wise inserts (baseline)
//   MODE 1: nested tests, single-insert group body
//   MODE 2: per-lane hit mask over the 8 groups, then one shared group body per set bit -- replaces
//           the chain of 8 dependent test-and-branch steps by independent compares plus a short loop
template <int MODE = 0>
__device__ __forceinline__ void scan32(Top2& s, const uint32_t* r, int col0) {
  float g[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    g[k] = fminf(fminf(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1])),
                 fminf(__uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3])));
  const float cm = fminf(fminf(fminf(g[0], g[1]), fminf(g[2], g[3])),
                         fminf(fminf(g[4], g[5]), fminf(g[6], g[7])));
  if (cm < s.m2) {
    if constexpr (MODE == 2) {
      uint32_t mask = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) mask |= (g[k] < s.m2 ? 1u : 0u) << k;
      do {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        float v0, v1, v2, v3, gk;
        switch (k) {
#define PM_CASE(K)                                                                               \
  case K:                                                                                        \
    v0 = __uint_as_float(r[4 * K]); v1 = __uint_as_float(r[4 * K + 1]);                           \
    v2 = __uint_as_float(r[4 * K + 2]); v3 = __uint_as_float(r[4 * K + 3]); gk = g[K];            \
    break;
          PM_CASE(0) PM_CASE(1) PM_CASE(2) PM_CASE(3) PM_CASE(4) PM_CASE(5) PM_CASE(6)
          default:
            v0 = __uint_as_float(r[28]); v1 = __uint_as_float(r[29]);
            v2 = __uint_as_float(r[30]); v3 = __uint_as_float(r[31]); gk = g[7];
            break;
#undef PM_CASE
        }
        if (gk < s.m2) {                      // an earlier insert of this loop may have tightened m2
          const int col = col0 + 4 * k;
          int pos = 3;
          pos = v2 == gk ? 2 : pos;
          pos = v1 == gk ? 1 : pos;
          pos = v0 == gk ? 0 : pos;
          const float second = fminf(fmaxf(fminf(v0, v1), fminf(v2, v3)),
                                     fminf(fmaxf(v0, v1), fmaxf(v2, v3)));
          const float nm2 = fmaxf(gk, s.m1);
          if (second < nm2) {
            group4_generic(&s, v0, v1, v2, v3, col);
          } else {
            const bool first = gk < s.m1;
            const int c = col + pos;
            s.i2 = first ? s.i1 : c;
            s.i1 = first ? c : s.i1;
            s.m2 = nm2;
            s.m1 = fminf(gk, s.m1);
          }
        }
      } while (mask);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (g[k] < s.m2) {
          if constexpr (MODE == 1) {
            group4_insert(s, __uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]),
                          __uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]), g[k], col0 + 4 * k);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) top2_push(s, __uint_as_float(r[4 * k + e]), col0 + 4 * k + e);
          }
        }
      }
    }
  }
}

// Minimum of 16 consecutive columns (branch-free; 3-input min where the compiler finds it).
__device__ __forceinline__ float min16(const uint32_t* r) {
  float g[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    g[k] = fminf(fminf(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1])),
                 fminf(__uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3])));
  return fminf(fminf(g[0], g[1]), fminf(g[2], g[3]));
}
// Values-only running state of the fast epilogue (EPI 5): smallest value, second smallest chunk
// minimum, base column of the first 16-column chunk that attained the smallest value.  No branches.
struct Fast2 {
  float m1, m2;
  int b1;
};
__device__ __forceinline__ void fast_update(Fast2& s, float cm, int cbase) {
  const float t = fmaxf(cm, s.m1);
  s.b1 = cm < s.m1 ? cbase : s.b1;                   // strict: the earliest chunk keeps a tie
  s.m1 = fminf(cm, s.m1);
  s.m2 = fminf(s.m2, t);
}

template <int EPI, int SCAN = (EPI == 3 ? 2 : 0)>
__global__ void __launch_bounds__(tc_threads(EPI), 1)
l2_top2_tc_kernel(const __grid_constant__ CUtensorMap q_main, const __grid_constant__ CUtensorMap q_ext,
                  const __grid_constant__ CUtensorMap t_main, const __grid_constant__ CUtensorMap t_ext,
                  const int32_t* __restrict__ qnorm, const PairJob* __restrict__ jobs, int n_jobs,
                  int tiles_per_job, int2* __restrict__ knn_idx, float2* __restrict__ knn_dist,
                  int stride, float* __restrict__ debug_dump) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + TC_SMEM_A;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_SMEM_A + TC_SMEM_B);
  uint64_t* a_full = bars + 0;
  uint64_t* a_empty = bars + 1;
  uint64_t* b_full = bars + 2;                      // [TC_STAGES]
  uint64_t* b_empty = bars + 2 + TC_STAGES;         // [TC_STAGES]
  uint64_t* acc_full = bars + 2 + 2 * TC_STAGES;    // [2 stages][2 m-tiles]
  uint64_t* acc_empty = acc_full + 4;               // [2][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);
  float4* xchg = reinterpret_cast<float4*>(smem + TC_SMEM_A + TC_SMEM_B + 256);
  constexpr uint32_t kEpiArrivals = EPI >= 3 ? 8 : 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&q_main); tma_prefetch_desc(&q_ext);
    tma_prefetch_desc(&t_main); tma_prefetch_desc(&t_ext);
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kEpiArrivals); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TC_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = n_jobs * tiles_per_job;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      uint32_t ai = 0, bi = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int jb = item / tiles_per_job, r = item - jb * tiles_per_job;
        const PairJob job = jobs[jb];
        if (r * TC_ROWS >= job.nq) continue;
        wait_bounded(a_empty, (ai & 1) ^ 1);
        mbar_expect_tx(a_full, TC_SMEM_A);
#pragma unroll
        for (int m = 0; m < TC_MT; ++m) {
          uint8_t* dst = sA + m * TC_TILE_BYTES;
          const int row = job.q_row + r * TC_ROWS + m * TC_BM;
          tma_load_2d(dst, &q_main, 0, row, a_full);
          tma_load_2d(dst + TC_ATOM_BYTES, &q_main, 64, row, a_full);
          tma_load_2d(dst + 2 * TC_ATOM_BYTES, &q_ext, TC_DIM, row, a_full);
        }
        ++ai;
        const int n_tiles = (job.nt + TC_BN - 1) / TC_BN;
        for (int n = 0; n < n_tiles; ++n, ++bi) {
          const uint32_t st = bi % TC_STAGES;
          wait_bounded(&b_empty[st], ((bi / TC_STAGES) & 1) ^ 1);
          mbar_expect_tx(&b_full[st], TC_TILE_BYTES);
          uint8_t* dst = sB + st * TC_TILE_BYTES;
          const int row = job.t_row + n * TC_BN;
          tma_load_2d(dst, &t_main, 0, row, &b_full[st]);
          tma_load_2d(dst + TC_ATOM_BYTES, &t_main, 64, row, &b_full[st]);
          tma_load_2d(dst + 2 * TC_ATOM_BYTES, &t_ext, TC_DIM, row, &b_full[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    // The whole warp walks the loop (all values stay warp-uniform -> uniform registers, no
    // R2UR round trips); one elected lane issues the MMAs and commits.
    uint32_t ai = 0, bi = 0, ti = 0;
    constexpr uint32_t HI128 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO | version | SWIZZLE_128B
    constexpr uint32_t HI32 = (256u >> 4) | (1u << 14) | (6u << 29);     // SBO | version | SWIZZLE_32B
    const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int jb = item / tiles_per_job, r = item - jb * tiles_per_job;
      const int job_nq = jobs[jb].nq, job_nt = jobs[jb].nt;
      if (r * TC_ROWS >= job_nq) continue;
      wait_bounded(a_full, ai & 1);
      ++ai;
      const int n_tiles = (job_nt + TC_BN - 1) / TC_BN;
      for (int n = 0; n < n_tiles; ++n, ++bi, ++ti) {
        const uint32_t st = bi % TC_STAGES;
        const uint32_t as = ti & 1, use = ti >> 1;
        wait_bounded(&b_full[st], (bi / TC_STAGES) & 1);
        const uint32_t b_lo = b_lo0 + st * (TC_TILE_BYTES >> 4);
#pragma unroll
        for (int m = 0; m < TC_MT; ++m) {
          wait_bounded(&acc_empty[as * 2 + m], (use & 1) ^ 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + m * (TC_TILE_BYTES >> 4);
            const uint32_t d_tmem = tmem_base + (as * 2 + m) * TC_BN;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t off = ((k >> 2) * TC_ATOM_BYTES + (k & 3) * 32) >> 4;
              umma_f16(d_tmem, (static_cast<uint64_t>(HI128) << 32) | (a_lo + off),
                       (static_cast<uint64_t>(HI128) << 32) | (b_lo + off), TC_IDESC, k > 0 ? 1u : 0u);
            }
            constexpr uint32_t xoff = (2 * TC_ATOM_BYTES) >> 4;
            umma_f16(d_tmem, (static_cast<uint64_t>(HI32) << 32) | (a_lo + xoff),
                     (static_cast<uint64_t>(HI32) << 32) | (b_lo + xoff), TC_IDESC, 1u);
            umma_commit(&acc_full[as * 2 + m]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&b_empty[st]);
        __syncwarp();
      }
      if (elect_one()) umma_commit(a_empty);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ======================================= epilogue =======================================
    const int m = ((warp - 4) >> 2) & 1, quarter = warp & 3;
    const int half = (warp - 4) >> 3;                 // EPI 3 only: which 64-column half of a tile
    uint32_t ti = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int jb = item / tiles_per_job, r = item - jb * tiles_per_job;
      const PairJob job = jobs[jb];
      if (r * TC_ROWS >= job.nq) continue;
      const int row = r * TC_ROWS + m * TC_BM + quarter * 32 + lane;
      Top2 s;
      s.m1 = s.m2 = __int_as_float(0x7f800000);
      s.i1 = s.i2 = -1;
      Fast2 fs;
      fs.m1 = fs.m2 = __int_as_float(0x7f800000);
      fs.b1 = -1;
      const int n_tiles = (job.nt + TC_BN - 1) / TC_BN;
      for (int n = 0; n < n_tiles; ++n, ++ti) {
        const uint32_t as = ti & 1, use = ti >> 1;
        // (the producer and MMA warps keep bounded waits: a protocol bug still traps there)
        if constexpr (EPI >= 3) mbar_wait_hint(&acc_full[as * 2 + m], use & 1);
        else wait_bounded(&acc_full[as * 2 + m], use & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               (as * 2 + m) * TC_BN;
        const int valid = job.nt - n * TC_BN;          // columns of this tile that exist
        const int col_base = n * TC_BN;
        auto release = [&]() {                         // accumulator drained -> hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[as * 2 + m]);
        };
        if constexpr (EPI == 5) {
          const uint32_t t0 = taddr + half * 64;
          const int c0 = col_base + half * 64;
          uint32_t v[32];
          if (valid >= TC_BN) {
            tmem_ld_32x32b_x32(t0, v);
            fast_update(fs, min16(v), c0);
            fast_update(fs, min16(v + 16), c0 + 16);
            tmem_ld_32x32b_x32(t0 + 32, v);
            tc_fence_before();
            if (lane == 0) mbar_arrive(&acc_empty[as * 2 + m]);
            fast_update(fs, min16(v), c0 + 32);
            fast_update(fs, min16(v + 16), c0 + 48);
          } else {
            const int lim = valid - half * 64;
            tmem_ld_32x32b_x32(t0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e >= lim) v[e] = 0x7f800000u;
            if (lim > 0) fast_update(fs, min16(v), c0);
            if (lim > 16) fast_update(fs, min16(v + 16), c0 + 16);
            tmem_ld_32x32b_x32(t0 + 32, v);
            tc_fence_before();
            if (lane == 0) mbar_arrive(&acc_empty[as * 2 + m]);
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (32 + e >= lim) v[e] = 0x7f800000u;
            if (lim > 32) fast_update(fs, min16(v), c0 + 32);
            if (lim > 48) fast_update(fs, min16(v + 16), c0 + 48);
          }
        } else if constexpr (EPI == 4) {
          // PROBE ONLY (results are garbage): hand the accumulator straight back, to time the
          // TMA + MMA pipeline without any epilogue work.
          tc_fence_before();
          if (lane == 0) mbar_arrive(&acc_empty[as * 2 + m]);
        } else if constexpr (EPI == 3) {
          // lean path: this warp owns 64 of the tile's 128 columns, two 32-column loads; the
          // masked variant only runs for a ragged last tile (warp-uniform branch).
          const uint32_t t0 = taddr + half * 64;
          const int c0 = col_base + half * 64;
          uint32_t v[32];
          if (valid >= TC_BN) {
            tmem_ld_32x32b_x32(t0, v);
            scan32<SCAN>(s, v, c0);
            tmem_ld_32x32b_x32(t0 + 32, v);
            tc_fence_before();
            if (lane == 0) mbar_arrive(&acc_empty[as * 2 + m]);
            scan32<SCAN>(s, v, c0 + 32);
          } else {
            const int lim = valid - half * 64;           // columns of this half that exist
            tmem_ld_32x32b_x32(t0, v);
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e >= lim) v[e] = 0x7f800000u;
            scan32(s, v, c0);
            tmem_ld_32x32b_x32(t0 + 32, v);
            tc_fence_before();
            if (lane == 0) mbar_arrive(&acc_empty[as * 2 + m]);
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (32 + e >= lim) v[e] = 0x7f800000u;
            scan32(s, v, c0 + 32);
          }
        } else if constexpr (EPI == 0) {
#pragma unroll
          for (int c = 0; c < TC_BN / 32; ++c) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld_32x32b_x32(taddr + c * 32, v);      // includes tcgen05.wait::ld
            if (c == TC_BN / 32 - 1) release();
            if (debug_dump != nullptr && item == 0 && n == 0) {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                debug_dump[static_cast<size_t>(m * TC_BM + quarter * 32 + lane) * TC_BN + c * 32 + e] =
                    __uint_as_float(v[e]);
            }
            if (valid < (c + 1) * 32) {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (c * 32 + e >= valid) v[e] = 0x7f800000u;
            }
            scan32(s, v, col_base + c * 32);
          }
        } else {
          uint32_t va[64], vb[64];
          __syncwarp();
          tmem_ld_32x32b_x64_async(taddr, va);
          if constexpr (EPI == 2) {
            tmem_ld_32x32b_x64_async(taddr + 64, vb);   // both halves in flight
            tmem_wait_pin(va);
            tmem_wait_pin(vb);
            release();
          } else {
            tmem_wait_pin(va);
          }
          if (valid < 64) {
#pragma unroll
            for (int e = 0; e < 64; ++e)
              if (e >= valid) va[e] = 0x7f800000u;
          }
          if constexpr (EPI == 1) tmem_ld_32x32b_x64_async(taddr + 64, vb);   // lands while va is scanned
          scan32(s, &va[0], col_base);
          scan32(s, &va[32], col_base + 32);
          if constexpr (EPI == 1) {
            tmem_wait_pin(vb);
            release();
          }
          if (valid < 128) {
#pragma unroll
            for (int e = 0; e < 64; ++e)
              if (64 + e >= valid) vb[e] = 0x7f800000u;
          }
          scan32(s, &vb[0], col_base + 64);
          scan32(s, &vb[32], col_base + 96);
        }
      }
      if constexpr (EPI == 5) {
        float4* slot = xchg + ((m * 4 + quarter) * 32 + lane);
        const int bar_id = 1 + m * 4 + quarter;
        if (half == 1) *slot = make_float4(fs.m1, __int_as_float(fs.b1), fs.m2, 0.f);
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        if (half == 0) {
          const float4 o = *slot;
          const int ob = __float_as_int(o.y);
          const bool take = ob >= 0 && (fs.b1 < 0 || o.x < fs.m1 || (o.x == fs.m1 && ob < fs.b1));
          const float hi = fmaxf(fs.m1, o.x);
          fs.m2 = fminf(fminf(fs.m2, o.z), hi);
          fs.m1 = fminf(fs.m1, o.x);
          fs.b1 = take ? ob : fs.b1;
          if (row < job.nq) {
            const float na = static_cast<float>(qnorm[job.q_row + row]);
            const size_t o2 = static_cast<size_t>(jb) * stride + row;
            knn_idx[o2] = make_int2(fs.b1, -2);                       // -2: "values only, run the fix-up"
            knn_dist[o2] = make_float2(__fadd_rn(fs.m1, na), __fadd_rn(fs.m2, na));   // exact d^2 / bound
          }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      } else if constexpr (EPI >= 3) {
        // the two column-split warps of this (m-tile, quarter) merge their top-2 through smem
        float4* slot = xchg + ((m * 4 + quarter) * 32 + lane);
        const int bar_id = 1 + m * 4 + quarter;
        if (half == 1) *slot = make_float4(s.m1, __int_as_float(s.i1), s.m2, __int_as_float(s.i2));
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        if (half == 0) {
          const float4 o = *slot;
          const float om1 = o.x, om2 = o.z;
          const int oi1 = __float_as_int(o.y), oi2 = __float_as_int(o.w);
          // ordered by (value, index); each list is already sorted and indices are distinct
          auto less = [](float va, int ia, float vb, int ib) { return va < vb || (va == vb && ia < ib); };
          Top2 t;
          if (oi1 >= 0 && (s.i1 < 0 || less(om1, oi1, s.m1, s.i1))) {
            t.m1 = om1; t.i1 = oi1;
            if (oi2 >= 0 && (s.i1 < 0 || less(om2, oi2, s.m1, s.i1))) { t.m2 = om2; t.i2 = oi2; }
            else { t.m2 = s.m1; t.i2 = s.i1; }
          } else {
            t.m1 = s.m1; t.i1 = s.i1;
            if (oi1 >= 0 && (s.i2 < 0 || less(om1, oi1, s.m2, s.i2))) { t.m2 = om1; t.i2 = oi1; }
            else { t.m2 = s.m2; t.i2 = s.i2; }
          }
          s = t;
        }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      }
      if (EPI != 5 && (EPI < 3 || half == 0) && row < job.nq) {
        const float na = static_cast<float>(qnorm[job.q_row + row]);
        int2 oi;
        float2 od;
        oi.x = s.i1; oi.y = s.i2;
        od.x = s.i1 < 0 ? s.m1 : __fsqrt_rn(fmaxf(__fadd_rn(s.m1, na), 0.f));
        od.y = s.i2 < 0 ? s.m2 : __fsqrt_rn(fmaxf(__fadd_rn(s.m2, na), 0.f));
        const size_t o = static_cast<size_t>(jb) * stride + row;
        knn_idx[o] = oi;
        knn_dist[o] = od;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

cudaError_t tc_configure() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(l2_top2_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
}

cudaError_t launch_l2_tc(const TcMaps& maps, const int32_t* qnorm, const PairJob* jobs, int n_jobs,
                         int max_nq, int2* idx, float2* dist, int stride, int num_sms,
                         float* debug_dump, int epi, cudaStream_t st) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  const int tiles_per_job = (max_nq + TC_ROWS - 1) / TC_ROWS;
  const int n_items = n_jobs * tiles_per_job;
  const int grid = n_items < num_sms ? n_items : num_sms;
#define PM_TC_LAUNCH(E)                                                                         \
  l2_top2_tc_kernel<E><<<grid, tc_threads(E), TC_SMEM_BYTES, st>>>(                             \
      maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, qnorm, jobs, n_jobs, tiles_per_job, idx, \
      dist, stride, debug_dump)
  if (debug_dump != nullptr || epi == 0) PM_TC_LAUNCH(0);
  else if (epi == 1) PM_TC_LAUNCH(1);
  else if (epi == 2) PM_TC_LAUNCH(2);
  else if (epi == 3) PM_TC_LAUNCH(3);
  else if (epi == 4) PM_TC_LAUNCH(4);
  else PM_TC_LAUNCH(5);
#undef PM_TC_LAUNCH
  return cudaGetLastError();
}

}  // namespace pm
