// l2f_fixup.cu -- exact fp32 re-rank behind the real-valued tensor-core search (l2_tc2.cu MODE 3).
//
// The tcgen05 kernel scores every (query, train) pair of a real-valued descriptor set (SuperPoint:
// 256 floats of unit norm, FeatureSuperPoint.cpp:195-205) with fp16 operands:
//     score(a, b) ~= |b|^2 - 2 a.b + 2         (= d^2 - |a|^2 + 2, always > 0 for |a|,|b| <= ~1)
// and keeps, per query row, the SIX smallest 16-column chunk minima as keys (score with the chunk id in
// the low 10 mantissa bits).  This kernel turns that into the exact answer of the brute-force search the
// reference's knnMatch stands for (FeatureMatcher.cpp:48-49; exact arbiter = sum_k (a_k - b_k)^2 in fp32,
// k ascending, fused multiply-add -- bit-identical to l2_top2_simt_kernel):
//   * every column of a candidate chunk is re-evaluated exactly;
//   * a rigorous bound eps on |key - exact score| (fp16 operand rounding 2^-9 |a||b|, tensor-core
//     accumulation, key truncation 2^-13, fp32 evaluation error) gives a LOWER bound on the exact d^2 of
//     every column that was not re-evaluated: lb(K_next) = K_next - eps + |a|^2 - 2;
//   * chunks are evaluated in key order until the requested facts are certain:
//       L2F_NEED_RATIO    nearest index + outcome of Lowe's test d1 < ratio * d2 (FeatureMatcher.cpp:55)
//       L2F_NEED_NEAREST  nearest index (cross-check direction)
//       L2F_NEED_FULL     (idx1, idx2, d1, d2) as cv::DMatch rows
//   * a row that six chunks cannot certify is scanned exhaustively by the whole block (exact, rare).
// Ties resolve to the lowest train index (cv::BFMatcher's rule): certification is strict (<), so an
// unevaluated column can never tie with a reported one.
#include "common.cuh"
#include "kernels.h"

namespace pm {

static constexpr int FF_THREADS = 128;            // 8 half-warps
static constexpr int FF_HW = FF_THREADS / 16;
static constexpr int FF_SPAN = 128;               // query rows per block
// block size of the s8-prefilter variant.  128 threads x 64 registers = 8 K: three blocks fit next to the SuperPoint tensor
// kernel.  256 threads (16 K, the footprint of the selection and RANSAC blocks, so that no block size is favoured when a
// hole opens) measured the same: 61.7 k vs 63.4 k pairs/s.
#ifndef PM_FFP_THREADS
#define PM_FFP_THREADS 128
#endif
static constexpr int FFP_THREADS = PM_FFP_THREADS;
static constexpr int FF_MAXDIM = 256;
static constexpr unsigned int FF_IDMASK = 0x3FFu;

struct FfBound {          // per query row: exact_d2(col) is within [K - eps(K) + c, K + eps(K) + c]
  double c, e0;
  __device__ __forceinline__ double eps(float K) const { return 1.02 * (e0 + 1.2207031e-4 * K) + 1e-7; }
  __device__ __forceinline__ float lb(float K) const {            // rounded down, >= 0; +inf stays +inf
    if (K == __int_as_float(0x7f800000)) return K;
    const double v = static_cast<double>(K) - eps(K) + c;
    return v > 0.0 ? __double2float_rd(v) : 0.f;
  }
  __device__ __forceinline__ float ub(float K) const {
    if (K == __int_as_float(0x7f800000)) return K;
    const double v = static_cast<double>(K) + eps(K) + c;
    return v > 0.0 ? __double2float_ru(v) : 0.f;
  }
};

// e_mode 0: fp16 operand forms (l2_tc2.cu MODE 3, kind::f16); 1: rows quantised to s8 with scale 254 (KIND 3);
// 2: as 1, train norms taken as exactly 1 (unit-norm images, no norm K-step)
__device__ __forceinline__ FfBound ff_bound(float na2, float nb2max, int dim, int e_mode) {
  FfBound b;
  const double na = sqrt(static_cast<double>(na2)) * (1.0 + 1e-6);
  const double nb = sqrt(static_cast<double>(nb2max)) * (1.0 + 1e-6);
  const double sD = sqrt(static_cast<double>(dim));
  const double u24 = 5.9604644775390625e-8;          // 2^-24
  const double e_op = 1.002 * 1.953125e-3 * na * nb + 1.01 * u24 * sD * (na + nb);      // fp16 operands
  const double e_acc = 1.52587890625e-5 * (nb * nb + 2.0 * na * nb + 2.0);              // 2^-16: accumulation
  const double e_nrm = 3.814697265625e-6 * (na * na + nb * nb);                         // 2^-18: fp32 norms, fp16 split
  const double e_f32 = 1.01 * (dim + 3) * u24 * (na + nb) * (na + nb);                  // exact side is fp32 too
  b.e0 = e_op + e_acc + e_nrm + e_f32;
  if (e_mode >= 1) {
    // q = rint(254 x): |q - 254 x| <= 1/2 per element, so |q_a.q_b - 254^2 a.b| <= 127 (|a|_1 + |b|_1) + D / 4
    // <= 127 sqrt(D) (|a| + |b|) + D / 4; the score is 2 / 254^2 times the integer accumulator (exact), whose
    // norm term was rounded to an integer (1/2) from an fp32 norm (2^-22 relative incl. its accumulation)
    const double S = 254.0;
    const double e_q = (2.0 / (S * S)) * (0.5 * S * sD * (na + nb) + 0.25 * dim + 0.75);
    const double e_n = 2.4e-7 * (dim + 3) * (nb * nb + 2.0) + 4.0 * u24 * (nb * nb + 2.0 * na * nb + 2.0);
    b.e0 = 1.001 * e_q + e_n + e_f32;
    // e_mode 2: the candidate stage scored every train row as if |b|^2 were exactly 1 (l2_i8x2_kernel NX); the rows are
    // within L2S8_UNIT_TOL of that (checked at ingest, fp32 norms: their own error is part of e_n)
    if (e_mode == 2) b.e0 += 1.001 * static_cast<double>(L2S8_UNIT_TOL);
  }
  b.c = static_cast<double>(na2) - 2.0;
  return b;
}

// exact fp32 squared distance in the SIMT kernel's operation order (l2_simt.cu: d = a - b; acc = fma(d, d, acc))
__device__ __forceinline__ float ff_dist(const float* __restrict__ qs, const float* __restrict__ trow, int dim) {
  float acc = 0.f;
#pragma unroll 4
  for (int k = 0; k < dim; k += 4) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(trow + k));
    const float4 a = *reinterpret_cast<const float4*>(qs + k);
    float d = a.x - b.x; acc = fmaf(d, d, acc);
    d = a.y - b.y; acc = fmaf(d, d, acc);
    d = a.z - b.z; acc = fmaf(d, d, acc);
    d = a.w - b.w; acc = fmaf(d, d, acc);
  }
  return acc;
}

// The same chain, abandoned once it reaches thr (checked every 64 dimensions): every partial sum of the chain is a
// lower bound of its final value (fp32 round-to-nearest is monotone and the addends are >= 0), so a column whose
// partial sum reached thr is >= thr for certain.  Returns false (and no value) in that case.
__device__ __forceinline__ bool ff_dist_below(const float* __restrict__ qs, const float* __restrict__ trow, int dim,
                                              float thr, float* out) {
  float acc = 0.f;
  for (int k0 = 0; k0 < dim; k0 += 64) {
#pragma unroll 4
    for (int k = k0; k < k0 + 64; k += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(trow + k));
      const float4 a = *reinterpret_cast<const float4*>(qs + k);
      float d = a.x - b.x; acc = fmaf(d, d, acc);
      d = a.y - b.y; acc = fmaf(d, d, acc);
      d = a.z - b.z; acc = fmaf(d, d, acc);
      d = a.w - b.w; acc = fmaf(d, d, acc);
    }
    if (acc >= thr) return false;
  }
  *out = acc;
  return true;
}

// One 16-column chunk by a half-warp, lane l = column cb + l, same fp32 chain per column as ff_dist -- but the rows
// travel through shared memory: the half-warp fetches each row as contiguous 256-byte pieces (one-row-per-lane loads
// touch 32 cache lines per warp instruction and had the kernel bound by L1 request throughput) and each lane then
// runs its own chain out of the staged tile.  The kernel is LATENCY bound (one row per half-warp at a time, every
// fetch a round trip to L2), so the tile is fetched in few, long segments: 64 dimensions of every column first,
// then 64 more of the columns still alive, then -- once at most 8 columns are left -- 128 dimensions at a time.
// With THR a column stops at the end of the segment in which its partial sum reaches thr (see ff_dist_below) and its
// row is no longer fetched: a typical first chunk (one true match among 15 other columns) costs three round trips
// instead of the eight of a fixed 32-dimension pipeline.  Returns true when the lane holds a finished chain (*out);
// false for abandoned / absent columns.  q_pending: the query row is still on its way into qs (cp.async by this
// half-warp): it lands with the first segment.
static constexpr int FF_TILE = 16 * 68;           // floats per half-warp: 16 rows x (64 + 4) or 8 rows x (128 + 4)
__device__ __forceinline__ void ff_cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void ff_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ff_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool THR>
__device__ __forceinline__ bool ff_chunk_rows(float* __restrict__ ts, const float* __restrict__ qs,
                                              const float* __restrict__ tbase, int cb, int ncol, int dim, int l,
                                              unsigned hmask, float thr, float* out) {
  float acc = 0.f;
  bool alive = l < ncol;
  const int shift = (hmask & 1u) ? 0 : 16;
  const float* src = tbase + static_cast<size_t>(cb) * dim;
  unsigned am = (__ballot_sync(hmask, alive) >> shift) & 0xFFFFu;
  __syncwarp(hmask);                                 // the tile is free (previous chunk consumed)
  int k0 = 0;
  while (k0 < dim && am) {
    const bool wide = __popc(am) <= 8 && dim - k0 >= 128;
    const int len = wide ? 128 : 64, tld = len + 4;
    for (unsigned m = am; m; m &= m - 1) {
      const int r = __ffs(m) - 1;
      const int slot = wide ? __popc(am & ((1u << r) - 1u)) : r;
      const float* g = src + static_cast<size_t>(r) * dim + k0;
      float* d = ts + slot * tld;
      ff_cp_async16(d + 4 * l, g + 4 * l);
      if (wide) ff_cp_async16(d + 64 + 4 * l, g + 64 + 4 * l);
    }
    ff_cp_commit();
    ff_cp_wait<0>();                                 // (also the query row of this half-warp, if it was pending)
    __syncwarp(hmask);
    if (alive) {
      const int slot = wide ? __popc(am & ((1u << l) - 1u)) : l;
      const float* tl = ts + slot * tld;
#pragma unroll 4
      for (int k = 0; k < len; k += 4) {
        const float4 b = *reinterpret_cast<const float4*>(tl + k);
        const float4 a = *reinterpret_cast<const float4*>(qs + k0 + k);
        float d = a.x - b.x; acc = fmaf(d, d, acc);
        d = a.y - b.y; acc = fmaf(d, d, acc);
        d = a.z - b.z; acc = fmaf(d, d, acc);
        d = a.w - b.w; acc = fmaf(d, d, acc);
      }
      if (THR && acc >= thr) alive = false;
    }
    k0 += len;
    am = (__ballot_sync(hmask, alive) >> shift) & 0xFFFFu;      // also: everyone is done with the tile
  }
  ff_cp_wait<0>();                                   // the query row, when no column existed at all
  __syncwarp(hmask);
  *out = acc;
  return alive;
}

__device__ __forceinline__ void ff_merge(unsigned long long& m1, unsigned long long& m2, unsigned long long o1,
                                         unsigned long long o2) {
  const unsigned long long lo = m1 < o1 ? m1 : o1;
  const unsigned long long hi = m1 < o1 ? o1 : m1;
  const unsigned long long s = m2 < o2 ? m2 : o2;
  m1 = lo;
  m2 = hi < s ? hi : s;
}

__device__ __forceinline__ void ff_write_exact(int2* knn_idx, float2* knn_dist, size_t o, unsigned long long e1,
                                               unsigned long long e2) {
  const float inf = __int_as_float(0x7f800000);
  int2 oi;
  float2 od;
  oi.x = e1 == KEY_NONE64 ? -1 : static_cast<int>(e1 & 0xFFFFFFFFull);
  oi.y = e2 == KEY_NONE64 ? -1 : static_cast<int>(e2 & 0xFFFFFFFFull);
  od.x = e1 == KEY_NONE64 ? inf : __fsqrt_rn(__uint_as_float(static_cast<unsigned int>(e1 >> 32)));
  od.y = e2 == KEY_NONE64 ? inf : __fsqrt_rn(__uint_as_float(static_cast<unsigned int>(e2 >> 32)));
  knn_idx[o] = oi;
  knn_dist[o] = od;
}

// PRE (rows quantised to s8, ratio-test rows, e_mode 1): before any fp32 work the columns of the first chunk are
// screened with the QUANTISED rows the tensor kernel itself multiplied (sq8 / st8, 256 bytes per row instead of 1 KB):
// lane l recomputes the integer dot product of its column with dp4a, which gives score'(c) = |b|^2 - 2 q_a.q_b / 254^2
// + 2 within the same certified eps as the keys, hence a lower bound lb(score') of the exact d^2 of that column.  A
// column with lb >= thr is "abandoned" before it starts (exactly the guarantee ff_dist_below gives: d^2 >= thr); the
// one or two columns that survive run their fp32 chains straight from global memory.  No staged tiles: 14.5 KB of
// shared memory and <= 64 registers, so that THREE blocks fit next to the tensor kernel (the staged variant: one).
// Opt-in (see launch_l2f_fixup): it did not change the step and is slower alone.
template <bool PRE>
__global__ void __launch_bounds__(PRE ? FFP_THREADS : FF_THREADS, PRE ? (1024 / FFP_THREADS) : 5)
l2f_fixup_kernel(const float* __restrict__ raw, const float* __restrict__ fnorm, int dim,
                 const PairJob* __restrict__ jobs, int2* __restrict__ knn_idx, float2* __restrict__ knn_dist,
                 const float2* __restrict__ extra, int stride, float ratio, int need,
                 unsigned long long* __restrict__ counters, int e_mode, const uint8_t* __restrict__ q8,
                 const uint8_t* __restrict__ t8, const uint8_t* __restrict__ flags, int nkeys) {
  // flags != nullptr: only the rows l2f_rerank1_kernel left open (flag 1) are processed, all others hold final results;
  // nkeys: how many of the six key slots the candidate stage filled (3 behind l2f_rerank1_kernel)
  constexpr int THREADS = PRE ? FFP_THREADS : FF_THREADS, HW = THREADS / 16, SPAN = THREADS;
  __shared__ __align__(16) float qs[HW][FF_MAXDIM];
  __shared__ __align__(16) uint4 qs8[PRE ? HW : 1][FF_MAXDIM / 16];       // PRE: the query rows as s8
  extern __shared__ __align__(16) float ts_dyn[];       // !PRE: one FF_TILE per half-warp (staged tiles)
  __shared__ float keys_s[SPAN][6];
  __shared__ int list[SPAN];
  __shared__ int ovf[SPAN];
  __shared__ unsigned long long red[HW][2];
  __shared__ int cnt, n_ovf;

  const PairJob jb = jobs[blockIdx.y];
  const int span0 = blockIdx.x * SPAN;
  if (span0 >= jb.nq) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int hw = tid >> 4, l = lane & 15;
  const unsigned hmask = (lane & 16) ? 0xffff0000u : 0x0000ffffu;
  const size_t base = static_cast<size_t>(blockIdx.y) * stride;
  const float inf = __int_as_float(0x7f800000);
  const float* tbase = raw + static_cast<size_t>(jb.t_row) * dim;

  // ---- pass 1: one thread per row; close the rows whose outcome the keys already decide -----------
  if (tid == 0) { cnt = 0; n_ovf = 0; }
  __syncthreads();
  {
    const int row = span0 + tid;
    if (row < jb.nq && (flags == nullptr || flags[base + row] != 0)) {
      const float2 k12 = knn_dist[base + row];
      const int2 k34 = knn_idx[base + row];
      const float2 k56 = extra[base + row];
      const float K[6] = {k12.x, k12.y, __int_as_float(k34.x), __int_as_float(k34.y), k56.x, k56.y};
      bool need_row = K[0] != inf;                                // no train rows at all: no neighbours
      if (need_row && need == L2F_NEED_RATIO) {
        const FfBound bd = ff_bound(fnorm[jb.q_row + row], jb.t_maxn, dim, e_mode);
        // true d1^2 >= lb(K1), true d2^2 <= ub(K2): the test fails for good when even these cannot pass
        need_row = !(__fsqrt_rn(bd.lb(K[0])) >= __fmul_rn(ratio, __fsqrt_rn(bd.ub(K[1]))));
      }
      if (need_row) {
        const int e = atomicAdd(&cnt, 1);
        list[e] = tid;
#pragma unroll
        for (int j = 0; j < 6; ++j) keys_s[tid][j] = K[j];
      } else {
        knn_idx[base + row] = make_int2(-1, -1);
        knn_dist[base + row] = make_float2(inf, inf);
      }
    }
  }
  __syncthreads();
  const int n_need = cnt;
  unsigned int n_chunks = 0;
  float worst = 0.f;

  // ---- pass 2: one half-warp per surviving row, lane = column of the chunk under evaluation ---------
  for (int e = hw; e < n_need; e += HW) {
    const int r = list[e];
    const int row = span0 + r;
    const float* qrow = raw + (static_cast<size_t>(jb.q_row) + row) * dim;
    __syncwarp(hmask);                                // the previous row's chains are done with qs[hw]
    const int kp8 = dim + 32;
    if (PRE) {
      for (int k = 4 * l; k < dim; k += 64)
        *reinterpret_cast<float4*>(&qs[hw][k]) = __ldg(reinterpret_cast<const float4*>(qrow + k));
      if (16 * l < dim)
        qs8[hw][l] = __ldg(reinterpret_cast<const uint4*>(q8 + (static_cast<size_t>(jb.q_row) + row) * kp8) + l);
      __syncwarp(hmask);
    } else {
      for (int k = 4 * l; k < dim; k += 64) ff_cp_async16(&qs[hw][k], qrow + k);      // lands with the first segment
    }
    const FfBound bd = ff_bound(fnorm[jb.q_row + row], jb.t_maxn, dim, e_mode);
    unsigned long long e1 = KEY_NONE64, e2 = KEY_NONE64;
    bool done = false;
    // first chunk of a ratio-test row: columns are abandoned as soon as their partial sum shows they cannot matter
    // (see thr below).  A row that the first chunk cannot close re-evaluates it in full (rare) and goes on as before.
    bool early = need == L2F_NEED_RATIO;
    for (int j = 0; j < nkeys && !done; ++j) {
      const float Kj = keys_s[r][j];
      if (Kj == inf) break;                              // (unreachable: the previous lbn was +inf)
      const int cb = static_cast<int>(__float_as_uint(Kj) & FF_IDMASK) * 16;
      const int ncol = min(16, jb.nt - cb);
      unsigned long long k1 = KEY_NONE64, k2 = KEY_NONE64;
      const bool use_thr = early && j == 0;
      // abandon threshold of the first chunk: a column is of no further interest once it is >= lbn (the bound of
      // everything outside the chunk) or >= ub(K1) / ratio^2 (then Lowe's test passes against it whatever the
      // nearest distance turns out to be).  The threshold only saves work: every decision below is re-tested with
      // `low`, the bound the abandoned columns actually satisfy.
      float thr = inf;
      if (use_thr) {
        const float r2 = __fmul_rd(ratio, ratio);
        thr = fminf(bd.lb(keys_s[r][1]), r2 > 0.f ? __fmul_ru(__fdiv_ru(bd.ub(Kj), r2), 1.0001f) : inf);
      }
      bool gave_up = false;
      {
        float d2 = 0.f;
        bool have;
        if (PRE) {
          bool alive = l < ncol;
          if (use_thr && alive) {
            const uint4* tq = reinterpret_cast<const uint4*>(t8 + (static_cast<size_t>(jb.t_row) + cb + l) * kp8);
            int dotneg = 0;                                  // q_a . (-q_b): the train form stores -q
            for (int w = 0; w < (dim >> 4); ++w) {
              const uint4 b = __ldg(tq + w);
              const uint4 a = qs8[hw][w];
              dotneg = __dp4a(static_cast<int>(a.x), static_cast<int>(b.x), dotneg);
              dotneg = __dp4a(static_cast<int>(a.y), static_cast<int>(b.y), dotneg);
              dotneg = __dp4a(static_cast<int>(a.z), static_cast<int>(b.z), dotneg);
              dotneg = __dp4a(static_cast<int>(a.w), static_cast<int>(b.w), dotneg);
            }
            const double sc = static_cast<double>(fnorm[jb.t_row + cb + l]) + 2.0 +
                              static_cast<double>(dotneg) * (2.0 / (254.0 * 254.0));
            if (bd.lb(__double2float_rd(sc)) >= thr) alive = false;          // d^2 of this column >= thr for certain
          }
          have = false;
          if (alive) {
            const float* trow = tbase + static_cast<size_t>(cb + l) * dim;
            if (use_thr) {
              have = ff_dist_below(qs[hw], trow, dim, thr, &d2);
            } else {
              d2 = ff_dist(qs[hw], trow, dim);
              have = true;
            }
          }
        } else {
          have = use_thr ? ff_chunk_rows<true>(ts_dyn + hw * FF_TILE, qs[hw], tbase, cb, ncol, dim, l, hmask, thr, &d2)
                         : ff_chunk_rows<false>(ts_dyn + hw * FF_TILE, qs[hw], tbase, cb, ncol, dim, l, hmask, thr, &d2);
        }
        gave_up = l < ncol && !have;
        if (have)
          k1 = (static_cast<unsigned long long>(__float_as_uint(d2)) << 32) | static_cast<unsigned int>(cb + l);
      }
      const bool any_gave_up = use_thr && __any_sync(hmask, gave_up);
      const float low = any_gave_up ? thr : inf;         // every abandoned column is >= low
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) {
        const unsigned long long o1 = __shfl_xor_sync(hmask, k1, off, 16);
        const unsigned long long o2 = __shfl_xor_sync(hmask, k2, off, 16);
        ff_merge(k1, k2, o1, o2);
      }
      ++n_chunks;
      if (k1 != KEY_NONE64) {  // how tight is the bound?  |key - exact chunk minimum score| / eps  (statistics only;
                               // an abandoned column is >= thr > every value kept, so k1 is still the chunk minimum)
        const float ex = __uint_as_float(static_cast<unsigned int>(k1 >> 32));
        const double err = fabs(static_cast<double>(Kj) - (static_cast<double>(ex) - bd.c));
        worst = fmaxf(worst, static_cast<float>(err / bd.eps(Kj)));
      }
      ff_merge(e1, e2, k1, k2);
      const float lbn = bd.lb(keys_s[r][j < nkeys - 1 ? j + 1 : nkeys - 1]);     // every column not evaluated so far is >= lbn
      const float d1sq = e1 == KEY_NONE64 ? inf : __uint_as_float(static_cast<unsigned int>(e1 >> 32));
      const float d2sq = e2 == KEY_NONE64 ? inf : __uint_as_float(static_cast<unsigned int>(e2 >> 32));
      const size_t o = base + row;
      if (need == L2F_NEED_FULL) {
        if (d2sq < lbn || lbn == inf) { done = true; if (l == 0) ff_write_exact(knn_idx, knn_dist, o, e1, e2); }
      } else if (need == L2F_NEED_NEAREST) {
        if (d1sq < lbn || lbn == inf) {
          done = true;
          if (l == 0) {
            knn_idx[o] = make_int2(static_cast<int>(e1 & 0xFFFFFFFFull), -1);
            knn_dist[o] = make_float2(__fsqrt_rn(d1sq), inf);
          }
        }
      } else {
        const float D1 = __fsqrt_rn(d1sq);
        const float rest = fminf(lbn, low);              // every column that was not evaluated in full is >= rest
        if ((lbn == inf && !any_gave_up) || (d1sq < rest && d2sq < rest)) {
          done = true;                                   // both neighbours certain
          if (l == 0) ff_write_exact(knn_idx, knn_dist, o, e1, e2);
        } else if (d1sq < rest && D1 < __fmul_rn(ratio, __fsqrt_rn(fminf(rest, d2sq)))) {
          done = true;                                   // nearest certain, true d2 >= min(rest, d2sq): passes whatever it is
          if (l == 0) {
            knn_idx[o] = make_int2(static_cast<int>(e1 & 0xFFFFFFFFull), 0x7ffffffe);
            knn_dist[o] = make_float2(D1, __fsqrt_rn(fminf(rest, d2sq)));
          }
        } else if (__fsqrt_rn(fminf(d1sq, rest)) >= __fmul_rn(ratio, __fsqrt_rn(d2sq))) {
          done = true;                                   // true d1 >= min(d1sq, rest), true d2 <= d2sq: fails
          if (l == 0) { knn_idx[o] = make_int2(-1, -1); knn_dist[o] = make_float2(inf, inf); }
        }
      }
      if (!done && any_gave_up) {                        // uniform over the half-warp: redo this chunk in full
        early = false;
        e1 = e2 = KEY_NONE64;
        --n_chunks;
        --j;
      }
    }
    if (!done && l == 0) ovf[atomicAdd(&n_ovf, 1)] = r;
  }
  __syncthreads();

  // ---- pass 3: rows the six chunks could not certify -- exhaustive exact scan by the whole block ----
  const int n_over = n_ovf;
  for (int e = 0; e < n_over; ++e) {
    const int row = span0 + ovf[e];
    const float* qrow = raw + (static_cast<size_t>(jb.q_row) + row) * dim;
    __syncthreads();
    for (int k = 4 * tid; k < dim; k += 4 * THREADS)
      *reinterpret_cast<float4*>(&qs[0][k]) = __ldg(reinterpret_cast<const float4*>(qrow + k));
    __syncthreads();
    unsigned long long m1 = KEY_NONE64, m2 = KEY_NONE64;
    for (int col = tid; col < jb.nt; col += THREADS) {
      const float d2 = ff_dist(qs[0], tbase + static_cast<size_t>(col) * dim, dim);
      const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(d2)) << 32) | static_cast<unsigned int>(col);
      const unsigned long long hi = key > m1 ? key : m1;
      m1 = key < m1 ? key : m1;
      m2 = m2 < hi ? m2 : hi;
    }
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
      const unsigned long long o1 = __shfl_xor_sync(hmask, m1, off, 16);
      const unsigned long long o2 = __shfl_xor_sync(hmask, m2, off, 16);
      ff_merge(m1, m2, o1, o2);
    }
    if (l == 0) { red[hw][0] = m1; red[hw][1] = m2; }
    __syncthreads();
    if (tid == 0) {
      for (int h = 1; h < HW; ++h) ff_merge(m1, m2, red[h][0], red[h][1]);
      ff_write_exact(knn_idx, knn_dist, base + row, m1, m2);
    }
  }

  if (counters != nullptr) {
    if (tid == 0) {
      if (n_need) atomicAdd(&counters[0], static_cast<unsigned long long>(n_need));
      if (n_over) atomicAdd(&counters[2], static_cast<unsigned long long>(n_over));
    }
    if (l == 0) {
      if (n_chunks) atomicAdd(&counters[1], static_cast<unsigned long long>(n_chunks));
      const unsigned int wb = __float_as_uint(worst);
      if (wb > static_cast<unsigned int>(counters[3])) atomicMax(&counters[3], static_cast<unsigned long long>(wb));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// l2f_rerank1_kernel -- the re-rank behind l2_i8x2_kernel MODE 4 (batched loop, ratio-test rows): ONE exact column per
// candidate row.  Input per query row (written by the tensor kernel): k1 = (acc << 13) | column of the smallest
// approximate score, k2 / k3 = the next chunk keys, w2 = second smallest acc inside k1's chunk.  Hence
//     K1 = score(k1),   K2 = score(min(w2, k2 >> 13)) = the second smallest approximate score of the whole row,
// and with the certified bound of ff_bound(e_mode 1): every column but c1 has an exact d^2 >= lb(K2).
//   level 0  the keys alone show that Lowe's test cannot pass                      -> closed, no neighbours reported
//            (a candidate row whose minimum the tensor kernel did not refine -- w2 == 0x7fffffff -- goes to level 2)
//   level 1  e1 = exact d^2 of column c1 (fp32 chain of ff_dist, bit-identical to the SIMT kernel):
//              e1 < lb(K2) and sqrt(e1) < ratio * sqrt(lb(K2))   -> nearest certain, test passes whatever d2 is
//              sqrt(min(e1, lb(K2))) >= ratio * sqrt(max(e1, ub(K2)))   -> the test fails whichever column is nearest
//   level 2  anything else (the outcome hinges on the uncertainty of K2; rare) -> flag 1 and the three chunk keys in
//            the float format of l2f_fixup_kernel, which then evaluates whole chunks / scans the row.
// One thread per row, 64 registers, no shared memory: several blocks fit next to the persistent tensor kernel.
static constexpr int R1_THREADS = 128;
static constexpr int R1_COLBITS = 13;
static constexpr int R1_ACC_NONE = 0x7fffffff >> R1_COLBITS;
__device__ __forceinline__ float r1_score(int acc) { return static_cast<float>(acc) * (2.f / (254.f * 254.f)); }
__device__ __forceinline__ float r1_float_key(int key) {        // chunk key -> the float key format of l2f_fixup_kernel
  if (key == 0x7fffffff) return __int_as_float(0x7f800000);
  const float sc = r1_score(key >> R1_COLBITS);
  return __uint_as_float((__float_as_uint(sc) & 0xFFFFFC00u) | static_cast<uint32_t>((key & ((1 << R1_COLBITS) - 1)) >> 4));
}
__global__ void __launch_bounds__(R1_THREADS, 8)
l2f_rerank1_kernel(const float* __restrict__ raw, const float* __restrict__ fnorm, int dim,
                   const PairJob* __restrict__ jobs, int2* __restrict__ knn_idx, float2* __restrict__ knn_dist,
                   float2* __restrict__ extra, uint8_t* __restrict__ flags, int stride, float ratio,
                   unsigned long long* __restrict__ counters, int e_mode) {
  const PairJob jb = jobs[blockIdx.y];
  const int row = blockIdx.x * R1_THREADS + threadIdx.x;
  if (blockIdx.x * R1_THREADS >= jb.nq) return;
  const float inf = __int_as_float(0x7f800000);
  int n_cand = 0, n_open = 0;
  float worst = 0.f;
  if (row < jb.nq) {
    const size_t o = static_cast<size_t>(blockIdx.y) * stride + row;
    const int2 k12 = knn_idx[o];
    const float2 kw = knn_dist[o];
    const int k1 = k12.x, k2 = k12.y, k3 = __float_as_int(kw.x), w2 = __float_as_int(kw.y);
    int2 oi = make_int2(-1, -1);
    float2 od = make_float2(inf, inf);
    uint8_t flag = 0;
    if (k1 != 0x7fffffff) {                                  // (no train rows at all: no neighbours)
      // refined: the tensor kernel resolved the column of the minimum and the second smallest score of its chunk; else
      // k1 carries the chunk's base column and the second smallest score of the row is only known to be <= score(k2)
      const bool refined = w2 != 0x7fffffff;
      const int acc2 = min(refined ? w2 : R1_ACC_NONE, k2 >> R1_COLBITS);
      const float K1 = r1_score(k1 >> R1_COLBITS), K2 = acc2 >= R1_ACC_NONE ? inf : r1_score(acc2);
      const FfBound bd = ff_bound(fnorm[jb.q_row + row], jb.t_maxn, dim, e_mode);
      // level 0: true d1^2 >= lb(K1), true d2^2 <= ub(K2) (K2 is at least an upper bound of the second smallest score)
      if (__fsqrt_rn(bd.lb(K1)) >= __fmul_rn(ratio, __fsqrt_rn(bd.ub(K2)))) {
        // closed: Lowe's test cannot pass
      } else if (!refined) {
        ++n_cand;
        flag = 1;
        od = make_float2(r1_float_key(k1), r1_float_key(k2));
        oi = make_int2(__float_as_int(r1_float_key(k3)), __float_as_int(inf));
        extra[o] = make_float2(inf, inf);
      } else {
        ++n_cand;
        const int c1 = k1 & ((1 << R1_COLBITS) - 1);
        const float4* qr = reinterpret_cast<const float4*>(raw + (static_cast<size_t>(jb.q_row) + row) * dim);
        const float4* tr = reinterpret_cast<const float4*>(raw + (static_cast<size_t>(jb.t_row) + c1) * dim);
        float e1 = 0.f;
#pragma unroll 8
        for (int k = 0; k < (dim >> 2); ++k) {               // the SIMT kernel's chain: d = a - b; acc = fma(d, d, acc)
          const float4 a = __ldg(qr + k), b = __ldg(tr + k);
          float d = a.x - b.x; e1 = fmaf(d, d, e1);
          d = a.y - b.y; e1 = fmaf(d, d, e1);
          d = a.z - b.z; e1 = fmaf(d, d, e1);
          d = a.w - b.w; e1 = fmaf(d, d, e1);
        }
        worst = static_cast<float>(fabs(static_cast<double>(K1) - (static_cast<double>(e1) - bd.c)) / bd.eps(K1));
        const float D1 = __fsqrt_rn(e1);
        const float rest = bd.lb(K2);                        // every column but c1 is >= rest
        if (K2 == inf) {                                     // the train image has a single row
          oi = make_int2(c1, -1);
          od = make_float2(D1, inf);
        } else if (e1 < rest && D1 < __fmul_rn(ratio, __fsqrt_rn(rest))) {
          oi = make_int2(c1, 0x7ffffffe);                    // nearest certain; true d2 >= sqrt(rest): passes whatever it is
          od = make_float2(D1, __fsqrt_rn(rest));
        } else if (__fsqrt_rn(fminf(e1, rest)) >= __fmul_rn(ratio, __fsqrt_rn(fmaxf(e1, bd.ub(K2))))) {
          // true d1^2 >= min(e1, rest) and true d2^2 <= max(e1, ub(K2)) whichever column is the nearest: fails
        } else {
          flag = 1;
          ++n_open;
          od = make_float2(r1_float_key(k1), r1_float_key(k2));
          oi = make_int2(__float_as_int(r1_float_key(k3)), __float_as_int(inf));
          extra[o] = make_float2(inf, inf);
        }
      }
    }
    knn_idx[o] = oi;
    knn_dist[o] = od;
    flags[o] = flag;
  }
  if (counters != nullptr) {
    n_cand = __reduce_add_sync(0xffffffffu, n_cand);
    n_open = __reduce_add_sync(0xffffffffu, n_open);
    const unsigned int wb = __reduce_max_sync(0xffffffffu, __float_as_uint(worst));
    if ((threadIdx.x & 31) == 0) {
      if (n_cand) { atomicAdd(&counters[0], static_cast<unsigned long long>(n_cand)); atomicAdd(&counters[1], static_cast<unsigned long long>(n_cand)); }
      if (wb > static_cast<unsigned int>(counters[3])) atomicMax(&counters[3], static_cast<unsigned long long>(wb));
    }
    (void)n_open;
  }
}

cudaError_t launch_l2f_rerank1(const float* raw, const float* fnorm, int dim, const PairJob* jobs, int n_jobs, int max_nq,
                               int2* idx, float2* dist, float2* extra, uint8_t* flags, int stride, float ratio,
                               unsigned long long* counters, const uint8_t* q8, const uint8_t* t8, cudaStream_t st, int e_mode) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  if (dim <= 0 || dim > FF_MAXDIM || (dim & 63)) return cudaErrorInvalidValue;
  l2f_rerank1_kernel<<<dim3((max_nq + R1_THREADS - 1) / R1_THREADS, n_jobs), R1_THREADS, 0, st>>>(
      raw, fnorm, dim, jobs, idx, dist, extra, flags, stride, ratio, counters, e_mode);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // the rows left open (flag 1), with three keys: the small-footprint variant, so that it too runs next to the tensor kernel
  l2f_fixup_kernel<true><<<dim3((max_nq + FFP_THREADS - 1) / FFP_THREADS, n_jobs), FFP_THREADS, 0, st>>>(
      raw, fnorm, dim, jobs, idx, dist, extra, stride, ratio, L2F_NEED_RATIO, counters, e_mode, q8, t8, flags, 3);
  return cudaGetLastError();
}

static constexpr int FF_DYN_SMEM = FF_HW * FF_TILE * static_cast<int>(sizeof(float));   // 34 KB
cudaError_t l2f_configure() {
  cudaError_t e = cudaFuncSetAttribute(l2f_fixup_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_DYN_SMEM);
  if (e != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2f_fixup_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(l2f_fixup_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

// q8 / t8: the s8 forms of the rows (pack_float_kernel), rows of dim + 32 bytes; with them, e_mode 1 and need =
// L2F_NEED_RATIO the prefiltering variant runs (the caller passes them only when asked to: api.cu, opt_prefilter).  The staged variant is the
// default: alone it needs 0.99 ms per 256 pairs against 1.27 ms (the survivors' one-row-per-lane loads have few
// requests in flight), and next to the tensor kernel the step is the same with either (63.7 k vs 62.6 k pairs/s).
cudaError_t launch_l2f_fixup(const float* raw, const float* fnorm, int dim, const PairJob* jobs, int n_jobs,
                             int max_nq, int2* idx, float2* dist, const float2* extra, int stride, float ratio,
                             int need, unsigned long long* counters, cudaStream_t st, int e_mode, const uint8_t* q8,
                             const uint8_t* t8) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  if (dim <= 0 || dim > FF_MAXDIM || (dim & 63)) return cudaErrorInvalidValue;
  if (e_mode == 1 && need == L2F_NEED_RATIO && q8 != nullptr && t8 != nullptr)
    l2f_fixup_kernel<true><<<dim3((max_nq + FFP_THREADS - 1) / FFP_THREADS, n_jobs), FFP_THREADS, 0, st>>>(raw, fnorm, dim, jobs, idx, dist, extra, stride, ratio, need, counters,
                                                        e_mode, q8, t8, nullptr, 6);
  else
    l2f_fixup_kernel<false><<<dim3((max_nq + FF_SPAN - 1) / FF_SPAN, n_jobs), FF_THREADS, FF_DYN_SMEM, st>>>(raw, fnorm, dim, jobs, idx, dist, extra, stride, ratio, need,
                                                                   counters, e_mode, nullptr, nullptr, nullptr, 6);
  return cudaGetLastError();
}

}  // namespace pm
