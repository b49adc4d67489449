// l2_fixup.cu -- exact index / second-neighbour recovery for the branch-free tensor-core epilogue.
//
// The values-only tcgen05 kernels (l2_tc.cu EPI 5, l2_tc2.cu MODE 2) keep, per query row, only VALUES:
// the smallest accumulator m1, the second smallest 16-column chunk minimum m2' and the base column of
// the first chunk that attained m1.  That is enough to decide almost everything:
//   * m1 is the exact nearest distance (the global minimum is the minimum of the chunk minima);
//   * the true second-nearest value is min(m2', second smallest value inside the winning chunk),
//     because any other chunk's non-minimal element is >= that chunk's minimum >= m2';
//   * the nearest index is the lowest column of the winning chunk whose distance equals m1 (the
//     winning chunk is the earliest one attaining m1, so this is the lowest index overall -- the
//     tie rule of cv::BFMatcher).
// This kernel recomputes the 16 distances of the winning chunk exactly (u8 dp4a: |a-b|^2 =
// |a|^2 + |b|^2 - 2 a.b in integers) for the rows that can still pass Lowe's ratio test
// (FeatureMatcher.cpp:55) -- or for all rows when the caller needs every nearest index -- and
// rewrites the row in the (idx1, idx2 | dist1, dist2) form the selection kernel consumes.
// Pass 1 closes the rows that cannot pass and lists the others; in pass 2 a half-warp owns a row: lane l
// of the half computes the distance to column cb + l.
#include "common.cuh"
#include "kernels.h"

namespace pm {

static constexpr int FX_WARPS = 4;
static constexpr int FX_SPAN = 256;  // rows per block
static constexpr int FX_LD = 33;     // words per staged train row (32 + 1 pad: conflict-free column reads)

__global__ void __launch_bounds__(FX_WARPS * 32)
l2_fixup_kernel(const uint32_t* __restrict__ u8desc, const int32_t* __restrict__ qnorm,
                const PairJob* __restrict__ jobs, int2* __restrict__ knn_idx,
                float2* __restrict__ knn_dist, int stride, float ratio, int all_rows) {
  __shared__ uint32_t stage[FX_WARPS * 2][16 * FX_LD];
  __shared__ int list[FX_SPAN];
  __shared__ int cnt;
  const PairJob jb = jobs[blockIdx.y];
  const int span0 = blockIdx.x * FX_SPAN;
  if (span0 >= jb.nq) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int hw = tid >> 4, l = lane & 15;
  const unsigned hmask = (lane & 16) ? 0xffff0000u : 0x0000ffffu;
  const size_t base = static_cast<size_t>(blockIdx.y) * stride;
  const float inf = __int_as_float(0x7f800000);

  // pass 1: one thread per row decides (coalesced reads); rows that cannot pass the ratio test with the
  // bound on d2 fail for good (the true d2 is <= the bound) and are closed here
  if (tid == 0) cnt = 0;
  __syncthreads();
  for (int r = tid; r < FX_SPAN; r += FX_WARPS * 32) {
    const int row = span0 + r;
    if (row >= jb.nq) break;
    const int2 id = knn_idx[base + row];
    if (id.y != -2) continue;                          // not produced by a values-only kernel
    const float2 dd = knn_dist[base + row];
    bool need = id.x >= 0;
    if (need && !all_rows) need = __fsqrt_rn(dd.x) < __fmul_rn(ratio, __fsqrt_rn(dd.y));
    if (need) list[atomicAdd(&cnt, 1)] = row;
    else { knn_idx[base + row] = make_int2(-1, -1); knn_dist[base + row] = make_float2(inf, inf); }
  }
  __syncthreads();
  const int n_need = cnt;

  // pass 2: one half-warp per surviving row
  uint32_t* st = stage[hw];
  for (int e = hw; e < n_need; e += FX_WARPS * 2) {
    const int row = list[e];
    const int cb = knn_idx[base + row].x;
    const float bound = knn_dist[base + row].y;        // second smallest chunk minimum (as d^2)
    const int ncol = min(16, jb.nt - cb);
    const uint4* src = reinterpret_cast<const uint4*>(u8desc + (static_cast<size_t>(jb.t_row) + cb) * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = i * 16 + l;                        // 16-byte piece: train row f / 8, words (f % 8) * 4 ..
      const int r = f >> 3, w = (f & 7) * 4;
      uint4 x = make_uint4(0, 0, 0, 0);
      if (r < ncol) x = __ldg(src + f);
      uint32_t* d = st + r * FX_LD + w;
      d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w;
    }
    const uint32_t* qs = u8desc + (static_cast<size_t>(jb.q_row) + row) * 32;
    const uint32_t q0 = __ldg(qs + l), q1 = __ldg(qs + 16 + l);
    __syncwarp(hmask);
    unsigned int dot = 0;
#pragma unroll
    for (int w = 0; w < 16; ++w) {
      const uint32_t a0 = __shfl_sync(hmask, q0, w, 16);
      const uint32_t a1 = __shfl_sync(hmask, q1, w, 16);
      dot = __dp4a(a0, st[l * FX_LD + w], dot);
      dot = __dp4a(a1, st[l * FX_LD + 16 + w], dot);
    }
    __syncwarp(hmask);
    unsigned int key = 0xFFFFFFFFu;                    // (d^2 << 4 | lane): d^2 <= 2 * 128 * 255^2 < 2^24
    if (l < ncol) {
      const int na = qnorm[jb.q_row + row];
      const int nb = qnorm[jb.t_row + cb + l];
      const int d2 = na + nb - 2 * static_cast<int>(dot);
      key = (static_cast<unsigned int>(d2) << 4) | static_cast<unsigned int>(l);
    }
    const unsigned int k1 = __reduce_min_sync(hmask, key);
    const unsigned int k2 = __reduce_min_sync(hmask, key == k1 ? 0xFFFFFFFFu : key);
    if (l == 0) {
      const float d1 = static_cast<float>(k1 >> 4);
      float d2 = bound;
      int i2 = 0x7ffffffe;                             // somewhere outside the winning chunk
      if (k2 != 0xFFFFFFFFu && static_cast<float>(k2 >> 4) <= d2) {
        d2 = static_cast<float>(k2 >> 4);
        i2 = cb + static_cast<int>(k2 & 15u);
      }
      int2 oi;
      float2 od;
      oi.x = cb + static_cast<int>(k1 & 15u);
      oi.y = d2 < inf ? i2 : -1;                       // fewer than two train rows: no second neighbour
      od.x = __fsqrt_rn(d1);
      od.y = d2 < inf ? __fsqrt_rn(d2) : inf;
      knn_idx[base + row] = oi;
      knn_dist[base + row] = od;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// I8 form (l2_tc2.cu KIND 2).  The tensor kernel tracked D' = 127*sum(a) - a.b + floor(|b|^2 / 2) instead of
// the distance itself:  |a - b|^2 = qoff[a] + 2 D' + (|b|^2 & 1).  Per query row it left m1 = the smallest D',
// m2' = the second smallest chunk minimum of D' (chunks of TC_I8_CHUNK columns) and the base column of the first chunk attaining m1.
// With lo(x) = qoff + 2x every distance outside the winning chunk is >= lo(m2'), and the smallest of them is
// lo(m2') or lo(m2') + 1.  After the winning chunk is recomputed exactly (d1 = its smallest distance, c2 = its
// second smallest):
//   * lo(m2') > d1  (always the case unless m2' == m1): the nearest neighbour is in the winning chunk, exact,
//     lowest index on ties;
//   * the second distance is c2 when c2 <= lo(m2'), else one of lo(m2'), lo(m2') + 1 -- the row is final when
//     Lowe's test gives the same answer for both (or when only the nearest index is needed);
//   * otherwise (m2' == m1, or a ratio test that hinges on one unit of d^2) the lane group rescans the whole
//     train image exactly.  Counted in counters[2]; measured: a handful of rows per million.
// CH = columns per chunk = lanes per group (16: half-warp, 32: warp); lane l of the group owns column cb + l.
template <int CH>
__device__ __forceinline__ unsigned int fx_chunk_key(uint32_t* st, const uint32_t* __restrict__ u8desc,
                                                     const int32_t* __restrict__ qnorm, const PairJob& jb, int cb,
                                                     int l, unsigned gmask, uint32_t q0, uint32_t q1, int na) {
  const int ncol = min(CH, jb.nt - cb);
  const uint4* src = reinterpret_cast<const uint4*>(u8desc + (static_cast<size_t>(jb.t_row) + cb) * 32);
  __syncwarp(gmask);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int f = i * CH + l;                          // 16-byte piece: train row f / 8, words (f % 8) * 4 ..
    const int r = f >> 3, w = (f & 7) * 4;
    uint4 x = make_uint4(0, 0, 0, 0);
    if (r < ncol) x = __ldg(src + f);
    uint32_t* d = st + r * FX_LD + w;
    d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w;
  }
  __syncwarp(gmask);
  unsigned int dot = 0;
  if (CH == 16) {
#pragma unroll
    for (int w = 0; w < 16; ++w) {
      const uint32_t a0 = __shfl_sync(gmask, q0, w, 16);
      const uint32_t a1 = __shfl_sync(gmask, q1, w, 16);
      dot = __dp4a(a0, st[l * FX_LD + w], dot);
      dot = __dp4a(a1, st[l * FX_LD + 16 + w], dot);
    }
  } else {
#pragma unroll
    for (int w = 0; w < 32; ++w) dot = __dp4a(__shfl_sync(gmask, q0, w, 32), st[l * FX_LD + w], dot);
  }
  unsigned int key = 0xFFFFFFFFu;                      // (d^2 << 5 | lane): d^2 < 2^24
  if (l < ncol) {
    const int nb = qnorm[jb.t_row + cb + l];
    key = (static_cast<unsigned int>(na + nb - 2 * static_cast<int>(dot)) << 5) | static_cast<unsigned int>(l);
  }
  return key;
}

static constexpr int FXI_INF = 0x7f800000;   // T2I_INF of l2_tc2.cu

// Two-stage evaluation of a 32-column chunk (one warp, lane l owns column cb + l).  Stage 1 reads only the first
// 32 of the 128 bytes of every candidate row and forms the partial distance over those 32 dimensions, a lower
// bound of the full one.  With d1_hi = the largest value the nearest distance can take (the tensor kernel left
// it up to one unit), a candidate is dropped when its partial distance exceeds d1_hi AND Lowe's test passes even
// with d1_hi against that partial distance: such a column can neither be the nearest nor make the test fail, so
// the outcome of the test and the nearest index computed from the survivors are exact (the stored second
// distance may then be larger than the true one -- only its comparison is used downstream).  When every nearest
// index is wanted (all_rows) only the first condition applies.  Stage 2 evaluates the survivors in full, one
// coalesced 128-byte row at a time.  Typical SIFT row with a true match: 1 survivor, 1.1 KB read instead of 4 KB.
#ifndef PM_I8_PREFIX
#define PM_I8_PREFIX 1
#endif
__device__ __forceinline__ unsigned int fx_chunk_key_prefix(const uint32_t* __restrict__ u8desc,
                                                            const int32_t* __restrict__ qnorm, const PairJob& jb,
                                                            int cb, int l, uint32_t qw, const uint32_t* qs, int na,
                                                            int d1_hi, float ratio, int all_rows, int* elim_lo) {
  const int ncol = min(32, jb.nt - cb);
  const uint32_t* trow = u8desc + (static_cast<size_t>(jb.t_row) + cb) * 32;
  uint4 b0 = make_uint4(0, 0, 0, 0), b1 = b0;
  if (l < ncol) {
    const uint4* bp = reinterpret_cast<const uint4*>(trow + l * 32);
    b0 = __ldg(bp); b1 = __ldg(bp + 1);
  }
  const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(qs)), a1 = __ldg(reinterpret_cast<const uint4*>(qs) + 1);
  unsigned int pa = 0, pb = 0, pab = 0;
#define PM_FX_W(A, B) pa = __dp4a(A, A, pa); pb = __dp4a(B, B, pb); pab = __dp4a(A, B, pab);
  PM_FX_W(a0.x, b0.x) PM_FX_W(a0.y, b0.y) PM_FX_W(a0.z, b0.z) PM_FX_W(a0.w, b0.w)
  PM_FX_W(a1.x, b1.x) PM_FX_W(a1.y, b1.y) PM_FX_W(a1.z, b1.z) PM_FX_W(a1.w, b1.w)
#undef PM_FX_W
  const int part = static_cast<int>(pa + pb) - 2 * static_cast<int>(pab);
  bool surv = l < ncol;
  if (surv && part > d1_hi)
    surv = !all_rows && !(__fsqrt_rn(static_cast<float>(d1_hi)) < __fmul_rn(ratio, __fsqrt_rn(static_cast<float>(part))));
  unsigned int alive = __ballot_sync(0xffffffffu, surv);
  // smallest partial distance among the dropped columns (a lower bound of their distances that still passes the test)
  *elim_lo = static_cast<int>(__reduce_min_sync(0xffffffffu, (l < ncol && !surv) ? static_cast<unsigned int>(part) : 0x7f800000u));
  unsigned int key = 0xFFFFFFFFu;
  while (alive) {
    const int j = __ffs(alive) - 1;
    alive &= alive - 1;
    const uint32_t w = __ldg(trow + j * 32 + l);
    const unsigned int dot = __reduce_add_sync(0xffffffffu, __dp4a(qw, w, 0u));
    if (l == j)
      key = (static_cast<unsigned int>(na + qnorm[jb.t_row + cb + j] - 2 * static_cast<int>(dot)) << 5) |
            static_cast<unsigned int>(l);
  }
  return key;
}

template <int CH>
__global__ void __launch_bounds__(FX_WARPS * 32)
l2_fixup_i8_kernel(const uint32_t* __restrict__ u8desc, const int32_t* __restrict__ qnorm,
                   const int32_t* __restrict__ qoff, const PairJob* __restrict__ jobs, int2* __restrict__ knn_idx,
                   float2* __restrict__ knn_dist, int stride, float ratio, int all_rows,
                   unsigned long long* __restrict__ counters) {
  constexpr int NG = FX_WARPS * 32 / CH;               // row groups per block
  __shared__ uint32_t stage[NG][CH * FX_LD];
  __shared__ int list[FX_SPAN];
  __shared__ int cnt;
  const PairJob jb = jobs[blockIdx.y];
  const int span0 = blockIdx.x * FX_SPAN;
  if (span0 >= jb.nq) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid / CH, l = tid % CH;
  const unsigned gmask = CH == 32 ? 0xffffffffu : ((lane & 16) ? 0xffff0000u : 0x0000ffffu);
  const size_t base = static_cast<size_t>(blockIdx.y) * stride;
  const float inf = __int_as_float(0x7f800000);

  // pass 1: rows that cannot pass the ratio test even with the most favourable reading of (m1, m2') are closed
  if (tid == 0) cnt = 0;
  __syncthreads();
  for (int r = tid; r < FX_SPAN; r += FX_WARPS * 32) {
    const int row = span0 + r;
    if (row >= jb.nq) break;
    const int2 id = knn_idx[base + row];
    if (id.y != -3) continue;                          // not produced by the i8 kernel
    const float2 dd = knn_dist[base + row];
    const int m1 = __float_as_int(dd.x), m2 = __float_as_int(dd.y);
    bool need = id.x >= 0;
    if (need && !all_rows && m2 != FXI_INF) {
      const int qo = qoff[jb.q_row + row];
      const float lo1 = static_cast<float>(max(qo + 2 * m1, 0));
      const float hi2 = static_cast<float>(qo + 2 * m2 + 1);
      need = __fsqrt_rn(lo1) < __fmul_rn(ratio, __fsqrt_rn(hi2));
    }
    if (need) list[atomicAdd(&cnt, 1)] = row;
    else { knn_idx[base + row] = make_int2(-1, -1); knn_dist[base + row] = make_float2(inf, inf); }
  }
  __syncthreads();
  const int n_need = cnt;

  // pass 2: one lane group per surviving row
  uint32_t* st = stage[grp];
  unsigned int n_rescan = 0;
  for (int e = grp; e < n_need; e += NG) {
    const int row = list[e];
    const int cb = knn_idx[base + row].x;
    const int m2 = __float_as_int(knn_dist[base + row].y);
    const int na = qnorm[jb.q_row + row];
    const uint32_t* qs = u8desc + (static_cast<size_t>(jb.q_row) + row) * 32;
    const uint32_t q0 = __ldg(qs + l), q1 = CH == 16 ? __ldg(qs + 16 + l) : 0u;
    unsigned int key;
    int elim_lo = FXI_INF;                             // FXI_INF: no column was dropped by the prefix filter
    if (CH == 32 && PM_I8_PREFIX) {
      const int m1 = __float_as_int(knn_dist[base + row].x);
      key = fx_chunk_key_prefix(u8desc, qnorm, jb, cb, l, q0, qs, na, qoff[jb.q_row + row] + 2 * m1 + 1, ratio, all_rows,
                                &elim_lo);
    } else {
      key = fx_chunk_key<CH>(st, u8desc, qnorm, jb, cb, l, gmask, q0, q1, na);
    }
    const unsigned int k1 = __reduce_min_sync(gmask, key);
    const unsigned int k2 = __reduce_min_sync(gmask, key == k1 ? 0xFFFFFFFFu : key);
    const int d1 = static_cast<int>(k1 >> 5);
    int i1 = cb + static_cast<int>(k1 & 31u), i2 = -1, d2 = FXI_INF;
    bool rescan = false;
    if (m2 == FXI_INF) {                               // a single chunk: everything is in it
      if (k2 != 0xFFFFFFFFu) { d2 = static_cast<int>(k2 >> 5); i2 = cb + static_cast<int>(k2 & 31u); }
      else if (elim_lo != FXI_INF) { d2 = elim_lo; i2 = 0x7ffffffe; }   // second neighbour among the dropped columns:
                                                       // it exists and Lowe's test passes against it whatever it is
    } else {
      const int others = qoff[jb.q_row + row] + 2 * m2;        // smallest outside distance is others or others + 1
      const int c2 = k2 != 0xFFFFFFFFu ? static_cast<int>(k2 >> 5) : FXI_INF;
      if (others <= d1) {
        rescan = true;
      } else if (c2 <= others) {
        d2 = c2; i2 = cb + static_cast<int>(k2 & 31u);
      } else {
        d2 = others; i2 = 0x7ffffffe;                  // somewhere outside the winning chunk
        if (!all_rows) {
          const float s1 = __fsqrt_rn(static_cast<float>(d1));
          const bool p_lo = s1 < __fmul_rn(ratio, __fsqrt_rn(static_cast<float>(others)));
          const bool p_hi = s1 < __fmul_rn(ratio, __fsqrt_rn(static_cast<float>(min(c2, others + 1))));
          rescan = p_lo != p_hi;
        }
      }
    }
    if (rescan) {                                      // uniform over the group
      ++n_rescan;
      unsigned long long b1 = KEY_NONE64, b2 = KEY_NONE64;     // (d^2 << 32 | column), per lane
      for (int c = 0; c < jb.nt; c += CH) {
        const unsigned int k = fx_chunk_key<CH>(st, u8desc, qnorm, jb, c, l, gmask, q0, q1, na);
        if (k != 0xFFFFFFFFu) {
          const unsigned long long kk = (static_cast<unsigned long long>(k >> 5) << 32) | static_cast<unsigned int>(c + l);
          if (kk < b1) { b2 = b1; b1 = kk; } else if (kk < b2) b2 = kk;
        }
      }
#pragma unroll
      for (int off = CH / 2; off >= 1; off >>= 1) {
        const unsigned long long o1 = __shfl_xor_sync(gmask, b1, off, CH);
        const unsigned long long o2 = __shfl_xor_sync(gmask, b2, off, CH);
        const unsigned long long lo = b1 < o1 ? b1 : o1, hi = b1 < o1 ? o1 : b1;
        const unsigned long long m = b2 < o2 ? b2 : o2;
        b1 = lo; b2 = hi < m ? hi : m;
      }
      i1 = static_cast<int>(b1 & 0xFFFFFFFFu);
      const int dd1 = static_cast<int>(b1 >> 32);
      if (b2 != KEY_NONE64) { d2 = static_cast<int>(b2 >> 32); i2 = static_cast<int>(b2 & 0xFFFFFFFFu); }
      else { d2 = FXI_INF; i2 = -1; }
      if (l == 0) {
        knn_idx[base + row] = make_int2(i1, i2);
        knn_dist[base + row] = make_float2(__fsqrt_rn(static_cast<float>(dd1)),
                                           i2 >= 0 ? __fsqrt_rn(static_cast<float>(d2)) : inf);
      }
    } else if (l == 0) {
      knn_idx[base + row] = make_int2(i1, d2 != FXI_INF ? i2 : -1);
      knn_dist[base + row] = make_float2(__fsqrt_rn(static_cast<float>(d1)),
                                         d2 != FXI_INF ? __fsqrt_rn(static_cast<float>(d2)) : inf);
    }
  }
  if (counters) {
    if (tid == 0 && n_need) atomicAdd(&counters[0], static_cast<unsigned long long>(n_need));
    if (l == 0 && n_rescan) atomicAdd(&counters[2], static_cast<unsigned long long>(n_rescan));
  }
}

cudaError_t launch_l2_fixup_i8(const uint32_t* u8desc, const int32_t* qnorm, const int32_t* qoff, const PairJob* jobs,
                               int n_jobs, int max_nq, int2* idx, float2* dist, int stride, float ratio, int all_rows,
                               unsigned long long* counters, cudaStream_t st) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  const int bx = (max_nq + FX_SPAN - 1) / FX_SPAN;
  dim3 grid(bx, n_jobs);
  l2_fixup_i8_kernel<TC_I8_CHUNK><<<grid, FX_WARPS * 32, 0, st>>>(u8desc, qnorm, qoff, jobs, idx, dist, stride, ratio,
                                                                  all_rows, counters);
  return cudaGetLastError();
}

cudaError_t fixup_configure() {
  cudaError_t e = cudaFuncSetAttribute(l2_fixup_i8_kernel<TC_I8_CHUNK>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(l2_fixup_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

cudaError_t launch_l2_fixup(const uint32_t* u8desc, const int32_t* qnorm, const PairJob* jobs, int n_jobs,
                            int max_nq, int reversed, int2* idx, float2* dist, int stride, float ratio,
                            int all_rows, cudaStream_t st) {
  (void)reversed;   // callers pass already-swapped jobs for the reverse search
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  const int bx = (max_nq + FX_SPAN - 1) / FX_SPAN;
  dim3 grid(bx, n_jobs);
  l2_fixup_kernel<<<grid, FX_WARPS * 32, 0, st>>>(u8desc, qnorm, jobs, idx, dist, stride, ratio, all_rows);
  return cudaGetLastError();
}

}  // namespace pm
