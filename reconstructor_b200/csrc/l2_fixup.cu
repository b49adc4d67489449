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

cudaError_t fixup_configure() {
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
