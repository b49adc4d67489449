// l2_fixup.cu -- exact index / second-neighbour recovery for the branch-free tensor-core epilogue.
//
// The fast tcgen05 kernel (l2_tc.cu, EPI 5) keeps, per query row, only VALUES: the smallest
// accumulator m1, the second smallest 32-column chunk minimum m2' and the base column of the first
// chunk that attained m1.  That is enough to decide almost everything:
//   * m1 is the exact nearest distance (the global minimum is the minimum of the chunk minima);
//   * the true second-nearest value is min(m2', second smallest value inside the winning chunk),
//     because any other chunk's non-minimal element is >= that chunk's minimum >= m2';
//   * the nearest index is the lowest column of the winning chunk whose distance equals m1 (the
//     winning chunk is the earliest one attaining m1, so this is the lowest index overall -- the
//     tie rule of cv::BFMatcher).
// This kernel recomputes the 32 distances of the winning chunk exactly (u8 dp4a: |a-b|^2 =
// |a|^2 + |b|^2 - 2 a.b in integers) for the rows that can still pass Lowe's ratio test
// (FeatureMatcher.cpp:55) -- or for all rows when the caller needs every nearest index -- and
// rewrites the row in the (idx1, idx2 | dist1, dist2) form the selection kernel consumes.
#include "common.cuh"
#include "kernels.h"

namespace pm {

static constexpr int FX_WARPS = 4;
static constexpr int FX_LD = 33;     // words per staged train row (32 + 1 pad: conflict-free column reads)

__global__ void __launch_bounds__(FX_WARPS * 32)
l2_fixup_kernel(const uint32_t* __restrict__ u8desc, const int32_t* __restrict__ qnorm,
                const PairJob* __restrict__ jobs, int reversed, int2* __restrict__ knn_idx,
                float2* __restrict__ knn_dist, int stride, float ratio, int all_rows) {
  __shared__ uint32_t stage[FX_WARPS][32 * FX_LD];
  const PairJob jb = jobs[blockIdx.y];
  const int q_row = reversed ? jb.t_row : jb.q_row, t_row = reversed ? jb.q_row : jb.t_row;
  const int nq = reversed ? jb.nt : jb.nq, nt = reversed ? jb.nq : jb.nt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = static_cast<size_t>(blockIdx.y) * stride;
  uint32_t* st = stage[warp];
  const float inf = __int_as_float(0x7f800000);

  for (int row = blockIdx.x * FX_WARPS + warp; row < nq; row += gridDim.x * FX_WARPS) {
    const int2 id = knn_idx[base + row];
    if (id.y != -2) continue;                          // not produced by the fast kernel
    const float2 dd = knn_dist[base + row];            // (d1^2, upper bound of d2^2), exact integers
    const int cb = id.x;
    int2 oi = make_int2(-1, -1);
    float2 od = make_float2(inf, inf);
    bool need = cb >= 0;
    if (need && !all_rows) {
      // the true d2 is <= the bound, so a row that fails with the bound fails for good
      need = __fsqrt_rn(dd.x) < __fmul_rn(ratio, __fsqrt_rn(dd.y));
    }
    if (need) {                                        // warp-uniform
      // stage the 32 train rows of the winning chunk (4 KB, coalesced 16-byte loads)
      const int ncol = min(32, nt - cb);
      const uint4* src = reinterpret_cast<const uint4*>(u8desc + (static_cast<size_t>(t_row) + cb) * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int f = i * 32 + lane;                   // 16-byte chunk index: row = f / 8, word = (f % 8) * 4
        const int r = f >> 3, w = (f & 7) * 4;
        uint4 x = make_uint4(0, 0, 0, 0);
        if (r < ncol) x = __ldg(src + f);
        uint32_t* d = st + r * FX_LD + w;
        d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w;
      }
      const uint32_t qw = __ldg(u8desc + (static_cast<size_t>(q_row) + row) * 32 + lane);
      __syncwarp();
      unsigned int dot = 0;
#pragma unroll
      for (int w = 0; w < 32; ++w) {
        const uint32_t a = __shfl_sync(0xffffffffu, qw, w);
        dot = __dp4a(a, st[lane * FX_LD + w], dot);
      }
      __syncwarp();
      const int na = qnorm[q_row + row];
      unsigned int key = 0xFFFFFFFFu;                  // (d^2 << 5 | lane): d^2 <= 2 * 128 * 255^2 < 2^24
      if (lane < ncol) {
        const int nb = qnorm[t_row + cb + lane];
        const int d2 = na + nb - 2 * static_cast<int>(dot);
        key = (static_cast<unsigned int>(d2) << 5) | static_cast<unsigned int>(lane);
      }
      const unsigned int k1 = __reduce_min_sync(0xffffffffu, key);
      const unsigned int k2 = __reduce_min_sync(0xffffffffu, key == k1 ? 0xFFFFFFFFu : key);
      const float d1 = static_cast<float>(k1 >> 5);
      float d2 = dd.y;                                 // second smallest chunk minimum
      int i2 = 0x7ffffffe;                             // somewhere outside the winning chunk
      if (k2 != 0xFFFFFFFFu && static_cast<float>(k2 >> 5) <= d2) {
        // (<=: on equal values the in-chunk element is reported; only its value matters downstream)
        d2 = static_cast<float>(k2 >> 5);
        i2 = cb + static_cast<int>(k2 & 31u);
      }
      oi.x = cb + static_cast<int>(k1 & 31u);
      oi.y = d2 < inf ? i2 : -1;                       // fewer than two train rows: no second neighbour
      od.x = __fsqrt_rn(d1);
      od.y = d2 < inf ? __fsqrt_rn(d2) : inf;
    }
    if (lane == 0) {
      knn_idx[base + row] = oi;
      knn_dist[base + row] = od;
    }
  }
}

cudaError_t launch_l2_fixup(const uint32_t* u8desc, const int32_t* qnorm, const PairJob* jobs, int n_jobs,
                            int max_nq, int reversed, int2* idx, float2* dist, int stride, float ratio,
                            int all_rows, cudaStream_t st) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  const int bx = min(64, (max_nq + FX_WARPS - 1) / FX_WARPS);
  dim3 grid(bx, n_jobs);
  l2_fixup_kernel<<<grid, FX_WARPS * 32, 0, st>>>(u8desc, qnorm, jobs, reversed, idx, dist, stride, ratio, all_rows);
  return cudaGetLastError();
}

}  // namespace pm
