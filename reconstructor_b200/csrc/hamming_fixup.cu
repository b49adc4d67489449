// hamming_fixup.cu -- exact index / second-neighbour recovery behind the tensor-core Hamming search.
//
// l2_top2_tc2_kernel<.., MODE 2, KIND f8f6f4> keeps per query row only VALUES (see l2_fixup.cu for the argument):
// the smallest Hamming distance m1, the second smallest 16-column chunk minimum m2' and the base column of the
// earliest chunk attaining m1.  This kernel recomputes the 16 distances of that chunk with XOR + popc on the
// original bits for the rows that can still pass Lowe's ratio test (FeatureMatcher.cpp:55; Hamming distances are
// compared as they are, no square root) -- or for all rows (cross-check direction) -- and rewrites the row as
// (idx1, idx2 | dist1, dist2) with cv::BFMatcher(NORM_HAMMING)'s tie rule (lowest train index).
#include "common.cuh"
#include "kernels.h"

namespace pm {

static constexpr int HF_THREADS = 128;
static constexpr int HF_SPAN = 256;

// CH = columns per chunk = lanes per row group (16 / 32); INTS: the tensor kernel left integer values with marker -3
// (kind::i8 two-set kernel, 0x7f800000 = none) instead of floats with marker -2.
template <int W, int CH = 16, bool INTS = false>
__global__ void __launch_bounds__(HF_THREADS)
hamming_fixup_kernel(const uint32_t* __restrict__ bits, const PairJob* __restrict__ jobs, int2* __restrict__ knn_idx,
                     float2* __restrict__ knn_dist, int stride, float ratio, int all_rows) {
  __shared__ int list[HF_SPAN];
  __shared__ int cnt;
  const PairJob jb = jobs[blockIdx.y];
  const int span0 = blockIdx.x * HF_SPAN;
  if (span0 >= jb.nq) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int hw = tid / CH, l = tid % CH;
  const unsigned hmask = CH == 32 ? 0xffffffffu : ((lane & 16) ? 0xffff0000u : 0x0000ffffu);
  const size_t base = static_cast<size_t>(blockIdx.y) * stride;
  const float inf = __int_as_float(0x7f800000);

  if (tid == 0) cnt = 0;
  __syncthreads();
  for (int r = tid; r < HF_SPAN; r += HF_THREADS) {
    const int row = span0 + r;
    if (row >= jb.nq) break;
    const int2 id = knn_idx[base + row];
    if (id.y != (INTS ? -3 : -2)) continue;
    float2 dd = knn_dist[base + row];
    if (INTS) {
      const int m1 = __float_as_int(dd.x), m2 = __float_as_int(dd.y);
      dd.x = m1 == 0x7f800000 ? inf : static_cast<float>(m1);
      dd.y = m2 == 0x7f800000 ? inf : static_cast<float>(m2);
      knn_dist[base + row] = dd;                        // pass 2 reads the bound as a float
    }
    bool need = id.x >= 0;
    if (need && !all_rows) need = dd.x < __fmul_rn(ratio, dd.y);      // true d2 <= the bound: fails for good otherwise
    if (need) list[atomicAdd(&cnt, 1)] = row;
    else { knn_idx[base + row] = make_int2(-1, -1); knn_dist[base + row] = make_float2(inf, inf); }
  }
  __syncthreads();
  const int n_need = cnt;

  for (int e = hw; e < n_need; e += HF_THREADS / CH) {
    const int row = list[e];
    const int cb = knn_idx[base + row].x;
    const float bound = knn_dist[base + row].y;
    const int ncol = min(CH, jb.nt - cb);
    unsigned int key = 0xFFFFFFFFu;                    // (distance << 5 | lane)
    if (l < ncol) {
      const uint4* q = reinterpret_cast<const uint4*>(bits + (static_cast<size_t>(jb.q_row) + row) * W);
      const uint4* t = reinterpret_cast<const uint4*>(bits + (static_cast<size_t>(jb.t_row) + cb + l) * W);
      int d = 0;
#pragma unroll
      for (int w = 0; w < W / 4; ++w) {
        const uint4 a = __ldg(q + w), b = __ldg(t + w);
        d += __popc(a.x ^ b.x) + __popc(a.y ^ b.y) + __popc(a.z ^ b.z) + __popc(a.w ^ b.w);
      }
      key = (static_cast<unsigned int>(d) << 5) | static_cast<unsigned int>(l);
    }
    const unsigned int k1 = __reduce_min_sync(hmask, key);
    const unsigned int k2 = __reduce_min_sync(hmask, key == k1 ? 0xFFFFFFFFu : key);
    if (l == 0) {
      float d2 = bound;
      int i2 = 0x7ffffffe;                             // somewhere outside the winning chunk
      if (k2 != 0xFFFFFFFFu && static_cast<float>(k2 >> 5) <= d2) {
        d2 = static_cast<float>(k2 >> 5);
        i2 = cb + static_cast<int>(k2 & 31u);
      }
      knn_idx[base + row] = make_int2(cb + static_cast<int>(k1 & 31u), d2 < inf ? i2 : -1);
      knn_dist[base + row] = make_float2(static_cast<float>(k1 >> 5), d2);
    }
  }
}

cudaError_t hamming_fixup_configure() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(hamming_fixup_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(hamming_fixup_kernel<8, 32, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(hamming_fixup_kernel<8, 32, false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(hamming_fixup_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

cudaError_t launch_hamming_fixup(const uint32_t* bits, int words, const PairJob* jobs, int n_jobs, int max_nq,
                                 int2* idx, float2* dist, int stride, float ratio, int all_rows, cudaStream_t st,
                                 int ints) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  dim3 grid((max_nq + HF_SPAN - 1) / HF_SPAN, n_jobs);
  if (ints == 2 && words == 8)
    hamming_fixup_kernel<8, 32, false><<<grid, HF_THREADS, 0, st>>>(bits, jobs, idx, dist, stride, ratio, all_rows);
  else if (ints == 2 && words == 16)
    hamming_fixup_kernel<16, 32, false><<<grid, HF_THREADS, 0, st>>>(bits, jobs, idx, dist, stride, ratio, all_rows);
  else if (ints == 1 && words == 8)
    hamming_fixup_kernel<8, 32, true><<<grid, HF_THREADS, 0, st>>>(bits, jobs, idx, dist, stride, ratio, all_rows);
  else if (ints)
    return cudaErrorInvalidValue;
  else if (words == 8)
    hamming_fixup_kernel<8><<<grid, HF_THREADS, 0, st>>>(bits, jobs, idx, dist, stride, ratio, all_rows);
  else if (words == 16)
    hamming_fixup_kernel<16><<<grid, HF_THREADS, 0, st>>>(bits, jobs, idx, dist, stride, ratio, all_rows);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace pm
