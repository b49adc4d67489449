// l2_simt.cu -- K3: exact 2-NN under the L2 norm on the fp32 ALUs, sum_k (a_k - b_k)^2.
//
// This is the generic float path (SuperPoint 256-d unit-norm, FeatureSuperPoint.cpp:195-205, or
// any real-valued descriptor) and the re-rank / cross-check engine for the tensor-core path:
// for integer-valued descriptors every partial sum is an integer < 2^24, so the result is the
// exact d^2 and sqrtf() of it is bit-identical to cv::BFMatcher's DMatch::distance.
//
// Replaces knnMatch at Mapper/libMapper/FeatureMatcher.cpp:48-49 (brute force instead of FLANN).
//
// Tiling: a block owns 128 query rows and walks over the train rows in tiles of 128; 256
// threads, each an 8x8 micro-tile with interleaved rows/columns (row = ty + 16 i, col = tx + 16 j)
// so that the shared-memory reads of a k-slice are conflict free with a row stride of 17 floats.
// The running top-2 of every (thread,row) is a pair of 64-bit keys (float bits << 32 | column):
// unsigned min/max on the key orders by distance first, then by lowest index.
#include "common.cuh"
#include "kernels.h"

namespace pm {

static constexpr int LS_TILE = 128;
static constexpr int LS_KC = 16;
static constexpr int LS_LD = LS_KC + 1;
static constexpr int LS_THREADS = 256;

__device__ __forceinline__ void top2_insert(unsigned long long key, unsigned long long& m1,
                                            unsigned long long& m2) {
  const unsigned long long hi = key > m1 ? key : m1;
  m1 = key < m1 ? key : m1;
  m2 = m2 < hi ? m2 : hi;
}

__global__ void __launch_bounds__(LS_THREADS, 2)
l2_top2_simt_kernel(const float* __restrict__ desc, int dim, const PairJob* __restrict__ jobs,
                    int2* __restrict__ knn_idx, float2* __restrict__ knn_dist, int stride) {
  const PairJob job = jobs[blockIdx.y];
  const int q0 = blockIdx.x * LS_TILE;
  if (q0 >= job.nq) return;

  __shared__ float As[LS_TILE * LS_LD];
  __shared__ float Bs[LS_TILE * LS_LD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const float* qbase = desc + static_cast<size_t>(job.q_row) * dim;
  const float* tbase = desc + static_cast<size_t>(job.t_row) * dim;

  unsigned long long m1[8], m2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m1[i] = KEY_NONE64; m2[i] = KEY_NONE64; }

  for (int t0 = 0; t0 < job.nt; t0 += LS_TILE) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < dim; k0 += LS_KC) {
      // global -> shared: 128 rows x 16 floats for both operands, float4 along k
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int f = tid + LS_THREADS * u;
        const int row = f >> 2, kq = (f & 3) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (k0 + kq < dim) {
          const int qr = min(q0 + row, job.nq - 1);
          const int tr = min(t0 + row, job.nt - 1);
          a = __ldg(reinterpret_cast<const float4*>(qbase + static_cast<size_t>(qr) * dim + k0 + kq));
          b = __ldg(reinterpret_cast<const float4*>(tbase + static_cast<size_t>(tr) * dim + k0 + kq));
        }
        float* pa = &As[row * LS_LD + kq];
        float* pb = &Bs[row * LS_LD + kq];
        pa[0] = a.x; pa[1] = a.y; pa[2] = a.z; pa[3] = a.w;
        pb[0] = b.x; pb[1] = b.y; pb[2] = b.z; pb[3] = b.w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < LS_KC; ++k) {
        float a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = As[(ty + 16 * i) * LS_LD + k];
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = Bs[(tx + 16 * j) * LS_LD + k];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = a[i] - b[j];
            acc[i][j] = fmaf(d, d, acc[i][j]);
          }
      }
      __syncthreads();
    }

#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = t0 + tx + 16 * j;
      if (col < job.nt) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const unsigned long long key =
              (static_cast<unsigned long long>(__float_as_uint(acc[i][j])) << 32) |
              static_cast<unsigned int>(col);
          top2_insert(key, m1[i], m2[i]);
        }
      }
    }
  }

  // merge across the 16 threads (tx) that share a row: they are 16 consecutive lanes
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
      const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, m1[i], off);
      const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, m2[i], off);
      const unsigned long long lo = m1[i] < o1 ? m1[i] : o1;
      const unsigned long long hi = m1[i] < o1 ? o1 : m1[i];
      const unsigned long long s = m2[i] < o2 ? m2[i] : o2;
      m1[i] = lo;
      m2[i] = hi < s ? hi : s;
    }
    const int row = q0 + ty + 16 * i;
    if (tx == 0 && row < job.nq) {
      int2 oi;
      float2 od;
      const float inf = __int_as_float(0x7f800000);
      oi.x = m1[i] == KEY_NONE64 ? -1 : static_cast<int>(m1[i] & 0xFFFFFFFFull);
      oi.y = m2[i] == KEY_NONE64 ? -1 : static_cast<int>(m2[i] & 0xFFFFFFFFull);
      od.x = m1[i] == KEY_NONE64 ? inf : __fsqrt_rn(__uint_as_float(static_cast<unsigned int>(m1[i] >> 32)));
      od.y = m2[i] == KEY_NONE64 ? inf : __fsqrt_rn(__uint_as_float(static_cast<unsigned int>(m2[i] >> 32)));
      const size_t o = static_cast<size_t>(blockIdx.y) * stride + row;
      knn_idx[o] = oi;
      knn_dist[o] = od;
    }
  }
}

cudaError_t launch_l2_simt(const float* desc, int dim, const PairJob* jobs, int n_jobs, int max_nq,
                           int2* idx, float2* dist, int stride, cudaStream_t st) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  if (dim <= 0 || (dim & 3)) return cudaErrorInvalidValue;
  dim3 grid((max_nq + LS_TILE - 1) / LS_TILE, n_jobs);
  l2_top2_simt_kernel<<<grid, LS_THREADS, 0, st>>>(desc, dim, jobs, idx, dist, stride);
  return cudaGetLastError();
}

}  // namespace pm
