// hamming.cu -- K1: exact 2-NN under the Hamming norm for binary descriptors (ORB, 256 bit).
//
// Replaces the kNN step of FlannMatcher::matchFeatures (Mapper/libMapper/FeatureMatcher.cpp:48-49)
// for binary descriptors; parity target is cv::BFMatcher(NORM_HAMMING).knnMatch(k=2): exact
// integer distances, ties -> lowest train index.
//
// Mapping: one thread owns RQ query descriptors in registers (W x RQ words).  Train descriptors
// stream through shared memory in tiles fetched by the TMA unit as 1-D bulk copies
// (cp.async.bulk -> mbarrier), double buffered; every lane reads the same train row, so the
// shared-memory reads are 128-bit broadcasts.  Per (q,t): W XOR + popc (optionally a 3-stage
// carry-save reduction that trades 3 of the 8 POPC for 6 LOP3), then a branch-free running
// top-2 on the packed key (dist << 16 | train index) -- min/max on the key makes the lowest
// index win ties without any compare on the index.
//
// Algorithmic work: nq * nt * W popc32 per pair.
#include "common.cuh"
#include "kernels.h"

namespace pm {

static constexpr int HB_THREADS = 128;
static constexpr int HB_TILE = 256;   // train rows per shared-memory stage
static constexpr int HB_STAGES = 2;

template <int W, bool CSA>
__device__ __forceinline__ uint32_t hamming_words(const uint32_t (&q)[W], const uint32_t (&t)[W]) {
  if constexpr (CSA && W == 8) {
    uint32_t x[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) x[w] = q[w] ^ t[w];
    // carry-save adders: sum = a^b^c, carry = maj(a,b,c)
    const uint32_t s0 = x[0] ^ x[1] ^ x[2], c0 = (x[0] & x[1]) | (x[2] & (x[0] | x[1]));
    const uint32_t s1 = x[3] ^ x[4] ^ x[5], c1 = (x[3] & x[4]) | (x[5] & (x[3] | x[4]));
    const uint32_t s2 = s0 ^ s1 ^ x[6], c2 = (s0 & s1) | (x[6] & (s0 | s1));
    const int ones = __popc(s2) + __popc(x[7]);
    const int twos = __popc(c0) + __popc(c1) + __popc(c2);
    return static_cast<uint32_t>(ones + 2 * twos);
  } else {
    int d = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) d += __popc(q[w] ^ t[w]);
    return static_cast<uint32_t>(d);
  }
}

template <int W, int RQ, bool CSA>
__global__ void __launch_bounds__(HB_THREADS)
hamming_top2_kernel(const uint32_t* __restrict__ bits, const PairJob* __restrict__ jobs,
                    int2* __restrict__ knn_idx, float2* __restrict__ knn_dist, int stride) {
  const PairJob job = jobs[blockIdx.y];
  const int q0 = blockIdx.x * (HB_THREADS * RQ);
  if (q0 >= job.nq) return;

  __shared__ __align__(128) uint32_t tile[HB_STAGES][HB_TILE * W];
  __shared__ __align__(8) uint64_t full[HB_STAGES];

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < HB_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const uint32_t* tbase = bits + static_cast<size_t>(job.t_row) * W;
  const int n_tiles = (job.nt + HB_TILE - 1) / HB_TILE;
  auto issue = [&](int tile_i) {
    const int rows = min(HB_TILE, job.nt - tile_i * HB_TILE);
    const uint32_t bytes = static_cast<uint32_t>(rows) * W * 4u;
    uint64_t* bar = &full[tile_i % HB_STAGES];
    mbar_expect_tx(bar, bytes);
    bulk_g2s(&tile[tile_i % HB_STAGES][0], tbase + static_cast<size_t>(tile_i) * HB_TILE * W, bytes,
             bar);
  };
  if (tid == 0) {
    for (int s = 0; s < HB_STAGES && s < n_tiles; ++s) issue(s);
  }

  // query rows -> registers (coalesced 128-bit loads; rows past nq are clamped, never stored)
  uint32_t q[RQ][W];
  uint32_t m1[RQ], m2[RQ];
#pragma unroll
  for (int r = 0; r < RQ; ++r) {
    const int row = min(q0 + r * HB_THREADS + tid, job.nq - 1);
    const uint4* src = reinterpret_cast<const uint4*>(bits + (static_cast<size_t>(job.q_row) + row) * W);
#pragma unroll
    for (int v = 0; v < W / 4; ++v) {
      const uint4 x = __ldg(src + v);
      q[r][4 * v + 0] = x.x; q[r][4 * v + 1] = x.y; q[r][4 * v + 2] = x.z; q[r][4 * v + 3] = x.w;
    }
    m1[r] = KEY_NONE32; m2[r] = KEY_NONE32;
  }

  for (int ti = 0; ti < n_tiles; ++ti) {
    const int stage = ti % HB_STAGES;
    mbar_wait(&full[stage], (ti / HB_STAGES) & 1);
    const int rows = min(HB_TILE, job.nt - ti * HB_TILE);
    const uint4* trow = reinterpret_cast<const uint4*>(&tile[stage][0]);
    const uint32_t tkey0 = static_cast<uint32_t>(ti * HB_TILE);
#pragma unroll 2
    for (int j = 0; j < rows; ++j) {
      uint32_t t[W];
#pragma unroll
      for (int v = 0; v < W / 4; ++v) {
        const uint4 x = trow[j * (W / 4) + v];
        t[4 * v + 0] = x.x; t[4 * v + 1] = x.y; t[4 * v + 2] = x.z; t[4 * v + 3] = x.w;
      }
      const uint32_t tkey = tkey0 + j;
#pragma unroll
      for (int r = 0; r < RQ; ++r) {
        const uint32_t key = hamming_words<W, CSA>(q[r], t) * 65536u + tkey;
        const uint32_t hi = max(key, m1[r]);
        m1[r] = min(key, m1[r]);
        m2[r] = min(m2[r], hi);
      }
    }
    __syncthreads();                                   // stage fully consumed by every warp
    if (tid == 0 && ti + HB_STAGES < n_tiles) issue(ti + HB_STAGES);
  }

#pragma unroll
  for (int r = 0; r < RQ; ++r) {
    const int row = q0 + r * HB_THREADS + tid;
    if (row < job.nq) {
      int2 oi;
      float2 od;
      oi.x = m1[r] == KEY_NONE32 ? -1 : static_cast<int>(m1[r] & 0xFFFFu);
      oi.y = m2[r] == KEY_NONE32 ? -1 : static_cast<int>(m2[r] & 0xFFFFu);
      od.x = m1[r] == KEY_NONE32 ? __int_as_float(0x7f800000) : static_cast<float>(m1[r] >> 16);
      od.y = m2[r] == KEY_NONE32 ? __int_as_float(0x7f800000) : static_cast<float>(m2[r] >> 16);
      const size_t o = static_cast<size_t>(blockIdx.y) * stride + row;
      knn_idx[o] = oi;
      knn_dist[o] = od;
    }
  }
}

template <int W>
static cudaError_t launch_w(const uint32_t* bits, const PairJob* jobs, int n_jobs, int max_nq,
                            int2* idx, float2* dist, int stride, int variant, cudaStream_t st) {
  constexpr int RQ = 4;
  dim3 grid((max_nq + HB_THREADS * RQ - 1) / (HB_THREADS * RQ), n_jobs);
  if (variant == 1 && W == 8)
    hamming_top2_kernel<W, RQ, true><<<grid, HB_THREADS, 0, st>>>(bits, jobs, idx, dist, stride);
  else
    hamming_top2_kernel<W, RQ, false><<<grid, HB_THREADS, 0, st>>>(bits, jobs, idx, dist, stride);
  return cudaGetLastError();
}

cudaError_t launch_hamming_top2(const uint32_t* bits, int words, const PairJob* jobs, int n_jobs,
                                int max_nq, int2* idx, float2* dist, int stride, int variant,
                                cudaStream_t st) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  switch (words) {
    case 4: return launch_w<4>(bits, jobs, n_jobs, max_nq, idx, dist, stride, variant, st);
    case 8: return launch_w<8>(bits, jobs, n_jobs, max_nq, idx, dist, stride, variant, st);
    case 16: return launch_w<16>(bits, jobs, n_jobs, max_nq, idx, dist, stride, variant, st);
    default: return cudaErrorInvalidValue;
  }
}

// ---- popc pipe micro-benchmark: the roofline denominator MEASURED_PEAKS.json lacks -----------
__global__ void __launch_bounds__(256) popc_peak_kernel(uint32_t* out, int iters) {
  uint32_t a0 = threadIdx.x * 2654435761u + blockIdx.x, a1 = a0 ^ 0x9e3779b9u, a2 = a0 * 3u + 1u,
           a3 = ~a0, a4 = a0 + 77u, a5 = a1 * 5u, a6 = a2 ^ 0xdeadbeefu, a7 = a3 + 12345u;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = __popc(a0) + a1; a1 = __popc(a1) + a2; a2 = __popc(a2) + a3; a3 = __popc(a3) + a4;
      a4 = __popc(a4) + a5; a5 = __popc(a5) + a6; a6 = __popc(a6) + a7; a7 = __popc(a7) + a0;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

cudaError_t launch_popc_peak(uint32_t* out, int blocks, int iters, cudaStream_t st) {
  popc_peak_kernel<<<blocks, 256, 0, st>>>(out, iters);
  return cudaGetLastError();
}

}  // namespace pm
