// select.cu -- K4: Lowe ratio test + uniqueness / cross-check + ordered compaction.
//
// Replaces the loop at Mapper/libMapper/FeatureMatcher.cpp:51-64:
//     if (knn[i][0].distance < ratioThresh * knn[i][1].distance)       // float, strict, :55
//         if (trainIdx not taken by an earlier query) matches[q] = t;  // first wins, :58-62
// "first query wins" in ascending query order == for every train index keep the MINIMUM query
// index among the ratio-test survivors, which is an atomicMin instead of a serial scan.
// The survivors are compacted in ascending query order (std::map iteration order at
// SequentialReconstructor.cpp:243-247) and their pixel coordinates gathered as float
// (featuresToCvPoints, utils.cpp:165-177) for the epipolar filter.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace pm {

static constexpr int SEL_THREADS = 512;    // 16 K registers per block: co-resident with the persistent tensor kernel

__global__ void __launch_bounds__(SEL_THREADS)
ratio_unique_compact_kernel(const PairJob* __restrict__ jobs, const int2* __restrict__ knn_idx,
                            const float2* __restrict__ knn_dist, const int2* __restrict__ rev_idx,
                            const int32_t* __restrict__ xy, int stride, float ratio, int mode,
                            int32_t* __restrict__ owner, int32_t* __restrict__ match_q,
                            int32_t* __restrict__ match_t, float2* __restrict__ pts1,
                            float2* __restrict__ pts2, int32_t* __restrict__ count) {
  const int slot = blockIdx.x;
  const PairJob job = jobs[slot];
  const size_t base = static_cast<size_t>(slot) * stride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = SEL_THREADS / 32;
  __shared__ int warp_cnt[NW];
  __shared__ int run_base;

  int32_t* own = owner + base;
  if (mode == 0) {
    for (int t = tid; t < job.nt; t += SEL_THREADS) own[t] = 0x7fffffff;
    __syncthreads();
    for (int q = tid; q < job.nq; q += SEL_THREADS) {
      const int2 id = knn_idx[base + q];
      const float2 d = knn_dist[base + q];
      if (id.x >= 0 && id.y >= 0 && d.x < __fmul_rn(ratio, d.y)) atomicMin(&own[id.x], q);
    }
  }
  if (tid == 0) run_base = 0;
  __syncthreads();

  for (int c0 = 0; c0 < job.nq; c0 += SEL_THREADS) {
    const int q = c0 + tid;
    bool keep = false;
    int t = -1;
    if (q < job.nq) {
      const int2 id = knn_idx[base + q];
      const float2 d = knn_dist[base + q];
      t = id.x;
      keep = id.x >= 0 && id.y >= 0 && d.x < __fmul_rn(ratio, d.y);
      if (keep) {
        if (mode == 0) keep = own[t] == q;
        else if (mode == 1) keep = rev_idx[base + t].x == q;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int lane_off = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      const int v = lane < NW ? warp_cnt[lane] : 0;
      int s = v;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, s, off);
        if (lane >= off) s += o;
      }
      if (lane < NW) warp_cnt[lane] = s - v;  // exclusive prefix over the warps
    }
    __syncthreads();
    const int pos = run_base + warp_cnt[warp] + lane_off;
    if (keep) {
      match_q[base + pos] = q;
      match_t[base + pos] = t;
      float2 p1 = make_float2(0.f, 0.f), p2 = p1;
      if (xy) {
        const int2 a = *reinterpret_cast<const int2*>(xy + 2 * (static_cast<size_t>(job.q_row) + q));
        const int2 b = *reinterpret_cast<const int2*>(xy + 2 * (static_cast<size_t>(job.t_row) + t));
        p1 = make_float2(static_cast<float>(a.x), static_cast<float>(a.y));
        p2 = make_float2(static_cast<float>(b.x), static_cast<float>(b.y));
      }
      pts1[base + pos] = p1;
      pts2[base + pos] = p2;
    }
    __syncthreads();
    // last thread of the chunk knows the chunk total: exclusive offset of warp 31 + its count
    if (tid == SEL_THREADS - 1) run_base = run_base + warp_cnt[NW - 1] + __popc(bal);
    __syncthreads();
  }
  if (tid == 0) count[slot] = run_base;
}

cudaError_t launch_select(const PairJob* jobs, int n_jobs, const int2* knn_idx,
                          const float2* knn_dist, const int2* rev_idx, const int32_t* xy,
                          int stride, float ratio, int mode, int32_t* owner, int32_t* match_q,
                          int32_t* match_t, float2* pts1, float2* pts2, int32_t* count,
                          cudaStream_t st) {
  if (n_jobs <= 0) return cudaSuccess;
  ratio_unique_compact_kernel<<<n_jobs, SEL_THREADS, 0, st>>>(jobs, knn_idx, knn_dist, rev_idx, xy,
                                                              stride, ratio, mode, owner, match_q,
                                                              match_t, pts1, pts2, count);
  return cudaGetLastError();
}

}  // namespace pm

namespace pm {

// Exclusive scan of the per-slot counts (one block) and gather of the per-slot slabs into the
// contiguous CSR staging arrays that are copied to the host with one D2H each.
// 256 threads (8 K registers): a 1024-thread block (32 K registers) does not fit next to a persistent tensor kernel, so
// it ran only at the boundary between two of them -- every batch's tail stalled there, and behind the SuperPoint kernel
// (2.8 ms per launch) the tails fell so far behind that the kNN stream ran out of slots (PM_TRACE stage events).
static constexpr int SC_THREADS = 256;
__global__ void __launch_bounds__(SC_THREADS)
scan_counts_kernel(const int32_t* __restrict__ count, int n_jobs, int64_t* __restrict__ offsets) {
  constexpr int NW = SC_THREADS / 32;
  __shared__ long long warp_sum[NW];
  __shared__ long long carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n_jobs; c0 += SC_THREADS) {
    const int i = c0 + tid;
    const long long v = i < n_jobs ? count[i] : 0;
    long long s = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const long long o = __shfl_up_sync(0xffffffffu, s, off);
      if (lane >= off) s += o;
    }
    if (lane == 31) warp_sum[warp] = s;
    __syncthreads();
    if (warp == 0) {
      const long long w = lane < NW ? warp_sum[lane] : 0;
      long long t = w;
#pragma unroll
      for (int off = 1; off < NW; off <<= 1) {
        const long long o = __shfl_up_sync(0xffffffffu, t, off);
        if (lane >= off) t += o;
      }
      if (lane < NW) warp_sum[lane] = t - w;
    }
    __syncthreads();
    const long long excl = carry + warp_sum[warp] + (s - v);
    if (i < n_jobs) offsets[i] = excl;
    __syncthreads();
    if (tid == SC_THREADS - 1) carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) offsets[n_jobs] = carry;
}

__global__ void __launch_bounds__(256)
gather_slabs_kernel(const int32_t* __restrict__ count, int stride, const int32_t* __restrict__ match_q,
                    const int32_t* __restrict__ match_t, const uint8_t* __restrict__ mask,
                    const int64_t* __restrict__ offsets, int32_t* __restrict__ out_q,
                    int32_t* __restrict__ out_t, uint8_t* __restrict__ out_mask) {
  const int slot = blockIdx.x;
  const int n = count[slot];
  const size_t src = static_cast<size_t>(slot) * stride;
  const size_t dst = static_cast<size_t>(offsets[slot]);
  for (int i = threadIdx.x; i < n; i += 256) {
    out_q[dst + i] = match_q[src + i];
    out_t[dst + i] = match_t[src + i];
    out_mask[dst + i] = mask[src + i];
  }
}

cudaError_t launch_compact(const int32_t* count, int n_jobs, int stride, const int32_t* match_q,
                           const int32_t* match_t, const uint8_t* mask, int64_t* offsets,
                           int32_t* out_q, int32_t* out_t, uint8_t* out_mask, cudaStream_t st) {
  if (n_jobs <= 0) return cudaSuccess;
  scan_counts_kernel<<<1, SC_THREADS, 0, st>>>(count, n_jobs, offsets);
  gather_slabs_kernel<<<n_jobs, 256, 0, st>>>(count, stride, match_q, match_t, mask, offsets, out_q,
                                              out_t, out_mask);
  return cudaGetLastError();
}

cudaError_t select_configure() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(ratio_unique_compact_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(scan_counts_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(gather_slabs_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

// Copies between pinned host memory and device memory by a KERNEL (zero-copy accesses over PCIe) instead of the copy
// engines.  cudaMemcpyAsync copies of a stream are executed by the copy engine the driver bound that stream to, in the
// order they were enqueued on that engine -- and pm_set_images_async enqueues every image upload up front.  Measured
// (PM_TRACE, 100 x 8192 SIFT): the few KB of job lists / per-pair metadata of a slot whose stream shared its engine
// with the ingest stream waited behind ~50 uploads still queued, and the next batch of that slot started 5.8 ms late
// (end-to-end step 40 ms instead of 36.6).  Job lists and per-pair metadata therefore move by these kernels; the
// compacted matches (MBs) stay on the copy engine.
__global__ void copy_words_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n_words) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_words;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
}
__global__ void copy_bytes_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
}

cudaError_t launch_copy_pinned(const void* src, void* dst, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return cudaSuccess;
  const bool words = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | bytes) & 3) == 0;
  const size_t n = words ? bytes >> 2 : bytes;
  const unsigned grid = static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 592));
  if (words)
    copy_words_kernel<<<grid, 256, 0, st>>>(static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst), n);
  else
    copy_bytes_kernel<<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), n);
  return cudaGetLastError();
}

}  // namespace pm
