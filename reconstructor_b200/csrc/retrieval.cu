// retrieval.cu -- pair pre-selection by global-descriptor retrieval.
//
// The reference pairs every image with every other one (FakeImgMatcher::match, Mapper/libMapper/ImageMatcher.cpp:6-24 --
// "a temporary solution for img matching", ImageMatcher.h:26-28) and lists "image matcher (apply some image retrieval ...)"
// as a todo (README.md:40).  This is that plugin's device side, on the descriptors the handle already holds:
//   1. global descriptor of an image = sum over its keypoints of the L2-normalised local descriptor (float rows), or of the
//      +-1 vector of the bits (binary rows; an exact integer count), normalised to unit length            (global_desc_kernel)
//   2. cosine similarity of every image pair, fp64, k ascending                                           (similarity_kernel)
//   3. per image the top_k most similar other images, ties to the lower image index                       (topk_kernel)
// The host (api.cu) turns the directed lists into the canonical pair list (i < j, union of both directions).  With
// top_k >= n_images - 1 the result is FakeImgMatcher's all-pairs list.  Every reduction runs in a fixed order, so the result
// is deterministic; the CPU restatement of the parity tests sums in another order (agreement to ~1e-15 on the scores).
#include "common.cuh"
#include "kernels.h"

namespace pm {

static constexpr int GD_THREADS = 256;      // 8 warps, one row per warp at a time
static constexpr int GD_MAXDIM = 512;

// float rows: [rows][dim] fp32.  One block per image; warp w visits rows w, w + 8, ...; lane l owns elements l, l + 32, ...
__global__ void __launch_bounds__(GD_THREADS)
global_desc_float_kernel(const float* __restrict__ raw, int dim, const int2* __restrict__ img /* (first row, n) */,
                         double* __restrict__ gdesc /* [n_images][dim] */) {
  __shared__ double part[GD_THREADS / 32][GD_MAXDIM];
  __shared__ double red[GD_THREADS / 32];
  const int2 im = img[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (dim + 31) / 32;                       // <= 16
  double acc[GD_MAXDIM / 32];
#pragma unroll
  for (int e = 0; e < GD_MAXDIM / 32; ++e) acc[e] = 0.0;
  for (int r = warp; r < im.y; r += GD_THREADS / 32) {
    const float* row = raw + (static_cast<size_t>(im.x) + r) * dim;
    double v[GD_MAXDIM / 32];
    double n2 = 0.0;
#pragma unroll
    for (int e = 0; e < GD_MAXDIM / 32; ++e) {
      const int d = lane + 32 * e;
      v[e] = (e < per && d < dim) ? static_cast<double>(row[d]) : 0.0;
      n2 += v[e] * v[e];
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, off);
    const double inv = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
#pragma unroll
    for (int e = 0; e < GD_MAXDIM / 32; ++e) acc[e] += v[e] * inv;
  }
#pragma unroll
  for (int e = 0; e < GD_MAXDIM / 32; ++e)
    if (lane + 32 * e < dim) part[warp][lane + 32 * e] = acc[e];
  __syncthreads();
  double n2 = 0.0;
  for (int d = threadIdx.x; d < dim; d += GD_THREADS) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < GD_THREADS / 32; ++w) s += part[w][d];        // fixed order
    part[0][d] = s;
    n2 += s * s;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, off);
  if (lane == 0) red[warp] = n2;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < GD_THREADS / 32; ++w) tot += red[w];
  const double inv = tot > 0.0 ? 1.0 / sqrt(tot) : 0.0;
  for (int d = threadIdx.x; d < dim; d += GD_THREADS) gdesc[static_cast<size_t>(blockIdx.x) * dim + d] = part[0][d] * inv;
}

// binary rows: [rows][words] u32, bit b of a row lives in word b / 32.  g_b = (#ones at bit b) * 2 - n, an exact integer.
__global__ void __launch_bounds__(GD_THREADS)
global_desc_bits_kernel(const uint32_t* __restrict__ bits, int words, const int2* __restrict__ img,
                        double* __restrict__ gdesc /* [n_images][32 * words] */) {
  __shared__ int cnt[GD_MAXDIM];
  __shared__ double red[GD_THREADS / 32];
  const int2 im = img[blockIdx.x];
  const int dim = 32 * words;
  for (int d = threadIdx.x; d < dim; d += GD_THREADS) cnt[d] = 0;
  __syncthreads();
  // thread t owns bit position t (and t + 256 for 512-bit rows): no atomics, integer sums are order-free anyway
  for (int d = threadIdx.x; d < dim; d += GD_THREADS) {
    int c = 0;
    const int w = d >> 5, b = d & 31;
    for (int r = 0; r < im.y; ++r) c += (bits[(static_cast<size_t>(im.x) + r) * words + w] >> b) & 1u;
    cnt[d] = 2 * c - im.y;
  }
  __syncthreads();
  double n2 = 0.0;
  for (int d = threadIdx.x; d < dim; d += GD_THREADS) n2 += static_cast<double>(cnt[d]) * cnt[d];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = n2;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < GD_THREADS / 32; ++w) tot += red[w];
  const double inv = tot > 0.0 ? 1.0 / sqrt(tot) : 0.0;
  for (int d = threadIdx.x; d < dim; d += GD_THREADS) gdesc[static_cast<size_t>(blockIdx.x) * dim + d] = cnt[d] * inv;
}

__global__ void __launch_bounds__(256)
similarity_kernel(const double* __restrict__ g, int n, int dim, double* __restrict__ sim /* [n][n] */) {
  const int j = blockIdx.x * 16 + (threadIdx.x & 15), i = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (i >= n || j >= n) return;
  const double* a = g + static_cast<size_t>(i) * dim;
  const double* b = g + static_cast<size_t>(j) * dim;
  double s = 0.0;
  for (int k = 0; k < dim; ++k) s += a[k] * b[k];
  sim[static_cast<size_t>(i) * n + j] = s;
}

// One block per image: k rounds of "largest remaining score, lowest index on ties", never the image itself.
__global__ void __launch_bounds__(128)
topk_kernel(double* __restrict__ sim, int n, int k, int32_t* __restrict__ out /* [n][k] */) {
  __shared__ double bs[128];
  __shared__ int bi[128];
  const int i = blockIdx.x;
  double* row = sim + static_cast<size_t>(i) * n;
  const double NEG = -1e300;
  if (threadIdx.x == 0) row[i] = NEG;
  __syncthreads();
  for (int round = 0; round < k; ++round) {
    double best = NEG;
    int arg = 0x7fffffff;
    for (int j = threadIdx.x; j < n; j += 128) {
      const double v = row[j];
      if (v > best || (v == best && j < arg)) { best = v; arg = j; }
    }
    bs[threadIdx.x] = best; bi[threadIdx.x] = arg;
    __syncthreads();
    for (int off = 64; off >= 1; off >>= 1) {
      if (threadIdx.x < off) {
        const double v = bs[threadIdx.x + off];
        const int a = bi[threadIdx.x + off];
        if (v > bs[threadIdx.x] || (v == bs[threadIdx.x] && a < bi[threadIdx.x])) { bs[threadIdx.x] = v; bi[threadIdx.x] = a; }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const bool ok = bs[0] > NEG && bi[0] < n;
      out[static_cast<size_t>(i) * k + round] = ok ? bi[0] : -1;
      if (ok) row[bi[0]] = NEG;
    }
    __syncthreads();
  }
}

cudaError_t launch_retrieval(const float* raw, const uint32_t* bits, int dim, int words, const int2* d_img, int n_images,
                             int top_k, double* d_gdesc, double* d_sim, int32_t* d_topk, cudaStream_t st) {
  if (n_images <= 0) return cudaSuccess;
  const int gdim = bits ? 32 * words : dim;
  if (gdim > GD_MAXDIM) return cudaErrorInvalidValue;
  if (bits) global_desc_bits_kernel<<<n_images, GD_THREADS, 0, st>>>(bits, words, d_img, d_gdesc);
  else global_desc_float_kernel<<<n_images, GD_THREADS, 0, st>>>(raw, dim, d_img, d_gdesc);
  const dim3 grid((n_images + 15) / 16, (n_images + 15) / 16);
  similarity_kernel<<<grid, 256, 0, st>>>(d_gdesc, n_images, gdim, d_sim);
  (void)top_k; (void)d_topk;
  return cudaGetLastError();
}
// marks the entries it takes in d_sim (the caller copies the matrix out first if it wants it)
cudaError_t launch_topk(double* d_sim, int n_images, int top_k, int32_t* d_topk, cudaStream_t st) {
  if (n_images <= 0 || top_k <= 0) return cudaSuccess;
  topk_kernel<<<n_images, 128, 0, st>>>(d_sim, n_images, top_k, d_topk);
  return cudaGetLastError();
}

}  // namespace pm
