// pack.cu -- one-pass ingest of an image's descriptors into the device layouts of the kNN kernels.
//
// Replaces featDescToCV (Mapper/libMapper/FeatureMatcher.cpp:11-25), which the reference re-runs
// for both images of every pair; here it runs once per image (pm_set_image).
//
// SIFT-like rows (128 values, integer-valued in [0,255]; FeatureDetector.cpp:20-24) are written
// in the two operand forms of the tcgen05 kernel, 144 fp16 per row:
//   query form  [ -2*a_0 .. -2*a_127 | 1, 2048, 2048, 0 x13 ]
//   train form  [    b_0 ..    b_127 | lo, mid, 2048*hi, 0 x13 ]   with |b|^2 = lo + 2048*mid + 2048*2048*hi
// so that the MMA accumulator equals |b|^2 - 2 a.b exactly (see l2_tc.cu).
//
// The same rows are also written in the BYTE forms of the kind::i8 kernel (l2_tc2.cu KIND 2), 160 bytes per row:
//   query form  [ a_0 .. a_127 (u8)        | 1, 255 x31 (u8) ]
//   train form  [ 127 - b_0 .. (s8)        | r, e_1 .. e_31 (u8) ]   with floor(|b|^2 / 2) = r + 255 * sum(e_j)
// and qoff[row] = |a|^2 - 254 * sum(a), so that |a - b|^2 = 2 * accumulator + (|b|^2 & 1) + qoff[a].
// flag bit 1 is raised when a row's norm does not fit the 31 digits (|b|^2 > 4,032,059; impossible for
// SIFT, whose rows have norm ~512): such an image stays on the fp16 form.
#include "common.cuh"
#include "kernels.h"

namespace pm {

__global__ void __launch_bounds__(256)
pack_sift_kernel(const float* __restrict__ raw_f32, const uint8_t* __restrict__ raw_u8, int n,
                 __half* __restrict__ qf, __half* __restrict__ tf, int32_t* __restrict__ qnorm,
                 float* __restrict__ raw_out, uint32_t* __restrict__ u8_out, int* __restrict__ not_integral,
                 uint8_t* __restrict__ iq, uint8_t* __restrict__ it, int32_t* __restrict__ qoff,
                 volatile uint8_t* __restrict__ host_flags) {
  // host_flags (asynchronous ingest): the two facts are stored straight into the image's pinned record (byte 0: some
  // value is not an integer in [0, 255]; byte 1: a squared norm beyond the byte form's range) -- no flag memset and no
  // device-to-host copy per image; every writer stores the same value, the record is read after the image's event
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  float v[4];
  if (raw_u8) {
    const uchar4 u = *reinterpret_cast<const uchar4*>(raw_u8 + static_cast<size_t>(row) * TC_DIM + 4 * lane);
    v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
  } else {
    const float4 f = *reinterpret_cast<const float4*>(raw_f32 + static_cast<size_t>(row) * TC_DIM + 4 * lane);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
  if (raw_out)
    *reinterpret_cast<float4*>(raw_out + static_cast<size_t>(row) * TC_DIM + 4 * lane) =
        make_float4(v[0], v[1], v[2], v[3]);
  // byte copy of the row (4 values per word) for the integer fix-up kernel
  const uint32_t w =
      static_cast<uint32_t>(static_cast<int>(fminf(fmaxf(v[0], 0.f), 255.f))) |
      (static_cast<uint32_t>(static_cast<int>(fminf(fmaxf(v[1], 0.f), 255.f))) << 8) |
      (static_cast<uint32_t>(static_cast<int>(fminf(fmaxf(v[2], 0.f), 255.f))) << 16) |
      (static_cast<uint32_t>(static_cast<int>(fminf(fmaxf(v[3], 0.f), 255.f))) << 24);
  u8_out[static_cast<size_t>(row) * 32 + lane] = w;
  bool bad = false;
  int nrm = 0, sum = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    bad |= !(v[e] >= 0.f && v[e] <= 255.f && v[e] == rintf(v[e]));
    const int iv = static_cast<int>(v[e]);
    nrm += iv * iv;
    sum += iv;
  }
  if (__any_sync(0xffffffffu, bad)) {
    if (lane == 0) { if (host_flags) host_flags[0] = 1; else atomicOr(not_integral, 1); }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    nrm += __shfl_xor_sync(0xffffffffu, nrm, off);
    sum += __shfl_xor_sync(0xffffffffu, sum, off);
  }
  if (iq) {
    uint8_t* qrow = iq + static_cast<size_t>(row) * TC_I8_ROW;
    uint8_t* trow = it + static_cast<size_t>(row) * TC_I8_ROW;
    reinterpret_cast<uint32_t*>(qrow)[lane] = w;
    // 127 - b as a two's-complement byte
    const uint32_t b0 = w & 0xFFu, b1 = (w >> 8) & 0xFFu, b2 = (w >> 16) & 0xFFu, b3 = w >> 24;
    reinterpret_cast<uint32_t*>(trow)[lane] = ((127u - b0) & 0xFFu) | (((127u - b1) & 0xFFu) << 8) |
                                              (((127u - b2) & 0xFFu) << 16) | (((127u - b3) & 0xFFu) << 24);
    const int h = nrm >> 1, qd = h / 255, r = h - 255 * qd;
    int e = 0;
    if (lane == 0) e = r;
    else e = min(255, max(0, qd - 255 * (lane - 1)));
    qrow[TC_DIM + lane] = lane == 0 ? 1 : 255;
    trow[TC_DIM + lane] = static_cast<uint8_t>(e);
    if (lane == 0) {
      qoff[row] = nrm - 254 * sum;
      if (qd > 31 * 255) { if (host_flags) host_flags[1] = 1; else atomicOr(not_integral, 2); }
    }
  }

  __half2 q01 = __floats2half2_rn(-2.f * v[0], -2.f * v[1]);
  __half2 q23 = __floats2half2_rn(-2.f * v[2], -2.f * v[3]);
  __half2 t01 = __floats2half2_rn(v[0], v[1]);
  __half2 t23 = __floats2half2_rn(v[2], v[3]);
  __half* qrow = qf + static_cast<size_t>(row) * TC_KPAD;
  __half* trow = tf + static_cast<size_t>(row) * TC_KPAD;
  reinterpret_cast<__half2*>(qrow)[2 * lane] = q01;
  reinterpret_cast<__half2*>(qrow)[2 * lane + 1] = q23;
  reinterpret_cast<__half2*>(trow)[2 * lane] = t01;
  reinterpret_cast<__half2*>(trow)[2 * lane + 1] = t23;
  if (lane < 16) {
    float qe = 0.f, te = 0.f;
    if (lane == 0) { qe = 1.f; te = static_cast<float>(nrm & 2047); }
    else if (lane == 1) { qe = 2048.f; te = static_cast<float>((nrm >> 11) & 2047); }
    else if (lane == 2) { qe = 2048.f; te = 2048.f * static_cast<float>(nrm >> 22); }
    qrow[TC_DIM + lane] = __float2half_rn(qe);
    trow[TC_DIM + lane] = __float2half_rn(te);
  }
  if (lane == 0) qnorm[row] = nrm;
}

cudaError_t launch_pack_sift(const float* raw_f32, const uint8_t* raw_u8, int n, __half* qf,
                             __half* tf, int32_t* qnorm, float* raw_out, uint32_t* u8_out,
                             int* not_integral, uint8_t* iq, uint8_t* it, int32_t* qoff, cudaStream_t st,
                             uint8_t* host_flags) {
  if (n <= 0) return cudaSuccess;
  pack_sift_kernel<<<(n + 7) / 8, 256, 0, st>>>(raw_f32, raw_u8, n, qf, tf, qnorm, raw_out, u8_out,
                                                not_integral, iq, it, qoff, host_flags);
  return cudaGetLastError();
}

// Real-valued rows (SuperPoint: 256 floats of unit norm, FeatureSuperPoint.cpp:195-205; or any float
// descriptor of 128 / 256 elements with moderate magnitudes) -> fp16 operand forms of dim + 16 halfs:
//   query form [ -2*a | 1, 1, 1, 1, 0 x12 ]      train form [ b | p0, p1, p2, 2, 0 x12 ]
// with |b|^2 = p0 + p1 + p2 up to 2^-33 (three fp16 pieces of the fp32 norm) and +2 keeping every score
// positive.  stats[0] = max |x| (float bits), stats[1] = max |row|^2 (float bits), stats[2] = non-finite seen,
// stats[3] = max | |row|^2 - 1 | (float bits).
// The same rows quantised for kind::i8 (l2_tc2.cu KIND 3), dim + 32 bytes per row, when q8 != nullptr:
//   query form [ q_k = rint(254 x_k) (s8) | 1, 255 x31 (u8) ]      train form [ -q_k (s8) | digits of h, base 255 ]
// with h = rint(254^2 (|b|^2 / 2 + 1)) from the fp32 norm.  Valid when max |x| <= 0.5 (stats[0]).
__global__ void __launch_bounds__(256)
pack_float_kernel(const float* __restrict__ raw, int n, int dim, __half* __restrict__ qh, __half* __restrict__ th,
                  float* __restrict__ fnorm, unsigned int* __restrict__ stats, uint8_t* __restrict__ q8,
                  uint8_t* __restrict__ t8) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const int kp = dim + 16;
  const float* src = raw + static_cast<size_t>(row) * dim;
  __half* qrow = qh + static_cast<size_t>(row) * kp;
  __half* trow = th + static_cast<size_t>(row) * kp;
  float nrm = 0.f, amax = 0.f;
  bool bad = false;
  for (int k = 4 * lane; k < dim; k += 128) {
    const float4 f = *reinterpret_cast<const float4*>(src + k);
    const float v[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      bad |= !(fabsf(v[e]) <= 3.0e38f);
      amax = fmaxf(amax, fabsf(v[e]));
      nrm = fmaf(v[e], v[e], nrm);
    }
    reinterpret_cast<__half2*>(qrow + k)[0] = __floats2half2_rn(-2.f * v[0], -2.f * v[1]);
    reinterpret_cast<__half2*>(qrow + k)[1] = __floats2half2_rn(-2.f * v[2], -2.f * v[3]);
    reinterpret_cast<__half2*>(trow + k)[0] = __floats2half2_rn(v[0], v[1]);
    reinterpret_cast<__half2*>(trow + k)[1] = __floats2half2_rn(v[2], v[3]);
    if (q8) {
      uint32_t wq = 0, wt = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int q = max(-127, min(127, __float2int_rn(254.f * v[e])));
        wq |= (static_cast<uint32_t>(q) & 0xFFu) << (8 * e);
        wt |= (static_cast<uint32_t>(-q) & 0xFFu) << (8 * e);
      }
      *reinterpret_cast<uint32_t*>(q8 + static_cast<size_t>(row) * (dim + 32) + k) = wq;
      *reinterpret_cast<uint32_t*>(t8 + static_cast<size_t>(row) * (dim + 32) + k) = wt;
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    nrm += __shfl_xor_sync(0xffffffffu, nrm, off);
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane < 16) {
    float qe = 0.f, te = 0.f;
    const float p0 = __half2float(__float2half_rn(nrm));
    const float r1 = nrm - p0;
    const float p1 = __half2float(__float2half_rn(r1));
    const float p2 = r1 - p1;
    if (lane == 0) { qe = 1.f; te = p0; }
    else if (lane == 1) { qe = 1.f; te = p1; }
    else if (lane == 2) { qe = 1.f; te = p2; }
    else if (lane == 3) { qe = 1.f; te = 2.f; }
    qrow[dim + lane] = __float2half_rn(qe);
    trow[dim + lane] = __float2half_rn(te);
  }
  if (q8) {
    const int h = __float2int_rn(254.f * 254.f * (0.5f * fminf(nrm, 4.f) + 1.f));
    const int qd = h / 255, r = h - 255 * qd;                    // h <= 3 * 254^2: qd <= 759 < 31 * 255
    q8[static_cast<size_t>(row) * (dim + 32) + dim + lane] = lane == 0 ? 1 : 255;
    t8[static_cast<size_t>(row) * (dim + 32) + dim + lane] =
        static_cast<uint8_t>(lane == 0 ? r : min(255, max(0, qd - 255 * (lane - 1))));
  }
  if (lane == 0) {
    fnorm[row] = nrm;
    atomicMax(&stats[0], __float_as_uint(amax));
    atomicMax(&stats[1], __float_as_uint(nrm));
    atomicMax(&stats[3], __float_as_uint(fabsf(nrm - 1.f)));     // how far the image is from unit-norm rows
    if (bad) atomicOr(&stats[2], 1u);
  }
}

cudaError_t launch_pack_float(const float* raw, int n, int dim, __half* qh, __half* th, float* fnorm,
                              unsigned int* stats, uint8_t* q8, uint8_t* t8, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  pack_float_kernel<<<(n + 7) / 8, 256, 0, st>>>(raw, n, dim, qh, th, fnorm, stats, q8, t8);
  return cudaGetLastError();
}

// Binary descriptors (ORB: 32 bytes, FeatureDetector.cpp:9,19) -> E4M3 operand rows of 32*W + 32 bytes for the
// tensor-core Hamming search: Hamming(a, b) = |a| + |b| - 2 a.b, a dense contraction over {0,1} values.
//   query form [ bit ? -2.0 : 0 | 256, 16, 1, 0 x29 ]     train form [ bit ? 1.0 : 0 | p2, p1, p0, 0 x29 ]
// with |b| = 256 p2 + 16 p1 + p0 (p1, p0 <= 15, p2 <= 2: all exactly representable in E4M3), so the fp32
// accumulator holds the exact integer |b| - 2 a.b.  One warp per row; popc[row] = |row|.
__device__ __forceinline__ unsigned long long spread_bits8(unsigned int b) {     // byte i = bit i of b (0 / 1)
  const unsigned long long t = (static_cast<unsigned long long>(b) * 0x0101010101010101ull) & 0x8040201008040201ull;
  return ((t + 0x7F7F7F7F7F7F7F7Full) >> 7) & 0x0101010101010101ull;
}
// With q8 / t8 (256-bit rows only) the same rows are also written as bytes for kind::i8 (l2_i8x2_kernel<.., 2>):
//   query form [ bit (u8 0 / 1) | 1, 255 x31 (u8) ]     train form [ bit ? -1 : +1 (s8) | r, q, 0 x30 (u8) ], |b| = r + 255 q
// so that the s32 accumulator is sum a (1 - 2 b) + |b| = |a| + |b| - 2 a.b, the Hamming distance itself.
// With q4 / t4 (256-bit rows only) also as E2M1 values, two per byte, rows of 160 bytes, for kind::mxf4
// (l2_i8x2_kernel<.., 1, 1>): 256 values = 128 bytes = ONE 128-byte K atom, then a 64-value norm block:
//   query form [ bit ? 1 : 0 | 6 x9, 1 x2, 6 x14, 2, 0 x38 ]      train form [ bit ? -1 : +1 | d_0 .. d_10, 6 x14, 4, 0 x38 ]
// with |b| = 36 n + 6 u + v: d_0..6 = 6 for the first n slots, d_7 + d_8 = u, d_9 + d_10 = v (every digit one of the
// E2M1 values 0, 1, 2, 3, 4, 6), so that the fp32 accumulator is sum a (1 - 2 b) + |b| + 512 = the Hamming distance plus
// the constant of the bias slots 11..25 (what puts the packed-pair kernel's accumulators into one binade, l2_tc2.cu).
// 512-bit rows (|b| <= 512): the same with NS6 = 14 slots of 36 (query weights 6 x16, 1 x2).
template <int NS6 = 7>
__device__ __forceinline__ uint32_t fp4_norm_digit(int slot, int cnt) {       // E2M1 code of train digit `slot`
  const int n36 = min(cnt / 36, NS6), r = cnt - 36 * n36, u = r / 6, v = r - 6 * u;
  int d = 0;
  if (slot < NS6) d = slot < n36 ? 6 : 0;
  else if (slot == NS6) d = min(u, 4);
  else if (slot == NS6 + 1) d = u - min(u, 4);
  else if (slot == NS6 + 2) d = min(v, 4);
  else if (slot == NS6 + 3) d = v - min(v, 4);
  // value -> code: 0 -> 0, 1 -> 2, 2 -> 4, 3 -> 5, 4 -> 6, 6 -> 7
  return d == 0 ? 0u : (d == 1 ? 2u : (d == 2 ? 4u : (d == 3 ? 5u : (d == 4 ? 6u : 7u))));
}
__global__ void __launch_bounds__(256)
pack_bits_kernel(const uint32_t* __restrict__ bits, int n, int words, uint8_t* __restrict__ qb,
                 uint8_t* __restrict__ tb, int32_t* __restrict__ popc, uint8_t* __restrict__ q8,
                 uint8_t* __restrict__ t8, uint8_t* __restrict__ q4, uint8_t* __restrict__ t4,
                 uint8_t* __restrict__ t4x) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const int kp = 32 * words + 32;
  const uint32_t w = lane < words ? bits[static_cast<size_t>(row) * words + lane] : 0u;
  int cnt = __popc(w);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  uint8_t* qrow = qb + static_cast<size_t>(row) * kp;
  uint8_t* trow = tb + static_cast<size_t>(row) * kp;
  // lane l expands byte j of every word it is handed: 8 bits -> 8 operand bytes
  for (int j = lane; j < 4 * words; j += 32) {
    const uint32_t wj = __shfl_sync(0xffffffffu, w, j >> 2);
    const unsigned long long ones = spread_bits8((wj >> (8 * (j & 3))) & 0xFFu);
    *reinterpret_cast<unsigned long long*>(qrow + 8 * j) = ones * 0xC0ull;      // -2.0
    *reinterpret_cast<unsigned long long*>(trow + 8 * j) = ones * 0x38ull;      //  1.0
    if (q8) {
      *reinterpret_cast<unsigned long long*>(q8 + static_cast<size_t>(row) * kp + 8 * j) = ones;
      // bit 1 -> 0xFF (-1), bit 0 -> 0x01 (+1):  0x01 + bit * 0xFE per byte
      *reinterpret_cast<unsigned long long*>(t8 + static_cast<size_t>(row) * kp + 8 * j) = 0x0101010101010101ull + ones * 0xFEull;
    }
  }
  if (q4) {                                              // words == 8 / 16: byte of the row -> 8 four-bit values
    const int row_b = words == 16 ? TC_FP4_ROW2 : TC_FP4_ROW;
    uint8_t* q4row = q4 + static_cast<size_t>(row) * row_b;
    uint8_t* t4row = t4 + static_cast<size_t>(row) * row_b;
    for (int j = lane; j < 4 * words; j += 32) {
      const uint32_t wj = __shfl_sync(0xffffffffu, w, j >> 2);
      const uint32_t byte = (wj >> (8 * (j & 3))) & 0xFFu;
      uint32_t nib = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) nib |= ((byte >> k) & 1u) << (4 * k);
      *reinterpret_cast<uint32_t*>(q4row + 4 * j) = nib * 0x2u;                   // 1.0 = 0b0010
      *reinterpret_cast<uint32_t*>(t4row + 4 * j) = 0x22222222u + nib * 0x8u;     // -1.0 = 0b1010
    }
    // norm block: byte L = slots 2L (low nibble), 2L + 1 (high nibble)
    if (words == 16) {
      q4row[256 + lane] = lane < 8 ? 0x77 : (lane == 8 ? 0x22 : 0x00);
      t4row[256 + lane] = static_cast<uint8_t>(fp4_norm_digit<14>(2 * lane, cnt) | (fp4_norm_digit<14>(2 * lane + 1, cnt) << 4));
    } else {
      // slots 11..25: the BIAS of the packed-pair kernel (l2_i8x2_kernel PK), 14 x 6 * 6 + 2 * 4 = 512 in every product of a
      // query row with a train row; the one-row kernel subtracts it when it writes a distance
      const uint8_t tb = lane == 5 ? 0x70 : (lane >= 6 && lane <= 11 ? 0x77 : (lane == 12 ? 0x67 : 0x00));
      // the query weights once more in the second K block (slots 32..): the packed-pair kernel multiplies them with the
      // norm block of the second train row there (t_ext2x96); every train row keeps zeros in its own second block
      const int l16 = lane & 15;
      const uint8_t qb = l16 == 5 ? 0x72 : (l16 >= 6 && l16 <= 11 ? 0x77 : (l16 == 12 ? 0x47 : 0x00));
      q4row[128 + lane] = l16 < 4 ? 0x77 : (l16 == 4 ? 0x27 : qb);
      t4row[128 + lane] =
          static_cast<uint8_t>(fp4_norm_digit<7>(2 * lane, cnt) | (fp4_norm_digit<7>(2 * lane + 1, cnt) << 4) | tb);
      // t4x (32 bytes per row): the norm block the packed-pair kernel multiplies in ONE K-step for both train rows of an
      // accumulator column -- first K block = this row, second K block = row + 192 of the same image (bias only past its end).
      // Only rows in the first half of a 384-row tile are ever read (l2_i8x2_kernel PK).
      if (t4x && row % 384 < 192) {
        int cnt2 = 0;
        const bool has2 = row + 192 < n;
        if (has2 && lane < 8) cnt2 = __popc(bits[static_cast<size_t>(row + 192) * 8 + lane]);
#pragma unroll
        for (int off = 4; off >= 1; off >>= 1) cnt2 += __shfl_xor_sync(0xffffffffu, cnt2, off);
        cnt2 = __shfl_sync(0xffffffffu, cnt2, 0);
        const int l16 = lane & 15, c = lane < 16 ? cnt : cnt2;
        const uint8_t tb16 = l16 == 5 ? 0x70 : (l16 >= 6 && l16 <= 11 ? 0x77 : (l16 == 12 ? 0x67 : 0x00));
        const uint8_t d = static_cast<uint8_t>(fp4_norm_digit<7>(2 * l16, c) | (fp4_norm_digit<7>(2 * l16 + 1, c) << 4) | tb16);
        // past the end of the image the second block still carries the bias (c = 0 there): whatever row the main K-steps
        // find 192 rows further on, the low field stays within 512 +- 256 and never borrows from the high field
        t4x[static_cast<size_t>(row) * 32 + lane] = d;
      }
    }
  }
  if (q8) {
    q8[static_cast<size_t>(row) * kp + 32 * words + lane] = lane == 0 ? 1 : 255;
    t8[static_cast<size_t>(row) * kp + 32 * words + lane] =
        static_cast<uint8_t>(lane == 0 ? cnt % 255 : (lane == 1 ? cnt / 255 : 0));
  }
  {
    const uint8_t e4m3_int[16] = {0x00, 0x38, 0x40, 0x44, 0x48, 0x4A, 0x4C, 0x4E,
                                  0x50, 0x51, 0x52, 0x53, 0x54, 0x55, 0x56, 0x57};
    uint8_t qe = 0, te = 0;
    if (lane == 0) { qe = 0x78; te = e4m3_int[(cnt >> 8) & 15]; }          // 256
    else if (lane == 1) { qe = 0x58; te = e4m3_int[(cnt >> 4) & 15]; }     // 16
    else if (lane == 2) { qe = 0x38; te = e4m3_int[cnt & 15]; }            // 1
    qrow[32 * words + lane] = qe;
    trow[32 * words + lane] = te;
  }
  if (lane == 0) popc[row] = cnt;
}

cudaError_t launch_pack_bits(const uint32_t* bits, int n, int words, uint8_t* qb, uint8_t* tb, int32_t* popc,
                             uint8_t* q8, uint8_t* t8, uint8_t* q4, uint8_t* t4, cudaStream_t st, uint8_t* t4x) {
  if (n <= 0) return cudaSuccess;
  if (words != 8) q8 = t8 = nullptr;
  if (words != 8 && words != 16) q4 = t4 = nullptr;
  if (words != 8) t4x = nullptr;
  pack_bits_kernel<<<(n + 7) / 8, 256, 0, st>>>(bits, n, words, qb, tb, popc, q8, t8, q4, t4, t4x);
  return cudaGetLastError();
}

__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = static_cast<float>(src[i]);
}

// fp32 rows that hold byte values -> bytes (wire form of the collective ingest); *not_integral is set when a value is not
// an integer in [0, 255]
__global__ void f32_to_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, size_t n, int* __restrict__ not_integral) {
  const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float4 v = *reinterpret_cast<const float4*>(src + i);
  const float f[4] = {v.x, v.y, v.z, v.w};
  uint32_t out = 0;
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float r = rintf(f[k]);
    bad |= !(r == f[k] && r >= 0.f && r <= 255.f);
    out |= (static_cast<uint32_t>(fminf(fmaxf(r, 0.f), 255.f)) & 0xFFu) << (8 * k);
  }
  *reinterpret_cast<uint32_t*>(dst + i) = out;
  if (bad) atomicOr(not_integral, 1);
}
cudaError_t launch_f32_to_u8(const float* src, uint8_t* dst, size_t n, int* not_integral, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  if (n & 3) return cudaErrorInvalidValue;
  f32_to_u8_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, st>>>(src, dst, n, not_integral);
  return cudaGetLastError();
}

cudaError_t launch_u8_to_f32(const uint8_t* src, float* dst, size_t n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  u8_to_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(src, dst, n);
  return cudaGetLastError();
}

}  // namespace pm
