// ransac.cu -- K5/K6: batched fundamental-matrix RANSAC, one CTA per image pair, one warp per
// hypothesis for the inlier count.
//
// Replaces GeometricFilter::estimateFundamental (Mapper/libMapper/GeometricFilter.cpp:39-61),
// i.e. cv::findFundamentalMat(p1, p2, mask) with its defaults (FM_RANSAC, 3 px, 0.99, 1000):
//   * sampler: OpenCV's fixed-seed 64-bit multiply-with-carry stream, duplicate re-draw,
//     last-point collinearity reject -- inherently sequential, one thread replays it;
//   * minimal solver: un-normalised 7-point (7x9 null space -> cubic -> up to 3 models), one
//     half-warp per hypothesis of the round (row-per-lane Householder QR), fp64;
//   * score: symmetric epipolar distance (or Sampson), fp64 rounded to fp32 and compared with
//     (float)(t*t); one warp per hypothesis, lanes stride over the matches;
//   * selection: (iteration, model)-ordered scan, "strictly more inliers replaces", adaptive
//     iteration bound -- identical to the sequential loop because hypotheses at or beyond
//     the shrunken bound are discarded.
// Rounds of 8, 16, then 32 iterations so easy pairs (95 % inliers need ~4) stop early.
//
// The arithmetic order follows the CPU filter of the parity tests operation for operation;
// this translation unit MUST be compiled with -fmad=false (no fp contraction).
#include <cfloat>
#if defined(PM_RANSAC_PROFILE) || defined(PM_RANSAC_PARANOID)
#include <cstdio>
#endif
#ifdef PM_RANSAC_PROFILE      // development build: per-phase cycle counts of the first CTAs (tools/ransac_prof.py)
#define PM_PHASE(acc) do { const long long t1_ = clock64(); acc += t1_ - t0_; t0_ = t1_; } while (0)
#else
#define PM_PHASE(acc) do { } while (0)
#endif

#include "common.cuh"
#include "kernels.h"

namespace pm {

#ifndef PM_RS_THREADS
#define PM_RS_THREADS 128
#endif
static constexpr int RS_THREADS = PM_RS_THREADS;   // 128: 16 K registers per block, co-resident with the persistent tensor kernel
static constexpr int RS_ROUND = 32;
#ifndef PM_RS_PPT
#define PM_RS_PPT 4
#endif
[[maybe_unused]] static constexpr int RS_PPT = PM_RS_PPT;        // match-major build: matches per thread and sweep of the scoring loop

struct MwcRng {
  unsigned long long s;
  __device__ __forceinline__ unsigned int next() {
    s = static_cast<unsigned long long>(static_cast<unsigned int>(s)) * 4164903690ull + (s >> 32);
    return static_cast<unsigned int>(s);
  }
  __device__ __forceinline__ int uniform(int a, int b) {
    return a == b ? a : static_cast<int>(next() % static_cast<unsigned int>(b - a)) + a;
  }
};

__device__ bool last_point_collinear(const float2* p) {
  const int i = 6;
  for (int j = 0; j < i; ++j) {
    const double dx1 = static_cast<double>(__fsub_rn(p[j].x, p[i].x));
    const double dy1 = static_cast<double>(__fsub_rn(p[j].y, p[i].y));
    for (int k = 0; k < j; ++k) {
      const double dx2 = static_cast<double>(__fsub_rn(p[k].x, p[i].x));
      const double dy2 = static_cast<double>(__fsub_rn(p[k].y, p[i].y));
      if (fabs(dx2 * dy1 - dy2 * dx1) <=
          static_cast<double>(FLT_EPSILON) * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
        return true;
    }
  }
  return false;
}

// Draws one 7-subset exactly like the sequential sampler; returns false after max_attempts.
// The RNG stream is consumed index by index (duplicate re-draw), the 14 point loads of a complete
// subset are then issued together -- same stream, same decisions, one memory latency instead of seven.
__device__ bool get_subset(const float2* __restrict__ p1, const float2* __restrict__ p2, int n,
                           MwcRng& rng, int max_attempts, int* idx) {
  int iters = 0, i = 0;
  float2 s1[7], s2[7];
  for (; iters < max_attempts; ++iters) {
    for (i = 0; i < 7 && iters < max_attempts;) {
      const int v = rng.uniform(0, n);
      int j = 0;
      for (; j < i; ++j)
        if (v == idx[j]) break;
      if (j < i) continue;
      idx[i] = v;
      ++i;
    }
    if (i == 7) {
#pragma unroll
      for (int k = 0; k < 7; ++k) { s1[k] = p1[idx[k]]; s2[k] = p2[idx[k]]; }
      if (last_point_collinear(s1) || last_point_collinear(s2)) continue;
    }
    break;
  }
  return i == 7 && iters < max_attempts;
}

__device__ int update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = p < 0 ? 0 : (p > 1 ? 1 : p);
  ep = ep < 0 ? 0 : (ep > 1 ? 1 : ep);
  double num = 1 - p;
  if (num < DBL_MIN) num = DBL_MIN;
  double den = 1 - pow(1 - ep, static_cast<double>(model_points));
  if (den < DBL_MIN) return 0;
  num = log(num);
  den = log(den);
  if (den >= 0 || -num >= max_iters * (-den)) return max_iters;
  return static_cast<int>(llrint(num / den));
}

__device__ int solve_cubic(const double* c, double* x) {
  double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
  int n = 0;
  double x0 = 0, x1 = 0, x2 = 0;
  if (a0 == 0) {
    if (a1 == 0) {
      if (a2 == 0) n = a3 == 0 ? -1 : 0;
      else { x0 = -a3 / a2; n = 1; }
    } else {
      double d = a2 * a2 - 4 * a1 * a3;
      if (d >= 0) {
        d = sqrt(d);
        const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (fabs(q1) > fabs(q2)) { x0 = q1 / a1; x1 = a3 / q1; }
        else { x0 = q2 / a1; x1 = a3 / q2; }
        n = d > 0 ? 2 : 1;
      }
    }
  } else {
    a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    const double Qcubed = Q * Q * Q;
    double d = Qcubed - R * R;
    if (d > 0) {
      const double theta = acos(R / sqrt(Qcubed));
      const double sqrtQ = sqrt(Q);
      const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
      x0 = t0 * cos(t1) - t2;
      x1 = t0 * cos(t1 + (2. * 3.14159265358979323846 / 3)) - t2;
      x2 = t0 * cos(t1 + (4. * 3.14159265358979323846 / 3)) - t2;
      n = 3;
    } else if (d == 0) {
      if (R >= 0) { x0 = -2 * pow(R, 1. / 3) - a1 / 3; x1 = pow(R, 1. / 3) - a1 / 3; }
      else { x0 = 2 * pow(-R, 1. / 3) - a1 / 3; x1 = -pow(-R, 1. / 3) - a1 / 3; }
      n = x0 == x1 ? 1 : 2;
    } else {
      d = sqrt(-d);
      double e = pow(d + fabs(R), 1. / 3);
      if (R > 0) e = -e;
      x0 = (e + Q / e) - a1 * (1. / 3);
      n = 1;
    }
  }
  x[0] = x0; x[1] = x1; x[2] = x2;
  return n;
}

// 7-point solver; m1/m2 are the 7 sampled correspondences.  F receives up to 3 models.
__device__ int seven_point(const float2* m1, const float2* m2, double* F) {
  // B = A^T (9 x 7); Householder vectors are kept in place (column k, rows k..8).
  double B[9][7], vn2[7];
  for (int i = 0; i < 7; ++i) {
    const double x0 = m1[i].x, y0 = m1[i].y, x1 = m2[i].x, y1 = m2[i].y;
    B[0][i] = x1 * x0; B[1][i] = x1 * y0; B[2][i] = x1;
    B[3][i] = y1 * x0; B[4][i] = y1 * y0; B[5][i] = y1;
    B[6][i] = x0;      B[7][i] = y0;      B[8][i] = 1.0;
  }
  for (int k = 0; k < 7; ++k) {
    double nrm2 = 0;
    for (int i = k; i < 9; ++i) nrm2 += B[i][k] * B[i][k];
    const double nrm = sqrt(nrm2);
    if (!(nrm > 0)) return 0;
    const double alpha = B[k][k] > 0 ? -nrm : nrm;
    B[k][k] -= alpha;                                  // column k now holds v_k
    double s2 = 0;
    for (int i = k; i < 9; ++i) s2 += B[i][k] * B[i][k];
    vn2[k] = s2;
    if (!(s2 > 0)) return 0;
    for (int j = k + 1; j < 7; ++j) {
      double s = 0;
      for (int i = k; i < 9; ++i) s += B[i][k] * B[i][j];
      const double f = 2 * s / s2;
      for (int i = k; i < 9; ++i) B[i][j] -= f * B[i][k];
    }
  }
  double f1[9], f2[9];
  for (int e = 0; e < 2; ++e) {
    double y[9];
    for (int i = 0; i < 9; ++i) y[i] = 0.0;
    y[7 + e] = 1.0;
    for (int k = 6; k >= 0; --k) {
      double s = 0;
      for (int i = k; i < 9; ++i) s += B[i][k] * y[i];
      const double f = 2 * s / vn2[k];
      for (int i = k; i < 9; ++i) y[i] -= f * B[i][k];
    }
    for (int i = 0; i < 9; ++i) (e == 0 ? f1 : f2)[i] = y[i];
  }
  for (int i = 0; i < 9; ++i) f1[i] -= f2[i];

  double c[4], r[3];
  double t0 = f2[4] * f2[8] - f2[5] * f2[7];
  double t1 = f2[3] * f2[8] - f2[5] * f2[6];
  double t2 = f2[3] * f2[7] - f2[4] * f2[6];
  c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
  c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 -
         f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
         f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
         f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
         f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) -
         f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
         f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
  t0 = f1[4] * f1[8] - f1[5] * f1[7];
  t1 = f1[3] * f1[8] - f1[5] * f1[6];
  t2 = f1[3] * f1[7] - f1[4] * f1[6];
  c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 -
         f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
         f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
         f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
         f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) -
         f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
         f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
  c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;

  const int n = solve_cubic(c, r);
  if (n < 1 || n > 3) return 0;
  for (int k = 0; k < n; ++k) {
    double lambda = r[k], mu = 1.0;
    const double s = f1[8] * r[k] + f2[8];
    double* Fk = F + 9 * k;
    if (fabs(s) > DBL_EPSILON) { mu = 1. / s; lambda *= mu; Fk[8] = 1.0; }
    else Fk[8] = 0.0;
    for (int i = 0; i < 8; ++i) Fk[i] = f1[i] * lambda + f2[i] * mu;
  }
  return n;
}

// ---- cooperative 7-point solver: one 16-lane half-warp per hypothesis ---------------------------
// Lane i (0..8) holds row i of B = A^T in registers (static column indices after unrolling), so the
// Householder sweeps touch no local memory.  Every inner product is still accumulated in the order
// i = k..8 (the partial products are exchanged with shuffles and added by every lane in that order),
// so all values are bit-identical to the scalar seven_point() above and to the CPU filter.
template <int K>
__device__ __forceinline__ double seq_sum(double p, unsigned mask) {
  double s = __shfl_sync(mask, p, K, 16);
#pragma unroll
  for (int i = K + 1; i < 9; ++i) s += __shfl_sync(mask, p, i, 16);
  return s;
}

template <int K>
__device__ __forceinline__ void qr_step(double (&B)[7], double (&vn2)[7], int li, unsigned mask, bool& ok) {
  const double x = B[K];
  const double nrm2 = seq_sum<K>(x * x, mask);
  const double nrm = sqrt(nrm2);
  ok = ok && (nrm > 0);
  const double bkk = __shfl_sync(mask, x, K, 16);
  const double alpha = bkk > 0 ? -nrm : nrm;
  if (li == K) B[K] -= alpha;                              // column K now holds v_K in rows K..8
  const double s2 = seq_sum<K>(B[K] * B[K], mask);
  vn2[K] = s2;
  ok = ok && (s2 > 0);
#pragma unroll
  for (int j = K + 1; j < 7; ++j) {
    const double sd = seq_sum<K>(B[K] * B[j], mask);
    const double f = 2 * sd / s2;
    if (li >= K) B[j] -= f * B[K];
  }
}

template <int K>
__device__ __forceinline__ void back_step(const double (&B)[7], const double (&vn2)[7], double& y, int li,
                                          unsigned mask) {
  const double sd = seq_sum<K>(B[K] * y, mask);
  const double f = 2 * sd / vn2[K];
  if (li >= K) y -= f * B[K];
}

// idx: the 7 sampled match indices (shared memory); F receives up to 3 models; returns their number
// (valid in every lane of the half-warp).
__device__ int seven_point_coop(const float2* __restrict__ p1, const float2* __restrict__ p2, const int* idx,
                                double* F, int li, unsigned mask) {
  double B[7], vn2[7];
#pragma unroll
  for (int c = 0; c < 7; ++c) {
    const float2 a = p1[idx[c]], b = p2[idx[c]];
    const double x0 = a.x, y0 = a.y, x1 = b.x, y1 = b.y;
    const int r3 = li / 3, c3 = li - 3 * r3;
    const double fa = r3 == 0 ? x1 : (r3 == 1 ? y1 : 1.0);
    const double fb = c3 == 0 ? x0 : (c3 == 1 ? y0 : 1.0);
    B[c] = li < 9 ? fa * fb : 0.0;                         // x*1.0 == x: same values as the scalar rows
  }
  bool ok = true;
  qr_step<0>(B, vn2, li, mask, ok); qr_step<1>(B, vn2, li, mask, ok); qr_step<2>(B, vn2, li, mask, ok);
  qr_step<3>(B, vn2, li, mask, ok); qr_step<4>(B, vn2, li, mask, ok); qr_step<5>(B, vn2, li, mask, ok);
  qr_step<6>(B, vn2, li, mask, ok);
  double y1v = li == 7 ? 1.0 : 0.0, y2v = li == 8 ? 1.0 : 0.0;
  back_step<6>(B, vn2, y1v, li, mask); back_step<5>(B, vn2, y1v, li, mask); back_step<4>(B, vn2, y1v, li, mask);
  back_step<3>(B, vn2, y1v, li, mask); back_step<2>(B, vn2, y1v, li, mask); back_step<1>(B, vn2, y1v, li, mask);
  back_step<0>(B, vn2, y1v, li, mask);
  back_step<6>(B, vn2, y2v, li, mask); back_step<5>(B, vn2, y2v, li, mask); back_step<4>(B, vn2, y2v, li, mask);
  back_step<3>(B, vn2, y2v, li, mask); back_step<2>(B, vn2, y2v, li, mask); back_step<1>(B, vn2, y2v, li, mask);
  back_step<0>(B, vn2, y2v, li, mask);
  y1v -= y2v;                                              // f1 -= f2
  double f1[9], f2[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { f1[i] = __shfl_sync(mask, y1v, i, 16); f2[i] = __shfl_sync(mask, y2v, i, 16); }
  if (!ok) return 0;                                       // uniform within the half-warp

  double c[4], r[3];
  double t0 = f2[4] * f2[8] - f2[5] * f2[7];
  double t1 = f2[3] * f2[8] - f2[5] * f2[6];
  double t2 = f2[3] * f2[7] - f2[4] * f2[6];
  c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
  c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 -
         f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
         f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
         f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
         f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) -
         f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
         f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
  t0 = f1[4] * f1[8] - f1[5] * f1[7];
  t1 = f1[3] * f1[8] - f1[5] * f1[6];
  t2 = f1[3] * f1[7] - f1[4] * f1[6];
  c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 -
         f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
         f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
         f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
         f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) -
         f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
         f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
  c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
  const int n = solve_cubic(c, r);
  if (n < 1 || n > 3) return 0;
  if (li == 0) {
    for (int k = 0; k < n; ++k) {
      double lambda = r[k], mu = 1.0;
      const double sc = f1[8] * r[k] + f2[8];
      double* Fk = F + 9 * k;
      if (fabs(sc) > DBL_EPSILON) { mu = 1. / sc; lambda *= mu; Fk[8] = 1.0; }
      else Fk[8] = 0.0;
      for (int i = 0; i < 8; ++i) Fk[i] = f1[i] * lambda + f2[i] * mu;
    }
  }
  return n;
}

// Inlier test of cv::findFundamentalMat's RANSAC: (float)err <= (float)(thr * thr), err computed in fp64
// without contraction (computeError) -- literal_inlier() below, the arbiter.
//
// Almost every point is far from the threshold, so the hot loop uses a CONSERVATIVE classifier first:
// n = d^2 and g = a^2 + b^2 evaluated with fused multiply-adds and no division;
//     n <  thr*(1-m)*g for both terms          -> inlier for sure
//     n >= thr*(1+2.5e-7)*(1+m)*g for either   -> outlier for sure          (1 / 0 / 2 = undecided)
// The margin m = 1e-9 + 4e-14 * max|coordinate| covers (a) the float rounding of err (< 1.2e-7 relative,
// the 2.5e-7 factor), (b) the fp64 rounding of n * (1/g) (< 4e-16), (c) the difference between the fused
// and the literal evaluation of d: a few ulps of the largest term, amplified by the cancellation
// |x a| + |y b| + |c| <= ~2 (|x|+|y|) sqrt(g) at the decision boundary |d| = 3 sqrt(g).  Undecided points
// (~1e-7 of all) go through literal_inlier(); the outcome is therefore identical by construction.
// g = 0 / inf / NaN need no special case: the strict '<' fails for g = 0 and every comparison with NaN is
// false, which lands in the literal formula's own behaviour (NaN compares false = outlier).
struct Pt4 { double x1, y1, x2, y2; };
template <int MODE>
__device__ __forceinline__ int classify(const double* __restrict__ F, const Pt4& p, double lo, double hi) {
  const double a2 = fma(F[0], p.x1, fma(F[1], p.y1, F[2]));
  const double b2 = fma(F[3], p.x1, fma(F[4], p.y1, F[5]));
  const double c2 = fma(F[6], p.x1, fma(F[7], p.y1, F[8]));
  const double g2 = fma(a2, a2, b2 * b2);
  const double d2 = fma(p.x2, a2, fma(p.y2, b2, c2));
  const double a1 = fma(F[0], p.x2, fma(F[3], p.y2, F[6]));
  const double b1 = fma(F[1], p.x2, fma(F[4], p.y2, F[7]));
  const double g1 = fma(a1, a1, b1 * b1);
  const double n2 = d2 * d2;
  if (MODE == 1) {
    const double g = g1 + g2;
    return n2 < lo * g ? 1 : (n2 >= hi * g ? 0 : 2);
  }
  const double c1 = fma(F[2], p.x2, fma(F[5], p.y2, F[8]));
  const double d1 = fma(p.x1, a1, fma(p.y1, b1, c1));
  const double n1 = d1 * d1;
  const bool in = n1 < lo * g1 && n2 < lo * g2;
  const bool out = n1 >= hi * g1 || n2 >= hi * g2;
  return in ? 1 : (out ? 0 : 2);
}
__device__ __noinline__ int literal_inlier(const double* F, float2 q1, float2 q2, int mode, float thr) {
  const double x1 = q1.x, y1 = q1.y, x2 = q2.x, y2 = q2.y;
  double a = F[0] * x1 + F[1] * y1 + F[2];
  double b = F[3] * x1 + F[4] * y1 + F[5];
  double c = F[6] * x1 + F[7] * y1 + F[8];
  const double g2 = a * a + b * b;
  const double d2 = x2 * a + y2 * b + c;
  a = F[0] * x2 + F[3] * y2 + F[6];
  b = F[1] * x2 + F[4] * y2 + F[7];
  c = F[2] * x2 + F[5] * y2 + F[8];
  const double g1 = a * a + b * b;
  const double d1 = x1 * a + y1 * b + c;
  if (mode == 1) return static_cast<float>(d2 * d2 / (g1 + g2)) <= thr ? 1 : 0;
  const double s2 = 1. / g2, s1 = 1. / g1;
  const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
  return static_cast<float>(e1 > e2 ? e1 : e2) <= thr ? 1 : 0;
}
__device__ __forceinline__ int inlier_of(const double* F, float2 q1, float2 q2, int mode, float thr, double lo,
                                         double hi) {
  const Pt4 p{q1.x, q1.y, q2.x, q2.y};
  const int c = mode == 1 ? classify<1>(F, p, lo, hi) : classify<0>(F, p, lo, hi);
  return c == 2 ? literal_inlier(F, q1, q2, mode, thr) : c;
}

// ---- (optional, -DPM_RS_FP32=1) the hot classifier in fp32, certified by A-PRIORI per-model error bounds ----------------
// The fp64 pipe is the bottleneck of outlier-heavy pairs (ncu: half of the kernel's instructions are DFMA / DMUL /
// DSETP, B200 issues them at half the fp32 rate).  Pixel coordinates are exact in fp32, so only F is rounded; with
// u = 2^-24, X = the largest |coordinate| of the pair and |F| the magnitudes, every fused fp32 evaluation obeys
//     |a^ - a| <= 4u A,  A = (|F0| + |F1|) X + |F2|   (likewise B, C, and A1, B1 for the transposed products)
//     |d^ - d| <= dd  = 16u (X (A + B) + C)            d = x2' F x1 -- ONE numerator: both point-to-line distances share it
//     |g^ - g| <= dg  = 24u (A^2 + B^2)                 g = a^2 + b^2, per image
// (derivation: forward error of the fma chains with factor-2 slack; the literal fp64 formula differs from the exact
// values by ~1e-13 of the same sums, inside that slack).  These are constants of the MODEL, computed once per
// hypothesis (model_f32), not running bounds per point.  A point is
//     an inlier for sure   if (|d^| + dd)^2 <  thr (1 - 4e-6) (g^ - dg) for both images
//     an outlier for sure  if |d^| > 2 dd (no cancellation in the next term) and (|d^| - dd)^2 >= thr (1 + 4e-6) (g^ + dg)
//                          for either image
// (the 4e-6 covers the float rounding of err and of these comparisons), otherwise it goes through literal_inlier(),
// the arbiter: undecided points are those within ~1e-3 of the threshold, ~1e-4 of all.  NaN / inf / g = 0 fail both
// tests and land in the literal formula's own behaviour.
#ifdef PM_RANSAC_PROFILE
__device__ unsigned long long g_prof_iters = 0, g_prof_lit_warps = 0, g_prof_lit_lanes = 0;
#endif
struct ModelF32 {
  float F[9];
  float dd2, dd_sq, dg1t_lo, dg2t_lo, dg1t_hi, dg2t_hi;          // 2 dd, dd^2, thr_lo dg_k, thr_hi dg_k
};
__device__ __forceinline__ void model_f32(const double* F, float X, float thr, ModelF32& o) {
  const float U = 5.9604645e-8f, UP = 1.0001f;
  float f[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { o.F[i] = static_cast<float>(F[i]); f[i] = fabsf(o.F[i]) * UP; }
  const float A2 = (f[0] + f[1]) * X + f[2], B2 = (f[3] + f[4]) * X + f[5], C2 = (f[6] + f[7]) * X + f[8];
  const float A1 = (f[0] + f[3]) * X + f[6], B1 = (f[1] + f[4]) * X + f[7];
  const float dd = 16.f * U * UP * (X * (A2 + B2) + C2);
  const float dg2 = 24.f * U * UP * (A2 * A2 + B2 * B2), dg1 = 24.f * U * UP * (A1 * A1 + B1 * B1);
  const float tlo = thr * (1.f - 4e-6f), thi = thr * (1.f + 4e-6f);
  o.dd2 = 2.f * dd; o.dd_sq = dd * dd * UP;
  o.dg1t_lo = tlo * dg1 * UP; o.dg2t_lo = tlo * dg2 * UP; o.dg1t_hi = thi * dg1 * UP; o.dg2t_hi = thi * dg2 * UP;
}
template <int MODE>
__device__ __forceinline__ int classify32(const ModelF32& M, float x1, float y1, float x2, float y2, float tlo, float thi) {
  const float a2 = fmaf(M.F[0], x1, fmaf(M.F[1], y1, M.F[2]));
  const float b2 = fmaf(M.F[3], x1, fmaf(M.F[4], y1, M.F[5]));
  const float c2 = fmaf(M.F[6], x1, fmaf(M.F[7], y1, M.F[8]));
  const float g2 = fmaf(a2, a2, b2 * b2);
  const float d = fabsf(fmaf(x2, a2, fmaf(y2, b2, c2)));
  const float a1 = fmaf(M.F[0], x2, fmaf(M.F[3], y2, M.F[6]));
  const float b1 = fmaf(M.F[1], x2, fmaf(M.F[4], y2, M.F[7]));
  const float g1 = fmaf(a1, a1, b1 * b1);
  const float up = fmaf(d, d + M.dd2, M.dd_sq);                    // (|d| + dd)^2, rounded; slack in tlo
  const float dn = fmaf(d, d - M.dd2, M.dd_sq);                    // (|d| - dd)^2
  if (MODE == 1) {
    const float g = g1 + g2;
    const bool in = up < fmaf(tlo, g, -(M.dg1t_lo + M.dg2t_lo));
    const bool out = d > M.dd2 && dn >= fmaf(thi, g, M.dg1t_hi + M.dg2t_hi);
    return in ? 1 : (out ? 0 : 2);
  }
  const bool in = up < fmaf(tlo, g1, -M.dg1t_lo) && up < fmaf(tlo, g2, -M.dg2t_lo);
  const bool out = d > M.dd2 && (dn >= fmaf(thi, g1, M.dg1t_hi) || dn >= fmaf(thi, g2, M.dg2t_hi));
  return in ? 1 : (out ? 0 : 2);
}

// Symmetric-epipolar mode, two steps (rs_score_kernel): a point is an outlier as soon as ONE of its two distances is, and
// for a wrong model almost every point is far from its line in image 2.  Step 1 evaluates that side only (line F x1,
// numerator d, g2) and its certain-outlier test -- the same inequality as in classify32<0>; only when some lane of the
// warp cannot be dismissed does step 2 add the other side and the remaining tests.  Same decisions as classify32<0>.
struct Side2 { float d, g2, dn; bool out; };
__device__ __forceinline__ Side2 classify32_side2(const ModelF32& M, float x1, float y1, float x2, float y2, float thi) {
  Side2 r;
  const float a2 = fmaf(M.F[0], x1, fmaf(M.F[1], y1, M.F[2]));
  const float b2 = fmaf(M.F[3], x1, fmaf(M.F[4], y1, M.F[5]));
  const float c2 = fmaf(M.F[6], x1, fmaf(M.F[7], y1, M.F[8]));
  r.g2 = fmaf(a2, a2, b2 * b2);
  r.d = fabsf(fmaf(x2, a2, fmaf(y2, b2, c2)));
  r.dn = fmaf(r.d, r.d - M.dd2, M.dd_sq);                          // (|d| - dd)^2
  r.out = r.d > M.dd2 && r.dn >= fmaf(thi, r.g2, M.dg2t_hi);
  return r;
}
__device__ __forceinline__ int classify32_rest(const ModelF32& M, const Side2& s, float x2, float y2, float tlo, float thi) {
  const float a1 = fmaf(M.F[0], x2, fmaf(M.F[3], y2, M.F[6]));
  const float b1 = fmaf(M.F[1], x2, fmaf(M.F[4], y2, M.F[7]));
  const float g1 = fmaf(a1, a1, b1 * b1);
  const float up = fmaf(s.d, s.d + M.dd2, M.dd_sq);                // (|d| + dd)^2
  const bool in = up < fmaf(tlo, g1, -M.dg1t_lo) && up < fmaf(tlo, s.g2, -M.dg2t_lo);
  const bool out = s.out || (s.d > M.dd2 && s.dn >= fmaf(thi, g1, M.dg1t_hi));
  return in ? 1 : (out ? 0 : 2);
}

// Adds, for every model of the round, the number of inliers among this thread's PPT points to its warp's slots sCntW.
// PM_RS_FP32 = 1 selects the fp32 classifier above.  Measured (B200, 100 x 8192 SIFT, half of the keypoints displaced: 1000
// iterations per pair): 37.9 k pairs/s against 36.9 k with the fp64 classifier, 0 mismatches in the paranoid build --
// the scoring loop is bound by the latency of its ~260-instruction body with four resident warps per scheduler, not by
// the fp64 pipe, so the default stays the fp64 classifier that two rounds of parity runs have exercised.
#ifndef PM_RS_FP32
#define PM_RS_FP32 0
#endif
template <int MODE, int PPT>
__device__ __forceinline__ void score_models(const double (*sF)[27], const ModelF32 (*sMf)[3], const int* sNm, int (*sCntW)[3],
                                             const unsigned* sDead, int gen, const float2 (&q1)[PPT],
                                             const float2 (&q2)[PPT], const bool (&valid)[PPT], float thr, double lo,
                                             double hi, int lane) {
  const float tlo = thr * (1.f - 4e-6f), thi = thr * (1.f + 4e-6f);
  for (int k = 0; k < gen; ++k) {
    const int nm = sNm[k];
    const unsigned dead = sDead[k];
    for (int m = 0; m < nm; ++m) {
      if ((dead >> m) & 1u) continue;                  // cannot beat the best model any more (see the caller)
      int cls[PPT];
#if PM_RS_FP32
      const ModelF32& Mf = sMf[k][m];
#pragma unroll
      for (int u = 0; u < PPT; ++u) cls[u] = classify32<MODE>(Mf, q1[u].x, q1[u].y, q2[u].x, q2[u].y, tlo, thi);
#else
      double F[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) F[i] = sF[k][9 * m + i];
#pragma unroll
      for (int u = 0; u < PPT; ++u) cls[u] = classify<MODE>(F, Pt4{q1[u].x, q1[u].y, q2[u].x, q2[u].y}, lo, hi);
      (void)sMf; (void)tlo; (void)thi;
#endif
#ifdef PM_RANSAC_PARANOID      // development build: every decided point is re-checked against the literal formula
#pragma unroll
      for (int u = 0; u < PPT; ++u)
        if (valid[u] && cls[u] != 2 && literal_inlier(&sF[k][9 * m], q1[u], q2[u], MODE, thr) != cls[u])
          printf("RANSAC PARANOID MISMATCH k %d m %d cls %d pt (%g,%g)-(%g,%g)\n", k, m, cls[u], q1[u].x, q1[u].y,
                 q2[u].x, q2[u].y);
      if (lane == 0 && k == 0 && m == 0 && blockIdx.x == 0 && threadIdx.x == 0) printf("paranoid build active\n");
#endif
      int good = 0;
#ifdef PM_RANSAC_PROFILE
      {
        int und = 0;
#pragma unroll
        for (int u = 0; u < PPT; ++u) und += (cls[u] == 2 && valid[u]) ? 1 : 0;
        const int tot_und = __reduce_add_sync(0xffffffffu, und);
        if (lane == 0) { atomicAdd(&g_prof_iters, 1ull); if (tot_und) { atomicAdd(&g_prof_lit_warps, 1ull); atomicAdd(&g_prof_lit_lanes, static_cast<unsigned long long>(tot_und)); } }
      }
#endif
#pragma unroll
      for (int u = 0; u < PPT; ++u) {
#if !defined(PM_RS_EXPERIMENT) || PM_RS_EXPERIMENT != 1
        if (cls[u] == 2 && valid[u]) cls[u] = literal_inlier(&sF[k][9 * m], q1[u], q2[u], MODE, thr);
#else
        if (cls[u] == 2) cls[u] = 0;      // timing experiment only: no literal path
#endif
        good += valid[u] ? cls[u] : 0;
      }
      // per-warp slot, plain read-modify-write by lane 0: no atomic, no vote (the tail of this loop body is a chain of
      // dependent warp-wide operations that one resident warp per scheduler cannot hide)
      const int total = __reduce_add_sync(0xffffffffu, good);
      if (lane == 0) sCntW[k][m] += total;
    }
  }
}

// ---- model-major scoring (PM_RS_MM = 1, the default) --------------------------------------------------------------
// Lane = MODEL (its fp32 entries and certified margins stay in registers), the matches of a staged chunk are
// shared-memory broadcasts: no per-model reduction, no per-model shared-memory traffic, ~30 fp32 instructions per
// (model, match) instead of ~65 half-rate fp64 ones.  The classifier is classify32() above (a-priori per-model error
// bounds); a point it cannot certify goes through literal_inlier(), the arbiter, so every count is the literal one.
// A chunk of RS_CHUNK matches is split over the warps of the block; every warp visits all model groups (32 models each).
// Early drop as before: a model whose count so far plus all matches not yet visited cannot exceed the best count from
// before the round is skipped once all models of its group are in that state (its partial count stays <= the bound,
// so the selection ignores it exactly as it would ignore the full count).
#ifndef PM_RS_MM
#define PM_RS_MM 1
#endif
static constexpr int RS_CHUNK = 256;
template <int MODE>
__device__ __forceinline__ void score_chunk_mm(const double (*sF)[27], const ModelF32* sMf, const int* sList, int nmod,
                                               int (*sCnt)[3], const float4* sPts, int j0, int j1, int bound, int rest,
                                               bool first, float thr, int lane) {
  const float tlo = thr * (1.f - 4e-6f), thi = thr * (1.f + 4e-6f);
  for (int g0 = 0; g0 < nmod; g0 += 32) {
    const int t = g0 + lane;
    const bool act = t < nmod;
    const int w = act ? sList[t] : 0;
    const int k = w / 3, m = w - 3 * k;
    const bool dead = !act || (!first && sCnt[k][m] + rest <= bound);
    if (__all_sync(0xffffffffu, dead)) continue;
    const ModelF32 Mf = sMf[act ? t : 0];
    int cnt = 0;
    int j = j0;
    for (; j + 4 <= j1; j += 4) {
      int cls[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 q = sPts[j + u];
        cls[u] = classify32<MODE>(Mf, q.x, q.y, q.z, q.w, tlo, thi);
      }
#ifdef PM_RANSAC_PARANOID      // development build: every decided point is re-checked against the literal formula
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 q = sPts[j + u];
        if (act && cls[u] != 2 && literal_inlier(&sF[k][9 * m], make_float2(q.x, q.y), make_float2(q.z, q.w), MODE, thr) != cls[u])
          printf("RANSAC PARANOID MISMATCH (model-major) k %d m %d cls %d pt (%g,%g)-(%g,%g)\n", k, m, cls[u], q.x, q.y, q.z, q.w);
      }
      if (blockIdx.x == 0 && threadIdx.x == 0 && j == 0 && g0 == 0 && first) printf("paranoid build active (model-major)\n");
#endif
#ifdef PM_RANSAC_PROFILE
      {
        int und = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) und += (cls[u] == 2 && act) ? 1 : 0;
        const int tot_und = __reduce_add_sync(0xffffffffu, und);
        if (lane == 0) { atomicAdd(&g_prof_iters, 4ull); if (tot_und) { atomicAdd(&g_prof_lit_warps, 1ull); atomicAdd(&g_prof_lit_lanes, static_cast<unsigned long long>(tot_und)); } }
      }
#endif
      if (((cls[0] | cls[1] | cls[2] | cls[3]) & 2) && act) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (cls[u] == 2) {
            const float4 q = sPts[j + u];
            cls[u] = literal_inlier(&sF[k][9 * m], make_float2(q.x, q.y), make_float2(q.z, q.w), MODE, thr);
          }
      }
      cnt += (cls[0] & 1) + (cls[1] & 1) + (cls[2] & 1) + (cls[3] & 1);
    }
    for (; j < j1; ++j) {
      const float4 q = sPts[j];
      int c = classify32<MODE>(Mf, q.x, q.y, q.z, q.w, tlo, thi);
      if (c == 2 && act) c = literal_inlier(&sF[k][9 * m], make_float2(q.x, q.y), make_float2(q.z, q.w), MODE, thr);
      cnt += c & 1;
    }
    if (act && cnt) atomicAdd(&sCnt[k][m], cnt);
  }
}

// Index subset of one iteration WITHOUT the collinearity test (checked in parallel afterwards): the
// sequential stream with duplicate re-draw.  uniform(0, n) = next() % n; the modulo is taken with a
// precomputed reciprocal (inv = floor((2^32 - 1) / n): the quotient estimate is at most 2 short).
__device__ __forceinline__ void draw7(MwcRng& rng, unsigned int n, unsigned int inv, int* idx) {
  int v[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    while (true) {
      const unsigned int raw = rng.next();
      unsigned int r = raw - __umulhi(raw, inv) * n;
      while (r >= n) r -= n;
      bool dup = false;
#pragma unroll
      for (int j = 0; j < i; ++j) dup |= v[j] == static_cast<int>(r);
      if (!dup) { v[i] = static_cast<int>(r); break; }
    }
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) idx[i] = v[i];
}

// (j, k) pair number t of the 15 pairs k < j < 6 tested against the last point (checkSubset)
__device__ __forceinline__ bool pair_collinear(const float2* __restrict__ p, const int* idx, int t) {
  int j = 1, k = t;
  while (k >= j) { k -= j; ++j; }
  const float2 pi = p[idx[6]], pj = p[idx[j]], pk = p[idx[k]];
  const double dx1 = static_cast<double>(__fsub_rn(pj.x, pi.x));
  const double dy1 = static_cast<double>(__fsub_rn(pj.y, pi.y));
  const double dx2 = static_cast<double>(__fsub_rn(pk.x, pi.x));
  const double dy2 = static_cast<double>(__fsub_rn(pk.y, pi.y));
  return fabs(dx2 * dy1 - dy2 * dx1) <=
         static_cast<double>(FLT_EPSILON) * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2));
}

// ---- PM_SAMPLER_PHILOX: the subset of iteration k is a pure function of (pair key, k) ---------------------------
// Philox4x32-10 with counter (k, block, 0, 'PMRF'); word w -> index floor(w * n / 2^32); slots filled in order,
// duplicates skipped, a complete subset checked like OpenCV's checkSubset (last point, both images), on rejection
// the filling restarts with the next words.  The CPU filter of the parity tests follows the same rules.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t lo0 = 0xD2511F53u * c[0], hi0 = __umulhi(0xD2511F53u, c[0]);
    const uint32_t lo1 = 0xCD9E8D57u * c[2], hi1 = __umulhi(0xCD9E8D57u, c[2]);
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ bool philox_subset(const float2* __restrict__ p1, const float2* __restrict__ p2, int n,
                              unsigned long long key, int iter, int* idx) {
  int filled = 0;
  int v7[7];
  float2 s1[7], s2[7];
  for (uint32_t blk = 0; blk < 256; ++blk) {
    uint32_t c[4] = {static_cast<uint32_t>(iter), blk, 0u, 0x504D5246u};
    philox4x32_10(c, static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32));
    for (int w = 0; w < 4; ++w) {
      const int v = static_cast<int>(__umulhi(c[w], static_cast<uint32_t>(n)));
      bool dup = false;
      for (int j = 0; j < filled; ++j) dup |= v7[j] == v;
      if (dup) continue;
      v7[filled] = v;
      s1[filled] = p1[v]; s2[filled] = p2[v];
      if (++filled == 7) {
        if (!last_point_collinear(s1) && !last_point_collinear(s2)) {
          for (int j = 0; j < 7; ++j) idx[j] = v7[j];
          return true;
        }
        filled = 0;
      }
    }
  }
  return false;
}

// ---- optional refit: normalised 8-point over the inliers of the winner (cv::findFundamentalMat FM_8POINT) --------
// Cyclic Jacobi, fixed sweep order (the CPU filter of the parity tests follows it operation for operation).
__device__ void jacobi_sym(double* A, double* V, int n, int sweeps) {
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = i == j ? 1.0 : 0.0;
  for (int s = 0; s < sweeps; ++s) {
    double off = 0;
    for (int p = 0; p < n; ++p) for (int q = p + 1; q < n; ++q) off += A[p * n + q] * A[p * n + q];
    if (off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - sn * akq;
          A[k * n + q] = sn * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - sn * aqk;
          A[q * n + k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - sn * vkq;
          V[k * n + q] = sn * vkp + c * vkq;
        }
      }
  }
}
__device__ __forceinline__ double block_sum(double v, double* red, int tid) {     // red: [RS_THREADS / 32 + 1]
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    double s = 0;
    for (int w = 0; w < RS_THREADS / 32; ++w) s += red[w];
    red[RS_THREADS / 32] = s;
  }
  __syncthreads();
  return red[RS_THREADS / 32];
}
// Whole block; msk selects the points.  Returns true and writes F (shared or global, 9 doubles) on success.
__device__ bool eight_point_refit(const float2* __restrict__ p1, const float2* __restrict__ p2, const uint8_t* msk,
                                  int M, double* Fout, double* sA /* [81 + 81 + 8] shared */, int tid) {
  double* red = sA + 162;
  double cnt_d = 0, sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0;
  for (int i = tid; i < M; i += RS_THREADS)
    if (msk[i]) { cnt_d += 1.0; sx1 += p1[i].x; sy1 += p1[i].y; sx2 += p2[i].x; sy2 += p2[i].y; }
  const double cnt = block_sum(cnt_d, red, tid);
  if (cnt < 8.0) return false;
  const double t = 1.0 / cnt;
  const double cx1 = block_sum(sx1, red, tid) * t, cy1 = block_sum(sy1, red, tid) * t;
  const double cx2 = block_sum(sx2, red, tid) * t, cy2 = block_sum(sy2, red, tid) * t;
  double d1 = 0, d2 = 0;
  for (int i = tid; i < M; i += RS_THREADS)
    if (msk[i]) {
      const double x1 = p1[i].x - cx1, y1 = p1[i].y - cy1, x2 = p2[i].x - cx2, y2 = p2[i].y - cy2;
      d1 += sqrt(x1 * x1 + y1 * y1);
      d2 += sqrt(x2 * x2 + y2 * y2);
    }
  double sc1 = block_sum(d1, red, tid) * t, sc2 = block_sum(d2, red, tid) * t;
  if (sc1 < static_cast<double>(FLT_EPSILON) || sc2 < static_cast<double>(FLT_EPSILON)) return false;
  sc1 = sqrt(2.0) / sc1; sc2 = sqrt(2.0) / sc2;
  // upper triangle of the 9 x 9 normal matrix, 45 entries per thread, block-reduced entry by entry
  double a[45];
#pragma unroll
  for (int e = 0; e < 45; ++e) a[e] = 0.0;
  for (int i = tid; i < M; i += RS_THREADS)
    if (msk[i]) {
      const double x1 = (p1[i].x - cx1) * sc1, y1 = (p1[i].y - cy1) * sc1;
      const double x2 = (p2[i].x - cx2) * sc2, y2 = (p2[i].y - cy2) * sc2;
      const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
      int e = 0;
#pragma unroll
      for (int j = 0; j < 9; ++j)
#pragma unroll
        for (int k = j; k < 9; ++k) a[e++] += r[j] * r[k];
    }
  {
    int e = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j)
#pragma unroll
      for (int k = j; k < 9; ++k) {
        const double v = block_sum(a[e++], red, tid);
        if (tid == 0) { sA[j * 9 + k] = v; sA[k * 9 + j] = v; }
      }
  }
  __syncthreads();
  __shared__ int sOk;
  if (tid == 0) {
    double* A = sA;
    double* V = sA + 81;
    jacobi_sym(A, V, 9, 60);
    int lo = 0, nz = 0;
    for (int i = 0; i < 9; ++i) {
      if (A[i * 9 + i] < A[lo * 9 + lo]) lo = i;
      if (fabs(A[i * 9 + i]) < DBL_EPSILON) ++nz;
    }
    sOk = nz > 1 ? 0 : 1;
    if (sOk) {
      double F0[9], G[9], W[9];
      for (int i = 0; i < 9; ++i) F0[i] = V[i * 9 + lo];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double sacc = 0;
        for (int k = 0; k < 3; ++k) sacc += F0[k * 3 + i] * F0[k * 3 + j];
        G[i * 3 + j] = sacc;
      }
      jacobi_sym(G, W, 3, 60);
      int l3 = 0;
      for (int i = 1; i < 3; ++i) if (G[i * 3 + i] < G[l3 * 3 + l3]) l3 = i;
      const double v[3] = {W[0 * 3 + l3], W[1 * 3 + l3], W[2 * 3 + l3]};
      for (int i = 0; i < 3; ++i) {
        const double fv = F0[i * 3] * v[0] + F0[i * 3 + 1] * v[1] + F0[i * 3 + 2] * v[2];
        for (int j = 0; j < 3; ++j) F0[i * 3 + j] -= fv * v[j];
      }
      const double T1[9] = {sc1, 0, -sc1 * cx1, 0, sc1, -sc1 * cy1, 0, 0, 1};
      const double T2[9] = {sc2, 0, -sc2 * cx2, 0, sc2, -sc2 * cy2, 0, 0, 1};
      double X[9], Fr[9];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double sacc = 0;
        for (int k = 0; k < 3; ++k) sacc += T2[k * 3 + i] * F0[k * 3 + j];
        X[i * 3 + j] = sacc;
      }
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double sacc = 0;
        for (int k = 0; k < 3; ++k) sacc += X[i * 3 + k] * T1[k * 3 + j];
        Fr[i * 3 + j] = sacc;
      }
      if (fabs(Fr[8]) > static_cast<double>(FLT_EPSILON)) { const double inv = 1.0 / Fr[8]; for (int i = 0; i < 9; ++i) Fr[i] *= inv; }
      for (int i = 0; i < 9; ++i) Fout[i] = Fr[i];
    }
  }
  __syncthreads();
  return sOk != 0;
}

// PHILOX is a template parameter and the refit a kernel of its own (fmat_refit8_kernel): compiled into the one hot
// kernel, the sampler's local arrays and the refit's Jacobi sweeps cost the OpenCV-replay path 16x (measured).
#ifndef PM_RS_MINBLOCKS
#define PM_RS_MINBLOCKS (512 / RS_THREADS)
#endif
template <bool PHILOX>
__global__ void __launch_bounds__(RS_THREADS, PM_RS_MINBLOCKS)
fmat_ransac_kernel(const float2* __restrict__ pts1, const float2* __restrict__ pts2,
                   const int32_t* __restrict__ count, int stride, RansacDev prm,
                   uint8_t* __restrict__ mask, double* __restrict__ F_out,
                   int32_t* __restrict__ status, int32_t* __restrict__ n_inliers,
                   int32_t* __restrict__ iters_out, const PairJob* __restrict__ jobs,
                   int cut_iters, RsState* __restrict__ hand) {
  const int slot = blockIdx.x;
  const size_t base = static_cast<size_t>(slot) * stride;
  const float2* p1 = pts1 + base;
  const float2* p2 = pts2 + base;
  uint8_t* msk = mask + base;
  const int M = count[slot];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (hand && tid == 0) { hand[slot].active = 0; hand[slot].handed = 0; }

  __shared__ double sF[RS_ROUND][27];
#if PM_RS_MM
  __shared__ ModelF32 sMm[RS_ROUND * 3];                // the live models of the round in fp32 + their certified margins
  __shared__ float4 sPts[RS_CHUNK];                     // staged matches (x1, y1, x2, y2)
  __shared__ int sList[RS_ROUND * 3];                   // live model t -> 3 * hypothesis + solution
  __shared__ int sNmod;
#else
  __shared__ ModelF32 sMf[PM_RS_FP32 ? RS_ROUND : 1][3];                // PM_RS_FP32 (match-major build): fp32 models + margins
  __shared__ int sCntW[RS_THREADS / 32][RS_ROUND][3];   // match-major build: per-warp partial counts
  __shared__ unsigned sDead[RS_ROUND];                  // bit m: model m of the hypothesis is out of the race
#endif
  __shared__ double bestF[9];
  __shared__ int sSub[RS_ROUND][7];
  __shared__ int sNm[RS_ROUND];
  __shared__ int sCnt[RS_ROUND][3];
  __shared__ int sGen, sStop, sIter, sNiters, sBest;

  if (!prm.do_filter || M < prm.min_matches) {
    for (int i = tid; i < M; i += RS_THREADS) msk[i] = 1;
    if (tid == 0) {
      status[slot] = 0; n_inliers[slot] = M; iters_out[slot] = 0;
      for (int i = 0; i < 9; ++i) F_out[9 * slot + i] = 0.0;
    }
    return;
  }

  if (M == 7) {                       // direct 7-point, mask all ones, first solution kept
    if (tid == 0) {
      double Fm[27];
      float2 a[7], b[7];
      for (int i = 0; i < 7; ++i) { a[i] = p1[i]; b[i] = p2[i]; }
      const int ns = seven_point(a, b, Fm);
      for (int i = 0; i < 9; ++i) F_out[9 * slot + i] = ns > 0 ? Fm[i] : 0.0;
      status[slot] = ns > 0 ? 1 : 2;
      n_inliers[slot] = ns > 0 ? 7 : 0;
      iters_out[slot] = 1;
      for (int i = 0; i < 7; ++i) msk[i] = ns > 0 ? 1 : 0;
    }
    return;
  }

  __shared__ unsigned long long sRng;
  __shared__ unsigned long long sState[RS_ROUND + 1];   // RNG state in front of every subset of the round
  __shared__ int sBad;
  if (tid == 0) { sRng = ~0ull; sIter = 0; sNiters = prm.max_iters; sBest = 0; sStop = 0; }
  __syncthreads();
  const unsigned int inv_m = 0xFFFFFFFFu / static_cast<unsigned int>(M);
  const unsigned long long pkey = jobs ? ((static_cast<unsigned long long>(jobs[slot].seed_hi) << 32) | jobs[slot].seed_lo)
                                       : prm.seed;
  // margin of the conservative classifier (see classify()): scaled by the largest |coordinate| of the pair
  __shared__ float sCmax;
  if (tid == 0) sCmax = 0.f;
  __syncthreads();
  {
    float cm = 0.f;
    for (int i = tid; i < M; i += RS_THREADS) {
      const float2 u = p1[i], v = p2[i];
      cm = fmaxf(cm, fmaxf(fmaxf(fabsf(u.x), fabsf(u.y)), fmaxf(fabsf(v.x), fabsf(v.y))));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, off));
    if (lane == 0 && cm > 0.f) atomicMax(reinterpret_cast<int*>(&sCmax), __float_as_int(cm));   // non-negative floats
  }
  __syncthreads();
  const double margin = 1e-9 + 4e-14 * static_cast<double>(sCmax);
  const double thr_d = prm.thr, thr_lo = thr_d * (1.0 - margin), thr_hi = thr_d * (1.0 + 2.5e-7) * (1.0 + margin);

#ifdef PM_RANSAC_PROFILE
  long long tS = 0, tV = 0, tC = 0, tU = 0, t0_ = clock64();
#endif
  int round_size = 8;
  while (true) {
    // ---- sample: the sequential index stream (one thread), then the collinearity test of every
    //      subset in parallel; a rejected subset (rare) re-plays the stream from its start -------------
    if constexpr (PHILOX) {
      // free-running sampler: thread g draws the subset of iteration sIter + g on its own (collinearity test included)
      if (tid == 0) sBad = 0x7fffffff;
      __syncthreads();
      const int want = min(round_size, sNiters - sIter);
      if (tid < want && !philox_subset(p1, p2, M, pkey, sIter + tid, sSub[tid])) atomicMin(&sBad, tid);
      __syncthreads();
      if (tid == 0) {
        sGen = sBad < want ? sBad : (want > 0 ? want : 0);       // no admissible subset: stop there (OpenCV's rule)
        if (sBad < want) sStop = 1;
        sBad = 0x7fffffff;
      }
      __syncthreads();
    } else {
    if (tid == 0) {
      MwcRng rng{sRng};
      int g = 0;
      for (; g < round_size; ++g) {
        if (sIter + g >= sNiters) break;
        sState[g] = rng.s;
        draw7(rng, static_cast<unsigned int>(M), inv_m, sSub[g]);
      }
      sGen = g;
      sRng = rng.s;
      sBad = 0x7fffffff;
    }
    __syncthreads();
    for (int w = tid; w < sGen * 30; w += RS_THREADS) {
      const int h = w / 30, t = w - 30 * h;
      if (pair_collinear(t < 15 ? p1 : p2, sSub[h], t < 15 ? t : t - 15)) atomicMin(&sBad, h);
    }
    __syncthreads();
    }
    if (sBad < sGen) {
      if (tid == 0) {
        MwcRng rng{sState[sBad]};
        int g = sBad;
        for (; g < round_size; ++g) {
          if (sIter + g >= sNiters) break;
          if (!get_subset(p1, p2, M, rng, 10000, sSub[g])) { sStop = 1; break; }
        }
        sGen = g;
        sRng = rng.s;
      }
      __syncthreads();
    }
    const int gen = sGen;
    PM_PHASE(tS);
    if (gen == 0) break;
    // ---- solve: one half-warp per hypothesis (row-per-lane Householder QR) ------------------
    {
      const int hw = tid >> 4, li = tid & 15;
      const unsigned hmask = (lane & 16) ? 0xffff0000u : 0x0000ffffu;
      for (int h = hw; h < gen; h += RS_THREADS / 16) {
        const int nm = seven_point_coop(p1, p2, sSub[h], sF[h], li, hmask);
        if (li == 0) sNm[h] = nm;
      }
    }
    __syncthreads();
    PM_PHASE(tV);
    // ---- score: every thread owns four matches at a time and visits all models of the round (the model
    //      is a shared-memory broadcast, the points stay in registers as doubles) ---------------------
#if PM_RS_MM
    if (tid == 0) sNmod = 0;
    for (int w = tid; w < gen * 3; w += RS_THREADS) sCnt[w / 3][w % 3] = 0;
    __syncthreads();
    for (int w = tid; w < gen * 3; w += RS_THREADS)
      if (w % 3 < sNm[w / 3]) {
        const int t = atomicAdd(&sNmod, 1);              // any order: the counts go back to (hypothesis, solution)
        sList[t] = w;
        model_f32(&sF[w / 3][9 * (w % 3)], sCmax, prm.thr, sMm[t]);
      }
    __syncthreads();
    {
      const int nmod = sNmod;
      const int bound = sBest > 6 ? sBest : 6;
      for (int i0 = 0; i0 < M && nmod > 0; i0 += RS_CHUNK) {
        if (i0 > 0) __syncthreads();                     // the previous chunk is consumed, its counts are in sCnt
        const int n_here = min(RS_CHUNK, M - i0);
        for (int i = tid; i < n_here; i += RS_THREADS) {
          const float2 a = p1[i0 + i], b = p2[i0 + i];
          sPts[i] = make_float4(a.x, a.y, b.x, b.y);
        }
        __syncthreads();
        constexpr int SL = RS_CHUNK / (RS_THREADS / 32);
        const int j0 = warp * SL, j1 = min(j0 + SL, n_here);
        if (j0 < j1) {
          if (prm.residual_mode == 1) score_chunk_mm<1>(sF, sMm, sList, nmod, sCnt, sPts, j0, j1, bound, M - i0, i0 == 0, prm.thr, lane);
          else score_chunk_mm<0>(sF, sMm, sList, nmod, sCnt, sPts, j0, j1, bound, M - i0, i0 == 0, prm.thr, lane);
        }
      }
    }
    __syncthreads();
#else
    for (int w = tid; w < gen * 3; w += RS_THREADS) {
      sCnt[w / 3][w % 3] = 0;
#pragma unroll
      for (int wp = 0; wp < RS_THREADS / 32; ++wp) sCntW[wp][w / 3][w % 3] = 0;
#if PM_RS_FP32
      if (w % 3 < sNm[w / 3]) model_f32(&sF[w / 3][9 * (w % 3)], sCmax, prm.thr, sMf[w / 3][w % 3]);
#endif
    }
    for (int w = tid; w < gen; w += RS_THREADS) sDead[w] = 0u;
    __syncthreads();
    for (int i0 = 0; i0 < M; i0 += RS_PPT * RS_THREADS) {
      if (i0 > 0) {
        // A model replaces the best one only with STRICTLY more inliers (and more than 6).  Once its count so far
        // plus all the matches not yet visited cannot exceed the best count from before this round, its exact
        // count no longer matters: it is dropped from the remaining blocks.  Its partial count stays <= the
        // bound, so the selection below ignores it exactly as it would ignore the full count.
        __syncthreads();
        const int bound = sBest > 6 ? sBest : 6;
        for (int w = tid; w < gen * 3; w += RS_THREADS) {
          int c = 0;
#pragma unroll
          for (int wp = 0; wp < RS_THREADS / 32; ++wp) c += sCntW[wp][w / 3][w % 3];
          if (w % 3 < sNm[w / 3] && c + (M - i0) <= bound) atomicOr(&sDead[w / 3], 1u << (w % 3));
        }
        __syncthreads();
      }
      float2 q1[RS_PPT], q2[RS_PPT];
      bool valid[RS_PPT];
#pragma unroll
      for (int u = 0; u < RS_PPT; ++u) {
        const int i = i0 + u * RS_THREADS + tid;
        valid[u] = i < M;
        const int ic = valid[u] ? i : M - 1;
        q1[u] = p1[ic];
        q2[u] = p2[ic];
      }
      if (prm.residual_mode == 1) score_models<1, RS_PPT>(sF, sMf, sNm, sCntW[warp], sDead, gen, q1, q2, valid, prm.thr, thr_lo, thr_hi, lane);
      else score_models<0, RS_PPT>(sF, sMf, sNm, sCntW[warp], sDead, gen, q1, q2, valid, prm.thr, thr_lo, thr_hi, lane);
    }
    __syncthreads();
    for (int w = tid; w < gen * 3; w += RS_THREADS) {
      int c = 0;
#pragma unroll
      for (int wp = 0; wp < RS_THREADS / 32; ++wp) c += sCntW[wp][w / 3][w % 3];
      sCnt[w / 3][w % 3] = c;
    }
    __syncthreads();
#endif   // PM_RS_MM
    PM_PHASE(tC);
    // ---- select: (iteration, model)-ordered "strictly more inliers replaces" with the adaptive bound.
    //      Improvements are rare, so warp 0 jumps from one improving iteration to the next. -----------
    if (warp == 0) {
      const int nm = lane < gen ? sNm[lane] : 0;
      const int c0 = nm > 0 ? sCnt[lane][0] : -1, c1 = nm > 1 ? sCnt[lane][1] : -1, c2 = nm > 2 ? sCnt[lane][2] : -1;
      const int cmax = max(c0, max(c1, c2));
      int best = sBest, niters = sNiters;
      const int iter0 = sIter;
      int best_k = -1, best_m = -1;
      unsigned alive = 0xffffffffu;
      while (true) {
        const unsigned imp = __ballot_sync(0xffffffffu, cmax > (best > 6 ? best : 6)) & alive;
        if (!imp) break;
        const int k = __ffs(imp) - 1;
        if (iter0 + k >= niters) break;               // iteration k lies beyond the (shrunken) bound
        int nb = best, nn = niters, bm = -1;
        if (lane == k) {
          for (int m = 0; m < nm; ++m) {
            const int c = m == 0 ? c0 : (m == 1 ? c1 : c2);
            if (c > (nb > 6 ? nb : 6)) {
              nb = c; bm = m;
              nn = update_num_iters(prm.confidence, static_cast<double>(M - c) / M, 7, nn);
            }
          }
        }
        best = __shfl_sync(0xffffffffu, nb, k);
        niters = __shfl_sync(0xffffffffu, nn, k);
        best_m = __shfl_sync(0xffffffffu, bm, k);
        best_k = k;
        alive = k >= 31 ? 0u : ~((2u << k) - 1u);
      }
      if (best_k >= 0 && lane < 9) bestF[lane] = sF[best_k][9 * best_m + lane];
      if (lane == 0) {
        // the sequential loop stops at the first k with iter0 + k >= niters; every iteration up to the last
        // improving one was visited before the bound shrank
        int kend = niters - iter0;
        kend = kend < best_k + 1 ? best_k + 1 : kend;
        kend = kend > gen ? gen : kend;
        sBest = best; sNiters = niters; sIter = iter0 + kend;
        if (iter0 + kend >= niters) sStop = 1;
      }
    }
    __syncthreads();
    PM_PHASE(tU);
    if (sStop) break;
    if (hand && sIter >= cut_iters) {
      // Not finished after the first rounds: the remaining iterations run as device-wide stages (rs_*_kernel below),
      // where thousands of hypotheses of many pairs are solved and scored at once instead of 32 per resident block.
      if (tid == 0) {
        RsState& o = hand[slot];
        o.rng = sRng; o.key = pkey; o.iter = sIter; o.niters = sNiters; o.best = sBest; o.cmax = sCmax;
        o.gen = 0; o.halt = 0;
        for (int i = 0; i < 9; ++i) o.bestF[i] = sBest > 0 ? bestF[i] : 0.0;
        o.handed = 1; o.active = 1;
      }
      return;
    }
    round_size = round_size < RS_ROUND ? round_size * 2 : RS_ROUND;
  }
  __syncthreads();

#ifdef PM_RANSAC_PROFILE
  if (tid == 0 && slot < 3)
    printf("RANSAC slot %d M %d iters %d best %d | sample %lld solve %lld score %lld select %lld cycles | warp-iterations %llu with literal %llu undecided lanes %llu\n", slot, M,
           sIter, sBest, tS, tV, tC, tU, g_prof_iters, g_prof_lit_warps, g_prof_lit_lanes);
#endif
  const int best = sBest;
  if (best > 0) {
    double F[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = bestF[i];
    for (int i = tid; i < M; i += RS_THREADS)
      msk[i] = static_cast<uint8_t>(inlier_of(F, p1[i], p2[i], prm.residual_mode, prm.thr, thr_lo, thr_hi));
  } else {
    for (int i = tid; i < M; i += RS_THREADS) msk[i] = 0;
  }
  if (tid == 0) {
    status[slot] = best > 0 ? 1 : 2;
    n_inliers[slot] = best;
    iters_out[slot] = sIter;
    for (int i = 0; i < 9; ++i) F_out[9 * slot + i] = best > 0 ? bestF[i] : 0.0;
  }
}

// ==== staged continuation: iterations beyond the first rounds as device-wide kernels ==============================
// A pair whose adaptive bound is still far away after RS_CUT (24) iterations (few inliers: hundreds of iterations, ~2.3 models
// each, every one scored against every match) is latency-bound in the one-block-per-pair kernel above: 32 hypotheses
// per round, one resident block per SM next to the persistent kNN kernel.  Such pairs are handed over (RsState) and
// continue in "mega-rounds" of up to RS_MEGA iterations, each four small kernels over ALL handed-over pairs of the batch:
//   rs_sample_kernel   the same index stream (OpenCV MWC replay by one thread + parallel collinearity test + serial
//                      re-play of a rejected subset; or Philox, one subset per thread), RS_MEGA subsets at once;
//   rs_solve_kernel    one THREAD per hypothesis: the scalar 7-point solver with the 9 x 7 system in registers -- the
//                      same operations in the same order as the cooperative half-warp solver (bit-identical), but
//                      thousands of them in flight and no shuffles;
//   rs_score_kernel    lane = model, block = 32 models x all matches of the pair (score_chunk_mm's classifier);
//   rs_select_kernel   one warp per pair replays the (iteration, model)-ordered "strictly more inliers replaces" rule
//                      with the adaptive bound over the mega-round, 32 iterations at a time -- identical to the
//                      sequential loop because hypotheses at or beyond the shrunken bound are discarded.
// rs_finalize_kernel writes mask / F / status of the handed-over pairs.  Results are identical to the single-kernel
// path by construction (same subsets, same models, same counts, same selection order); tests compare both.
#ifndef PM_RS_CUT
#define PM_RS_CUT 24
#endif
static constexpr int RS_CUT = PM_RS_CUT;   // 8 + 16 iterations in the per-pair kernel (56 = 8 + 16 + 32: 69.7 k instead of 71.7 k pairs/s
                                           // at 50 % outliers, the same at 0 %)
static constexpr int RS_MEGA = 256;      // iterations per mega-round
static constexpr int RS_SOLVE_THREADS = 32;   // 152 registers per thread: small blocks pack next to the kNN kernel

struct RsWs {                            // views into the workspace of one slot
  RsState* state; int* sub; int* nm; int* cnt; int* list; int* handed; double* models;   // handed: [0] = n, [1 + k] = slot
};
__host__ __device__ inline RsWs rs_views(void* ws, int pairs) {
  RsWs v;
  char* p = static_cast<char*>(ws);
  v.state = reinterpret_cast<RsState*>(p);  p += sizeof(RsState) * static_cast<size_t>(pairs);
  v.models = reinterpret_cast<double*>(p);  p += sizeof(double) * 27 * RS_MEGA * static_cast<size_t>(pairs);
  v.sub = reinterpret_cast<int*>(p);        p += sizeof(int) * 7 * RS_MEGA * static_cast<size_t>(pairs);
  v.nm = reinterpret_cast<int*>(p);         p += sizeof(int) * RS_MEGA * static_cast<size_t>(pairs);
  v.cnt = reinterpret_cast<int*>(p);        p += sizeof(int) * 3 * RS_MEGA * static_cast<size_t>(pairs);
  v.list = reinterpret_cast<int*>(p);       p += sizeof(int) * 3 * RS_MEGA * static_cast<size_t>(pairs);
  v.handed = reinterpret_cast<int*>(p);
  return v;
}
size_t ransac_workspace_bytes(int pairs) {
  return (sizeof(RsState) + (sizeof(double) * 27 + sizeof(int) * (7 + 1 + 3 + 3)) * RS_MEGA + sizeof(int)) * static_cast<size_t>(pairs) + 64;
}

// The list of handed-over pairs of the batch (slot order): every staged kernel loops over it with a bounded grid, so a
// batch without such pairs (the common case) costs a handful of blocks per kernel instead of one per pair and stage.
__global__ void __launch_bounds__(256) rs_list_kernel(int n_jobs, RsWs ws) {
  __shared__ int sWarp[8];
  __shared__ int sBase;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) sBase = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n_jobs; b0 += 256) {
    const int slot = b0 + tid;
    const bool f = slot < n_jobs && ws.state[slot].handed != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) sWarp[warp] = __popc(bal);
    __syncthreads();
    int off = sBase;
    for (int w = 0; w < warp; ++w) off += sWarp[w];
    if (f) ws.handed[1 + off + __popc(bal & ((1u << lane) - 1u))] = slot;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += sWarp[w]; sBase += t; }
    __syncthreads();
  }
  if (tid == 0) ws.handed[0] = sBase;
}

template <bool PHILOX>
__device__ __forceinline__ void rs_sample_body(int slot, const float2* __restrict__ pts1, const float2* __restrict__ pts2,
                                               const int32_t* __restrict__ count, int stride, const RsWs& ws) {
  const int tid = threadIdx.x;
  RsState& st = ws.state[slot];
  if (!st.active) return;
  const size_t base = static_cast<size_t>(slot) * stride;
  const float2* p1 = pts1 + base;
  const float2* p2 = pts2 + base;
  const int M = count[slot];
  int* sub = ws.sub + static_cast<size_t>(slot) * RS_MEGA * 7;
  __shared__ int sSub[RS_MEGA][7];
  __shared__ unsigned long long sState[RS_MEGA + 1];
  __shared__ int sBad, sGen, sHalt;
  const int want = min(RS_MEGA, st.niters - st.iter);
  if (tid == 0) { sBad = 0x7fffffff; sHalt = 0; sGen = want > 0 ? want : 0; }
  __syncthreads();
  if constexpr (PHILOX) {
    for (int h = tid; h < want; h += RS_THREADS)
      if (!philox_subset(p1, p2, M, st.key, st.iter + h, sSub[h])) atomicMin(&sBad, h);
    __syncthreads();
    if (tid == 0 && sBad < want) { sGen = sBad; sHalt = 1; }      // no admissible subset: stop there (OpenCV's rule)
    __syncthreads();
  } else {
    // One thread replays the index stream (no point loads), all threads test the subsets for collinearity.  A rejected
    // subset b (a few per mega-round with integer pixel coordinates) is re-drawn by the sequential sampler, which also
    // yields the stream position behind it; the subsets after b are then drawn again from there and tested again.
    const unsigned int inv_m = 0xFFFFFFFFu / static_cast<unsigned int>(M);
    __shared__ int sFrom;
    if (tid == 0) { sFrom = 0; sState[0] = st.rng; }
    __syncthreads();
    while (true) {
      const int from = sFrom;
      if (tid == 0) {
        MwcRng rng{sState[from]};
        for (int g = from; g < want; ++g) {
          sState[g] = rng.s;
          draw7(rng, static_cast<unsigned int>(M), inv_m, sSub[g]);
        }
        sState[want > 0 ? want : 0] = rng.s;
        sBad = 0x7fffffff;
      }
      __syncthreads();
      for (int w = from * 30 + tid; w < want * 30; w += RS_THREADS) {
        const int h = w / 30, t = w - 30 * h;
        if (pair_collinear(t < 15 ? p1 : p2, sSub[h], t < 15 ? t : t - 15)) atomicMin(&sBad, h);
      }
      __syncthreads();
      const int bad = sBad;
      if (bad >= want) break;
      if (tid == 0) {
        MwcRng rng{sState[bad]};
        if (!get_subset(p1, p2, M, rng, 10000, sSub[bad])) { sHalt = 1; sGen = bad; }
        sState[bad + 1] = rng.s;
        sFrom = bad + 1;
      }
      __syncthreads();
      if (sHalt) break;
    }
    if (tid == 0) st.rng = sState[sHalt ? sGen : (want > 0 ? want : 0)];
    __syncthreads();
  }
  const int gen = sGen;
  for (int w = tid; w < gen * 7; w += RS_THREADS) sub[w] = sSub[w / 7][w % 7];
  if (tid == 0) { st.gen = gen; st.halt = sHalt; st.nmod = 0; }
}
template <bool PHILOX>
__global__ void __launch_bounds__(RS_THREADS)
rs_sample_kernel(const float2* __restrict__ pts1, const float2* __restrict__ pts2, const int32_t* __restrict__ count,
                 int stride, RsWs ws) {
  const int n = ws.handed[0];
  for (int li = blockIdx.x; li < n; li += gridDim.x) {
    rs_sample_body<PHILOX>(ws.handed[1 + li], pts1, pts2, count, stride, ws);
    __syncthreads();
  }
}

__device__ __forceinline__ void rs_solve_body(int slot, const float2* __restrict__ pts1, const float2* __restrict__ pts2,
                                              int stride, const RsWs& ws) {
  const RsState& st = ws.state[slot];
  const int h = blockIdx.x * RS_SOLVE_THREADS + threadIdx.x;
  if (!st.active || h >= st.gen) return;
  const size_t base = static_cast<size_t>(slot) * stride;
  const int* idx = ws.sub + (static_cast<size_t>(slot) * RS_MEGA + h) * 7;
  float2 m1[7], m2[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) { const int i = idx[k]; m1[k] = pts1[base + i]; m2[k] = pts2[base + i]; }
  double F[27];
  const int n = seven_point(m1, m2, F);
  double* out = ws.models + (static_cast<size_t>(slot) * RS_MEGA + h) * 27;
  for (int i = 0; i < 9 * n; ++i) out[i] = F[i];
  ws.nm[static_cast<size_t>(slot) * RS_MEGA + h] = n;
  if (n > 0) {                                         // live-model list of the pair, any order (counts go back by slot)
    const int at = atomicAdd(&ws.state[slot].nmod, n);
    int* list = ws.list + static_cast<size_t>(slot) * RS_MEGA * 3;
    for (int k = 0; k < n; ++k) list[at + k] = 3 * h + k;
  }
}
__global__ void __launch_bounds__(RS_SOLVE_THREADS)
rs_solve_kernel(const float2* __restrict__ pts1, const float2* __restrict__ pts2, int stride, RsWs ws) {
  const int n = ws.handed[0];
  for (int li = blockIdx.y; li < n; li += gridDim.y) rs_solve_body(ws.handed[1 + li], pts1, pts2, stride, ws);
}

#ifndef PM_RS_SCORE_MINB
#define PM_RS_SCORE_MINB 8
#endif
template <int MODE>
__device__ __forceinline__ void rs_score_body(int slot, const float2* __restrict__ pts1, const float2* __restrict__ pts2,
                                              const int32_t* __restrict__ count, int stride, float thr, const RsWs& ws) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const RsState& st = ws.state[slot];
  if (!st.active) return;
  const int nmod = st.nmod;
  if (blockIdx.x * 32 >= nmod) return;
  const int t = blockIdx.x * 32 + lane;
  const bool act = t < nmod;
  const int w = ws.list[static_cast<size_t>(slot) * RS_MEGA * 3 + (act ? t : blockIdx.x * 32)];   // 3 * hypothesis + solution
  const double* Fd = ws.models + static_cast<size_t>(slot) * RS_MEGA * 27 + 9 * w;
  __shared__ float4 sPts[RS_CHUNK];
  __shared__ int sC[32];
  if (tid < 32) sC[tid] = 0;
  ModelF32 Mf;
  model_f32(Fd, st.cmax, thr, Mf);
  if (!act) {
    // a lane without a model (last block of a pair only) holds constants for which every point is "an outlier for sure"
    // (|d| = 0 > 2 dd = -1 and (|d| - dd)^2 := 1 >= thr * 0 + 0), so the hot loop needs no test for it
#pragma unroll
    for (int i = 0; i < 9; ++i) Mf.F[i] = 0.f;
    Mf.dd2 = -1.f; Mf.dd_sq = 1.f; Mf.dg1t_lo = Mf.dg2t_lo = Mf.dg1t_hi = Mf.dg2t_hi = 0.f;
  }
  const float tlo = thr * (1.f - 4e-6f), thi = thr * (1.f + 4e-6f);
  const size_t base = static_cast<size_t>(slot) * stride;
  const float2* p1 = pts1 + base;
  const float2* p2 = pts2 + base;
  const int M = count[slot];
  // Early drop (exact): a model replaces the best one only with STRICTLY more inliers (and more than 6), and the best
  // count only grows during the mega-round.  Once a model's count so far plus all matches not yet visited cannot
  // exceed the best count from BEFORE the mega-round, its exact count no longer matters; when that holds for all 32
  // models of the block the remaining chunks are skipped (the partial counts stay <= the bound, so the selection
  // ignores them exactly as it would ignore the full counts).
  const int bound = st.best > 6 ? st.best : 6;
  for (int i0 = 0; i0 < M; i0 += RS_CHUNK) {
    __syncthreads();                                  // the previous chunk is consumed, its counts are in sC
    const bool dead_w = __all_sync(0xffffffffu, !act || sC[lane] + (M - i0) <= bound);
    const int n_here = min(RS_CHUNK, M - i0);
    for (int i = tid; i < n_here; i += RS_THREADS) {
      const float2 a = p1[i0 + i], b = p2[i0 + i];
      sPts[i] = make_float4(a.x, a.y, b.x, b.y);
    }
    if (__syncthreads_and(dead_w ? 1 : 0)) break;     // block-uniform
    if (dead_w) continue;
    int cnt = 0;
    constexpr int SL = RS_CHUNK / (RS_THREADS / 32);
    const int j0 = warp * SL, j1 = min(j0 + SL, n_here);
    if (MODE == 0) {
#pragma unroll 2
      for (int j = j0; j < j1; ++j) {
        const float4 q = sPts[j];
        const Side2 s2 = classify32_side2(Mf, q.x, q.y, q.z, q.w, thi);
#ifdef PM_RANSAC_PARANOID
        int c_chk = 0;
#endif
        if (__any_sync(0xffffffffu, !s2.out)) {                 // rare for wrong models: the other side + the inlier test
          int c = classify32_rest(Mf, s2, q.z, q.w, tlo, thi);
          if (c == 2 && act) c = literal_inlier(Fd, make_float2(q.x, q.y), make_float2(q.z, q.w), MODE, thr);
          cnt += c & 1;                                         // (lanes without a model always read "outlier for sure")
#ifdef PM_RANSAC_PARANOID
          c_chk = c;
#endif
        }
#ifdef PM_RANSAC_PARANOID
        if (act && literal_inlier(Fd, make_float2(q.x, q.y), make_float2(q.z, q.w), MODE, thr) != (c_chk & 1))
          printf("RANSAC PARANOID MISMATCH (staged) w %d cls %d pt (%g,%g)-(%g,%g)\n", w, c_chk, q.x, q.y, q.z, q.w);
        if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0 && i0 == 0 && j == 0) printf("paranoid build active (staged)\n");
#endif
      }
    } else {
      int j = j0;
      for (; j + 4 <= j1; j += 4) {
        int cls[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 q = sPts[j + u];
          cls[u] = classify32<MODE>(Mf, q.x, q.y, q.z, q.w, tlo, thi);
        }
#ifdef PM_RANSAC_PARANOID
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 q = sPts[j + u];
          if (act && cls[u] != 2 && literal_inlier(Fd, make_float2(q.x, q.y), make_float2(q.z, q.w), MODE, thr) != cls[u])
            printf("RANSAC PARANOID MISMATCH (staged) w %d cls %d pt (%g,%g)-(%g,%g)\n", w, cls[u], q.x, q.y, q.z, q.w);
        }
        if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0 && i0 == 0 && j == 0) printf("paranoid build active (staged)\n");
#endif
        if (((cls[0] | cls[1] | cls[2] | cls[3]) & 2) && act) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (cls[u] == 2) {
              const float4 q = sPts[j + u];
              cls[u] = literal_inlier(Fd, make_float2(q.x, q.y), make_float2(q.z, q.w), MODE, thr);
            }
        }
        cnt += (cls[0] & 1) + (cls[1] & 1) + (cls[2] & 1) + (cls[3] & 1);
      }
      for (; j < j1; ++j) {
        const float4 q = sPts[j];
        int c = classify32<MODE>(Mf, q.x, q.y, q.z, q.w, tlo, thi);
        if (c == 2 && act) c = literal_inlier(Fd, make_float2(q.x, q.y), make_float2(q.z, q.w), MODE, thr);
        cnt += c & 1;
      }
    }
    if (act && cnt) atomicAdd(&sC[lane], cnt);
  }
  __syncthreads();
  if (warp == 0 && act) ws.cnt[static_cast<size_t>(slot) * RS_MEGA * 3 + w] = sC[lane];
}
template <int MODE>
__global__ void __launch_bounds__(RS_THREADS, PM_RS_SCORE_MINB)
rs_score_kernel(const float2* __restrict__ pts1, const float2* __restrict__ pts2, const int32_t* __restrict__ count,
                int stride, float thr, RsWs ws) {
  const int n = ws.handed[0];
  for (int li = blockIdx.y; li < n; li += gridDim.y) {
    rs_score_body<MODE>(ws.handed[1 + li], pts1, pts2, count, stride, thr, ws);
    __syncthreads();
  }
}

__device__ __forceinline__ void rs_select_body(int slot, const int32_t* __restrict__ count, double confidence, const RsWs& ws) {
  const int lane = threadIdx.x;
  RsState& st = ws.state[slot];
  if (!st.active) return;
  const int M = count[slot];
  const int gen = st.gen;
  const int* nmv = ws.nm + static_cast<size_t>(slot) * RS_MEGA;
  const int* cntv = ws.cnt + static_cast<size_t>(slot) * RS_MEGA * 3;
  int best = st.best, niters = st.niters, iter = st.iter;
  int win_h = -1, win_m = -1;
  bool stop = gen == 0;
  for (int b0 = 0; b0 < gen && !stop; b0 += 32) {
    const int g32 = min(32, gen - b0);
    const int nm = lane < g32 ? nmv[b0 + lane] : 0;
    const int c0 = nm > 0 ? cntv[3 * (b0 + lane)] : -1, c1 = nm > 1 ? cntv[3 * (b0 + lane) + 1] : -1,
              c2 = nm > 2 ? cntv[3 * (b0 + lane) + 2] : -1;
    const int cmax = max(c0, max(c1, c2));
    const int iter0 = iter;
    int best_k = -1;
    unsigned alive = 0xffffffffu;
    while (true) {
      const unsigned imp = __ballot_sync(0xffffffffu, cmax > (best > 6 ? best : 6)) & alive;
      if (!imp) break;
      const int k = __ffs(imp) - 1;
      if (iter0 + k >= niters) break;               // iteration k lies beyond the (shrunken) bound
      int nb = best, nn = niters, bm = -1;
      if (lane == k) {
        for (int mm = 0; mm < nm; ++mm) {
          const int c = mm == 0 ? c0 : (mm == 1 ? c1 : c2);
          if (c > (nb > 6 ? nb : 6)) {
            nb = c; bm = mm;
            nn = update_num_iters(confidence, static_cast<double>(M - c) / M, 7, nn);
          }
        }
      }
      best = __shfl_sync(0xffffffffu, nb, k);
      niters = __shfl_sync(0xffffffffu, nn, k);
      win_m = __shfl_sync(0xffffffffu, bm, k);
      win_h = b0 + k;
      best_k = k;
      alive = k >= 31 ? 0u : ~((2u << k) - 1u);
    }
    int kend = niters - iter0;
    kend = kend < best_k + 1 ? best_k + 1 : kend;
    kend = kend > g32 ? g32 : kend;
    iter = iter0 + kend;
    if (iter >= niters) stop = true;
  }
  if (win_h >= 0 && lane < 9) st.bestF[lane] = ws.models[(static_cast<size_t>(slot) * RS_MEGA + win_h) * 27 + 9 * win_m + lane];
  if (lane == 0) {
    st.best = best; st.niters = niters; st.iter = iter;
    if (stop || st.halt) st.active = 0;
  }
}
__global__ void __launch_bounds__(32)
rs_select_kernel(const int32_t* __restrict__ count, double confidence, RsWs ws) {
  const int n = ws.handed[0];
  for (int li = blockIdx.x; li < n; li += gridDim.x) { rs_select_body(ws.handed[1 + li], count, confidence, ws); __syncwarp(); }
}

__device__ __forceinline__ void rs_finalize_body(int slot, const float2* __restrict__ pts1, const float2* __restrict__ pts2,
                                                 const int32_t* __restrict__ count, int stride, const RansacDev& prm,
                                                 uint8_t* __restrict__ mask, double* __restrict__ F_out,
                                                 int32_t* __restrict__ status, int32_t* __restrict__ n_inliers,
                                                 int32_t* __restrict__ iters_out, const RsWs& ws) {
  const int tid = threadIdx.x;
  const RsState& st = ws.state[slot];
  if (!st.handed) return;                         // finished in the per-pair kernel, which wrote its own outputs
  const size_t base = static_cast<size_t>(slot) * stride;
  const float2* p1 = pts1 + base;
  const float2* p2 = pts2 + base;
  uint8_t* msk = mask + base;
  const int M = count[slot];
  const double margin = 1e-9 + 4e-14 * static_cast<double>(st.cmax);
  const double thr_d = prm.thr, thr_lo = thr_d * (1.0 - margin), thr_hi = thr_d * (1.0 + 2.5e-7) * (1.0 + margin);
  const int best = st.best;
  if (best > 0) {
    double F[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = st.bestF[i];
    for (int i = tid; i < M; i += RS_THREADS)
      msk[i] = static_cast<uint8_t>(inlier_of(F, p1[i], p2[i], prm.residual_mode, prm.thr, thr_lo, thr_hi));
  } else {
    for (int i = tid; i < M; i += RS_THREADS) msk[i] = 0;
  }
  if (tid == 0) {
    status[slot] = best > 0 ? 1 : 2;
    n_inliers[slot] = best;
    iters_out[slot] = st.iter;
    for (int i = 0; i < 9; ++i) F_out[9 * slot + i] = best > 0 ? st.bestF[i] : 0.0;
  }
}
__global__ void __launch_bounds__(RS_THREADS)
rs_finalize_kernel(const float2* __restrict__ pts1, const float2* __restrict__ pts2, const int32_t* __restrict__ count,
                   int stride, RansacDev prm, uint8_t* __restrict__ mask, double* __restrict__ F_out,
                   int32_t* __restrict__ status, int32_t* __restrict__ n_inliers, int32_t* __restrict__ iters_out, RsWs ws) {
  const int n = ws.handed[0];
  for (int li = blockIdx.x; li < n; li += gridDim.x)
    rs_finalize_body(ws.handed[1 + li], pts1, pts2, count, stride, prm, mask, F_out, status, n_inliers, iters_out, ws);
}

// Optional refit (pm_params.refit_8point): F of every filtered pair with >= 8 inliers is replaced by the normalised
// 8-point estimate over its inliers; mask, counts and status stay those of the winning RANSAC hypothesis.
__global__ void __launch_bounds__(RS_THREADS)
fmat_refit8_kernel(const float2* __restrict__ pts1, const float2* __restrict__ pts2, const int32_t* __restrict__ count,
                   int stride, const uint8_t* __restrict__ mask, double* __restrict__ F_out,
                   const int32_t* __restrict__ status, const int32_t* __restrict__ n_inliers) {
  const int slot = blockIdx.x;
  if (status[slot] != 1 || n_inliers[slot] < 8) return;          // uniform over the block
  __shared__ double sA[81 + 81 + 8];
  __shared__ double sFr[9];
  const size_t base = static_cast<size_t>(slot) * stride;
  const int tid = threadIdx.x;
  const bool ok = eight_point_refit(pts1 + base, pts2 + base, mask + base, count[slot], sFr, sA, tid);
  if (ok && tid < 9) F_out[9 * slot + tid] = sFr[tid];
}

cudaError_t ransac_configure() {
  cudaError_t e = cudaFuncSetAttribute(fmat_ransac_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(fmat_ransac_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

int ransac_stage_cut() { return RS_CUT; }
int ransac_mega_rounds(int max_iters) {
  if (max_iters <= RS_CUT) return 0;
  const int n = (max_iters - RS_CUT + RS_MEGA - 1) / RS_MEGA;
  return n > 16 ? 0 : n;                      // very long bounds stay in the per-pair kernel
}

cudaError_t launch_ransac(const float2* pts1, const float2* pts2, const int32_t* count, int n_jobs,
                          int stride, const RansacDev& prm, uint8_t* mask, double* F,
                          int32_t* status, int32_t* n_inliers, int32_t* iters, cudaStream_t st,
                          const PairJob* jobs, void* workspace, int* n_launches) {
  if (n_jobs <= 0) return cudaSuccess;
  const int n_mega = (workspace && prm.do_filter) ? ransac_mega_rounds(prm.max_iters) : 0;
  RsWs ws{};
  if (n_mega > 0) ws = rs_views(workspace, n_jobs);
  RsState* hand = n_mega > 0 ? ws.state : nullptr;
  int launches = 1;
  if (prm.sampler == 1)
    fmat_ransac_kernel<true><<<n_jobs, RS_THREADS, 0, st>>>(pts1, pts2, count, stride, prm, mask, F, status, n_inliers,
                                                            iters, jobs, RS_CUT, hand);
  else
    fmat_ransac_kernel<false><<<n_jobs, RS_THREADS, 0, st>>>(pts1, pts2, count, stride, prm, mask, F, status, n_inliers,
                                                             iters, jobs, RS_CUT, hand);
  cudaError_t e = cudaGetLastError();
  // bounded grids: the kernels loop over the list of handed-over pairs
  const int gp = n_jobs;                                         // one block per pair and stage (serial latency per pair: no looping)
  const int gy = n_jobs < 128 ? n_jobs : 128;                    // x (hypothesis / model groups of a pair)
  if (n_mega > 0 && e == cudaSuccess) {
    rs_list_kernel<<<1, 256, 0, st>>>(n_jobs, ws);
    ++launches;
    e = cudaGetLastError();
  }
  for (int r = 0; r < n_mega && e == cudaSuccess; ++r) {
    if (prm.sampler == 1) rs_sample_kernel<true><<<gp, RS_THREADS, 0, st>>>(pts1, pts2, count, stride, ws);
    else rs_sample_kernel<false><<<gp, RS_THREADS, 0, st>>>(pts1, pts2, count, stride, ws);
    rs_solve_kernel<<<dim3(RS_MEGA / RS_SOLVE_THREADS, gy), RS_SOLVE_THREADS, 0, st>>>(pts1, pts2, stride, ws);
    if (prm.residual_mode == 1)
      rs_score_kernel<1><<<dim3(RS_MEGA * 3 / 32, gy), RS_THREADS, 0, st>>>(pts1, pts2, count, stride, prm.thr, ws);
    else
      rs_score_kernel<0><<<dim3(RS_MEGA * 3 / 32, gy), RS_THREADS, 0, st>>>(pts1, pts2, count, stride, prm.thr, ws);
    rs_select_kernel<<<gp, 32, 0, st>>>(count, prm.confidence, ws);
    launches += 4;
    e = cudaGetLastError();
  }
  if (n_mega > 0 && e == cudaSuccess) {
    rs_finalize_kernel<<<gp, RS_THREADS, 0, st>>>(pts1, pts2, count, stride, prm, mask, F, status, n_inliers, iters, ws);
    ++launches;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && prm.refit_8point && prm.do_filter) {
    fmat_refit8_kernel<<<n_jobs, RS_THREADS, 0, st>>>(pts1, pts2, count, stride, mask, F, status, n_inliers);
    ++launches;
    e = cudaGetLastError();
  }
  if (n_launches) *n_launches = launches;
  return e;
}

}  // namespace pm
