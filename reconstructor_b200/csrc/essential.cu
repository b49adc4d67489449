// essential.cu -- GeometricFilter::estimateEssential on the batched RANSAC machinery.
//
// Replaces cv::findEssentialMat(p1, p2, K1, dist1, K2, dist2) with its defaults (RANSAC, prob 0.999, threshold 1.0,
// 1000 iterations) as the reference calls it (Mapper/libMapper/GeometricFilter.cpp:26-31; caller
// SequentialReconstructor::chooseInitialPair, .cpp:355, once per reconstruction):
//   1. undistortPoints with each image's own camera (radial k1, k2 of PinholeCamera, Camera.h:113-123; OpenCV's five
//      fixed-point iterations), stored as float; mapped to the pixels of the mean camera K0 = (K1 + K2) / 2 in float
//      arithmetic (cv::transform); normalised (p - c0) / f0 in double; threshold / ((fx0 + fy0) / 2);
//   2. RANSAC of the point-set registrator findFundamentalMat uses as well -- OpenCV's fixed-seed multiply-with-carry
//      stream (or the Philox sampler), 5 distinct indices per sample, NO subset check for this model, "strictly more
//      inliers (and more than 4) replaces", adaptive stop -- with Nister's five-point solver (<= 10 models per sample)
//      and the Sampson residual (x2' E x1)^2 / (|E x1|_xy^2 + |E' x2|_xy^2), stored as float, compared with (float)(t*t).
// One CTA per pair; rounds of ES_ROUND samples: one thread replays the index stream, one thread per sample solves, all
// threads score, one thread walks the samples in order.  This call runs once per reconstruction, so the solver is
// written for clarity (dense exponent-indexed polynomials in local memory), not for speed.
// The reference never hands a mask to cv::findEssentialMat (GeometricFilter.cpp:25-33: `inliersCV` stays empty), so its
// inlierMatchIds comes back empty; the C ABI returns the mask nevertheless (the shim reproduces the reference's behaviour
// unless asked otherwise).
// Compiled with -fmad=false: the arithmetic follows the CPU filter of the parity tests operation for operation.
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace pm {

static constexpr int ES_THREADS = 128;
static constexpr int ES_ROUND = 32;

__device__ __constant__ int ES_MONO3[20][3] = {
    {3, 0, 0}, {0, 3, 0}, {2, 1, 0}, {1, 2, 0}, {2, 0, 1}, {2, 0, 0}, {0, 2, 1}, {0, 2, 0}, {1, 1, 1}, {1, 1, 0},
    {1, 0, 2}, {1, 0, 1}, {1, 0, 0}, {0, 1, 2}, {0, 1, 1}, {0, 1, 0}, {0, 0, 3}, {0, 0, 2}, {0, 0, 1}, {0, 0, 0}};

struct Poly3 { double c[4][4][4]; };      // total degree <= 3 in (x, y, z), dense over the exponents
__device__ void p_zero(Poly3& p) {
  for (int i = 0; i < 64; ++i) (&p.c[0][0][0])[i] = 0.0;
}
__device__ void p_lin(Poly3& p, double x, double y, double z, double w) {
  p_zero(p); p.c[1][0][0] = x; p.c[0][1][0] = y; p.c[0][0][1] = z; p.c[0][0][0] = w;
}
__device__ void p_mul(Poly3& r, const Poly3& a, const Poly3& b) {
  Poly3 t; p_zero(t);
  for (int a0 = 0; a0 < 4; ++a0) for (int a1 = 0; a1 + a0 < 4; ++a1) for (int a2 = 0; a2 + a1 + a0 < 4; ++a2) {
    const double va = a.c[a0][a1][a2];
    if (va == 0.0) continue;
    for (int b0 = 0; b0 + a0 < 4; ++b0) for (int b1 = 0; b1 + a1 + b0 + a0 < 4; ++b1)
      for (int b2 = 0; b2 + a2 + b1 + a1 + b0 + a0 < 4; ++b2)
        t.c[a0 + b0][a1 + b1][a2 + b2] += va * b.c[b0][b1][b2];
  }
  r = t;
}
__device__ void p_axpy(Poly3& r, double s, const Poly3& a) {
  for (int i = 0; i < 64; ++i) (&r.c[0][0][0])[i] += s * (&a.c[0][0][0])[i];
}

__device__ bool null_space_5x9(const double (*Q)[9], double (*N)[9]) {
  double B[9][5], vn2[5];
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 9; ++j) B[j][i] = Q[i][j];
  for (int k = 0; k < 5; ++k) {
    double nrm2 = 0;
    for (int i = k; i < 9; ++i) nrm2 += B[i][k] * B[i][k];
    const double nrm = sqrt(nrm2);
    if (!(nrm > 0)) return false;
    const double alpha = B[k][k] > 0 ? -nrm : nrm;
    B[k][k] -= alpha;
    double s2 = 0;
    for (int i = k; i < 9; ++i) s2 += B[i][k] * B[i][k];
    vn2[k] = s2;
    if (!(s2 > 0)) return false;
    for (int j = k + 1; j < 5; ++j) {
      double s = 0;
      for (int i = k; i < 9; ++i) s += B[i][k] * B[i][j];
      const double f = 2 * s / s2;
      for (int i = k; i < 9; ++i) B[i][j] -= f * B[i][k];
    }
  }
  for (int e = 0; e < 4; ++e) {
    double y[9];
    for (int i = 0; i < 9; ++i) y[i] = 0.0;
    y[5 + e] = 1.0;
    for (int k = 4; k >= 0; --k) {
      double s = 0;
      for (int i = k; i < 9; ++i) s += B[i][k] * y[i];
      const double f = 2 * s / vn2[k];
      for (int i = k; i < 9; ++i) y[i] -= f * B[i][k];
    }
    for (int i = 0; i < 9; ++i) N[e][i] = y[i];
  }
  return true;
}

// det(E) = 0 and 2 E E' E - trace(E E') E = 0 on E = x X + y Y + z Z + W: ten cubics over Nister's monomial order
__device__ void constraint_matrix(const double (*N)[9], double (*A)[20]) {
  Poly3 e[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    p_lin(e[i][j], N[0][3 * i + j], N[1][3 * i + j], N[2][3 * i + j], N[3][3 * i + j]);
  Poly3 eet[3][3], tr, t, row;
  p_zero(tr);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    p_zero(eet[i][j]);
    for (int k = 0; k < 3; ++k) { p_mul(t, e[i][k], e[j][k]); p_axpy(eet[i][j], 1.0, t); }
    if (i == j) p_axpy(tr, 1.0, eet[i][j]);
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    p_zero(row);
    for (int k = 0; k < 3; ++k) { p_mul(t, eet[i][k], e[k][j]); p_axpy(row, 2.0, t); }
    p_mul(t, tr, e[i][j]); p_axpy(row, -1.0, t);
    for (int m = 0; m < 20; ++m) A[3 * i + j][m] = row.c[ES_MONO3[m][0]][ES_MONO3[m][1]][ES_MONO3[m][2]];
  }
  Poly3 mm, d;
  p_zero(row);
  p_mul(mm, e[1][1], e[2][2]); p_mul(d, e[1][2], e[2][1]); p_axpy(mm, -1.0, d); p_mul(t, e[0][0], mm); p_axpy(row, 1.0, t);
  p_mul(mm, e[1][0], e[2][2]); p_mul(d, e[1][2], e[2][0]); p_axpy(mm, -1.0, d); p_mul(t, e[0][1], mm); p_axpy(row, -1.0, t);
  p_mul(mm, e[1][0], e[2][1]); p_mul(d, e[1][1], e[2][0]); p_axpy(mm, -1.0, d); p_mul(t, e[0][2], mm); p_axpy(row, 1.0, t);
  for (int m = 0; m < 20; ++m) A[9][m] = row.c[ES_MONO3[m][0]][ES_MONO3[m][1]][ES_MONO3[m][2]];
}

__device__ bool gauss_jordan_10x20(double (*A)[20]) {
  for (int c = 0; c < 10; ++c) {
    int p = c;
    for (int r = c + 1; r < 10; ++r) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
    if (!(fabs(A[p][c]) > 1e-300)) return false;
    if (p != c) for (int k = 0; k < 20; ++k) { const double t = A[c][k]; A[c][k] = A[p][k]; A[p][k] = t; }
    const double inv = 1.0 / A[c][c];
    for (int k = 0; k < 20; ++k) A[c][k] *= inv;
    for (int r = 0; r < 10; ++r) {
      if (r == c) continue;
      const double f = A[r][c];
      if (f == 0.0) continue;
      for (int k = 0; k < 20; ++k) A[r][k] -= f * A[c][k];
    }
  }
  return true;
}

// all roots of c[0] + c[1] z + ... + c[n] z^n: Durand-Kerner iteration (the method of cv::solvePoly)
__device__ int poly_roots(const double* c, int n, double* re, double* im) {
  while (n > 1 && !(fabs(c[n]) > DBL_EPSILON)) --n;
  double pr = 1, pi = 0;
  for (int i = 0; i < n; ++i) { re[i] = pr; im[i] = pi; const double t = pr - pi; pi = pr + pi; pr = t; }
  for (int iter = 0; iter < 500; ++iter) {
    double maxd = 0;
    for (int i = 0; i < n; ++i) {
      const double xr = re[i], xi = im[i];
      double nr = c[n], ni = 0, dr = c[n], di = 0;
      for (int j = 0; j < n; ++j) {
        double t = nr * xr - ni * xi + c[n - j - 1];
        ni = nr * xi + ni * xr; nr = t;
        if (j != i) {
          const double ar = xr - re[j], ai = xi - im[j];
          if (ar != 0 || ai != 0) { t = dr * ar - di * ai; di = dr * ai + di * ar; dr = t; }
        }
      }
      const double s = 1.0 / (dr * dr + di * di);
      const double qr = (nr * dr + ni * di) * s, qi = (ni * dr - nr * di) * s;
      re[i] = xr - qr; im[i] = xi - qi;
      const double d = sqrt(qr * qr + qi * qi);
      if (d > maxd) maxd = d;
    }
    if (maxd <= 0) break;
  }
  return n;
}

// m1, m2: the normalised points; idx: 5 indices.  Es receives up to 10 models (unit Frobenius norm, ascending root).
__device__ int five_point(const double2* __restrict__ m1, const double2* __restrict__ m2, const int* idx, double* Es) {
  double Q[5][9], N[4][9], A[10][20];
  for (int i = 0; i < 5; ++i) {
    const double2 a = m1[idx[i]], b = m2[idx[i]];
    const double x1 = a.x, y1 = a.y, x2 = b.x, y2 = b.y;
    Q[i][0] = x1 * x2; Q[i][1] = y1 * x2; Q[i][2] = x2;
    Q[i][3] = x1 * y2; Q[i][4] = y1 * y2; Q[i][5] = y2;
    Q[i][6] = x1;      Q[i][7] = y1;      Q[i][8] = 1.0;
  }
  if (!null_space_5x9(Q, N)) return 0;
  constraint_matrix(N, A);
  if (!gauss_jordan_10x20(A)) return 0;
  double B[3][13];
  for (int i = 0; i < 3; ++i) {
    const double* r1 = A[2 * i + 4];
    const double* r2 = A[2 * i + 5];
    double a[13], b[13];
    for (int k = 0; k < 13; ++k) { a[k] = 0.0; b[k] = 0.0; }
    for (int k = 0; k < 3; ++k) { a[1 + k] = r1[10 + k]; a[5 + k] = r1[13 + k]; b[k] = r2[10 + k]; b[4 + k] = r2[13 + k]; }
    for (int k = 0; k < 4; ++k) { a[9 + k] = r1[16 + k]; b[8 + k] = r2[16 + k]; }
    for (int k = 0; k < 13; ++k) B[i][k] = a[k] - b[k];
  }
  double P[3][3][5];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) for (int k = 0; k < 5; ++k) P[i][j][k] = 0.0;
    for (int k = 0; k < 4; ++k) { P[i][0][k] = B[i][3 - k]; P[i][1][k] = B[i][7 - k]; }
    for (int k = 0; k < 5; ++k) P[i][2][k] = B[i][12 - k];
  }
  double c[11];
  for (int k = 0; k < 11; ++k) c[k] = 0.0;
  const int perm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {2, 1, 0}, {1, 0, 2}};
  for (int p = 0; p < 6; ++p) {
    const double sg = p < 3 ? 1.0 : -1.0;
    const double* f0 = P[perm[p][0]][0];
    const double* f1 = P[perm[p][1]][1];
    const double* f2 = P[perm[p][2]][2];
    double t01[7];
    for (int k = 0; k < 7; ++k) t01[k] = 0.0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) t01[i + j] += f0[i] * f1[j];
    for (int i = 0; i < 7; ++i) for (int j = 0; j < 5; ++j) c[i + j] += sg * t01[i] * f2[j];
  }
  double re[10], im[10], zs[10];
  const int deg = poly_roots(c, 10, re, im);
  int nz = 0;
  for (int i = 0; i < deg; ++i) if (fabs(im[i]) <= 1e-10) zs[nz++] = re[i];
  for (int i = 1; i < nz; ++i) { const double v = zs[i]; int j = i - 1; while (j >= 0 && zs[j] > v) { zs[j + 1] = zs[j]; --j; } zs[j + 1] = v; }
  int count = 0;
  for (int r = 0; r < nz; ++r) {
    const double z1 = zs[r], z2 = z1 * z1, z3 = z2 * z1, z4 = z3 * z1;
    double bz[3][3];
    for (int j = 0; j < 3; ++j) {
      const double* br = B[j];
      bz[j][0] = br[0] * z3 + br[1] * z2 + br[2] * z1 + br[3];
      bz[j][1] = br[4] * z3 + br[5] * z2 + br[6] * z1 + br[7];
      bz[j][2] = br[8] * z4 + br[9] * z3 + br[10] * z2 + br[11] * z1 + br[12];
    }
    double best[3] = {0, 0, 0}, bn = -1;
    for (int a = 0; a < 3; ++a) for (int b = a + 1; b < 3; ++b) {
      const double v0 = bz[a][1] * bz[b][2] - bz[a][2] * bz[b][1], v1 = bz[a][2] * bz[b][0] - bz[a][0] * bz[b][2];
      const double v2 = bz[a][0] * bz[b][1] - bz[a][1] * bz[b][0];
      const double nn = v0 * v0 + v1 * v1 + v2 * v2;
      if (nn > bn) { bn = nn; best[0] = v0; best[1] = v1; best[2] = v2; }
    }
    if (!(bn > 0)) continue;
    const double inv = 1.0 / sqrt(bn);
    const double vx = best[0] * inv, vy = best[1] * inv, vw = best[2] * inv;
    if (fabs(vw) < 1e-10) continue;
    const double x = vx / vw, y = vy / vw;
    double E[9], nrm = 0;
    for (int k = 0; k < 9; ++k) { E[k] = N[0][k] * x + N[1][k] * y + N[2][k] * z1 + N[3][k]; nrm += E[k] * E[k]; }
    nrm = sqrt(nrm);
    if (!(nrm > 0)) continue;
    for (int k = 0; k < 9; ++k) Es[9 * count + k] = E[k] / nrm;
    ++count;
  }
  return count;
}

__device__ __forceinline__ int e_inlier(const double* E, double2 a, double2 b, float thr) {
  const double x1 = a.x, y1 = a.y, x2 = b.x, y2 = b.y;
  const double a0 = E[0] * x1 + E[1] * y1 + E[2], a1 = E[3] * x1 + E[4] * y1 + E[5], a2 = E[6] * x1 + E[7] * y1 + E[8];
  const double b0 = E[0] * x2 + E[3] * y2 + E[6], b1 = E[1] * x2 + E[4] * y2 + E[7];
  const double s = x2 * a0 + y2 * a1 + a2;
  return static_cast<float>(s * s / (a0 * a0 + a1 * a1 + b0 * b0 + b1 * b1)) <= thr ? 1 : 0;
}

__device__ int e_update_num_iters(double p, double ep, int model_points, int max_iters) {
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

__device__ __forceinline__ void e_philox(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t lo0 = 0xD2511F53u * c[0], hi0 = __umulhi(0xD2511F53u, c[0]);
    const uint32_t lo1 = 0xCD9E8D57u * c[2], hi1 = __umulhi(0xCD9E8D57u, c[2]);
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

struct EmatCam { double fx, fy, cx, cy, k1, k2; };
struct EmatDev {
  EmatCam c1, c2;
  double prob, threshold;
  int max_iters, sampler;
  unsigned long long seed;
};

// step 1 of the header: pixel -> normalised coordinate of the mean camera
__device__ __forceinline__ double2 e_normalize(float2 p, const EmatCam& own, double fx0, double fy0, double cx0, double cy0) {
  const double ifx = 1. / own.fx, ify = 1. / own.fy;
  double x = (static_cast<double>(p.x) - own.cx) * ifx, y = (static_cast<double>(p.y) - own.cy) * ify;
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    const double icdist = 1. / (1 + ((0 * r2 + own.k2) * r2 + own.k1) * r2);
    if (icdist < 0) { x = x0; y = y0; break; }
    x = x0 * icdist; y = y0 * icdist;
  }
  const float xf = static_cast<float>(x), yf = static_cast<float>(y);
  const float px = __fadd_rn(__fmul_rn(static_cast<float>(fx0), xf), static_cast<float>(cx0));
  const float py = __fadd_rn(__fmul_rn(static_cast<float>(fy0), yf), static_cast<float>(cy0));
  return make_double2((static_cast<double>(px) - cx0) / fx0, (static_cast<double>(py) - cy0) / fy0);
}

__global__ void __launch_bounds__(ES_THREADS)
emat_ransac_kernel(const float2* __restrict__ pts1, const float2* __restrict__ pts2, int M, EmatDev prm,
                   double2* __restrict__ m1, double2* __restrict__ m2, double* __restrict__ models /* [ES_ROUND][90] */,
                   uint8_t* __restrict__ mask, double* __restrict__ E_out, int32_t* __restrict__ status,
                   int32_t* __restrict__ n_inliers, int32_t* __restrict__ iters_out) {
  const int tid = threadIdx.x;
  __shared__ int sSub[ES_ROUND][5];
  __shared__ int sNm[ES_ROUND];
  __shared__ int sCnt[ES_ROUND][10];
  __shared__ double bestE[9];
  __shared__ int sGen, sStop, sIter, sNiters, sBest;
  __shared__ unsigned long long sRng;

  const double fx0 = 0.5 * (prm.c1.fx + prm.c2.fx), fy0 = 0.5 * (prm.c1.fy + prm.c2.fy);
  const double cx0 = 0.5 * (prm.c1.cx + prm.c2.cx), cy0 = 0.5 * (prm.c1.cy + prm.c2.cy);
  for (int i = tid; i < M; i += ES_THREADS) {
    m1[i] = e_normalize(pts1[i], prm.c1, fx0, fy0, cx0, cy0);
    m2[i] = e_normalize(pts2[i], prm.c2, fx0, fy0, cx0, cy0);
    mask[i] = 0;
  }
  if (tid == 0) { sRng = ~0ull; sIter = 0; sNiters = prm.max_iters; sBest = 0; sStop = 0; }
  __syncthreads();
  const double thr_n = prm.threshold / (0.5 * (fx0 + fy0));
  const float thr = static_cast<float>(thr_n * thr_n);

  if (M < 5) {
    if (tid == 0) {
      status[0] = 2; n_inliers[0] = 0; iters_out[0] = 0;
      for (int i = 0; i < 9; ++i) E_out[i] = 0.0;
    }
    return;
  }
  if (M == 5) {                        // one direct solve, first model kept, mask all ones (RANSAC run() with count == modelPoints)
    if (tid == 0) {
      const int idx[5] = {0, 1, 2, 3, 4};
      const int nm = five_point(m1, m2, idx, models);
      for (int i = 0; i < 9; ++i) E_out[i] = nm > 0 ? models[i] : 0.0;
      for (int i = 0; i < 5; ++i) mask[i] = nm > 0 ? 1 : 0;
      status[0] = nm > 0 ? 1 : 2; n_inliers[0] = nm > 0 ? 5 : 0; iters_out[0] = 1;
    }
    return;
  }

  while (true) {
    // ---- sample --------------------------------------------------------------------------------------
    if (prm.sampler == 1) {
      const int want = min(ES_ROUND, sNiters - sIter);
      if (tid == 0) sGen = want > 0 ? want : 0;
      if (tid < want) {
        int filled = 0, v5[5];
        for (uint32_t blk = 0; filled < 5 && blk < 256; ++blk) {
          uint32_t c[4] = {static_cast<uint32_t>(sIter + tid), blk, 0u, 0x504D5245u};
          e_philox(c, static_cast<uint32_t>(prm.seed), static_cast<uint32_t>(prm.seed >> 32));
          for (int w = 0; w < 4 && filled < 5; ++w) {
            const int v = static_cast<int>(__umulhi(c[w], static_cast<uint32_t>(M)));
            bool dup = false;
            for (int j = 0; j < filled; ++j) dup |= v5[j] == v;
            if (!dup) v5[filled++] = v;
          }
        }
        for (int j = 0; j < 5; ++j) sSub[tid][j] = v5[j];     // (5 distinct values exist for M >= 6 within a few words)
      }
    } else if (tid == 0) {
      unsigned long long s = sRng;
      int g = 0;
      for (; g < ES_ROUND && sIter + g < sNiters; ++g) {
        for (int i = 0; i < 5;) {
          s = static_cast<unsigned long long>(static_cast<unsigned int>(s)) * 4164903690ull + (s >> 32);
          const int v = static_cast<int>(static_cast<unsigned int>(s) % static_cast<unsigned int>(M));
          int j = 0;
          for (; j < i; ++j) if (sSub[g][j] == v) break;
          if (j < i) continue;
          sSub[g][i++] = v;
        }
      }
      sGen = g;
      sRng = s;
    }
    __syncthreads();
    const int gen = sGen;
    if (gen == 0) break;
    // ---- solve: one thread per sample ----------------------------------------------------------------------
    if (tid < gen) sNm[tid] = five_point(m1, m2, sSub[tid], models + 90 * tid);
    for (int w = tid; w < gen * 10; w += ES_THREADS) sCnt[w / 10][w % 10] = 0;
    __syncthreads();
    // ---- score: every thread strides over the matches, all models of the round -----------------------------------
    for (int g = 0; g < gen; ++g) {
      const int nm = sNm[g];
      for (int m = 0; m < nm; ++m) {
        double E[9];
        for (int k = 0; k < 9; ++k) E[k] = models[90 * g + 9 * m + k];
        int good = 0;
        for (int i = tid; i < M; i += ES_THREADS) good += e_inlier(E, m1[i], m2[i], thr);
        good = __reduce_add_sync(0xffffffffu, good);
        if ((tid & 31) == 0 && good) atomicAdd(&sCnt[g][m], good);
      }
    }
    __syncthreads();
    // ---- select: the sequential rule, sample-major / model-minor -------------------------------------------------
    if (tid == 0) {
      int best = sBest, niters = sNiters, g = 0;
      for (; g < gen && sIter + g < niters; ++g)
        for (int m = 0; m < sNm[g]; ++m) {
          const int c = sCnt[g][m];
          if (c > (best > 4 ? best : 4)) {
            best = c;
            for (int k = 0; k < 9; ++k) bestE[k] = models[90 * g + 9 * m + k];
            niters = e_update_num_iters(prm.prob, static_cast<double>(M - c) / M, 5, niters);
          }
        }
      sBest = best; sNiters = niters; sIter += g;
      if (sIter >= niters) sStop = 1;
    }
    __syncthreads();
    if (sStop) break;
  }
  __syncthreads();
  const int best = sBest;
  if (best > 0) {
    double E[9];
    for (int k = 0; k < 9; ++k) E[k] = bestE[k];
    for (int i = tid; i < M; i += ES_THREADS) mask[i] = static_cast<uint8_t>(e_inlier(E, m1[i], m2[i], thr));
  }
  if (tid == 0) {
    status[0] = best > 0 ? 1 : 2;
    n_inliers[0] = best;
    iters_out[0] = sIter;
    for (int i = 0; i < 9; ++i) E_out[i] = best > 0 ? bestE[i] : 0.0;
  }
}

cudaError_t launch_emat_ransac(const float2* pts1, const float2* pts2, int M, const double cam1[6], const double cam2[6],
                               double prob, double threshold, int max_iters, int sampler, unsigned long long seed,
                               void* scratch, uint8_t* mask, double* E, int32_t* status, int32_t* n_inliers,
                               int32_t* iters, cudaStream_t st) {
  EmatDev prm;
  prm.c1 = EmatCam{cam1[0], cam1[1], cam1[2], cam1[3], cam1[4], cam1[5]};
  prm.c2 = EmatCam{cam2[0], cam2[1], cam2[2], cam2[3], cam2[4], cam2[5]};
  prm.prob = prob; prm.threshold = threshold; prm.max_iters = max_iters; prm.sampler = sampler; prm.seed = seed;
  double2* m1 = static_cast<double2*>(scratch);
  double2* m2 = m1 + (M > 0 ? M : 1);
  double* models = reinterpret_cast<double*>(m2 + (M > 0 ? M : 1));
  emat_ransac_kernel<<<1, ES_THREADS, 0, st>>>(pts1, pts2, M, prm, m1, m2, models, mask, E, status, n_inliers, iters);
  return cudaGetLastError();
}
size_t emat_scratch_bytes(int M) { return 2 * sizeof(double2) * static_cast<size_t>(M > 0 ? M : 1) + sizeof(double) * 90 * ES_ROUND; }

}  // namespace pm
