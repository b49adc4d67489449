/*
 * pm_oracle.c -- CPU ORACLE (test infrastructure, see pm_oracle.h).
 *
 * Plain C restatement of the reference's per-pair matching body.  The arithmetic the
 * reference delegates to OpenCV (not vendored under the reference tree; pinned only as
 * "OpenCV >= 4.2" in Mapper/CMakeLists.txt:42) is restated from its published
 * algorithm and pinned against cv2 4.13.0 outputs in tests/golden/.
 *
 * Build: see oracle/Makefile.  -ffp-contract=off is REQUIRED: the hypothesis
 * arithmetic must not be fused so that the CUDA path (compiled with -fmad=false for
 * its fp64 solver) can reproduce it operation for operation.
 */
#include "pm_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * kNN (k = 2), exact brute force.
 * Replaces knnMatch at FeatureMatcher.cpp:48-49 (FLANN there; cv::BFMatcher is the
 * exact comparator named by the north star).  Only a strictly smaller distance
 * displaces an incumbent, so the lowest train index wins ties.
 * ---------------------------------------------------------------------------------- */
int orc_knn2_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int nbytes,
                     int32_t* idx, int32_t* dist) {
  if (nbytes <= 0 || (nbytes % 8) != 0) return -1;
  const int nw = nbytes / 8;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nq; ++i) {
    uint64_t a[16];
    memcpy(a, q + (size_t)i * nbytes, (size_t)nbytes);
    int32_t b1 = INT32_MAX, b2 = INT32_MAX, i1 = -1, i2 = -1;
    for (int j = 0; j < nt; ++j) {
      uint64_t b[16];
      memcpy(b, t + (size_t)j * nbytes, (size_t)nbytes);
      int32_t d = 0;
      for (int w = 0; w < nw; ++w) d += __builtin_popcountll(a[w] ^ b[w]);
      if (d < b1) { b2 = b1; i2 = i1; b1 = d; i1 = j; }
      else if (d < b2) { b2 = d; i2 = j; }
    }
    idx[2 * i] = i1; idx[2 * i + 1] = i2;
    dist[2 * i] = b1; dist[2 * i + 1] = b2;
  }
  return 0;
}

int orc_knn2_l2(const float* q, int nq, const float* t, int nt, int dim,
                int32_t* idx, double* d2) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nq; ++i) {
    const float* a = q + (size_t)i * dim;
    double b1 = INFINITY, b2 = INFINITY;
    int32_t i1 = -1, i2 = -1;
    for (int j = 0; j < nt; ++j) {
      const float* b = t + (size_t)j * dim;
      double s = 0.0;
      for (int k = 0; k < dim; ++k) {
        double df = (double)a[k] - (double)b[k];
        s += df * df;
      }
      if (s < b1) { b2 = b1; i2 = i1; b1 = s; i1 = j; }
      else if (s < b2) { b2 = s; i2 = j; }
    }
    idx[2 * i] = i1; idx[2 * i + 1] = i2;
    d2[2 * i] = b1; d2[2 * i + 1] = b2;
  }
  return 0;
}

int orc_best_query_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int nbytes,
                           int32_t* best_q) {
  if (nbytes <= 0 || (nbytes % 8) != 0) return -1;
  const int nw = nbytes / 8;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < nt; ++j) {
    uint64_t b[16];
    memcpy(b, t + (size_t)j * nbytes, (size_t)nbytes);
    int32_t best = INT32_MAX, bi = -1;
    for (int i = 0; i < nq; ++i) {
      uint64_t a[16];
      memcpy(a, q + (size_t)i * nbytes, (size_t)nbytes);
      int32_t d = 0;
      for (int w = 0; w < nw; ++w) d += __builtin_popcountll(a[w] ^ b[w]);
      if (d < best) { best = d; bi = i; }
    }
    best_q[j] = bi;
  }
  return 0;
}

int orc_best_query_l2(const float* q, int nq, const float* t, int nt, int dim,
                      int32_t* best_q) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < nt; ++j) {
    const float* b = t + (size_t)j * dim;
    double best = INFINITY;
    int32_t bi = -1;
    for (int i = 0; i < nq; ++i) {
      const float* a = q + (size_t)i * dim;
      double s = 0.0;
      for (int k = 0; k < dim; ++k) {
        double df = (double)a[k] - (double)b[k];
        s += df * df;
      }
      if (s < best) { best = s; bi = i; }
    }
    best_q[j] = bi;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------
 * Ratio test + uniqueness.  FeatureMatcher.cpp:51-64, ratioThresh FeatureMatcher.h:45.
 * The compare is on the float (non-squared) distances with a float product, strict <.
 * Rows with fewer than two neighbours produce no match (the reference reads
 * knnMatches[i][1] unchecked, :55; SURVEY "quirks").
 * ---------------------------------------------------------------------------------- */
int orc_ratio_unique(const int32_t* idx, const float* dist, int nq, int nt, float ratio,
                     int mode, const int32_t* best_q_of_train,
                     int32_t* out_q, int32_t* out_t) {
  uint8_t* taken = (uint8_t*)calloc((size_t)(nt > 0 ? nt : 1), 1);
  int n = 0;
  for (int i = 0; i < nq; ++i) {
    const int32_t t1 = idx[2 * i], t2 = idx[2 * i + 1];
    if (t1 < 0 || t2 < 0) continue;
    const float d1 = dist[2 * i], d2 = dist[2 * i + 1];
    const float rhs = ratio * d2;
    if (!(d1 < rhs)) continue;
    if (mode == ORC_UNIQUE_FIRST_WINS) {
      if (taken[t1]) continue;           /* std::find over matchedFeatIds, :58 */
      taken[t1] = 1;
    } else if (mode == ORC_MUTUAL_NN) {
      if (!best_q_of_train || best_q_of_train[t1] != i) continue;
    }
    out_q[n] = i; out_t[n] = t1; ++n;
  }
  free(taken);
  return n;
}

/* ------------------------------------------------------------------------------------
 * cv::findFundamentalMat restated (GeometricFilter.cpp:47).
 * ---------------------------------------------------------------------------------- */

/* OpenCV's cv::RNG: 64-bit multiply-with-carry; RANSAC seeds it with (uint64)-1. */
typedef struct { uint64_t s; } orc_rng;
static inline uint32_t rng_next(orc_rng* r) {
  r->s = (uint64_t)(uint32_t)r->s * 4164903690u + (r->s >> 32);
  return (uint32_t)r->s;
}
static inline int rng_uniform(orc_rng* r, int a, int b) {
  return a == b ? a : (int)(rng_next(r) % (uint32_t)(b - a)) + a;
}

static int last_point_collinear(const float* pts /* [7][2] */, int count) {
  const int i = count - 1;
  for (int j = 0; j < i; ++j) {
    /* float subtraction first, then widened (C++ float - float semantics) */
    const double dx1 = (double)(float)(pts[2 * j] - pts[2 * i]);
    const double dy1 = (double)(float)(pts[2 * j + 1] - pts[2 * i + 1]);
    for (int k = 0; k < j; ++k) {
      const double dx2 = (double)(float)(pts[2 * k] - pts[2 * i]);
      const double dy2 = (double)(float)(pts[2 * k + 1] - pts[2 * i + 1]);
      if (fabs(dx2 * dy1 - dy2 * dx1) <=
          (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
        return 1;
    }
  }
  return 0;
}

/* Draw one 7-subset; returns 1 on success. */
static int get_subset(const float* xy1, const float* xy2, int n, orc_rng* rng,
                      int max_attempts, int32_t idx[7], float s1[14], float s2[14]) {
  int iters = 0, i = 0;
  for (; iters < max_attempts; ++iters) {
    for (i = 0; i < 7 && iters < max_attempts;) {
      const int v = rng_uniform(rng, 0, n);
      int j = 0;
      for (; j < i; ++j) if (v == idx[j]) break;
      if (j < i) continue;
      idx[i] = v;
      s1[2 * i] = xy1[2 * v]; s1[2 * i + 1] = xy1[2 * v + 1];
      s2[2 * i] = xy2[2 * v]; s2[2 * i + 1] = xy2[2 * v + 1];
      ++i;
    }
    if (i == 7 && (last_point_collinear(s1, 7) || last_point_collinear(s2, 7))) continue;
    break;
  }
  return i == 7 && iters < max_attempts;
}

int orc_sample_subsets(const float* xy1, const float* xy2, int n, int iters, int32_t* out) {
  orc_rng rng = { (uint64_t)-1 };
  float s1[14], s2[14];
  int k = 0;
  for (; k < iters; ++k)
    if (!get_subset(xy1, xy2, n, &rng, 10000, out + 7 * k, s1, s2)) break;
  return k;
}

int orc_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = p < 0 ? 0 : (p > 1 ? 1 : p);
  ep = ep < 0 ? 0 : (ep > 1 ? 1 : ep);
  double num = 1 - p; if (num < DBL_MIN) num = DBL_MIN;
  double den = 1 - pow(1 - ep, (double)model_points);
  if (den < DBL_MIN) return 0;
  num = log(num); den = log(den);
  if (den >= 0 || -num >= max_iters * (-den)) return max_iters;
  return (int)lrint(num / den);          /* cvRound: round-half-even */
}

/* Cubic a0 x^3 + a1 x^2 + a2 x + a3 = 0, real roots in the order cv::solveCubic gives. */
int orc_solve_cubic(const double c[4], double x[3]) {
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
        double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
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
      x1 = t0 * cos(t1 + (2. * M_PI / 3)) - t2;
      x2 = t0 * cos(t1 + (4. * M_PI / 3)) - t2;
      n = 3;
    } else if (d == 0) {
      if (R >= 0) { x0 = -2 * pow(R, 1. / 3) - a1 / 3; x1 = pow(R, 1. / 3) - a1 / 3; }
      else { x0 = 2 * pow(-R, 1. / 3) - a1 / 3; x1 = -pow(-R, 1. / 3) - a1 / 3; }
      x2 = 0;
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

/* Two-dimensional null space of the 7x9 epipolar constraint matrix by Householder QR of
 * its transpose: A^T = Q R, null(A) = span(Q e7, Q e8).  OpenCV takes the last two right
 * singular vectors instead; any basis of the same plane yields the same model SET
 * ("7-point solver").  Fixed operation order, no pivoting: the CUDA
 * solver repeats these loops verbatim. */
static int null_space_7x9(const double A[7][9], double f1[9], double f2[9]) {
  double B[9][7], V[7][9], vn2[7];
  for (int i = 0; i < 9; ++i) for (int j = 0; j < 7; ++j) B[i][j] = A[j][i];
  for (int k = 0; k < 7; ++k) {
    double nrm2 = 0;
    for (int i = k; i < 9; ++i) nrm2 += B[i][k] * B[i][k];
    const double nrm = sqrt(nrm2);
    if (!(nrm > 0)) return 0;
    const double alpha = B[k][k] > 0 ? -nrm : nrm;
    for (int i = 0; i < 9; ++i) V[k][i] = i < k ? 0.0 : B[i][k];
    V[k][k] -= alpha;
    double s2 = 0;
    for (int i = k; i < 9; ++i) s2 += V[k][i] * V[k][i];
    vn2[k] = s2;
    if (!(s2 > 0)) return 0;
    for (int j = k + 1; j < 7; ++j) {
      double s = 0;
      for (int i = k; i < 9; ++i) s += V[k][i] * B[i][j];
      const double f = 2 * s / s2;
      for (int i = k; i < 9; ++i) B[i][j] -= f * V[k][i];
    }
  }
  for (int e = 0; e < 2; ++e) {
    double y[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    y[7 + e] = 1.0;
    for (int k = 6; k >= 0; --k) {
      double s = 0;
      for (int i = k; i < 9; ++i) s += V[k][i] * y[i];
      const double f = 2 * s / vn2[k];
      for (int i = k; i < 9; ++i) y[i] -= f * V[k][i];
    }
    memcpy(e == 0 ? f1 : f2, y, sizeof(y));
  }
  return 1;
}

/* Un-normalised 7-point solver.  F is [3][9]; returns the number of solutions. */
int orc_seven_point(const float* m1, const float* m2, double* F) {
  double A[7][9], f1[9], f2[9], c[4], r[3];
  for (int i = 0; i < 7; ++i) {
    const double x0 = m1[2 * i], y0 = m1[2 * i + 1];
    const double x1 = m2[2 * i], y1 = m2[2 * i + 1];
    A[i][0] = x1 * x0; A[i][1] = x1 * y0; A[i][2] = x1;
    A[i][3] = y1 * x0; A[i][4] = y1 * y0; A[i][5] = y1;
    A[i][6] = x0;      A[i][7] = y0;      A[i][8] = 1.0;
  }
  if (!null_space_7x9(A, f1, f2)) return 0;

  /* F ~ lambda*f1 + (1-lambda)*f2  =>  with f1 -= f2:  F = lambda*f1 + f2,
   * det(F) = c0 l^3 + c1 l^2 + c2 l + c3. */
  for (int i = 0; i < 9; ++i) f1[i] -= f2[i];

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

  const int n = orc_solve_cubic(c, r);
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

void orc_residuals(const double F[9], const float* m1, const float* m2, int n,
                   int residual_mode, float* err) {
  for (int i = 0; i < n; ++i) {
    const double x1 = m1[2 * i], y1 = m1[2 * i + 1];
    const double x2 = m2[2 * i], y2 = m2[2 * i + 1];
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
    if (residual_mode == ORC_RESID_SAMPSON) {
      err[i] = (float)(d2 * d2 / (g1 + g2));
    } else {
      const double s2 = 1. / g2, s1 = 1. / g1;
      const double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
      err[i] = (float)(e1 > e2 ? e1 : e2);
    }
  }
}

/* ---- Philox sampler ("PHILOX(seed,i,j) = free-running") -------------------------------------
 * The subset of iteration k is a pure function of (pair seed, k): counter = (k, block, 0, 'PMRF'), key = seed.
 * Each block yields four words; word w becomes the index floor(w * n / 2^32).  Slots are filled in order,
 * duplicates are skipped, a complete 7-subset is tested like OpenCV's checkSubset (last point against all
 * pairs, both images) and on rejection the filling restarts with the NEXT words of the same stream. */
void orc_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
uint64_t orc_pair_seed(uint64_t seed, int32_t i, int32_t j) {
  return splitmix64(seed ^ splitmix64(((uint64_t)(uint32_t)i << 32) | (uint32_t)j));
}
static int philox_subset(const float* xy1, const float* xy2, int n, uint64_t seed, int iter,
                         int32_t idx[7], float s1[14], float s2[14]) {
  int filled = 0;
  for (uint32_t blk = 0; blk < 256; ++blk) {
    uint32_t c[4] = { (uint32_t)iter, blk, 0u, 0x504D5246u };
    orc_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    for (int w = 0; w < 4; ++w) {
      const int v = (int)(((uint64_t)c[w] * (uint64_t)(uint32_t)n) >> 32);
      int j = 0;
      for (; j < filled; ++j) if (idx[j] == v) break;
      if (j < filled) continue;
      idx[filled] = v;
      s1[2 * filled] = xy1[2 * v]; s1[2 * filled + 1] = xy1[2 * v + 1];
      s2[2 * filled] = xy2[2 * v]; s2[2 * filled + 1] = xy2[2 * v + 1];
      if (++filled == 7) {
        if (!last_point_collinear(s1, 7) && !last_point_collinear(s2, 7)) return 1;
        filled = 0;
      }
    }
  }
  return 0;
}
int orc_philox_subset(const float* xy1, const float* xy2, int n, uint64_t seed, int iter, int32_t idx[7]) {
  float s1[14], s2[14];
  return philox_subset(xy1, xy2, n, seed, iter, idx, s1, s2);
}

/* ---- normalised 8-point refit --------------------------------------------------------------------------------
 * cv::findFundamentalMat(FM_8POINT) (run8Point, SURVEY App. A [R]); unreachable from the reference's own call
 * (GeometricFilter.cpp:47 uses the defaults), offered as the optional refit the north star names.
 * Cyclic Jacobi for the symmetric eigenproblems: fixed sweep order, so the CUDA kernel can follow it operation for
 * operation. */
static void jacobi_sym(double* A, double* V, int n, int sweeps) {     /* A: n x n symmetric (destroyed), V: eigenvectors in columns */
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
        for (int k = 0; k < n; ++k) {                   /* columns p, q of A */
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - sn * akq;
          A[k * n + q] = sn * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {                   /* rows p, q of A */
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

int orc_eight_point(const float* m1, const float* m2, int n, const uint8_t* mask, double F[9]) {
  int cnt = 0;
  double cx1 = 0, cy1 = 0, cx2 = 0, cy2 = 0;
  for (int i = 0; i < n; ++i) {
    if (mask && !mask[i]) continue;
    cx1 += m1[2 * i]; cy1 += m1[2 * i + 1]; cx2 += m2[2 * i]; cy2 += m2[2 * i + 1];
    ++cnt;
  }
  if (cnt < 8) return 0;
  const double t = 1.0 / cnt;
  cx1 *= t; cy1 *= t; cx2 *= t; cy2 *= t;
  double sc1 = 0, sc2 = 0;
  for (int i = 0; i < n; ++i) {
    if (mask && !mask[i]) continue;
    const double x1 = m1[2 * i] - cx1, y1 = m1[2 * i + 1] - cy1, x2 = m2[2 * i] - cx2, y2 = m2[2 * i + 1] - cy2;
    sc1 += sqrt(x1 * x1 + y1 * y1);
    sc2 += sqrt(x2 * x2 + y2 * y2);
  }
  sc1 *= t; sc2 *= t;
  if (sc1 < FLT_EPSILON || sc2 < FLT_EPSILON) return 0;
  sc1 = sqrt(2.0) / sc1; sc2 = sqrt(2.0) / sc2;
  double A[81], V[81];
  memset(A, 0, sizeof A);
  for (int i = 0; i < n; ++i) {
    if (mask && !mask[i]) continue;
    const double x1 = (m1[2 * i] - cx1) * sc1, y1 = (m1[2 * i + 1] - cy1) * sc1;
    const double x2 = (m2[2 * i] - cx2) * sc2, y2 = (m2[2 * i + 1] - cy2) * sc2;
    const double r[9] = { x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0 };
    for (int j = 0; j < 9; ++j) for (int k = 0; k < 9; ++k) A[j * 9 + k] += r[j] * r[k];
  }
  jacobi_sym(A, V, 9, 60);
  int lo = 0, nz = 0;
  for (int i = 0; i < 9; ++i) {
    if (A[i * 9 + i] < A[lo * 9 + lo]) lo = i;
    if (fabs(A[i * 9 + i]) < DBL_EPSILON) ++nz;
  }
  if (nz > 1) return 0;                                  /* rank < 8: degenerate configuration */
  double F0[9];
  for (int i = 0; i < 9; ++i) F0[i] = V[i * 9 + lo];
  /* rank 2: F0 <- F0 (I - v v^T), v = right singular vector of the smallest singular value */
  double G[9], W[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    double sacc = 0;
    for (int k = 0; k < 3; ++k) sacc += F0[k * 3 + i] * F0[k * 3 + j];
    G[i * 3 + j] = sacc;
  }
  jacobi_sym(G, W, 3, 60);
  int l3 = 0;
  for (int i = 1; i < 3; ++i) if (G[i * 3 + i] < G[l3 * 3 + l3]) l3 = i;
  const double v[3] = { W[0 * 3 + l3], W[1 * 3 + l3], W[2 * 3 + l3] };
  for (int i = 0; i < 3; ++i) {
    const double fv = F0[i * 3] * v[0] + F0[i * 3 + 1] * v[1] + F0[i * 3 + 2] * v[2];
    for (int j = 0; j < 3; ++j) F0[i * 3 + j] -= fv * v[j];
  }
  /* F = T2^T F0 T1, T = [s 0 -s cx; 0 s -s cy; 0 0 1] */
  const double T1[9] = { sc1, 0, -sc1 * cx1, 0, sc1, -sc1 * cy1, 0, 0, 1 };
  const double T2[9] = { sc2, 0, -sc2 * cx2, 0, sc2, -sc2 * cy2, 0, 0, 1 };
  double X[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    double sacc = 0;
    for (int k = 0; k < 3; ++k) sacc += T2[k * 3 + i] * F0[k * 3 + j];
    X[i * 3 + j] = sacc;
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    double sacc = 0;
    for (int k = 0; k < 3; ++k) sacc += X[i * 3 + k] * T1[k * 3 + j];
    F[i * 3 + j] = sacc;
  }
  if (fabs(F[8]) > FLT_EPSILON) { const double inv = 1.0 / F[8]; for (int i = 0; i < 9; ++i) F[i] *= inv; }
  return 1;
}

int orc_find_fundamental(const float* xy1, const float* xy2, int n,
                         const orc_ransac_params* prm, double* F, uint8_t* mask,
                         orc_ransac_trace* trace) {
  orc_ransac_trace tr = {0, 0, -1, -1, 0, 0};
  if (trace) *trace = tr;
  if (n < 7) return 0;
  if (n == 7) {
    const int ns = orc_seven_point(xy1, xy2, F);
    if (ns > 0) memset(mask, 1, (size_t)n);
    return ns;
  }
  const float thr = (float)(prm->threshold * prm->threshold);
  int niters = prm->max_iters;
  int best = 0;
  float* err = (float*)malloc(sizeof(float) * (size_t)n);
  uint8_t* cur = (uint8_t*)malloc((size_t)n);
  orc_rng rng = { (uint64_t)-1 };
  int iter = 0;
  for (; iter < niters; ++iter) {
    int32_t idx[7];
    float s1[14], s2[14];
    double Fm[27];
    const int got = prm->sampler == ORC_SAMPLER_PHILOX ? philox_subset(xy1, xy2, n, prm->seed, iter, idx, s1, s2)
                                                       : get_subset(xy1, xy2, n, &rng, 10000, idx, s1, s2);
    if (!got) {
      if (iter == 0) { free(err); free(cur); return 0; }
      break;
    }
    const int nm = orc_seven_point(s1, s2, Fm);
    for (int m = 0; m < nm; ++m) {
      orc_residuals(Fm + 9 * m, xy1, xy2, n, prm->residual_mode, err);
      int good = 0;
      for (int i = 0; i < n; ++i) { cur[i] = err[i] <= thr; good += cur[i]; }
      ++tr.models_tested;
      if (good > (best > 6 ? best : 6)) {
        best = good;
        memcpy(mask, cur, (size_t)n);
        memcpy(F, Fm + 9 * m, 9 * sizeof(double));
        tr.best_iter = iter; tr.best_model = m; tr.best_count = good;
        niters = orc_update_num_iters(prm->confidence, (double)(n - good) / n, 7, niters);
      }
    }
  }
  tr.iters_run = iter; tr.niters_final = niters;
  if (trace) *trace = tr;
  free(err); free(cur);
  if (best > 0 && prm->refit_8point && best >= 8) {
    double Fr[9];
    if (orc_eight_point(xy1, xy2, n, mask, Fr)) memcpy(F, Fr, sizeof Fr);
  }
  return best > 0 ? 1 : 0;
}

/* ------------------------------------------------------------------------------------
 * Per-pair body of SequentialReconstructor::matchFeatures (.cpp:213-276).
 * ---------------------------------------------------------------------------------- */
int orc_match_pair(int desc_kind, const void* desc1, const int32_t* xy1, int n1,
                   const void* desc2, const int32_t* xy2, int n2, int dim,
                   float ratio, int unique_mode, int min_matches,
                   const orc_ransac_params* prm,
                   int32_t* out_q, int32_t* out_t, int* n_putative, double F[9]) {
  int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)(n1 > 0 ? n1 : 1));
  float* dist = (float*)malloc(sizeof(float) * 2 * (size_t)(n1 > 0 ? n1 : 1));
  int32_t* bestq = NULL;
  if (desc_kind == 1) {
    int32_t* di = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)(n1 > 0 ? n1 : 1));
    orc_knn2_hamming((const uint8_t*)desc1, n1, (const uint8_t*)desc2, n2, dim, idx, di);
    for (int i = 0; i < 2 * n1; ++i) dist[i] = (float)di[i];
    free(di);
  } else {
    double* d2 = (double*)malloc(sizeof(double) * 2 * (size_t)(n1 > 0 ? n1 : 1));
    orc_knn2_l2((const float*)desc1, n1, (const float*)desc2, n2, dim, idx, d2);
    for (int i = 0; i < 2 * n1; ++i) dist[i] = (float)sqrt(d2[i]);
    free(d2);
  }
  if (unique_mode == ORC_MUTUAL_NN) {
    bestq = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n2 > 0 ? n2 : 1));
    if (desc_kind == 1)
      orc_best_query_hamming((const uint8_t*)desc1, n1, (const uint8_t*)desc2, n2, dim, bestq);
    else
      orc_best_query_l2((const float*)desc1, n1, (const float*)desc2, n2, dim, bestq);
  }
  int m = orc_ratio_unique(idx, dist, n1, n2, ratio, unique_mode, bestq, out_q, out_t);
  free(idx); free(dist); free(bestq);
  if (n_putative) *n_putative = m;
  if (F) memset(F, 0, 9 * sizeof(double));
  if (m < min_matches || !prm) return m;          /* the "else" branch, .cpp:270-276 */

  float* p1 = (float*)malloc(sizeof(float) * 2 * (size_t)m);
  float* p2 = (float*)malloc(sizeof(float) * 2 * (size_t)m);
  uint8_t* mask = (uint8_t*)calloc((size_t)m, 1);
  for (int i = 0; i < m; ++i) {                    /* featuresToCvPoints, utils.cpp:165-177 */
    p1[2 * i] = (float)xy1[2 * out_q[i]]; p1[2 * i + 1] = (float)xy1[2 * out_q[i] + 1];
    p2[2 * i] = (float)xy2[2 * out_t[i]]; p2[2 * i + 1] = (float)xy2[2 * out_t[i] + 1];
  }
  double Fs[27];
  const int ns = orc_find_fundamental(p1, p2, m, prm, Fs, mask, NULL);
  int kept = -1;
  if (ns > 0) {                                    /* empty F => pair dropped, .cpp:253-256 */
    kept = 0;
    for (int i = 0; i < m; ++i)
      if (mask[i]) { out_q[kept] = out_q[i]; out_t[kept] = out_t[i]; ++kept; }
    if (F) memcpy(F, Fs, 9 * sizeof(double));     /* cvMatToEigen3d keeps the first 3x3 */
  }
  free(p1); free(p2); free(mask);
  return kept;
}
