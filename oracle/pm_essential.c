/*
 * pm_essential.c -- CPU ORACLE of GeometricFilter::estimateEssential.  TEST INFRASTRUCTURE ONLY (see pm_oracle.h).
 *
 * Restates (paths relative to the reference tree)
 *     GeometricFilter::estimateEssential           Mapper/libMapper/GeometricFilter.cpp:10-37
 *       = cv::findEssentialMat(p1, p2, K1, dist1, K2, dist2)   (:26-31; defaults RANSAC, prob 0.999, threshold 1.0)
 *     called once per reconstruction from SequentialReconstructor::chooseInitialPair, .cpp:355
 * OpenCV (>= 4.5 for this overload; not vendored) carries the arithmetic; its published algorithm is restated here:
 *   1. undistortPoints with each camera (radial k1, k2 of PinholeCamera, Camera.h:113-123; 5 fixed-point iterations),
 *      result stored as float; both sets mapped back to pixels of the MEAN camera K0 = (K1 + K2) / 2 (float transform);
 *   2. findEssentialMat(p1, p2, K0): normalised coordinates (p - c0) / f0 in double, threshold / ((fx0 + fy0) / 2);
 *   3. RANSAC (the registrator of findFundamentalMat: fixed-seed MWC stream, "strictly more inliers replaces", adaptive
 *      stop) with modelPoints = 5, NO subset check, Nister's five-point solver (<= 10 models per sample), residual =
 *      Sampson distance (x2' E x1)^2 / (|E x1|_xy^2 + |E' x2|_xy^2) stored as float and compared with (float)(t * t).
 * What cannot be pinned bit for bit against cv2: the ORDER of the models of one sample (it follows OpenCV's SVD basis
 * of the null space and the iteration of its polynomial solver) -- it matters only when two models of one sample tie;
 * the goldens count those scenes.  The reference never passes the mask to cv::findEssentialMat (:25-33), so its
 * `inlierMatchIds` stays empty; the mask is computed here nevertheless (tests use it to pin the RANSAC).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pm_oracle.h"

/* monomials of a cubic in (x, y, z), Nister's order: the first ten are eliminated, the last ten are
 * x * (z^2, z, 1), y * (z^2, z, 1), (z^3, z^2, z, 1) */
static const int MONO3[20][3] = {
  {3,0,0},{0,3,0},{2,1,0},{1,2,0},{2,0,1},{2,0,0},{0,2,1},{0,2,0},{1,1,1},{1,1,0},
  {1,0,2},{1,0,1},{1,0,0},{0,1,2},{0,1,1},{0,1,0},{0,0,3},{0,0,2},{0,0,1},{0,0,0}};
static int mono3_index(int a, int b, int c) {
  for (int i = 0; i < 20; ++i) if (MONO3[i][0] == a && MONO3[i][1] == b && MONO3[i][2] == c) return i;
  return -1;
}
/* polynomials of total degree <= 3 in (x, y, z) as dense arrays over exponents [a][b][c], a, b, c <= 3 */
typedef struct { double c[4][4][4]; } poly3;
static void p_zero(poly3* p) { memset(p, 0, sizeof *p); }
static void p_lin(poly3* p, double x, double y, double z, double w) {
  p_zero(p); p->c[1][0][0] = x; p->c[0][1][0] = y; p->c[0][0][1] = z; p->c[0][0][0] = w;
}
static void p_mul(poly3* r, const poly3* a, const poly3* b) {      /* terms above degree 3 cannot occur in use */
  poly3 t; p_zero(&t);
  for (int a0 = 0; a0 < 4; ++a0) for (int a1 = 0; a1 + a0 < 4; ++a1) for (int a2 = 0; a2 + a1 + a0 < 4; ++a2) {
    const double va = a->c[a0][a1][a2];
    if (va == 0.0) continue;
    for (int b0 = 0; b0 + a0 < 4; ++b0) for (int b1 = 0; b1 + a1 + b0 + a0 < 4; ++b1)
      for (int b2 = 0; b2 + a2 + b1 + a1 + b0 + a0 < 4; ++b2)
        t.c[a0 + b0][a1 + b1][a2 + b2] += va * b->c[b0][b1][b2];
  }
  *r = t;
}
static void p_axpy(poly3* r, double s, const poly3* a) {
  for (int i = 0; i < 64; ++i) (&r->c[0][0][0])[i] += s * (&a->c[0][0][0])[i];
}

/* null space of the 5 x 9 system: Householder QR of Q^T (9 x 5); basis k = H1..H5 e_(5+k) */
static int null_space_5x9(const double Q[5][9], double N[4][9]) {
  double B[9][5], vn2[5];
  for (int i = 0; i < 5; ++i) for (int j = 0; j < 9; ++j) B[j][i] = Q[i][j];
  for (int k = 0; k < 5; ++k) {
    double nrm2 = 0;
    for (int i = k; i < 9; ++i) nrm2 += B[i][k] * B[i][k];
    const double nrm = sqrt(nrm2);
    if (!(nrm > 0)) return 0;
    const double alpha = B[k][k] > 0 ? -nrm : nrm;
    B[k][k] -= alpha;
    double s2 = 0;
    for (int i = k; i < 9; ++i) s2 += B[i][k] * B[i][k];
    vn2[k] = s2;
    if (!(s2 > 0)) return 0;
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
  return 1;
}

/* The ten cubic constraints det(E) = 0, 2 E E' E - trace(E E') E = 0 on E = x X + y Y + z Z + W: 10 x 20 */
static void constraint_matrix(const double N[4][9], double A[10][20]) {
  poly3 e[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    p_lin(&e[i][j], N[0][3 * i + j], N[1][3 * i + j], N[2][3 * i + j], N[3][3 * i + j]);
  poly3 eet[3][3], tr, t;
  p_zero(&tr);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    p_zero(&eet[i][j]);
    for (int k = 0; k < 3; ++k) { p_mul(&t, &e[i][k], &e[j][k]); p_axpy(&eet[i][j], 1.0, &t); }
    if (i == j) p_axpy(&tr, 1.0, &eet[i][j]);
  }
  poly3 rows[10];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    poly3* r = &rows[3 * i + j];
    p_zero(r);
    for (int k = 0; k < 3; ++k) { p_mul(&t, &eet[i][k], &e[k][j]); p_axpy(r, 2.0, &t); }
    p_mul(&t, &tr, &e[i][j]); p_axpy(r, -1.0, &t);
  }
  {   /* det(E) */
    poly3* r = &rows[9];
    poly3 m, d;
    p_zero(r);
    p_mul(&m, &e[1][1], &e[2][2]); p_mul(&d, &e[1][2], &e[2][1]); p_axpy(&m, -1.0, &d); p_mul(&t, &e[0][0], &m); p_axpy(r, 1.0, &t);
    p_mul(&m, &e[1][0], &e[2][2]); p_mul(&d, &e[1][2], &e[2][0]); p_axpy(&m, -1.0, &d); p_mul(&t, &e[0][1], &m); p_axpy(r, -1.0, &t);
    p_mul(&m, &e[1][0], &e[2][1]); p_mul(&d, &e[1][1], &e[2][0]); p_axpy(&m, -1.0, &d); p_mul(&t, &e[0][2], &m); p_axpy(r, 1.0, &t);
  }
  for (int r = 0; r < 10; ++r)
    for (int m = 0; m < 20; ++m) A[r][m] = rows[r].c[MONO3[m][0]][MONO3[m][1]][MONO3[m][2]];
  (void)mono3_index;
}

/* reduced row echelon form on the first ten columns (partial pivoting); 0 when singular */
static int gauss_jordan_10x20(double A[10][20]) {
  for (int c = 0; c < 10; ++c) {
    int p = c;
    for (int r = c + 1; r < 10; ++r) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
    if (!(fabs(A[p][c]) > 1e-300)) return 0;
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
  return 1;
}

/* all roots of c[0] + c[1] z + ... + c[n] z^n (Durand-Kerner / Weierstrass iteration, the method of cv::solvePoly:
 * start at powers of 1 + i, Gauss-Seidel updates); returns the degree actually solved */
static int poly_roots(const double* c, int n, double* re, double* im) {
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

/* Nister's five-point solver on normalised coordinates; E receives up to 10 unit-Frobenius-norm models */
int orc_five_point(const double* m1, const double* m2, double* Es) {
  double Q[5][9], N[4][9], A[10][20];
  for (int i = 0; i < 5; ++i) {
    const double x1 = m1[2 * i], y1 = m1[2 * i + 1], x2 = m2[2 * i], y2 = m2[2 * i + 1];
    Q[i][0] = x1 * x2; Q[i][1] = y1 * x2; Q[i][2] = x2;
    Q[i][3] = x1 * y2; Q[i][4] = y1 * y2; Q[i][5] = y2;
    Q[i][6] = x1;      Q[i][7] = y1;      Q[i][8] = 1.0;
  }
  if (!null_space_5x9(Q, N)) return 0;
  constraint_matrix(N, A);
  if (!gauss_jordan_10x20(A)) return 0;
  /* rows 4..9 hold <x^2 z>, <x^2>, <y^2 z>, <y^2>, <xyz>, <xy>: row(2i+4) - z * row(2i+5) is free of the eliminated
   * monomials.  B[i] = [x: z^3 z^2 z 1 | y: z^3 z^2 z 1 | 1: z^4 z^3 z^2 z 1] */
  double B[3][13];
  for (int i = 0; i < 3; ++i) {
    const double* r1 = A[2 * i + 4];
    const double* r2 = A[2 * i + 5];
    double a[13], b[13];
    memset(a, 0, sizeof a); memset(b, 0, sizeof b);
    for (int k = 0; k < 3; ++k) { a[1 + k] = r1[10 + k]; a[5 + k] = r1[13 + k]; b[k] = r2[10 + k]; b[4 + k] = r2[13 + k]; }
    for (int k = 0; k < 4; ++k) { a[9 + k] = r1[16 + k]; b[8 + k] = r2[16 + k]; }
    for (int k = 0; k < 13; ++k) B[i][k] = a[k] - b[k];
  }
  /* det of the 3 x 3 polynomial matrix [cubic cubic quartic] in z: degree 10.  Coefficients by convolution, lowest
   * power first (the B rows store the highest power first). */
  double P[3][3][5];
  memset(P, 0, sizeof P);
  for (int i = 0; i < 3; ++i) {
    for (int k = 0; k < 4; ++k) { P[i][0][k] = B[i][3 - k]; P[i][1][k] = B[i][7 - k]; }
    for (int k = 0; k < 5; ++k) P[i][2][k] = B[i][12 - k];
  }
  double c[11];
  memset(c, 0, sizeof c);
  static const int perm[6][3] = {{0,1,2},{1,2,0},{2,0,1},{0,2,1},{2,1,0},{1,0,2}};
  for (int p = 0; p < 6; ++p) {
    const double sg = p < 3 ? 1.0 : -1.0;
    const double* f0 = P[perm[p][0]][0];      /* row perm[p][col] supplies column col */
    const double* f1 = P[perm[p][1]][1];
    const double* f2 = P[perm[p][2]][2];
    double t01[7];
    memset(t01, 0, sizeof t01);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) t01[i + j] += f0[i] * f1[j];
    for (int i = 0; i < 7; ++i) for (int j = 0; j < 5; ++j) c[i + j] += sg * t01[i] * f2[j];
  }
  double re[10], im[10];
  const int deg = poly_roots(c, 10, re, im);
  /* real roots in ascending order (OpenCV's order follows its SVD basis and solver iteration: not reproducible) */
  double zs[10];
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
    /* null vector of bz: the largest of the three row cross products */
    double best[3] = {0, 0, 0}, bn = -1;
    for (int a = 0; a < 3; ++a) for (int b = a + 1; b < 3; ++b) {
      const double v[3] = { bz[a][1] * bz[b][2] - bz[a][2] * bz[b][1], bz[a][2] * bz[b][0] - bz[a][0] * bz[b][2],
                            bz[a][0] * bz[b][1] - bz[a][1] * bz[b][0] };
      const double nn = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
      if (nn > bn) { bn = nn; best[0] = v[0]; best[1] = v[1]; best[2] = v[2]; }
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

/* cv::undistortPoints + the transform to the mean camera + findEssentialMat's own normalisation (see the header) */
void orc_normalize_for_essential(const float* xy, int n, const orc_camera* own, const orc_camera* c1, const orc_camera* c2,
                                 double* out) {
  const double fx0 = 0.5 * (c1->fx + c2->fx), fy0 = 0.5 * (c1->fy + c2->fy);
  const double cx0 = 0.5 * (c1->cx + c2->cx), cy0 = 0.5 * (c1->cy + c2->cy);
  const double ifx = 1. / own->fx, ify = 1. / own->fy;
  const float m00 = (float)fx0, m02 = (float)cx0, m11 = (float)fy0, m12 = (float)cy0;
  for (int i = 0; i < n; ++i) {
    double x = ((double)xy[2 * i] - own->cx) * ifx, y = ((double)xy[2 * i + 1] - own->cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
      const double r2 = x * x + y * y;
      const double icdist = 1. / (1 + ((0 * r2 + own->k2) * r2 + own->k1) * r2);
      if (icdist < 0) { x = x0; y = y0; break; }
      x = x0 * icdist; y = y0 * icdist;
    }
    const float xf = (float)x, yf = (float)y;
    const float px = (float)((float)(m00 * xf) + m02), py = (float)((float)(m11 * yf) + m12);
    out[2 * i] = ((double)px - cx0) / fx0;
    out[2 * i + 1] = ((double)py - cy0) / fy0;
  }
}

void orc_essential_residuals(const double E[9], const double* m1, const double* m2, int n, float* err) {
  for (int i = 0; i < n; ++i) {
    const double x1 = m1[2 * i], y1 = m1[2 * i + 1], x2 = m2[2 * i], y2 = m2[2 * i + 1];
    const double a0 = E[0] * x1 + E[1] * y1 + E[2], a1 = E[3] * x1 + E[4] * y1 + E[5], a2 = E[6] * x1 + E[7] * y1 + E[8];
    const double b0 = E[0] * x2 + E[3] * y2 + E[6], b1 = E[1] * x2 + E[4] * y2 + E[7];
    const double s = x2 * a0 + y2 * a1 + a2;
    err[i] = (float)(s * s / (a0 * a0 + a1 * a1 + b0 * b0 + b1 * b1));
  }
}

typedef struct { uint64_t s; } e_rng;
static inline uint32_t e_next(e_rng* r) { r->s = (uint64_t)(uint32_t)r->s * 4164903690u + (r->s >> 32); return (uint32_t)r->s; }

int orc_find_essential(const float* xy1, const float* xy2, int n, const orc_camera* c1, const orc_camera* c2,
                       double prob, double threshold, int max_iters, int sampler, uint64_t seed,
                       double E[9], uint8_t* mask, orc_ransac_trace* trace) {
  orc_ransac_trace tr = {0, 0, -1, -1, 0, 0};
  if (trace) *trace = tr;
  memset(E, 0, 9 * sizeof(double));
  if (n < 5) return 0;
  double* m1 = (double*)malloc(sizeof(double) * 2 * (size_t)n);
  double* m2 = (double*)malloc(sizeof(double) * 2 * (size_t)n);
  orc_normalize_for_essential(xy1, n, c1, c1, c2, m1);
  orc_normalize_for_essential(xy2, n, c2, c1, c2, m2);
  const double thr_n = threshold / (0.5 * (0.5 * (c1->fx + c2->fx) + 0.5 * (c1->fy + c2->fy)));
  const float thr = (float)(thr_n * thr_n);
  double Es[90];
  int ret = 0;
  if (n == 5) {
    const int nm = orc_five_point(m1, m2, Es);
    if (nm > 0) { memcpy(E, Es, 9 * sizeof(double)); memset(mask, 1, 5); ret = 1; }
    free(m1); free(m2);
    return ret;
  }
  float* err = (float*)malloc(sizeof(float) * (size_t)n);
  uint8_t* cur = (uint8_t*)malloc((size_t)n);
  e_rng rng = { (uint64_t)-1 };
  int niters = max_iters, best = 0, iter = 0;
  for (; iter < niters; ++iter) {
    int idx[5];
    if (sampler == ORC_SAMPLER_PHILOX) {
      int filled = 0;
      for (uint32_t blk = 0; filled < 5 && blk < 256; ++blk) {
        uint32_t c[4] = { (uint32_t)iter, blk, 0u, 0x504D5245u };
        orc_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        for (int w = 0; w < 4 && filled < 5; ++w) {
          const int v = (int)(((uint64_t)c[w] * (uint64_t)(uint32_t)n) >> 32);
          int j = 0;
          for (; j < filled; ++j) if (idx[j] == v) break;
          if (j == filled) idx[filled++] = v;
        }
      }
      if (filled < 5) break;
    } else {
      for (int i = 0; i < 5;) {            /* getSubset: duplicate re-draw, no subset check for this model */
        const int v = (int)(e_next(&rng) % (uint32_t)n);
        int j = 0;
        for (; j < i; ++j) if (idx[j] == v) break;
        if (j < i) continue;
        idx[i++] = v;
      }
    }
    double s1[10], s2[10];
    for (int i = 0; i < 5; ++i) { s1[2 * i] = m1[2 * idx[i]]; s1[2 * i + 1] = m1[2 * idx[i] + 1]; s2[2 * i] = m2[2 * idx[i]]; s2[2 * i + 1] = m2[2 * idx[i] + 1]; }
    const int nm = orc_five_point(s1, s2, Es);
    for (int m = 0; m < nm; ++m) {
      orc_essential_residuals(Es + 9 * m, m1, m2, n, err);
      int good = 0;
      for (int i = 0; i < n; ++i) { cur[i] = err[i] <= thr; good += cur[i]; }
      ++tr.models_tested;
      if (good > (best > 4 ? best : 4)) {
        best = good;
        memcpy(mask, cur, (size_t)n);
        memcpy(E, Es + 9 * m, 9 * sizeof(double));
        tr.best_iter = iter; tr.best_model = m; tr.best_count = good;
        niters = orc_update_num_iters(prob, (double)(n - good) / n, 5, niters);
      }
    }
  }
  tr.iters_run = iter; tr.niters_final = niters;
  if (trace) *trace = tr;
  free(err); free(cur); free(m1); free(m2);
  return best > 0 ? 1 : 0;
}
