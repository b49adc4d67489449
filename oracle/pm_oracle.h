/*
 * pm_oracle.h -- CPU ORACLE for the exhaustive pair-matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under reconstructor_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker / the timed CPU arm.
 *
 * What it restates (paths relative to the reference tree):
 *   - FlannMatcher::matchFeatures           Mapper/libMapper/FeatureMatcher.cpp:32-65
 *       kNN(k=2) is restated as the EXACT brute-force search (the parity arbiter the
 *       north star names: cv::BFMatcher with the same norm), ratio test :55,
 *       first-come uniqueness :58-62.
 *   - GeometricFilter::estimateFundamental  Mapper/libMapper/GeometricFilter.cpp:39-61
 *       = cv::findFundamentalMat(p1, p2, mask) with defaults.  OpenCV (>= 4.2,
 *       Mapper/CMakeLists.txt:42; not vendored, no lockfile) carries the arithmetic;
 *       its published algorithm (fixed-seed MWC sampler, 7-point minimal solver,
 *       symmetric epipolar residual, strict-improvement update with adaptive stop)
 *       is restated in pm_oracle.c and pinned against cv2 4.13.0 outputs committed
 *       under tests/golden/ (see tests/golden/make_golden.py).
 *   - the per-pair body of SequentialReconstructor::matchFeatures
 *                                           Mapper/libMapper/SequentialReconstructor.cpp:213-276
 */
#ifndef PM_ORACLE_H_
#define PM_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_UNIQUE_FIRST_WINS = 0, ORC_MUTUAL_NN = 1, ORC_UNIQUE_NONE = 2 };
enum { ORC_RESID_SYMMETRIC_EPIPOLAR = 0, ORC_RESID_SAMPSON = 1 };

enum { ORC_SAMPLER_OPENCV_MWC = 0, ORC_SAMPLER_PHILOX = 1 };

typedef struct {
  double threshold;   /* pixels; OpenCV default 3.0          */
  double confidence;  /* OpenCV default 0.99                 */
  int max_iters;      /* OpenCV default 1000                 */
  int residual_mode;  /* ORC_RESID_*                         */
  int sampler;        /* ORC_SAMPLER_*: OpenCV's fixed-seed stream (default) or the free-running Philox
                         sampler of SURVEY App. B3 (subset of iteration k = f(seed, k), no sequential state) */
  int refit_8point;   /* != 0: F is re-estimated from the inliers of the winning hypothesis with the
                         normalised 8-point algorithm (the "8-point" of the north star; the mask is kept) */
  uint64_t seed;      /* Philox key of this pair (see orc_pair_seed)                                      */
} orc_ransac_params;

typedef struct {
  int iters_run;      /* value of the loop counter at exit        */
  int niters_final;   /* adaptive bound when the loop ended       */
  int best_iter;      /* iteration that produced the winner       */
  int best_model;     /* index of the winner inside its sample    */
  int best_count;     /* inliers of the winner                    */
  int models_tested;  /* total hypotheses scored                  */
} orc_ransac_trace;

/* Exact brute-force 2-NN, Hamming norm over nbytes-byte binary descriptors.
 * idx/dist are [nq][2]; missing neighbours (nt < 2) are idx = -1, dist = INT32_MAX.
 * Ties: lowest train index first (cv::BFMatcher behaviour). */
int orc_knn2_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int nbytes,
                     int32_t* idx, int32_t* dist);

/* Exact brute-force 2-NN, squared L2 accumulated in fp64 as sum (a-b)^2.
 * idx is [nq][2], d2 is [nq][2] (squared distances).  Ties: lowest index. */
int orc_knn2_l2(const float* q, int nq, const float* t, int nt, int dim,
                int32_t* idx, double* d2);

/* Best (top-1) query for every train row; used by the MUTUAL_NN mode. */
int orc_best_query_hamming(const uint8_t* q, int nq, const uint8_t* t, int nt, int nbytes,
                           int32_t* best_q);
int orc_best_query_l2(const float* q, int nq, const float* t, int nt, int dim,
                      int32_t* best_q);

/* Lowe ratio on the (non-squared) float distances + uniqueness.
 * dist is [nq][2] float as a DMatch would hold it.  Output pairs are in ascending
 * query order; returns the number of matches. */
int orc_ratio_unique(const int32_t* idx, const float* dist, int nq, int nt, float ratio,
                     int mode, const int32_t* best_q_of_train,
                     int32_t* out_q, int32_t* out_t);

/* cv::findFundamentalMat(p1, p2, mask) restated.  xy are [n][2] float.
 * Returns number of 3x3 solutions written to F (0 = failure => the reference drops
 * the pair, GeometricFilter.cpp:50-53).  n == 7: direct 7-point, mask all ones, up
 * to 3 stacked solutions in F[27].  n >= 8: RANSAC (OpenCV switches to LMedS for
 * 8 <= n < 15; that window is noise-determined and deliberately NOT emulated,
 * SURVEY Appendix B(3)). */
int orc_find_fundamental(const float* xy1, const float* xy2, int n,
                         const orc_ransac_params* prm, double* F, uint8_t* mask,
                         orc_ransac_trace* trace);

/* Sub-oracles exposed for direct pinning against cv2.solveCubic and for tests. */
int orc_solve_cubic(const double coeffs[4], double roots[3]);
int orc_seven_point(const float* xy1, const float* xy2, double* F /* [27] */);
void orc_residuals(const double F[9], const float* xy1, const float* xy2, int n,
                   int residual_mode, float* err);
int orc_update_num_iters(double p, double ep, int model_points, int max_iters);
/* Replays the sampler: writes iters x 7 indices; returns number of subsets produced. */
int orc_sample_subsets(const float* xy1, const float* xy2, int n, int iters, int32_t* out);

/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator behind ORC_SAMPLER_PHILOX;
 * exposed so that the published known-answer vectors pin it.  ctr is updated in place. */
void orc_philox4x32_10(uint32_t ctr[4], uint32_t key0, uint32_t key1);
/* Philox key of pair (i, j) under a global seed: the result of a pair must not depend on which GPU or
 * batch ran it. */
uint64_t orc_pair_seed(uint64_t seed, int32_t img_i, int32_t img_j);
/* Subset of iteration `iter` under the Philox sampler (7 indices); returns 1, or 0 when no admissible
 * subset was found in 1024 draws. */
int orc_philox_subset(const float* xy1, const float* xy2, int n, uint64_t pair_seed, int iter, int32_t idx[7]);
/* Normalised 8-point algorithm (Hartley) over the points with mask != 0 (mask NULL: all points), as
 * cv::findFundamentalMat(..., FM_8POINT) computes it: per-image centroid / mean-distance scaling, 9x9
 * normal matrix, smallest eigenvector, rank-2 enforcement, de-normalisation, F[8] = 1.  Returns 1 / 0. */
int orc_eight_point(const float* xy1, const float* xy2, int n, const uint8_t* mask, double F[9]);

/* ---- GeometricFilter::estimateEssential (GeometricFilter.cpp:10-37) = cv::findEssentialMat(p1, p2, K1, d1, K2, d2) with
 * its defaults (RANSAC, prob 0.999, threshold 1.0, 1000 iterations); restated in pm_essential.c. ---------------------- */
typedef struct { double fx, fy, cx, cy, k1, k2; } orc_camera;     /* PinholeCamera, Camera.h:127 */
/* Nister's five-point solver on 5 normalised correspondences (m: [5][2] doubles); Es receives up to 10 models. */
int orc_five_point(const double* m1, const double* m2, double* Es /* [90] */);
/* undistort with `own`, map to the mean camera of (c1, c2) in float, normalise in double: what the RANSAC sees. */
void orc_normalize_for_essential(const float* xy, int n, const orc_camera* own, const orc_camera* c1, const orc_camera* c2,
                                 double* out /* [n][2] */);
void orc_essential_residuals(const double E[9], const double* m1, const double* m2, int n, float* err);
/* Returns 1 and E (unit Frobenius norm, sign arbitrary) + mask, or 0 (n < 5 or no model). */
int orc_find_essential(const float* xy1, const float* xy2, int n, const orc_camera* c1, const orc_camera* c2,
                       double prob, double threshold, int max_iters, int sampler, uint64_t seed,
                       double E[9], uint8_t* mask, orc_ransac_trace* trace);

/* Whole per-pair body (match -> >=7 gate -> filter -> keep inliers).
 * desc_kind: 0 = float L2 (dim floats per row), 1 = binary Hamming (dim bytes per row).
 * Returns the number of surviving matches written to out_q/out_t, or -1 when the pair
 * is dropped because F estimation failed.  n_putative receives the pre-filter count. */
int orc_match_pair(int desc_kind, const void* desc1, const int32_t* xy1, int n1,
                   const void* desc2, const int32_t* xy2, int n2, int dim,
                   float ratio, int unique_mode, int min_matches,
                   const orc_ransac_params* prm,
                   int32_t* out_q, int32_t* out_t, int* n_putative, double F[9]);

#ifdef __cplusplus
}
#endif
#endif
