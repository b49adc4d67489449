/*
 * pairmatch_b200.h -- C ABI of the B200-native exhaustive pair matcher + epipolar filter.
 *
 * Drop-in boundary for ONE hot path of smileyenot983/reconstructor (paths below are relative
 * to the reference tree):
 *
 *   SequentialReconstructor::matchFeatures(bool)   Mapper/libMapper/SequentialReconstructor.cpp:199-279
 *     -> FeatureMatcher::matchFeatures             Mapper/libMapper/FeatureMatcher.h:18-22,
 *        (FlannMatcher impl)                       Mapper/libMapper/FeatureMatcher.cpp:32-65
 *     -> GeometricFilter::estimateFundamental      Mapper/libMapper/GeometricFilter.h:33-35,
 *                                                  Mapper/libMapper/GeometricFilter.cpp:39-61
 *
 * The reference has no FFI (it is one C++ process); these are the entry points its plugin
 * classes would bind (see INTEGRATION.md for the C++ shim a maintainer adds).  Plain C
 * linkage, POD arguments only, no exceptions cross this boundary, no CPU fallback: every
 * compute entry point fails with PM_ERR_NO_DEVICE when there is no CUDA device.
 *
 * Threading: every entry point taking a pm_handle is thread-safe (the reference calls its
 * plugins from up to 4 OpenMP threads on one shared object, SequentialReconstructor.cpp:202).
 */
#ifndef PAIRMATCH_B200_H_
#define PAIRMATCH_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pm_context* pm_handle;

/* Status codes (0 = ok, negative = error; text via pm_last_error). */
enum {
  PM_OK = 0,
  PM_ERR_INVALID = -1,      /* bad argument / inconsistent descriptor shape          */
  PM_ERR_NO_DEVICE = -2,    /* no usable CUDA device (there is NO CPU fallback)      */
  PM_ERR_CUDA = -3,         /* CUDA runtime / driver error                           */
  PM_ERR_OOM = -4,          /* host or device allocation failed                      */
  PM_ERR_STATE = -5,        /* e.g. image id not set                                 */
  PM_ERR_UNSUPPORTED = -6   /* valid request this build does not implement           */
};

/* Descriptor element kinds.
 *   PM_DESC_F32      dim floats per row; what FeatDesc::desc holds (datatypes.h:70-71).
 *                    Rows that are integer-valued in [0,255] with dim == 128 (SIFT,
 *                    FeatureDetector.cpp:20-24) take the exact fp16 tensor-core path.
 *   PM_DESC_U8_BITS  dim BITS per row (dim/8 bytes), Hamming norm (ORB, 256 bits).
 *   PM_DESC_U8       dim bytes per row, integer-valued L2 (SIFT shipped as bytes). */
enum { PM_DESC_F32 = 0, PM_DESC_U8_BITS = 1, PM_DESC_U8 = 2 };

/* Uniqueness after the ratio test.
 *   PM_UNIQUE_FIRST_WINS  reference semantics, FeatureMatcher.cpp:58-62: in ascending query
 *                         order the first query claiming a train index keeps it.
 *   PM_MUTUAL_NN          cross-check: keep (q,t) iff q is also t's nearest query.
 *   PM_UNIQUE_NONE        ratio test only. */
enum { PM_UNIQUE_FIRST_WINS = 0, PM_MUTUAL_NN = 1, PM_UNIQUE_NONE = 2 };

/* Residual of the epipolar filter.
 *   PM_RESID_SYMMETRIC_EPIPOLAR  what cv::findFundamentalMat uses (GeometricFilter.cpp:47):
 *                                max of the two squared point-to-epipolar-line distances.
 *   PM_RESID_SAMPSON             first-order geometric error (north-star wording). */
enum { PM_RESID_SYMMETRIC_EPIPOLAR = 0, PM_RESID_SAMPSON = 1 };

/* Hypothesis sampler.
 *   PM_SAMPLER_OPENCV_MWC  replays cv::findFundamentalMat's fixed-seed multiply-with-carry stream, which makes
 *                          inlier masks comparable with cv2 one to one (the default; GeometricFilter.cpp:47).
 *   PM_SAMPLER_PHILOX      free-running counter-based sampler: the subset of iteration k is a pure
 *                          function of (pm_params.seed, query image id, train image id, k) -- Philox4x32-10 -- so a
 *                          pair's result does not depend on the batch / device / rank that ran it and all subsets of
 *                          a round are drawn in parallel.  Everything else (7-point solver, residual, "strictly more
 *                          inliers replaces", adaptive stop) is unchanged.  Parity: the repo's CPU filter with the
 *                          same sampler (oracle/pm_oracle.c, ORC_SAMPLER_PHILOX). */
enum { PM_SAMPLER_OPENCV_MWC = 0, PM_SAMPLER_PHILOX = 1 };

/* Per-pair outcome (pm_csr_result.status / pm_pair_result.status). */
enum {
  PM_PAIR_UNFILTERED = 0,   /* fewer than min_matches putative matches (or filter off): all kept,
                               SequentialReconstructor.cpp:270-276                            */
  PM_PAIR_FILTERED = 1,     /* F estimated, inlier flags valid, .cpp:259-267                  */
  PM_PAIR_DROPPED = 2       /* F estimation failed: pair gets no entry, .cpp:253-256          */
};

typedef struct {
  float   ratio;              /* Lowe ratio, FeatureMatcher.h:45 (0.7)                      */
  int32_t unique_mode;        /* PM_UNIQUE_*                                                */
  int32_t min_matches;        /* >= gate before the filter, SequentialReconstructor.cpp:237 */
  int32_t do_filter;          /* matchFeatures(bool filter)                                 */
  double  ransac_threshold;   /* cv::findFundamentalMat defaults: 3.0 px                    */
  double  ransac_confidence;  /*                                  0.99                      */
  int32_t ransac_max_iters;   /*                                  1000                      */
  int32_t residual_mode;      /* PM_RESID_*                                                 */
  int32_t sampler;            /* PM_SAMPLER_*                                               */
  int32_t batch_pairs;        /* pairs per device batch, 0 = auto                           */
  int64_t reserve_keypoints;  /* device arena rows to preallocate, 0 = grow on demand       */
  int32_t debug_flags;        /* kernel-variant switches for parity cross-checks and probes;
                                 0 = product defaults (see enqueue_knn in csrc/api.cu)      */
  int32_t refit_8point;       /* != 0: F of a filtered pair (>= 8 inliers) is re-estimated from its
                                 inliers with the normalised 8-point algorithm (what
                                 cv::findFundamentalMat(FM_8POINT) computes); mask and counts
                                 stay those of the winning RANSAC hypothesis.  Not reachable from
                                 the reference's call (defaults), hence off by default.     */
  uint64_t seed;              /* PM_SAMPLER_PHILOX: global seed (per-pair key = f(seed, i, j)) */
  double  essential_confidence; /* cv::findEssentialMat defaults (GeometricFilter.cpp:26-31): 0.999 */
  double  essential_threshold;  /*                                                   1.0 px */
} pm_params;

/* One pair, caller-allocated outputs (capacity = number of query keypoints). */
typedef struct {
  int32_t  capacity;          /* in: length of q/t/inlier                                   */
  int32_t  n_matches;         /* out: putative matches after ratio + uniqueness             */
  int32_t  n_inliers;         /* out: matches that survive the filter                       */
  int32_t  status;            /* out: PM_PAIR_*                                             */
  int32_t  ransac_iters;      /* out: hypotheses iterations actually consumed               */
  int32_t* q;                 /* out: query indices, ascending                              */
  int32_t* t;                 /* out: train indices                                         */
  uint8_t* inlier;            /* out: 1 = kept                                              */
  double   F[9];              /* out: row-major, zeros unless PM_PAIR_FILTERED              */
} pm_pair_result;

/* Batched result: flat CSR over pairs, library-allocated (free with pm_free_result). */
typedef struct {
  int64_t  n_pairs;
  int32_t* pair_ij;           /* [n_pairs][2]  (query image, train image)                   */
  int64_t* offsets;           /* [n_pairs+1]   into q/t/inlier                              */
  int32_t* q;                 /* putative matches, ascending q inside a pair                */
  int32_t* t;
  uint8_t* inlier;            /* 1 = belongs to featureMatches[(i,j)]                       */
  double*  F;                 /* [n_pairs][9]                                               */
  int32_t* status;            /* [n_pairs]     PM_PAIR_*                                    */
  int32_t* n_inliers;         /* [n_pairs]                                                  */
  int32_t* ransac_iters;      /* [n_pairs]                                                  */
  double   device_ms;         /* CUDA-event time of the whole call on the device            */
  void*    owner_;            /* private                                                    */
} pm_csr_result;

typedef struct {
  int64_t pairs_matched;
  int64_t putative_matches;
  int64_t inlier_matches;
  int64_t kernel_launches;    /* launches of this library's own kernels                     */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  double  knn_ms;             /* CUDA-event time of the dominant kNN kernel, summed         */
  int64_t knn_launches;
  double  knn_work;           /* algorithmic work of those launches: FLOP (L2) or popc32    */
  int32_t device_id;
  int32_t n_images;
  /* real-valued tensor path (approximate fp16 scores + exact fp32 re-rank, l2f_fixup.cu)       */
  int64_t rerank_rows;        /* query rows whose candidate chunks were re-evaluated exactly */
  int64_t rerank_chunks;      /* 16-column chunks re-evaluated                               */
  int64_t rerank_overflow;    /* rows that needed the exhaustive exact scan                  */
  double  rerank_worst_err;   /* max observed |approx - exact| / certified bound (must be < 1) */
} pm_stats;

void        pm_default_params(pm_params* p);
/* device_ids == NULL / n_dev == 0: current device.  n_dev > 1: pairs are partitioned over the
 * devices, descriptors replicated on each, one host thread per device. */
int         pm_create(const pm_params* p, const int* device_ids, int n_dev, pm_handle* out);
int         pm_destroy(pm_handle h);
const char* pm_last_error(pm_handle h);          /* h may be NULL: last error of pm_create   */
const char* pm_version(void);

/* Packs + uploads one image once (replaces featDescToCV's per-pair cost,
 * FeatureMatcher.cpp:11-25).  desc: n rows of `dim` elements of `dtype`, row-major, host
 * memory.  xy: n x 2 int32 pixel coordinates (FeatCoord<int>, datatypes.h:12-25) or NULL.
 * All images of one handle must share dim and dtype. */
int pm_set_image(pm_handle h, int img_id, const void* desc, int n, int dim, int dtype,
                 const int32_t* xy);
/* Same, descriptors/xy already in DEVICE memory of the handle's device (ingest after an
 * NCCL all-gather of sharded extraction).  Single-device handles only. */
int pm_set_image_device(pm_handle h, int img_id, const void* d_desc, int n, int dim, int dtype,
                        const int32_t* d_xy);
/* pm_set_image (the once-per-image replacement of featDescToCV, FeatureMatcher.cpp:11-25) without the host
 * synchronisation: the upload and the packing kernels are queued on the
 * handle's ingest stream and the call returns.  desc / xy must stay valid and unchanged until
 * pm_sync_images() or a matching call that uses the image returns (pinned host memory makes the copy
 * truly asynchronous, so the first batches of pm_match_all_pairs overlap the upload of later images;
 * pageable memory is staged by the driver and behaves like pm_set_image).  pm_match_all_pairs(ALL)
 * visits every image against all earlier ones, so early batches need only the first images. */
int pm_set_image_async(pm_handle h, int img_id, const void* desc, int n, int dim, int dtype,
                       const int32_t* xy);
/* pm_set_image_async for a whole set in one call (descs[k]: ns[k] rows, xys may be NULL or hold NULLs). */
int pm_set_images_async(pm_handle h, int n_images, const int* img_ids, const void* const* descs, const int* ns,
                        int dim, int dtype, const int32_t* const* xys);
/* pm_set_image_device without the host synchronisation (same lifetime rule for the device buffers). */
int pm_set_image_device_async(pm_handle h, int img_id, const void* d_desc, int n, int dim, int dtype,
                              const int32_t* d_xy);
/* Waits until every asynchronously ingested image is resident. */
int pm_sync_images(pm_handle h);
int pm_num_keypoints(pm_handle h, int img_id);
/* Forgets an image and returns its device rows to the handle's free list (re-used by later pm_set_image calls of
 * at most that many keypoints).  PM_ERR_STATE when the id is not set. */
int pm_remove_image(pm_handle h, int img_id);

/* Collective ingest (SURVEY 8e "Collective"): feature extraction sharded over the ranks of one job, ONE process per
 * GPU, image k (0 <= k < n_images_total) extracted by rank k % n_ranks.  Every rank hands over only its own images
 * (ascending id order, n_keypoints rows each, host or device memory) and receives all of them: own images -> device ->
 * ncclAllGather over NVLink / NVSwitch, slot by slot (slot s = images s*R .. s*R+R-1) -> asynchronous ingest, so that
 * pm_match_all_pairs starts on the first images while the later ones are still being packed.  Nothing waits for the
 * device; pm_sync_images (or the first matching call that touches an image) does.
 *   wire_dtype  what travels: == dtype, or PM_DESC_U8 for PM_DESC_F32 rows of dim 128 that hold byte values (SIFT,
 *               FeatureDetector.cpp:20-24): 4x fewer bytes; the promise is checked on the device and reported by
 *               pm_sync_images (PM_ERR_INVALID).
 * The communicator is the library's own: rank 0 calls pm_comm_get_unique_id and ships the 128 bytes to the other
 * ranks by any means (MPI, a file, torch.distributed), then every rank calls pm_comm_init on its single-device
 * handle.  NCCL is bound at run time (dlopen of libnccl.so.2): PM_ERR_UNSUPPORTED when it is not installed. */
int pm_comm_get_unique_id(uint8_t id[128]);
int pm_comm_init(pm_handle h, const uint8_t id[128], int rank, int n_ranks);
int pm_ingest_allgather(pm_handle h, int n_images_total, int n_keypoints, int dim, int dtype, int wire_dtype,
                        const void* own_desc, const int32_t* own_xy, int own_on_device);

/* Raw 2-NN rows of pair (i -> j): what knnMatch(desc_i, desc_j, 2) returns as DMatch rows
 * (FeatureMatcher.cpp:49).  idx/dist are [n_i][2]; missing neighbours are idx -1, dist +inf.
 * dist is the float L2 norm (sqrt) or the Hamming count as float, like cv::DMatch::distance. */
int pm_knn_pair(pm_handle h, int img_i, int img_j, int32_t* idx, float* dist);

/* FeatureMatcher::matchFeatures for two resident images: kNN -> ratio -> uniqueness.
 * Fills n_matches/q/t only (status = PM_PAIR_UNFILTERED). */
int pm_match_pair(pm_handle h, int img_i, int img_j, pm_pair_result* out);
/* Same from raw host descriptors (what the per-pair virtual call hands over). */
int pm_match_descriptors(pm_handle h, const void* desc1, int n1, const void* desc2, int n2,
                         int dim, int dtype, pm_pair_result* out);

/* GeometricFilter::estimateFundamental: xy1/xy2 are M x 2 float pixel coordinates of
 * already matched points.  status: PM_PAIR_FILTERED, or PM_PAIR_DROPPED when no model was
 * found (the reference then returns Zero() and an empty mask, GeometricFilter.cpp:50-53). */
int pm_filter_pair_F(pm_handle h, const float* xy1, const float* xy2, int M,
                     double F[9], uint8_t* mask, int32_t* status, int32_t* iters);
/* The same with an explicit Philox key for this one call (filter_pair_F(h, xy1, xy2, M, seed, ...)):
 * with PM_SAMPLER_PHILOX the hypothesis stream is f(pair_key, iteration); with PM_SAMPLER_OPENCV_MWC the key is
 * ignored.  pm_filter_pair_F uses pm_params.seed as the key.  pm_pair_seed gives the key the batched loop derives
 * for pair (img_i, img_j), so a single pair can be re-run bit-identically. */
int pm_filter_pair_F_seeded(pm_handle h, const float* xy1, const float* xy2, int M, uint64_t pair_key,
                            double F[9], uint8_t* mask, int32_t* status, int32_t* iters);
uint64_t pm_pair_seed(uint64_t seed, int32_t img_i, int32_t img_j);

/* GeometricFilter::estimateEssential (GeometricFilter.h:23-27, GeometricFilter.cpp:10-37) = cv::findEssentialMat(p1, p2,
 * K1, dist1, K2, dist2) with its defaults: undistortion with each image's own camera, RANSAC (pm_params.
 * essential_confidence / essential_threshold / ransac_max_iters, sampler as for F) with Nister's five-point solver
 * and the Sampson residual.  The caller is SequentialReconstructor::chooseInitialPair (.cpp:355), once per
 * reconstruction.  E is row-major, of unit Frobenius norm, its sign arbitrary (as cv2's); status PM_PAIR_FILTERED or
 * PM_PAIR_DROPPED (M < 5 or no model: zeros).  The reference never passes a mask to cv::findEssentialMat
 * (GeometricFilter.cpp:25-33), so ITS inlierMatchIds stays empty; here the mask is returned when `mask` is not NULL. */
typedef struct { double fx, fy, cx, cy, k1, k2; } pm_camera;   /* PinholeCamera, Camera.h:127 */
int pm_filter_pair_E(pm_handle h, const float* xy1, const float* xy2, int M, const pm_camera* cam1,
                     const pm_camera* cam2, double E[9], uint8_t* mask, int32_t* status, int32_t* iters);

/* The whole pair body for one pair (match + gate + filter). */
int pm_match_filter_pair(pm_handle h, int img_i, int img_j, pm_pair_result* out);

/* The whole loop: pairs = [n_pairs][2] image ids (query, train).  pairs == NULL with n_pairs ==
 * PM_ALL_PAIRS: all i < j over the images set so far (the FakeImgMatcher list, ImageMatcher.cpp:6-24);
 * n_pairs == 0 is an empty list (an empty result), whatever `pairs` is.  Runs kNN -> ratio ->
 * uniqueness -> F-RANSAC batched on the device(s) and returns CSR arrays in host memory. */
#define PM_ALL_PAIRS ((int64_t)-1)
int pm_match_all_pairs(pm_handle h, const int32_t* pairs, int64_t n_pairs, pm_csr_result** out);
int pm_free_result(pm_csr_result* r);

/* Pair pre-selection by image retrieval: the plugin behind ImageMatcher::match (ImageMatcher.h:18-21)
 * that the reference only has as FakeImgMatcher -- every image with every other one, ImageMatcher.cpp:6-24 -- and lists
 * as a todo (README.md:40).  Every image of the handle gets a global descriptor (unit-length sum of its unit-length local
 * descriptors; for binary rows of the +-1 bit vectors) and is paired with its top_k most similar other images (cosine
 * similarity in fp64, ties to the lower image id).  The result is the canonical pair list pm_match_all_pairs takes:
 * [n_pairs][2], query = lower image id, the union of both directions, ordered like the implicit all-pairs list.
 * top_k <= 0 or >= n_images - 1: all pairs (= FakeImgMatcher).  scores (optional): receives the n_images x n_images
 * similarity matrix in ascending image-id order (caller-allocated).  Free the list with pm_free_pairs. */
int pm_select_pairs(pm_handle h, int top_k, int32_t** pairs_out, int64_t* n_pairs_out, double* scores);
/* The same over a subset of the handle's images (img_ids: n_ids distinct ids, any order; NULL = every image): what a
 * plugin sharing the handle with other users needs (the per-pair matcher keeps cached uploads under ids of its own).
 * scores is n_ids x n_ids in ascending id order.  PM_ERR_STATE when an id is not set. */
int pm_select_pairs_among(pm_handle h, const int32_t* img_ids, int n_ids, int top_k, int32_t** pairs_out,
                          int64_t* n_pairs_out, double* scores);
int pm_free_pairs(int32_t* pairs);

/* On-disk cache (the reference's README lists "save intermediate steps" as a todo).
 * One little-endian container format (layout: reconstructor_b200/cache.py) with a checksum; a file that
 * fails magic / version / size / checksum validation is rejected with PM_ERR_INVALID.
 *   pm_save_images / pm_load_images  every ingested image of the handle (descriptors as ingested + xy):
 *                                    resume without re-running extraction, or ship sharded extraction.
 *   pm_save_result / pm_load_result  a CSR result (featureMatches, SequentialReconstructor.h:226, in flat
 *                                    form); loading needs no device.  Free with pm_free_result.
 * Errors of the two handle-less calls are reported through pm_last_error(NULL). */
int pm_save_images(pm_handle h, const char* path);
int pm_load_images(pm_handle h, const char* path);
int pm_save_result(const pm_csr_result* r, const char* path);
int pm_load_result(const char* path, pm_csr_result** out);

int pm_get_stats(pm_handle h, pm_stats* out);
int pm_reset_stats(pm_handle h);

/* popc32 / fp32 pipe micro-benchmarks used to measure the roofline denominators that
 * MEASURED_PEAKS.json does not hold: returns ops per second. */
int pm_measure_popc_peak(pm_handle h, double* popc32_per_s);
/* Tensor-pipe peaks of the MMA kinds the kNN kernels use, measured live (TMA + tcgen05.mma only, operands resident in
 * shared memory, no epilogue; M = 256, N = 256 per CTA pair): FLOP/s (2 per multiply-add).  MEASURED_PEAKS.json holds
 * the bf16 figure only; the roofline fractions of bench.py divide by these same-kind numbers. */
enum { PM_PEAK_KIND_F16 = 0, PM_PEAK_KIND_I8 = 1, PM_PEAK_KIND_MXF4 = 2 };
int pm_measure_tensor_peak(pm_handle h, int kind, double* flop_per_s);

/* Development aid of the tensor path (raw accumulators of the first 256 x 128 block next to the kNN rows of a pair);
 * exported for the parity tests, not part of the drop-in surface. */
int pm_debug_tc_dump(pm_handle h, int img_i, int img_j, int32_t* idx, float* dist, float* acc256x128);

#ifdef __cplusplus
}
#endif
#endif
