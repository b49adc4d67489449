// kernels.h -- host-side launchers of the pair-matching kernels (internal to the library).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace pm {

struct RansacDev {
  double confidence;
  float thr;            // (float)(threshold * threshold)
  int max_iters;
  int residual_mode;
  int min_matches;
  int do_filter;
  int sampler;          // PM_SAMPLER_*
  int refit_8point;     // re-estimate F from the inliers of the winner (normalised 8-point)
  unsigned long long seed;   // Philox key when no per-pair jobs are given (pm_filter_pair_F_seeded)
};

// K1 -- hamming.cu
cudaError_t launch_hamming_top2(const uint32_t* bits, int words, const PairJob* jobs, int n_jobs,
                                int max_nq, int2* idx, float2* dist, int stride, int variant,
                                cudaStream_t st);
cudaError_t launch_popc_peak(uint32_t* out, int blocks, int iters, cudaStream_t st);

// K3 -- l2_simt.cu
cudaError_t launch_l2_simt(const float* desc, int dim, const PairJob* jobs, int n_jobs, int max_nq,
                           int2* idx, float2* dist, int stride, cudaStream_t st);

// K2 -- l2_tc.cu (tcgen05 / TMEM / TMA), integer-valued 128-d descriptors
static constexpr int TC_DIM = 128;        // descriptor length handled by the tensor path
static constexpr int TC_KPAD = 144;       // 128 + one K=16 step carrying the train-row norm
#ifndef PM_I8_CHUNK
#define PM_I8_CHUNK 32
#endif
static constexpr int TC_I8_CHUNK = PM_I8_CHUNK;   // columns per candidate chunk of the i8 epilogue (16 or 32)
static constexpr int TC_FP4_ROW = 160;    // 256-bit rows as E2M1 values: 128 bytes + a 64-value (32-byte) norm block
static constexpr int TC_FP4_ROW2 = 288;   // 512-bit rows: two 128-byte K atoms + the norm block
constexpr int tc_fp4_row(int words) { return words == 16 ? TC_FP4_ROW2 : TC_FP4_ROW; }
static constexpr int TC_I8_ROW = 160;     // byte form: 128 x u8/s8 + one K=32 step carrying floor(|b|^2 / 2)
struct TcMaps {
  CUtensorMap q_main, q_ext, t_main, t_ext;   // boxes of 128 rows
  CUtensorMap t_main96, t_ext96;              // boxes of 96 rows (pair kernel, 192-column tiles)
  CUtensorMap t_ext2x96;                      // 256-bit E2M1 forms: pair norm blocks (pack_bits_kernel t4x), 96-row boxes
};
cudaError_t tc_configure();               // one-time function attributes
cudaError_t launch_l2_tc(const TcMaps& maps, const int32_t* qnorm, const PairJob* jobs, int n_jobs,
                         int max_nq, int2* idx, float2* dist, int stride, int num_sms,
                         float* debug_dump, int epi, cudaStream_t st);

// K2, CTA-pair version -- l2_tc2.cu (tcgen05 cta_group::2)
cudaError_t tc2_configure();
cudaError_t launch_l2_tc2(const TcMaps& maps, const int32_t* qnorm, const PairJob* jobs, int n_jobs, int max_nq,
                          int2* idx, float2* dist, int stride, int num_sms, int variant, int mode,
                          cudaStream_t st);

// integer-valued 128-d rows as bytes on kind::i8 (l2_tc2.cu KIND 2) + l2_fixup_i8 (l2_fixup.cu).  maps: UINT8
// tensor maps over rows of TC_I8_ROW bytes.  counters: [0] rows re-evaluated, [2] rows rescanned exhaustively.
cudaError_t launch_l2i8_tc2(const TcMaps& maps, const PairJob* jobs, int n_jobs, int max_nq, int2* idx, float2* dist,
                            int stride, int num_sms, int probe, cudaStream_t st);
// two query row sets per cluster: half the L2 traffic of launch_l2i8_tc2 (l2_tc2.cu, l2_i8x2_kernel).
// variant: 0 = product (72-register build), 1 / 2 = timing probes, 3 = 64-register build
cudaError_t i8x2_configure();
cudaError_t launch_l2i8x2(const TcMaps& maps, const PairJob* jobs, int n_jobs, int max_nq, int2* idx, float2* dist,
                          int stride, int num_sms, int variant, cudaStream_t st);
cudaError_t launch_l2_fixup_i8(const uint32_t* u8desc, const int32_t* qnorm, const int32_t* qoff, const PairJob* jobs,
                               int n_jobs, int max_nq, int2* idx, float2* dist, int stride, float ratio, int all_rows,
                               unsigned long long* counters, cudaStream_t st);

// binary descriptors on the tensor cores: Hamming = |a| + |b| - 2 a.b with E4M3 {0,1} operands (l2_tc2.cu,
// kind::f8f6f4) + hamming_fixup.cu.  maps: UINT8 tensor maps over rows of 32*words + 32 bytes.
cudaError_t launch_ham_tc2(const TcMaps& maps, int words, const int32_t* qnorm, const PairJob* jobs, int n_jobs,
                           int max_nq, int2* idx, float2* dist, int stride, int num_sms, cudaStream_t st);
cudaError_t launch_pack_bits(const uint32_t* bits, int n, int words, uint8_t* qb, uint8_t* tb, int32_t* popc,
                             uint8_t* q8, uint8_t* t8, uint8_t* q4, uint8_t* t4, cudaStream_t st, uint8_t* t4x = nullptr);
// ints = 0: float (m1, m2') of 16-column chunks (marker -2, kind::f8f6f4 kernel); ints = 1: integer values of
// 32-column chunks (marker -3, kind::i8 two-set kernel, words == 8 only); ints = 2: float values of 32-column chunks
// (marker -2, kind::mxf4 two-set kernel, words == 8 only)
cudaError_t launch_hamming_fixup(const uint32_t* bits, int words, const PairJob* jobs, int n_jobs, int max_nq,
                                 int2* idx, float2* dist, int stride, float ratio, int all_rows, cudaStream_t st,
                                 int ints = 0);
// 256-bit rows as bytes on kind::i8, two query row sets per cluster (l2_tc2.cu, l2_i8x2_kernel<.., 2>)
cudaError_t launch_ham_i8x2(const TcMaps& maps, const PairJob* jobs, int n_jobs, int max_nq, int2* idx, float2* dist,
                            int stride, int num_sms, int probe, cudaStream_t st);
// 256-bit rows as E2M1 values on kind::mxf4, 192-column train tiles (l2_tc2.cu, l2_i8x2_kernel<.., 1, 1>); maps: rows of
// 160 bytes (128 bytes = 256 four-bit values + 32 bytes = 64-value norm block), t_main96 / t_ext96 boxes for the train side
cudaError_t launch_ham_fp4x2(const TcMaps& maps, const PairJob* jobs, int n_jobs, int max_nq, int2* idx, float2* dist,
                             int stride, int num_sms, int probe, cudaStream_t st, int words = 8, int packed = 0);
cudaError_t hamming_fixup_configure();

// real-valued rows on the tensor cores: l2_tc2.cu (MODE 3) + l2f_fixup.cu
cudaError_t launch_l2f_tc2(const TcMaps& maps, int dim, const PairJob* jobs, int n_jobs, int max_nq, int2* idx,
                           float2* dist, float2* extra, int stride, int num_sms, cudaStream_t st);
cudaError_t launch_pack_float(const float* raw, int n, int dim, __half* qh, __half* th, float* fnorm,
                              unsigned int* stats, uint8_t* q8, uint8_t* t8, cudaStream_t st);
// the same search with rows quantised to s8 (|x| <= L2S8_MAX_ABS) on kind::i8: half the K-steps; l2f_fixup with
// e_mode = 1 (the wider error bound of the quantised scores) follows.  maps: UINT8 maps over rows of dim + 32 bytes.
static constexpr float L2S8_MAX_ABS = 0.5f;
// rows with | |x|^2 - 1 | <= L2S8_UNIT_TOL everywhere: the s8 search drops the norm K-step (keys3 = 2, e_mode 2 in the re-rank)
static constexpr float L2S8_UNIT_TOL = 9.765625e-4f;
cudaError_t s8_configure();
cudaError_t launch_l2s8_tc2(const TcMaps& maps, int dim, const PairJob* jobs, int n_jobs, int max_nq, int2* idx,
                            float2* dist, float2* extra, int stride, int num_sms, cudaStream_t st);
// ... with two query row sets per cluster (l2_i8x2_kernel MODE 3): half the L2 traffic; the batched loop's default
// keys3 = 1 (default of the batched loop, needs nt <= L2S8_KEYS3_MAX_NT): per row (argmin key, 2nd, 3rd chunk key, second
// smallest score of the argmin's chunk) for l2f_rerank1; keys3 = 0: six chunk keys for l2f_fixup (MODE 3)
static constexpr int L2S8_KEYS3_MAX_NT = 8192;
cudaError_t launch_l2s8x2(const TcMaps& maps, int dim, const PairJob* jobs, int n_jobs, int max_nq, int2* idx,
                          float2* dist, float2* extra, int stride, int num_sms, cudaStream_t st, int keys3);
// exact fp32 re-rank of the candidate chunks (l2f_fixup.cu).  need: 0 = rows that can pass the ratio test
// (exact nearest index + exact test outcome), 1 = exact nearest index of every row, 2 = exact DMatch rows
enum { L2F_NEED_RATIO = 0, L2F_NEED_NEAREST = 1, L2F_NEED_FULL = 2 };
static constexpr int L2F_MAX_NT = 16384;   // chunk id = 10 bits of the key
static constexpr float L2F_MAX_NORM2 = 1.01f;   // rows must satisfy |x|^2 <= this (scores stay positive)
cudaError_t launch_l2f_fixup(const float* raw, const float* fnorm, int dim, const PairJob* jobs, int n_jobs,
                             int max_nq, int2* idx, float2* dist, const float2* extra, int stride, float ratio,
                             int need, unsigned long long* counters, cudaStream_t st, int e_mode = 0,
                             const uint8_t* q8 = nullptr, const uint8_t* t8 = nullptr);

// the re-rank behind launch_l2s8x2(keys3 = 1): one exact column per candidate row (l2f_rerank1_kernel), then the
// small-footprint l2f_fixup variant over the rows it left open.  flags: one byte per row of the batch (scratch).
cudaError_t launch_l2f_rerank1(const float* raw, const float* fnorm, int dim, const PairJob* jobs, int n_jobs, int max_nq,
                               int2* idx, float2* dist, float2* extra, uint8_t* flags, int stride, float ratio,
                               unsigned long long* counters, const uint8_t* q8, const uint8_t* t8, cudaStream_t st,
                               int e_mode = 1);

// shared-memory carve-out of the small tail kernels = the tensor kernel's, so that they can be co-resident
cudaError_t fixup_configure();
cudaError_t l2f_configure();
cudaError_t select_configure();
cudaError_t ransac_configure();

// l2_fixup.cu -- exact index / 2nd-neighbour recovery after the branch-free tensor epilogue
cudaError_t launch_l2_fixup(const uint32_t* u8desc, const int32_t* qnorm, const PairJob* jobs, int n_jobs,
                            int max_nq, int reversed, int2* idx, float2* dist, int stride, float ratio,
                            int all_rows, cudaStream_t st);

// pack.cu
cudaError_t launch_pack_sift(const float* raw_f32, const uint8_t* raw_u8, int n, __half* qf,
                             __half* tf, int32_t* qnorm, float* raw_out, uint32_t* u8_out,
                             int* not_integral, uint8_t* iq, uint8_t* it, int32_t* qoff, cudaStream_t st,
                             uint8_t* host_flags = nullptr);
cudaError_t launch_u8_to_f32(const uint8_t* src, float* dst, size_t n, cudaStream_t st);
cudaError_t launch_f32_to_u8(const float* src, uint8_t* dst, size_t n, int* not_integral, cudaStream_t st);

// K4 -- select.cu
cudaError_t launch_select(const PairJob* jobs, int n_jobs, const int2* knn_idx,
                          const float2* knn_dist, const int2* rev_idx, const int32_t* xy,
                          int stride, float ratio, int mode, int32_t* owner, int32_t* match_q,
                          int32_t* match_t, float2* pts1, float2* pts2, int32_t* count,
                          cudaStream_t st);
// exclusive scan of counts + gather of the per-slot slabs into contiguous arrays
cudaError_t launch_compact(const int32_t* count, int n_jobs, int stride, const int32_t* match_q,
                           const int32_t* match_t, const uint8_t* mask, int64_t* offsets,
                           int32_t* out_q, int32_t* out_t, uint8_t* out_mask, cudaStream_t st);

// K5/K6 -- ransac.cu
// jobs: per-pair Philox keys (PairJob::seed_lo/hi) of the batched loop, or nullptr (prm.seed is the key)
// workspace: ransac_workspace_bytes(n_jobs) of device memory for the staged continuation of pairs that need more than
// the first rounds (nullptr: everything stays in the per-pair kernel); n_launches receives the number of kernels queued.
struct RsState {            // per pair, between the kernels of the staged filter (ransac.cu)
  unsigned long long rng, key;
  double bestF[9];
  int iter, niters, best, gen;
  int active, handed, halt;
  float cmax;
  int nmod, pad;            // live models of the current mega-round (rs_solve_kernel appends to the list)
};
size_t ransac_workspace_bytes(int pairs);
int ransac_stage_cut();      // iterations that always run in the per-pair kernel
cudaError_t launch_ransac(const float2* pts1, const float2* pts2, const int32_t* count, int n_jobs,
                          int stride, const RansacDev& prm, uint8_t* mask, double* F,
                          int32_t* status, int32_t* n_inliers, int32_t* iters, cudaStream_t st,
                          const PairJob* jobs = nullptr, void* workspace = nullptr, int* n_launches = nullptr);

// essential.cu -- GeometricFilter::estimateEssential: five-point RANSAC of one pair (cam = fx, fy, cx, cy, k1, k2).
// scratch: emat_scratch_bytes(M) of device memory.
size_t emat_scratch_bytes(int M);
cudaError_t launch_emat_ransac(const float2* pts1, const float2* pts2, int M, const double cam1[6], const double cam2[6],
                               double prob, double threshold, int max_iters, int sampler, unsigned long long seed,
                               void* scratch, uint8_t* mask, double* E, int32_t* status, int32_t* n_inliers,
                               int32_t* iters, cudaStream_t st);

// retrieval.cu -- pair pre-selection: global descriptors, cosine similarity, top-k (d_img: (first row, n) per image)
cudaError_t launch_retrieval(const float* raw, const uint32_t* bits, int dim, int words, const int2* d_img, int n_images,
                             int top_k, double* d_gdesc, double* d_sim, int32_t* d_topk, cudaStream_t st);
cudaError_t launch_topk(double* d_sim, int n_images, int top_k, int32_t* d_topk, cudaStream_t st);

// tensor-pipe peak micro-benchmark (l2_tc2.cu, tensor_peak_kernel): kind 0 f16, 1 i8, 2 mxf4
cudaError_t launch_tensor_peak(int kind, int iters, int num_sms, double* flop_per_launch, cudaStream_t st);

// pinned host memory <-> device memory by a kernel (zero-copy), keeping the batched loop off the copy engines' queues
cudaError_t launch_copy_pinned(const void* pinned_src, void* dst, size_t bytes, cudaStream_t st);

}  // namespace pm
