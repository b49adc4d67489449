// api.cu -- host side of libpairmatch_b200.so: device arenas, batch scheduler, C ABI.
//
// Mirrors the reference's pair loop (Mapper/libMapper/SequentialReconstructor.cpp:199-279) as a
// batched device pipeline:   kNN (K1/K2/K3) -> ratio+uniqueness (K4) -> F-RANSAC (K5/K6) ->
// CSR compaction -> pinned D2H, several batches in flight on separate streams.
// There is no CPU fallback anywhere in this file: without a CUDA device every compute entry
// point returns PM_ERR_NO_DEVICE.
#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/pairmatch_b200.h"
#include "kernels.h"
#include "nccl_dyn.h"

namespace {

using namespace pm;

thread_local std::string g_create_error;

struct Image {
  int32_t row = 0;       // first arena row
  int32_t n = 0;         // keypoints
  int32_t cap = 0;       // rows reserved
  bool integral = false; // u8-valued rows (exact tensor path allowed)
  bool i8_ok = false;    // ... and every row norm fits the byte form's norm block (kind::i8 path allowed)
  bool unit_ok = false;  // finite real-valued rows with |x|^2 <= L2F_MAX_NORM2 and fp16 forms packed
  bool s8_ok = false;    // ... and max |x| <= L2S8_MAX_ABS: the quantised s8 forms are valid (kind::i8 path)
  bool unit1_ok = false; // ... and every squared row norm within L2S8_UNIT_TOL of 1: the search may drop the norm K-step
  float maxn = 0.f;      // largest squared row norm
  bool has_xy = false;
  // asynchronous ingest (pm_set_image_async): the upload + packing kernels are queued on the ingest stream and
  // the per-image facts (integral? unit norm?) arrive in a pinned record; resolved at first use
  bool pending = false;
  cudaEvent_t ready = nullptr;
  int rec = -1;          // index of the pinned record
};

constexpr int kTmpA = INT32_MIN, kTmpB = INT32_MIN + 1;

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline uint64_t pair_seed(uint64_t seed, int32_t i, int32_t j) {
  return splitmix64(seed ^ splitmix64((static_cast<uint64_t>(static_cast<uint32_t>(i)) << 32) | static_cast<uint32_t>(j)));
}

// Pinned host buffers are expensive to create, so results recycle them through a pool that
// outlives the handle if a result is freed late.
struct PinnedPool {
  std::mutex mu;
  std::vector<std::pair<void*, size_t>> free_list;
  void* acquire(size_t bytes, size_t* got) {
    {
      std::lock_guard<std::mutex> lk(mu);
      size_t best = free_list.size();
      for (size_t i = 0; i < free_list.size(); ++i)
        if (free_list[i].second >= bytes && (best == free_list.size() || free_list[i].second < free_list[best].second))
          best = i;
      if (best != free_list.size()) {
        auto pr = free_list[best];
        free_list.erase(free_list.begin() + best);
        *got = pr.second;
        return pr.first;
      }
    }
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    *got = bytes;
    return p;
  }
  void release(void* p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu);
    free_list.emplace_back(p, bytes);
  }
  ~PinnedPool() {
    for (auto& pr : free_list) cudaFreeHost(pr.first);
  }
};

struct Result {  // owner of a pm_csr_result
  pm_csr_result pub{};
  std::vector<int32_t> pair_ij, status, n_inliers, iters;
  std::vector<int64_t> offsets;
  std::vector<double> F;
  // match arrays live in pinned memory: the device compacted arrays are copied straight into them
  std::shared_ptr<PinnedPool> pool;
  int32_t *q = nullptr, *t = nullptr;
  uint8_t* inlier = nullptr;
  size_t q_bytes = 0, t_bytes = 0, inl_bytes = 0;
  int64_t cap = 0, size = 0;
  std::vector<int32_t> vq, vt;      // results loaded from a cache file live in plain host memory
  std::vector<uint8_t> vin;

  bool reserve(int64_t want) {      // caller guarantees no copy into the old buffers is in flight
    if (want <= cap) return true;
    const int64_t nc = std::max<int64_t>(want, cap * 2);
    size_t gq = 0, gt = 0, gi = 0;
    int32_t* nq = static_cast<int32_t*>(pool->acquire(4 * static_cast<size_t>(nc), &gq));
    int32_t* nt = static_cast<int32_t*>(pool->acquire(4 * static_cast<size_t>(nc), &gt));
    uint8_t* ni = static_cast<uint8_t*>(pool->acquire(static_cast<size_t>(nc), &gi));
    if (!nq || !nt || !ni) return false;
    if (size > 0) {
      std::memcpy(nq, q, 4 * static_cast<size_t>(size));
      std::memcpy(nt, t, 4 * static_cast<size_t>(size));
      std::memcpy(ni, inlier, static_cast<size_t>(size));
    }
    pool->release(q, q_bytes); pool->release(t, t_bytes); pool->release(inlier, inl_bytes);
    q = nq; t = nt; inlier = ni; q_bytes = gq; t_bytes = gt; inl_bytes = gi;
    cap = std::min<int64_t>({static_cast<int64_t>(gq / 4), static_cast<int64_t>(gt / 4), static_cast<int64_t>(gi)});
    return true;
  }
  ~Result() {
    if (pool) { pool->release(q, q_bytes); pool->release(t, t_bytes); pool->release(inlier, inl_bytes); }
  }
};

#define PM_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) return fail_cuda(e_, #call, __LINE__);                       \
  } while (0)

// ------------------------------------------------------------------------------------------------
struct Slot {
  int cap_pairs = 0, stride = 0;
  bool mutual = false;
  bool unit1 = false;                                   // the batch's kNN kernel ran without the norm K-step (e_mode 2 re-rank)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_done = nullptr, ev_k0 = nullptr, ev_k1 = nullptr, ev_jobs = nullptr, ev_knn = nullptr;
  cudaEvent_t ev_tr[3] = {nullptr, nullptr, nullptr};   // PM_TRACE: after the fix-up / selection / RANSAC of the batch
  // device
  PairJob *d_jobs = nullptr, *d_rjobs = nullptr;
  int2 *knn_idx = nullptr, *rev_idx = nullptr;
  float2 *knn_dist = nullptr, *rev_dist = nullptr;
  float2 *knn_extra = nullptr, *rev_extra = nullptr;   // 5th / 6th candidate keys of the real-valued tensor path
  uint8_t* rr_flag = nullptr;                           // rows l2f_rerank1_kernel left open (real-valued tensor path)
  int32_t *owner = nullptr, *match_q = nullptr, *match_t = nullptr, *count = nullptr;
  float2 *pts1 = nullptr, *pts2 = nullptr;
  uint8_t* mask = nullptr;
  double* F = nullptr;
  int32_t *status = nullptr, *n_inl = nullptr, *iters = nullptr;
  int64_t* offsets = nullptr;
  int32_t *out_q = nullptr, *out_t = nullptr;
  uint8_t* out_mask = nullptr;
  void* rs_ws = nullptr;                                // staged continuation of the epipolar filter (ransac.cu)
  // pinned host
  PairJob *h_jobs = nullptr, *h_rjobs = nullptr;
  int64_t* h_offsets = nullptr;
  int32_t *h_q = nullptr, *h_t = nullptr, *h_status = nullptr, *h_ninl = nullptr, *h_iters = nullptr,
          *h_count = nullptr;
  uint8_t* h_mask = nullptr;
  double* h_F = nullptr;
  // in-flight bookkeeping
  bool busy = false;
  int n_jobs = 0;
  int64_t first_pair = 0;
  double knn_work = 0;
  bool timed = false;

  void release() {
    auto fd = [](auto*& p) { if (p) { cudaFree(p); p = nullptr; } };
    auto fh = [](auto*& p) { if (p) { cudaFreeHost(p); p = nullptr; } };
    fd(d_jobs); fd(d_rjobs); fd(knn_idx); fd(rev_idx); fd(knn_dist); fd(rev_dist); fd(knn_extra); fd(rev_extra); fd(rr_flag); fd(owner);
    fd(match_q); fd(match_t); fd(count); fd(pts1); fd(pts2); fd(mask); fd(F); fd(status);
    fd(n_inl); fd(iters); fd(offsets); fd(out_q); fd(out_t); fd(out_mask); fd(rs_ws);
    fh(h_jobs); fh(h_rjobs); fh(h_offsets); fh(h_q); fh(h_t); fh(h_status); fh(h_ninl);
    fh(h_iters); fh(h_count); fh(h_mask); fh(h_F);
    cap_pairs = stride = 0;
  }
  void destroy() {
    release();
    if (stream) cudaStreamDestroy(stream);
    if (ev_done) cudaEventDestroy(ev_done);
    if (ev_k0) cudaEventDestroy(ev_k0);
    if (ev_k1) cudaEventDestroy(ev_k1);
    if (ev_jobs) cudaEventDestroy(ev_jobs);
    if (ev_knn) cudaEventDestroy(ev_knn);
    for (auto& e : ev_tr) if (e) { cudaEventDestroy(e); e = nullptr; }
    stream = nullptr; ev_done = ev_k0 = ev_k1 = ev_jobs = ev_knn = nullptr;
  }
};

struct DeviceCtx {
  int dev = 0;
  int num_sms = 148;
  pm_params prm{};
  std::string err;
  pm_stats stats{};

  // descriptor format of this handle (fixed by the first image)
  int dim = 0, dtype = -1, words = 0;
  bool tc_ready = false;

  // arenas, all indexed by row
  int64_t cap_rows = 0, used_rows = 0;
  float* raw = nullptr;        // [rows][dim]    fp32 (F32 / U8 dtypes)
  __half *qf = nullptr, *tf = nullptr;   // [rows][144] tensor-path operand forms
  int32_t* qnorm = nullptr;    // [rows]
  __half *fq = nullptr, *ft = nullptr;   // [rows][dim+16] operand forms of real-valued rows (dim 128 / 256)
  float* fnorm = nullptr;      // [rows] fp32 squared norms of real-valued rows
  TcMaps fmaps{};
  bool tcf_ready = false;
  uint8_t *sq8 = nullptr, *st8 = nullptr;   // [rows][dim+32] s8 operand forms of real-valued rows (kind::i8)
  TcMaps smaps{};
  bool tcs_ready = false;
  uint8_t *hq = nullptr, *ht = nullptr;  // [rows][32*words+32] E4M3 operand forms of binary rows (256 / 512 bit)
  TcMaps hmaps{};
  bool tch_ready = false;
  uint8_t *hq8 = nullptr, *ht8 = nullptr;   // [rows][288] byte forms of 256-bit rows (kind::i8 two-set kernel)
  TcMaps h8maps{};
  bool tch8_ready = false;
  uint8_t* ht4x = nullptr;                   // [rows][32] pair norm blocks of the 256-bit E2M1 forms (l2_i8x2_kernel PK)
  uint8_t *hq4 = nullptr, *ht4 = nullptr;   // [rows][160 / 288] E2M1 forms of 256- / 512-bit rows (kind::mxf4 two-set kernel)
  TcMaps h4maps{};                           // train boxes of 96 rows (192-column tiles)
  bool tch4_ready = false;
  // development switches, read from the environment ONCE in init() (never in the batch loop):
  bool result_by_ce = true;                  // PM_RESULT_COPY=kernel: compacted matches leave by copy kernels too
  bool trace_on = false;                     // PM_TRACE: device + host timeline of run_pairs
  bool opt_rampdown = true;                  // PM_RAMPDOWN=0 (development): the last batches keep the full size
  bool staged_hint = false;                  // the last retrieved batch held pairs that ran past the filter's first rounds
  int opt_slots = 0;                         // PM_SLOTS: batches in flight (0 = default)
  bool opt_prefilter = false;                // PM_L2F_PREFILTER: s8-prefilter variant of the re-rank
  std::chrono::steady_clock::time_point trace_t0{};
  unsigned int* d_fstats = nullptr;   // pack_float statistics of the image being ingested
  unsigned int* h_fstats = nullptr;   // pinned
  unsigned long long* d_l2f = nullptr;   // l2f_fixup counters
  uint32_t* u8d = nullptr;     // [rows][32]     byte copy of integer-valued 128-d rows (fix-up kernel)
  uint8_t *iq = nullptr, *it = nullptr;  // [rows][160] byte operand forms of integer-valued 128-d rows (kind::i8)
  int32_t* qoff = nullptr;     // [rows] |a|^2 - 254 sum(a) (i8 form)
  TcMaps imaps{};
  bool tci_ready = false;
  uint32_t* bits = nullptr;    // [rows][words]  (U8_BITS)
  int32_t* xy = nullptr;       // [rows][2]
  TcMaps maps{};
  int* d_flag = nullptr;       // not-integral flag
  int* h_flag = nullptr;       // pinned
  static constexpr int kRecInts = 8, kRecCap = 1 << 16;
  int* h_recs = nullptr;       // pinned [kRecCap][kRecInts]: {flag, fstats[4]} of asynchronously ingested images
  int next_rec = 0, n_pending = 0;   // records are recycled whenever no image is pending
  void* stage = nullptr;       // device staging for u8 uploads
  size_t stage_bytes = 0;

  std::unordered_map<int, Image> images;
  // released row ranges (pm_remove_image, images re-set with more keypoints), kept sorted by row and coalesced; a new
  // image takes the smallest range that fits before the arena is bumped
  std::vector<std::pair<int32_t, int32_t>> free_rows;   // (first row, rows)
  void release_rows(int32_t row, int32_t cap) {
    if (cap <= 0) return;
    if (static_cast<int64_t>(row) + cap == used_rows) { used_rows = row; }      // the arena's tail: just step back
    else free_rows.emplace_back(row, cap);
    std::sort(free_rows.begin(), free_rows.end());
    for (size_t k = 0; k + 1 < free_rows.size();) {
      if (free_rows[k].first + free_rows[k].second == free_rows[k + 1].first) {
        free_rows[k].second += free_rows[k + 1].second;
        free_rows.erase(free_rows.begin() + k + 1);
      } else ++k;
    }
    while (!free_rows.empty() && static_cast<int64_t>(free_rows.back().first) + free_rows.back().second == used_rows) {
      used_rows = free_rows.back().first;
      free_rows.pop_back();
    }
  }
  bool take_rows(int32_t n, int32_t* row, int32_t* cap) {
    size_t best = free_rows.size();
    for (size_t k = 0; k < free_rows.size(); ++k)
      if (free_rows[k].second >= n && (best == free_rows.size() || free_rows[k].second < free_rows[best].second)) best = k;
    if (best == free_rows.size()) return false;
    *row = free_rows[best].first;
    // a much larger range is split, the remainder stays free
    if (free_rows[best].second >= 2 * n && free_rows[best].second - n >= 256) {
      *cap = n;
      free_rows[best].first += n;
      free_rows[best].second -= n;
    } else {
      *cap = free_rows[best].second;
      free_rows.erase(free_rows.begin() + best);
    }
    return true;
  }
  int remove_image(int id) {
    auto it = images.find(id);
    if (it == images.end()) return fail(PM_ERR_STATE, "image id %d not set", id);
    PM_CUDA(cudaSetDevice(dev));
    // queued work may still read the rows: drain before they can be handed to another image
    PM_CUDA(cudaDeviceSynchronize());
    if (it->second.pending) { const int rc = resolve(it->second); if (rc != PM_OK) return rc; }
    if (it->second.ready) cudaEventDestroy(it->second.ready);
    release_rows(it->second.row, it->second.cap);
    images.erase(it);
    stats.n_images = static_cast<int32_t>(images.size());
    return PM_OK;
  }
  std::shared_ptr<PinnedPool> pool = std::make_shared<PinnedPool>();
  std::vector<Slot> slots;
  Slot single;
  cudaStream_t ingest = nullptr;
  cudaStream_t knn_stream = nullptr;   // all kNN kernels run back to back on this stream
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  float* d_dump = nullptr;
  void* e_scratch = nullptr;           // pm_filter_pair_E: normalised points + the models of a round
  size_t e_scratch_bytes = 0;

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
  int fail_cuda(cudaError_t e, const char* what, int line) {
    (void)cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? PM_ERR_OOM : PM_ERR_CUDA, "CUDA error '%s' at api.cu:%d: %s",
                cudaGetErrorString(e), line, what);
  }

  int init(int device, const pm_params& p) {
    dev = device;
    prm = p;
    PM_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop{};
    PM_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
      return fail(PM_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev,
                  prop.major, prop.minor);
    num_sms = prop.multiProcessorCount;
    trace_on = std::getenv("PM_TRACE") != nullptr;
    if (const char* e = std::getenv("PM_SLOTS")) opt_slots = std::max(1, std::min(32, std::atoi(e)));
    { const char* e = std::getenv("PM_RESULT_COPY"); result_by_ce = !(e && std::strcmp(e, "kernel") == 0); }
    opt_prefilter = std::getenv("PM_L2F_PREFILTER") != nullptr;
    { const char* e = std::getenv("PM_RAMPDOWN"); opt_rampdown = !(e && e[0] == '0'); }
    {
      // the persistent kNN kernels go first whenever SMs free up; the small tail kernels of earlier batches
      // (default priority) fill whatever registers / shared memory the kNN kernel leaves
      int prio_lo = 0, prio_hi = 0;
      PM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      PM_CUDA(cudaStreamCreateWithPriority(&knn_stream, cudaStreamNonBlocking, prio_hi));
      // the ingest stream too: the (single) NCCL kernel of a collective ingest must not starve behind the kNN kernels
      PM_CUDA(cudaStreamCreateWithPriority(&ingest, cudaStreamNonBlocking, prio_hi));
    }
    PM_CUDA(cudaEventCreate(&ev_a));
    PM_CUDA(cudaEventCreate(&ev_b));
    PM_CUDA(cudaMalloc(&d_flag, sizeof(int)));
    PM_CUDA(cudaMallocHost(&h_flag, sizeof(int)));
    PM_CUDA(cudaMallocHost(&h_recs, sizeof(int) * kRecInts * kRecCap));
    PM_CUDA(cudaMalloc(&d_fstats, 4 * sizeof(unsigned int)));
    PM_CUDA(cudaMallocHost(&h_fstats, 4 * sizeof(unsigned int)));
    PM_CUDA(cudaMalloc(&d_l2f, 4 * sizeof(unsigned long long)));
    PM_CUDA(cudaMemset(d_l2f, 0, 4 * sizeof(unsigned long long)));
    PM_CUDA(tc_configure());
    PM_CUDA(tc2_configure());
    PM_CUDA(i8x2_configure());
    PM_CUDA(s8_configure());
    PM_CUDA(fixup_configure());
    PM_CUDA(l2f_configure());
    PM_CUDA(select_configure());
    PM_CUDA(ransac_configure());
    PM_CUDA(hamming_fixup_configure());
    stats.device_id = dev;
    return PM_OK;
  }

  void shutdown() {
    cudaSetDevice(dev);
    cudaDeviceSynchronize();
    for (auto& s : slots) s.destroy();
    single.destroy();
    auto fd = [](auto*& p) { if (p) { cudaFree(p); p = nullptr; } };
    fd(raw); fd(qf); fd(tf); fd(qnorm); fd(u8d); fd(bits); fd(xy); fd(d_flag); fd(stage); fd(d_dump); fd(e_scratch);
    fd(fq); fd(ft); fd(fnorm); fd(d_fstats); fd(d_l2f); fd(hq); fd(ht); fd(iq); fd(it); fd(qoff); fd(sq8); fd(st8); fd(hq8); fd(ht8); fd(hq4); fd(ht4); fd(ht4x);
    if (comm) { nccl_api().CommDestroy(comm); comm = nullptr; }
    if (ag_own) cudaFree(ag_own);
    if (ag_all) cudaFree(ag_all);
    if (ag_flag) cudaFree(ag_flag);
    if (h_agflag) cudaFreeHost(h_agflag);
    h_agflag = nullptr;
    ag_own = ag_all = nullptr; ag_flag = nullptr; ag_own_bytes = ag_all_bytes = 0;
    if (h_flag) cudaFreeHost(h_flag);
    if (h_recs) cudaFreeHost(h_recs);
    for (auto& kv : images) if (kv.second.ready) cudaEventDestroy(kv.second.ready);
    if (h_fstats) cudaFreeHost(h_fstats);
    if (ingest) cudaStreamDestroy(ingest);
    if (knn_stream) cudaStreamDestroy(knn_stream);
    if (ev_a) cudaEventDestroy(ev_a);
    if (ev_b) cudaEventDestroy(ev_b);
  }

  // ---- tensor maps -----------------------------------------------------------------------
  bool float_tc_shape() const { return dtype == PM_DESC_F32 && (dim == 128 || dim == 256); }
  bool bits_tc_shape() const { return dtype == PM_DESC_U8_BITS && (words == 8 || words == 16); }

  int build_maps() {
    tc_ready = false;
    tcf_ready = false;
    tch_ready = false;
    tci_ready = false;
    tcs_ready = false;
    tch8_ready = false;
    tch4_ready = false;
    const bool sift_shape = dim == TC_DIM && (dtype == PM_DESC_F32 || dtype == PM_DESC_U8);
    if (!sift_shape && !float_tc_shape() && !bits_tc_shape()) return PM_OK;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
      return fail(PM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
    int kpad = TC_KPAD;
    const cuuint32_t estr[2] = {1, 1};
    auto mk = [&](CUtensorMap* m, void* base, cuuint32_t box_k, CUtensorMapSwizzle sw, cuuint32_t rows = 128) -> CUresult {
      const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kpad), static_cast<cuuint64_t>(cap_rows)};
      const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(kpad) * sizeof(__half)};
      const cuuint32_t box[2] = {box_k, rows};
      return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUresult r;
    int kb = 32 * words + 32;
    auto mkb = [&](CUtensorMap* m, void* base, cuuint32_t box_k, CUtensorMapSwizzle sw, cuuint32_t rows = 128) -> CUresult {
        const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kb), static_cast<cuuint64_t>(cap_rows)};
        const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(kb)};
        const cuuint32_t box[2] = {box_k, rows};
        return encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    if (bits_tc_shape()) {
      if ((r = mkb(&hmaps.q_main, hq, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
          (r = mkb(&hmaps.q_ext, hq, 32, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS ||
          (r = mkb(&hmaps.t_main, ht, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
          (r = mkb(&hmaps.t_ext, ht, 32, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS)
        return fail(PM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
      tch_ready = true;
      if (words == 8) {
        if ((r = mkb(&h8maps.q_main, hq8, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
            (r = mkb(&h8maps.q_ext, hq8, 32, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS ||
            (r = mkb(&h8maps.t_main, ht8, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
            (r = mkb(&h8maps.t_ext, ht8, 32, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS)
          return fail(PM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
        tch8_ready = true;
      }
      if (words == 8 || words == 16) {      // E2M1 forms on kind::mxf4: one 128-byte K atom per 256 bits
        kb = tc_fp4_row(words);
        if ((r = mkb(&h4maps.q_main, hq4, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
            (r = mkb(&h4maps.q_ext, hq4, 32, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS ||
            (r = mkb(&h4maps.t_main96, ht4, 128, CU_TENSOR_MAP_SWIZZLE_128B, 96)) != CUDA_SUCCESS ||
            (r = mkb(&h4maps.t_ext96, ht4, 32, CU_TENSOR_MAP_SWIZZLE_32B, 96)) != CUDA_SUCCESS)
          return fail(PM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
        if (words == 8) {      // packed-pair kernel: one 32-byte norm row for both train rows of an accumulator column
          kb = 32;
          if ((r = mkb(&h4maps.t_ext2x96, ht4x, 32, CU_TENSOR_MAP_SWIZZLE_32B, 96)) != CUDA_SUCCESS)
            return fail(PM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
        }
        tch4_ready = true;
      }
      return PM_OK;
    }
    if (float_tc_shape()) {
      kpad = dim + 16;
      if ((r = mk(&fmaps.q_main, fq, 64, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
          (r = mk(&fmaps.q_ext, fq, 16, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS ||
          (r = mk(&fmaps.t_main, ft, 64, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
          (r = mk(&fmaps.t_ext, ft, 16, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS)
        return fail(PM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
      tcf_ready = true;
      kpad = TC_KPAD;
      kb = dim + 32;
      if ((r = mkb(&smaps.q_main, sq8, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
          (r = mkb(&smaps.q_ext, sq8, 32, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS ||
          (r = mkb(&smaps.t_main, st8, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
          (r = mkb(&smaps.t_ext, st8, 32, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS)
        return fail(PM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
      tcs_ready = true;
    }
    if (!sift_shape) return PM_OK;
    if ((r = mk(&maps.q_main, qf, 64, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
        (r = mk(&maps.q_ext, qf, 16, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS ||
        (r = mk(&maps.t_main, tf, 64, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
        (r = mk(&maps.t_ext, tf, 16, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS ||
        (r = mk(&maps.t_main96, tf, 64, CU_TENSOR_MAP_SWIZZLE_128B, 96)) != CUDA_SUCCESS ||
        (r = mk(&maps.t_ext96, tf, 16, CU_TENSOR_MAP_SWIZZLE_32B, 96)) != CUDA_SUCCESS)
      return fail(PM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    tc_ready = true;
    kb = TC_I8_ROW;
    if ((r = mkb(&imaps.q_main, iq, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
        (r = mkb(&imaps.q_ext, iq, 32, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS ||
        (r = mkb(&imaps.t_main, it, 128, CU_TENSOR_MAP_SWIZZLE_128B)) != CUDA_SUCCESS ||
        (r = mkb(&imaps.t_ext, it, 32, CU_TENSOR_MAP_SWIZZLE_32B)) != CUDA_SUCCESS)
      return fail(PM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    tci_ready = true;
    return PM_OK;
  }

  // ---- arenas ----------------------------------------------------------------------------
  template <class T>
  int grow(T*& p, size_t elems_per_row, int64_t new_cap) {
    T* np = nullptr;
    const size_t bytes = static_cast<size_t>(new_cap) * elems_per_row * sizeof(T);
    PM_CUDA(cudaMalloc(&np, bytes));
    PM_CUDA(cudaMemsetAsync(np, 0, bytes, ingest));
    if (p && used_rows > 0)
      PM_CUDA(cudaMemcpyAsync(np, p, static_cast<size_t>(used_rows) * elems_per_row * sizeof(T),
                              cudaMemcpyDeviceToDevice, ingest));
    PM_CUDA(cudaStreamSynchronize(ingest));
    if (p) PM_CUDA(cudaFree(p));
    p = np;
    return PM_OK;
  }

  int ensure_rows(int64_t need) {
    if (need <= cap_rows) return PM_OK;
    int64_t nc = std::max<int64_t>({need, cap_rows * 2, prm.reserve_keypoints, 4096});
    nc = (nc + 255) / 256 * 256;
    // all in-flight work reads the arenas: drain before moving them
    PM_CUDA(cudaDeviceSynchronize());
    int rc;
    if (dtype == PM_DESC_U8_BITS) {
      if ((rc = grow(bits, words, nc)) != PM_OK) return rc;
      if (bits_tc_shape()) {
        if ((rc = grow(hq, 32 * words + 32, nc)) != PM_OK) return rc;
        if ((rc = grow(ht, 32 * words + 32, nc)) != PM_OK) return rc;
        if (words == 8) {
          if ((rc = grow(hq8, 32 * words + 32, nc)) != PM_OK) return rc;
          if ((rc = grow(ht8, 32 * words + 32, nc)) != PM_OK) return rc;
        }
        if (words == 8 || words == 16) {
          if ((rc = grow(hq4, tc_fp4_row(words), nc)) != PM_OK) return rc;
          if ((rc = grow(ht4, tc_fp4_row(words), nc)) != PM_OK) return rc;
          if (words == 8 && (rc = grow(ht4x, 32, nc)) != PM_OK) return rc;
        }
        if ((rc = grow(qnorm, 1, nc)) != PM_OK) return rc;
      }
    } else {
      if ((rc = grow(raw, dim, nc)) != PM_OK) return rc;
      if (dim == TC_DIM) {
        if ((rc = grow(qf, TC_KPAD, nc)) != PM_OK) return rc;
        if ((rc = grow(tf, TC_KPAD, nc)) != PM_OK) return rc;
        if ((rc = grow(qnorm, 1, nc)) != PM_OK) return rc;
        if ((rc = grow(u8d, 32, nc)) != PM_OK) return rc;
        if ((rc = grow(iq, TC_I8_ROW, nc)) != PM_OK) return rc;
        if ((rc = grow(it, TC_I8_ROW, nc)) != PM_OK) return rc;
        if ((rc = grow(qoff, 1, nc)) != PM_OK) return rc;
      }
      if (float_tc_shape()) {
        if ((rc = grow(fq, dim + 16, nc)) != PM_OK) return rc;
        if ((rc = grow(ft, dim + 16, nc)) != PM_OK) return rc;
        if ((rc = grow(fnorm, 1, nc)) != PM_OK) return rc;
        if ((rc = grow(sq8, dim + 32, nc)) != PM_OK) return rc;
        if ((rc = grow(st8, dim + 32, nc)) != PM_OK) return rc;
      }
    }
    if ((rc = grow(xy, 2, nc)) != PM_OK) return rc;
    cap_rows = nc;
    return build_maps();
  }

  // Waits for an asynchronously ingested image and reads its facts from the pinned record.
  int resolve(Image& im) {
    if (!im.pending) return PM_OK;
    const auto t_w0 = std::chrono::steady_clock::now();
    PM_CUDA(cudaEventSynchronize(im.ready));
    if (trace_on) {                            // development (PM_TRACE): when did the host see this image?
      const auto t_w1 = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[pm trace] host: image at row %d ready at %.3f ms (waited %.3f ms)\n", im.row,
                   std::chrono::duration<double, std::milli>(t_w1 - trace_t0).count(),
                   std::chrono::duration<double, std::milli>(t_w1 - t_w0).count());
    }
    const int* rec = h_recs + static_cast<size_t>(im.rec) * kRecInts;
    if (dtype != PM_DESC_U8_BITS && dim == TC_DIM) {
      im.integral = (rec[0] & 1) == 0;
      im.i8_ok = rec[0] == 0;
    }
    im.pending = false;
    if (--n_pending <= 0) { n_pending = 0; next_rec = 0; }     // (rec stays readable below: nothing is queued in between)
    if (float_tc_shape() && !im.integral) {
      if (dim == TC_DIM) {
        // 128-d rows that turned out not to be integer-valued: their fp16 forms were not packed on speculation
        int* rec2 = h_recs + static_cast<size_t>(im.rec) * kRecInts;
        const int kp = dim + 16;
        PM_CUDA(cudaMemsetAsync(d_fstats, 0, 4 * sizeof(unsigned int), ingest));
        PM_CUDA(launch_pack_float(raw + static_cast<size_t>(im.row) * dim, im.n, dim, fq + static_cast<size_t>(im.row) * kp,
                                  ft + static_cast<size_t>(im.row) * kp, fnorm + im.row, d_fstats,
                                  sq8 + static_cast<size_t>(im.row) * (dim + 32), st8 + static_cast<size_t>(im.row) * (dim + 32),
                                  ingest));
        ++stats.kernel_launches;
        PM_CUDA(cudaMemcpyAsync(rec2 + 1, d_fstats, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, ingest));
        PM_CUDA(cudaStreamSynchronize(ingest));
      }
      std::memcpy(&im.maxn, &rec[2], sizeof(float));
      im.unit_ok = rec[3] == 0 && im.maxn <= L2F_MAX_NORM2;
      float amax;
      std::memcpy(&amax, &rec[1], sizeof(float));
      im.s8_ok = im.unit_ok && amax <= L2S8_MAX_ABS;
      float dev1;
      std::memcpy(&dev1, &rec[4], sizeof(float));
      im.unit1_ok = im.s8_ok && dev1 <= L2S8_UNIT_TOL;
    }
    return PM_OK;
  }
  int resolve_all() {
    for (auto& kv : images) {
      const int rc = resolve(kv.second);
      if (rc != PM_OK) return rc;
    }
    return PM_OK;
  }

  int set_image(int id, const void* desc, int n, int dim_, int dtype_, const int32_t* xy_, bool on_device,
                bool async = false) {
    PM_CUDA(cudaSetDevice(dev));
    if (n < 0 || dim_ <= 0 || (n > 0 && !desc)) return fail(PM_ERR_INVALID, "set_image: bad arguments");
    if (dtype_ != PM_DESC_F32 && dtype_ != PM_DESC_U8_BITS && dtype_ != PM_DESC_U8)
      return fail(PM_ERR_INVALID, "set_image: unknown dtype %d", dtype_);
    if (dtype < 0) {
      if (dtype_ == PM_DESC_U8_BITS) {
        if (dim_ != 128 && dim_ != 256 && dim_ != 512)
          return fail(PM_ERR_UNSUPPORTED, "binary descriptors of %d bits (supported: 128, 256, 512)", dim_);
        words = dim_ / 32;
      } else if (dim_ % 4 != 0) {
        return fail(PM_ERR_UNSUPPORTED, "float descriptors need dim %% 4 == 0 (got %d)", dim_);
      }
      dim = dim_;
      dtype = dtype_;
    } else if (dim_ != dim || dtype_ != dtype) {
      // the reference only asserts this (FeatureMatcher.cpp:41-42, and compares a set with itself)
      return fail(PM_ERR_INVALID, "set_image: descriptor shape (dim %d, dtype %d) differs from the handle's (dim %d, dtype %d)",
                  dim_, dtype_, dim, dtype);
    }
    if (dtype == PM_DESC_U8_BITS && n > 65535)
      return fail(PM_ERR_UNSUPPORTED, "binary path packs the train index in 16 bits: n must be <= 65535");

    Image& im = images[id];
    if (im.pending) { const int rc = resolve(im); if (rc != PM_OK) return rc; }
    if (async && (n == 0 || next_rec >= kRecCap || (dtype == PM_DESC_U8 && dim != TC_DIM))) async = false;
    if (n > im.cap) {
      if (im.cap > 0) {
        // the old rows may still be read by queued work (per-pair calls are synchronous, so only after async ingest)
        PM_CUDA(cudaStreamSynchronize(ingest));
        release_rows(im.row, im.cap);
        im.cap = 0;
      }
      int32_t row = 0, cap = 0;
      if (take_rows(n, &row, &cap)) {
        im.row = row; im.cap = cap;
      } else {
        const int rc = ensure_rows(used_rows + n);
        if (rc != PM_OK) { images.erase(id); return rc; }
        im.row = static_cast<int32_t>(used_rows);
        im.cap = n;
        used_rows += n;
      }
    }
    im.n = n;
    im.has_xy = xy_ != nullptr;
    im.integral = false;
    im.i8_ok = false;
    im.unit_ok = false;
    im.s8_ok = false;
    im.unit1_ok = false;
    im.maxn = 0.f;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (n > 0) {
      if (dtype == PM_DESC_U8_BITS) {
        const size_t bytes = static_cast<size_t>(n) * words * 4;
        PM_CUDA(cudaMemcpyAsync(bits + static_cast<size_t>(im.row) * words, desc, bytes, kind, ingest));
        if (!on_device) stats.h2d_bytes += bytes;
        if (bits_tc_shape()) {
          const size_t kb = 32 * static_cast<size_t>(words) + 32;
          PM_CUDA(launch_pack_bits(bits + static_cast<size_t>(im.row) * words, n, words, hq + im.row * kb,
                                   ht + im.row * kb, qnorm + im.row, words == 8 ? hq8 + im.row * kb : nullptr,
                                   words == 8 ? ht8 + im.row * kb : nullptr,
                                   (words == 8 || words == 16) ? hq4 + static_cast<size_t>(im.row) * tc_fp4_row(words) : nullptr,
                                   (words == 8 || words == 16) ? ht4 + static_cast<size_t>(im.row) * tc_fp4_row(words) : nullptr, ingest,
                                   words == 8 ? ht4x + static_cast<size_t>(im.row) * 32 : nullptr));
          ++stats.kernel_launches;
        }
      } else {
        float* rdst = raw + static_cast<size_t>(im.row) * dim;
        const uint8_t* u8src = nullptr;
        if (dtype == PM_DESC_U8) {
          const size_t bytes = static_cast<size_t>(n) * dim;
          if (on_device) {
            u8src = static_cast<const uint8_t*>(desc);
          } else {
            if (bytes > stage_bytes) {
              if (stage) PM_CUDA(cudaFree(stage));
              stage = nullptr; stage_bytes = 0;
              PM_CUDA(cudaMalloc(&stage, bytes));
              stage_bytes = bytes;
            }
            PM_CUDA(cudaMemcpyAsync(stage, desc, bytes, cudaMemcpyHostToDevice, ingest));
            stats.h2d_bytes += bytes;
            u8src = static_cast<const uint8_t*>(stage);
          }
          if (dim != TC_DIM) {
            PM_CUDA(launch_u8_to_f32(u8src, rdst, bytes, ingest));
            ++stats.kernel_launches;
            im.integral = true;
          }
        } else {
          const size_t bytes = static_cast<size_t>(n) * dim * sizeof(float);
          PM_CUDA(cudaMemcpyAsync(rdst, desc, bytes, kind, ingest));
          if (!on_device) stats.h2d_bytes += bytes;
        }
        if (dim == TC_DIM) {
          // asynchronous ingest: the kernel stores the image's facts straight into its pinned record (zero-copy), so
          // there is no flag memset and no device-to-host copy per image (host-side enqueue cost of a 282-image
          // collective ingest: 5.4 ms before, the device idle behind it)
          uint8_t* host_flags = nullptr;
          if (async) {
            int* rec = h_recs + static_cast<size_t>(next_rec) * kRecInts;
            rec[0] = 0; rec[1] = rec[2] = rec[3] = rec[4] = 0;
            host_flags = reinterpret_cast<uint8_t*>(rec);
          } else {
            PM_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), ingest));
          }
          PM_CUDA(launch_pack_sift(u8src ? nullptr : rdst, u8src, n, qf + static_cast<size_t>(im.row) * TC_KPAD,
                                   tf + static_cast<size_t>(im.row) * TC_KPAD, qnorm + im.row,
                                   u8src ? rdst : nullptr, u8d + static_cast<size_t>(im.row) * 32, d_flag,
                                   iq + static_cast<size_t>(im.row) * TC_I8_ROW,
                                   it + static_cast<size_t>(im.row) * TC_I8_ROW, qoff + im.row, ingest, host_flags));
          ++stats.kernel_launches;
          if (!async) PM_CUDA(cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ingest));
        }
      }
      if (xy_) {
        PM_CUDA(cudaMemcpyAsync(xy + 2 * static_cast<size_t>(im.row), xy_, static_cast<size_t>(n) * 8, kind, ingest));
        if (!on_device) stats.h2d_bytes += static_cast<size_t>(n) * 8;
      }
    }
    if (async) {
      // no host synchronisation: facts go to this image's pinned record, an event marks completion
      int* rec = h_recs + static_cast<size_t>(next_rec) * kRecInts;
      if (!(dtype != PM_DESC_U8_BITS && dim == TC_DIM && n > 0)) { rec[0] = 0; rec[1] = rec[2] = rec[3] = rec[4] = 0; }
      // (128-d rows: the record was cleared before pack_sift_kernel, which writes its flags into it)
      if (float_tc_shape() && dtype == PM_DESC_F32 && dim != TC_DIM) {     // 128-d: packed in resolve() if needed
        const int kp = dim + 16;
        PM_CUDA(cudaMemsetAsync(d_fstats, 0, 4 * sizeof(unsigned int), ingest));
        PM_CUDA(launch_pack_float(raw + static_cast<size_t>(im.row) * dim, n, dim, fq + static_cast<size_t>(im.row) * kp,
                                  ft + static_cast<size_t>(im.row) * kp, fnorm + im.row, d_fstats,
                                  sq8 + static_cast<size_t>(im.row) * (dim + 32), st8 + static_cast<size_t>(im.row) * (dim + 32),
                                  ingest));
        ++stats.kernel_launches;
        PM_CUDA(cudaMemcpyAsync(rec + 1, d_fstats, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, ingest));
      }
      if (!im.ready) PM_CUDA(cudaEventCreateWithFlags(&im.ready, cudaEventDisableTiming));
      PM_CUDA(cudaEventRecord(im.ready, ingest));
      im.rec = next_rec++;
      im.pending = true;
      ++n_pending;
      stats.n_images = static_cast<int32_t>(images.size());
      return PM_OK;
    }
    PM_CUDA(cudaStreamSynchronize(ingest));   // caller's buffers are free to change on return
    if (n > 0 && dtype != PM_DESC_U8_BITS && dim == TC_DIM) {
      im.integral = (*h_flag & 1) == 0;
      im.i8_ok = *h_flag == 0;
    }
    if (n > 0 && float_tc_shape() && !im.integral) {
      // real-valued rows: fp16 operand forms + norms for the tensor-core search (l2_tc2.cu MODE 3)
      const int kp = dim + 16;
      PM_CUDA(cudaMemsetAsync(d_fstats, 0, 4 * sizeof(unsigned int), ingest));
      PM_CUDA(launch_pack_float(raw + static_cast<size_t>(im.row) * dim, n, dim, fq + static_cast<size_t>(im.row) * kp,
                                ft + static_cast<size_t>(im.row) * kp, fnorm + im.row, d_fstats,
                                  sq8 + static_cast<size_t>(im.row) * (dim + 32), st8 + static_cast<size_t>(im.row) * (dim + 32),
                                  ingest));
      ++stats.kernel_launches;
      PM_CUDA(cudaMemcpyAsync(h_fstats, d_fstats, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, ingest));
      PM_CUDA(cudaStreamSynchronize(ingest));
      std::memcpy(&im.maxn, &h_fstats[1], sizeof(float));
      im.unit_ok = h_fstats[2] == 0 && im.maxn <= L2F_MAX_NORM2;
      float amax;
      std::memcpy(&amax, &h_fstats[0], sizeof(float));
      im.s8_ok = im.unit_ok && amax <= L2S8_MAX_ABS;
      float dev1;
      std::memcpy(&dev1, &h_fstats[3], sizeof(float));
      im.unit1_ok = im.s8_ok && dev1 <= L2S8_UNIT_TOL;
    }
    stats.n_images = static_cast<int32_t>(images.size());
    return PM_OK;
  }

  // ---- scratch ---------------------------------------------------------------------------
  int ensure_slot(Slot& s, int pairs, int stride, bool mutual) {
    if (!s.stream) {
      PM_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
      PM_CUDA(cudaEventCreate(&s.ev_done));
      PM_CUDA(cudaEventCreate(&s.ev_k0));
      PM_CUDA(cudaEventCreate(&s.ev_k1));
      PM_CUDA(cudaEventCreateWithFlags(&s.ev_jobs, cudaEventDisableTiming));
      PM_CUDA(cudaEventCreateWithFlags(&s.ev_knn, cudaEventDisableTiming));
    }
    const bool extra_ok = !float_tc_shape() || (s.knn_extra && (!mutual || s.rev_extra));
    if (pairs <= s.cap_pairs && stride <= s.stride && (!mutual || s.mutual) && extra_ok) return PM_OK;
    PM_CUDA(cudaStreamSynchronize(s.stream));
    pairs = std::max(pairs, s.cap_pairs);
    stride = std::max(stride, s.stride);
    mutual = mutual || s.mutual;
    s.release();
    const size_t ps = static_cast<size_t>(pairs) * stride;
    PM_CUDA(cudaMalloc(&s.d_jobs, sizeof(PairJob) * pairs));
    PM_CUDA(cudaMalloc(&s.d_rjobs, sizeof(PairJob) * pairs));
    PM_CUDA(cudaMalloc(&s.knn_idx, sizeof(int2) * ps));
    PM_CUDA(cudaMalloc(&s.knn_dist, sizeof(float2) * ps));
    if (float_tc_shape()) PM_CUDA(cudaMalloc(&s.knn_extra, sizeof(float2) * ps));
    if (float_tc_shape()) PM_CUDA(cudaMalloc(&s.rr_flag, ps));
    if (mutual) {
      PM_CUDA(cudaMalloc(&s.rev_idx, sizeof(int2) * ps));
      PM_CUDA(cudaMalloc(&s.rev_dist, sizeof(float2) * ps));
      if (float_tc_shape()) PM_CUDA(cudaMalloc(&s.rev_extra, sizeof(float2) * ps));
    }
    PM_CUDA(cudaMalloc(&s.owner, 4 * ps));
    PM_CUDA(cudaMalloc(&s.match_q, 4 * ps));
    PM_CUDA(cudaMalloc(&s.match_t, 4 * ps));
    PM_CUDA(cudaMalloc(&s.pts1, 8 * ps));
    PM_CUDA(cudaMalloc(&s.pts2, 8 * ps));
    PM_CUDA(cudaMalloc(&s.mask, ps));
    PM_CUDA(cudaMalloc(&s.out_q, 4 * ps));
    PM_CUDA(cudaMalloc(&s.out_t, 4 * ps));
    PM_CUDA(cudaMalloc(&s.out_mask, ps));
    PM_CUDA(cudaMalloc(&s.count, 4 * pairs));
    PM_CUDA(cudaMalloc(&s.F, 72 * pairs));
    PM_CUDA(cudaMalloc(&s.status, 4 * pairs));
    PM_CUDA(cudaMalloc(&s.n_inl, 4 * pairs));
    PM_CUDA(cudaMalloc(&s.iters, 4 * pairs));
    PM_CUDA(cudaMalloc(&s.offsets, 8 * (pairs + 1)));
    PM_CUDA(cudaMalloc(&s.rs_ws, ransac_workspace_bytes(pairs)));
    PM_CUDA(cudaMallocHost(&s.h_jobs, sizeof(PairJob) * pairs));
    PM_CUDA(cudaMallocHost(&s.h_rjobs, sizeof(PairJob) * pairs));
    PM_CUDA(cudaMallocHost(&s.h_offsets, 8 * (pairs + 1)));
    PM_CUDA(cudaMallocHost(&s.h_F, 72 * pairs));
    PM_CUDA(cudaMallocHost(&s.h_status, 4 * pairs));
    PM_CUDA(cudaMallocHost(&s.h_ninl, 4 * pairs));
    PM_CUDA(cudaMallocHost(&s.h_iters, 4 * pairs));
    PM_CUDA(cudaMallocHost(&s.h_count, 4 * pairs));
    s.cap_pairs = pairs; s.stride = stride; s.mutual = mutual;
    return PM_OK;
  }

  RansacDev ransac_dev(bool do_filter) const {
    RansacDev r{};
    r.sampler = prm.sampler;
    r.refit_8point = prm.refit_8point;
    r.seed = prm.seed;
    r.confidence = prm.ransac_confidence;
    r.thr = static_cast<float>(prm.ransac_threshold * prm.ransac_threshold);
    r.max_iters = prm.ransac_max_iters;
    r.residual_mode = prm.residual_mode;
    r.min_matches = prm.min_matches;
    r.do_filter = do_filter ? 1 : 0;
    return r;
  }

  // Queues the kNN kernel(s) of a batch whose jobs are already in s.h_jobs[0..n).
  int enqueue_knn(Slot& s, int n, bool want_rev, bool timed, float* dump = nullptr, bool fast = false) {
    int max_nq = 0, max_nt = 0;
    bool all_integral = true, all_unit = true, all_i8 = true, all_s8 = true, all_u1 = true;
    double work = 0;
    for (int i = 0; i < n; ++i) {
      max_nq = std::max(max_nq, s.h_jobs[i].nq);
      max_nt = std::max(max_nt, s.h_jobs[i].nt);
      work += static_cast<double>(s.h_jobs[i].nq) * s.h_jobs[i].nt;
      s.h_rjobs[i] = PairJob{s.h_jobs[i].t_row, s.h_jobs[i].q_row, s.h_jobs[i].nt, s.h_jobs[i].nq,
                             s.h_jobs[i].t_maxn, s.h_jobs[i].q_maxn, s.h_jobs[i].seed_lo, s.h_jobs[i].seed_hi};
    }
    // by a kernel, not the copy engine: an H2D copy would wait behind every image upload still queued there
    PM_CUDA(launch_copy_pinned(s.h_jobs, s.d_jobs, sizeof(PairJob) * n, s.stream));
    ++stats.kernel_launches;
    if (want_rev) {
      PM_CUDA(launch_copy_pinned(s.h_rjobs, s.d_rjobs, sizeof(PairJob) * n, s.stream));
      ++stats.kernel_launches;
    }
    if (dtype != PM_DESC_U8_BITS) {
      for (int i = 0; i < n && all_integral; ++i) all_integral = job_integral[i];
      for (int i = 0; i < n && all_i8; ++i) all_i8 = job_i8[i];
      for (int i = 0; i < n && all_s8; ++i) all_s8 = job_s8[i];
      for (int i = 0; i < n && all_u1; ++i) all_u1 = job_u1[i];
      for (int i = 0; i < n && all_unit; ++i) all_unit = job_unit[i];
    }
    // The kNN kernels of all batches are serialised on one stream (they fill the machine anyway);
    // the slot's own stream carries the small tail kernels and the D2H, overlapping the next kNN.
    PM_CUDA(cudaEventRecord(s.ev_jobs, s.stream));
    PM_CUDA(cudaStreamWaitEvent(knn_stream, s.ev_jobs, 0));
    // debug_flags (kernel variants; 0 = product defaults):
    //   bit0      force the fp32 SIMT L2 kernel
    //   bit1      Hamming: plain 8-popc variant instead of the carry-save one (default)
    //   bits2-4   GENERAL tensor kernel (exact top-2 with indices; raw kNN rows / single-pair calls):
    //             0 single-CTA 16 epilogue warps (default), 1 x32 / 8 warps, 2 x64 overlapped, 3 x128,
    //             4 timing probe, 5 CTA pair 256-col tiles, 6 pair probe, 7 CTA pair 192-col tiles
    //   bit5      pair-192 timing probe
    //   bit6      batched loop uses the general kernel instead of values-only kernel + fix-up
    //   bits7-8   VALUES-ONLY kernel of the batched loop: 0 CTA pair 256-col (default), 1 single-CTA,
    //             2 CTA pair 192-col
    //   bit9      values-only CTA-pair kernel in its 64-register build (tail kernels co-resident)
    //   bit11     batched loop keeps the fp16 form (kind::f16) instead of the byte form (kind::i8, default)
    //   bits12-13 kind::i8 kernel: 1 / 2 = timing probes (no matches): TMA + MMA only / + accumulator loads without
    //             the reduction; 3 = the 64-register build
    //   bit16     256-bit rows: kind::f8f6f4 kernel (E4M3 {0,1} forms) instead of the two-set kernels; bit17 = TMA + MMA probe
    //   bit18     256-bit rows: kind::i8 two-set kernel (byte forms) instead of the kind::mxf4 kernels (E2M1 forms)
    //   bit15     real-valued rows: fp16 forms (kind::f16) in the batched loop instead of the s8 forms (kind::i8)
    //   bit14     kind::i8: one query row set per cluster (l2_top2_tc2_kernel) instead of two (l2_i8x2_kernel)
    //   bit19     real-valued rows, s8 forms: one query row set per cluster instead of two (six chunk keys + chunk re-rank)
    //   bit20     real-valued rows, s8 forms: two row sets, six chunk keys + chunk re-rank (instead of the argmin epilogue
    //             + one-column re-rank, the default)
    //   bit21     epipolar filter: every iteration in the one-block-per-pair kernel (no staged continuation, ransac.cu)
    //   bit22     epipolar filter: always queue the staged continuation (default: while recent batches needed it)
    //   bit23     real-valued rows of unit norm: keep the norm K-step of the s8 search (default: dropped, l2_i8x2_kernel NX)
    //   bit24     pm_ingest_allgather: all gathers first, then all ingests; bit25: head first, then ONE gather for the rest
    //             (default: chunks of doubling size, each gathered and ingested in turn)
    //   bit26     256-bit rows on kind::mxf4: one train row per accumulator column (the round-1 form) instead of two
    //             (l2_i8x2_kernel PK, the default)
    const int code = (prm.debug_flags >> 2) & 7;
    const int fcode = (prm.debug_flags >> 7) & 3;
    const int variant = ((prm.debug_flags >> 1) & 1) ^ 1;
    const bool use_tc = tc_ready && all_integral && !(prm.debug_flags & 1);
    const bool use_fast = use_tc && fast && !dump && !((prm.debug_flags >> 6) & 1);
    // integer-valued 128-d rows as bytes on kind::i8: 5 K-steps per tile instead of 9 (l2_tc2.cu KIND 2)
    const bool use_i8 = use_fast && tci_ready && all_i8 && fcode == 0 && !((prm.debug_flags >> 11) & 1);
    const int i8_probe = (prm.debug_flags >> 12) & 3;
    // real-valued rows (SuperPoint): approximate tensor-core scores + exact fp32 re-rank (l2f_fixup.cu)
    const bool use_tcf = !use_tc && tcf_ready && dtype == PM_DESC_F32 && all_unit && !dump &&
                         max_nt <= L2F_MAX_NT &&
                         (!want_rev || max_nq <= L2F_MAX_NT) && !(prm.debug_flags & 1);
    // ... with the rows quantised to s8 on kind::i8 (half the K-steps) when only the outcome of the ratio test and
    // the nearest index of passing rows are needed (batched loop without the cross-check); debug bit15 keeps fp16
    const bool use_tcs8 = use_tcf && tcs_ready && all_s8 && fast && !want_rev && !((prm.debug_flags >> 15) & 1);
    // ... and the epilogue reporting the exact argmin column + the second smallest score (one exact column per
    // candidate row in l2f_rerank1 instead of a 16-column chunk); 13 column bits: train images of <= 8192 rows
    const bool use_keys3 = use_tcs8 && max_nt <= L2S8_KEYS3_MAX_NT && !((prm.debug_flags >> 19) & 1) && !((prm.debug_flags >> 20) & 1);
    // ... without the norm K-step when every train row of the batch has unit norm (SuperPoint); debug bit23 keeps it
    const bool use_unit1 = use_keys3 && all_u1 && !((prm.debug_flags >> 23) & 1);
    s.unit1 = use_unit1;
    const int epi_of_code[5] = {3, 0, 1, 2, 4};
    // binary rows: Hamming = |a| + |b| - 2 a.b on the tensor cores (E4M3 {0,1} operands) in the batched loop;
    // debug_flags bit10 keeps the XOR/popc kernel there too (it always serves raw kNN rows / single pairs)
    const bool use_tch = dtype == PM_DESC_U8_BITS && tch_ready && fast && !dump && !((prm.debug_flags >> 10) & 1);
    // 256-bit rows as bytes on kind::i8 with two query row sets per cluster (half the L2 traffic of the E4M3 kernel);
    // debug bit16 keeps the kind::f8f6f4 kernel, bit17 = TMA + MMA timing probe of the i8 kernel
    const bool use_tch4 = use_tch && tch4_ready && !((prm.debug_flags >> 16) & 1) && !((prm.debug_flags >> 18) & 1);
    const bool use_tch8 = use_tch && tch8_ready && !((prm.debug_flags >> 16) & 1) && !use_tch4;
    auto knn_main = [&](const PairJob* jobs_d, int mq, int2* oi, float2* od, float2* ox) -> cudaError_t {
      if (use_tch8 || use_tch4) {
        const int probe = (prm.debug_flags >> 17) & 1;
        if (probe) {
          cudaError_t e = cudaMemsetAsync(oi, 0xFF, sizeof(int2) * static_cast<size_t>(n) * s.stride, knn_stream);
          if (e != cudaSuccess) return e;
        }
        if (use_tch4)
          return launch_ham_fp4x2(h4maps, jobs_d, n, mq, oi, od, s.stride, num_sms, probe, knn_stream, words,
                                  words == 8 && !((prm.debug_flags >> 26) & 1));
        return launch_ham_i8x2(h8maps, jobs_d, n, mq, oi, od, s.stride, num_sms, probe, knn_stream);
      }
      if (use_tch) return launch_ham_tc2(hmaps, words, qnorm, jobs_d, n, mq, oi, od, s.stride, num_sms, knn_stream);
      if (dtype == PM_DESC_U8_BITS)
        return launch_hamming_top2(bits, words, jobs_d, n, mq, oi, od, s.stride, variant, knn_stream);
      if (use_tcs8) {
        if ((prm.debug_flags >> 19) & 1)
          return launch_l2s8_tc2(smaps, dim, jobs_d, n, mq, oi, od, ox, s.stride, num_sms, knn_stream);
        return launch_l2s8x2(smaps, dim, jobs_d, n, mq, oi, od, ox, s.stride, num_sms, knn_stream, use_unit1 ? 2 : (use_keys3 ? 1 : 0));
      }
      if (use_tcf) return launch_l2f_tc2(fmaps, dim, jobs_d, n, mq, oi, od, ox, s.stride, num_sms, knn_stream);
      if (!use_tc) return launch_l2_simt(raw, dim, jobs_d, n, mq, oi, od, s.stride, knn_stream);
      if (use_i8) {
        if (i8_probe == 1 || i8_probe == 2) {     // probes write nothing: every row reads "no neighbour"
          cudaError_t e = cudaMemsetAsync(oi, 0xFF, sizeof(int2) * static_cast<size_t>(n) * s.stride, knn_stream);
          if (e != cudaSuccess) return e;
        }
        if ((prm.debug_flags >> 14) & 1)
          return launch_l2i8_tc2(imaps, jobs_d, n, mq, oi, od, s.stride, num_sms, i8_probe, knn_stream);
        return launch_l2i8x2(imaps, jobs_d, n, mq, oi, od, s.stride, num_sms, i8_probe, knn_stream);
      }
      if (use_fast) {
        if (fcode == 1) return launch_l2_tc(maps, qnorm, jobs_d, n, mq, oi, od, s.stride, num_sms, nullptr, 5, knn_stream);
        return launch_l2_tc2(maps, qnorm, jobs_d, n, mq, oi, od, s.stride, num_sms, fcode == 2,
                             (fcode == 0 && ((prm.debug_flags >> 9) & 1)) ? 4 : 2, knn_stream);
      }
      if (code >= 5 && !dump)
        return launch_l2_tc2(maps, qnorm, jobs_d, n, mq, oi, od, s.stride, num_sms, code == 7,
                             (code == 6 || ((prm.debug_flags >> 5) & 1)) ? 1 : 0, knn_stream);
      return launch_l2_tc(maps, qnorm, jobs_d, n, mq, oi, od, s.stride, num_sms, dump,
                          dump ? 0 : epi_of_code[code < 5 ? code : 0], knn_stream);
    };
    if (timed) PM_CUDA(cudaEventRecord(s.ev_k0, knn_stream));
    PM_CUDA(knn_main(s.d_jobs, max_nq, s.knn_idx, s.knn_dist, s.knn_extra));
    if (timed) PM_CUDA(cudaEventRecord(s.ev_k1, knn_stream));
    ++stats.kernel_launches;
    s.knn_work = dtype == PM_DESC_U8_BITS ? work * words : work * 2.0 * dim;
    s.timed = timed;
    if (want_rev) {
      PM_CUDA(knn_main(s.d_rjobs, max_nt, s.rev_idx, s.rev_dist, s.rev_extra));
      ++stats.kernel_launches;
    }
    PM_CUDA(cudaEventRecord(s.ev_knn, knn_stream));
    PM_CUDA(cudaStreamWaitEvent(s.stream, s.ev_knn, 0));
    if (use_i8) {
      PM_CUDA(launch_l2_fixup_i8(u8d, qnorm, qoff, s.d_jobs, n, max_nq, s.knn_idx, s.knn_dist, s.stride, prm.ratio, 0,
                                 d_l2f, s.stream));
      ++stats.kernel_launches;
      if (want_rev) {
        PM_CUDA(launch_l2_fixup_i8(u8d, qnorm, qoff, s.d_rjobs, n, max_nt, s.rev_idx, s.rev_dist, s.stride, prm.ratio, 1,
                                   d_l2f, s.stream));
        ++stats.kernel_launches;
      }
    } else if (use_fast) {
      // exact indices / second neighbour, on the slot's stream so that it overlaps the next batch's
      // tensor kernel: rows that can still pass the ratio test, and -- for the cross-check, which needs
      // the nearest query of EVERY train row -- all rows of the reversed search
      PM_CUDA(launch_l2_fixup(u8d, qnorm, s.d_jobs, n, max_nq, 0, s.knn_idx, s.knn_dist, s.stride, prm.ratio, 0, s.stream));
      ++stats.kernel_launches;
      if (want_rev) {
        PM_CUDA(launch_l2_fixup(u8d, qnorm, s.d_rjobs, n, max_nt, 0, s.rev_idx, s.rev_dist, s.stride, prm.ratio, 1, s.stream));
        ++stats.kernel_launches;
      }
    }
    if (use_tch) {
      PM_CUDA(launch_hamming_fixup(bits, words, s.d_jobs, n, max_nq, s.knn_idx, s.knn_dist, s.stride, prm.ratio, 0, s.stream,
                                   use_tch4 ? 2 : (use_tch8 ? 1 : 0)));
      ++stats.kernel_launches;
      if (want_rev) {
        PM_CUDA(launch_hamming_fixup(bits, words, s.d_rjobs, n, max_nt, s.rev_idx, s.rev_dist, s.stride, prm.ratio, 1, s.stream,
                                     use_tch4 ? 2 : (use_tch8 ? 1 : 0)));
        ++stats.kernel_launches;
      }
    }
    if (use_keys3) {
      PM_CUDA(launch_l2f_rerank1(raw, fnorm, dim, s.d_jobs, n, max_nq, s.knn_idx, s.knn_dist, s.knn_extra, s.rr_flag, s.stride,
                                 prm.ratio, d_l2f, sq8, st8, s.stream, s.unit1 ? 2 : 1));
      stats.kernel_launches += 2;
    } else if (use_tcf) {
      PM_CUDA(launch_l2f_fixup(raw, fnorm, dim, s.d_jobs, n, max_nq, s.knn_idx, s.knn_dist, s.knn_extra, s.stride,
                               prm.ratio, fast ? L2F_NEED_RATIO : L2F_NEED_FULL, d_l2f, s.stream, use_tcs8 ? 1 : 0,
                               use_tcs8 && opt_prefilter ? sq8 : nullptr, use_tcs8 && opt_prefilter ? st8 : nullptr));
      ++stats.kernel_launches;
      if (want_rev) {
        PM_CUDA(launch_l2f_fixup(raw, fnorm, dim, s.d_rjobs, n, max_nt, s.rev_idx, s.rev_dist, s.rev_extra, s.stride,
                                 prm.ratio, fast ? L2F_NEED_NEAREST : L2F_NEED_FULL, d_l2f, s.stream));
        ++stats.kernel_launches;
      }
    }
    if (trace_on && timed) {
      for (auto& e : s.ev_tr) if (!e) PM_CUDA(cudaEventCreate(&e));
      PM_CUDA(cudaEventRecord(s.ev_tr[0], s.stream));
    }
    return PM_OK;
  }
  std::vector<char> job_integral, job_unit, job_i8, job_s8, job_u1;   // per job of the batch being built

  // ---- collective ingest (SURVEY 8e "Collective", 2.2 C1): extraction sharded over ranks, one NCCL all-gather ----
  NcclComm comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  uint8_t *ag_own = nullptr, *ag_all = nullptr;       // device: this rank's images in wire form / the gathered set
  size_t ag_own_bytes = 0, ag_all_bytes = 0;
  int* ag_flag = nullptr;                             // not-integral flag of the f32 -> u8 wire conversion
  int* h_agflag = nullptr;                            // pinned copy

  int comm_init(const uint8_t id[128], int rank, int n_ranks) {
    PM_CUDA(cudaSetDevice(dev));
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(PM_ERR_INVALID, "pm_comm_init: rank %d of %d", rank, n_ranks);
    if (const char* why = nccl_api().load()) return fail(PM_ERR_UNSUPPORTED, "NCCL unavailable: %s", why);
    if (comm) { nccl_api().CommDestroy(comm); comm = nullptr; }
    NcclUniqueId uid;
    std::memcpy(uid.internal, id, sizeof uid.internal);
    const int rc = nccl_api().CommInitRank(&comm, n_ranks, uid, rank);
    if (rc != kNcclSuccess) { comm = nullptr; return fail(PM_ERR_CUDA, "ncclCommInitRank: %s", nccl_api().GetErrorString(rc)); }
    comm_rank = rank; comm_size = n_ranks;
    return PM_OK;
  }

  // Image k (0 <= k < n_total) was extracted on rank k % R; own_desc / own_xy hold this rank's images in ascending
  // id order (n_kp rows each).  Own images -> device (wire form) -> all-gather slot by slot (slot s = images
  // s*R .. s*R+R-1, which land contiguously in id order) -> asynchronous ingest from the gathered buffer.  Everything is
  // queued on the ingest stream; nothing here waits for the device.
  int ingest_allgather(int n_total, int n_kp, int dim_, int dtype_, int wire, const void* own_desc, const int32_t* own_xy,
                       bool own_on_device) {
    PM_CUDA(cudaSetDevice(dev));
    if (!comm) return fail(PM_ERR_STATE, "pm_ingest_allgather: no communicator (call pm_comm_init first)");
    if (n_total < 0 || n_kp <= 0 || dim_ <= 0) return fail(PM_ERR_INVALID, "pm_ingest_allgather: bad arguments");
    if (dtype_ != PM_DESC_F32 && dtype_ != PM_DESC_U8_BITS && dtype_ != PM_DESC_U8) return fail(PM_ERR_INVALID, "unknown dtype %d", dtype_);
    if (wire != dtype_ && !(wire == PM_DESC_U8 && dtype_ == PM_DESC_F32))
      return fail(PM_ERR_INVALID, "wire dtype %d not available for descriptors of dtype %d", wire, dtype_);
    const int R = comm_size, n_slots = (n_total + R - 1) / R;
    const int n_own = n_total > comm_rank ? (n_total - comm_rank + R - 1) / R : 0;
    if (n_own > 0 && !own_desc) return fail(PM_ERR_INVALID, "pm_ingest_allgather: own_desc is NULL");
    const size_t in_row = dtype_ == PM_DESC_F32 ? 4u * dim_ : (dtype_ == PM_DESC_U8 ? static_cast<size_t>(dim_) : static_cast<size_t>(dim_) / 8);
    const size_t w_row = wire == PM_DESC_F32 ? 4u * dim_ : (wire == PM_DESC_U8 ? static_cast<size_t>(dim_) : static_cast<size_t>(dim_) / 8);
    const bool has_xy = own_xy != nullptr || n_own == 0;
    const size_t img_in = in_row * n_kp, img_w = w_row * n_kp, img_xy = 8u * n_kp;
    const size_t slot_w = img_w + (has_xy ? img_xy : 0);          // wire bytes of one image: descriptors, then xy
    // staging: [n_slots] own images in wire form (+ room for the raw fp32 rows when they are converted on the device)
    const size_t own_need = static_cast<size_t>(n_slots) * slot_w + (wire != dtype_ ? static_cast<size_t>(n_slots) * img_in : 0);
    const size_t all_need = static_cast<size_t>(n_slots) * R * slot_w;
    if (own_need > ag_own_bytes || all_need > ag_all_bytes) {
      PM_CUDA(cudaStreamSynchronize(ingest));                      // an earlier ingest may still read the buffers
      if (own_need > ag_own_bytes) { if (ag_own) cudaFree(ag_own); ag_own = nullptr; ag_own_bytes = 0; PM_CUDA(cudaMalloc(&ag_own, own_need)); ag_own_bytes = own_need; }
      if (all_need > ag_all_bytes) { if (ag_all) cudaFree(ag_all); ag_all = nullptr; ag_all_bytes = 0; PM_CUDA(cudaMalloc(&ag_all, all_need)); ag_all_bytes = all_need; }
    }
    if (!ag_flag) PM_CUDA(cudaMalloc(&ag_flag, sizeof(int)));
    if (!h_agflag) PM_CUDA(cudaMallocHost(&h_agflag, sizeof(int)));
    const cudaMemcpyKind up = own_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    uint8_t* raw_stage = ag_own + static_cast<size_t>(n_slots) * slot_w;
    if (wire != dtype_) PM_CUDA(cudaMemsetAsync(ag_flag, 0, sizeof(int), ingest));
    // The wire: own images -> device (copy engine), one grouped all-gather per chunk of slots, asynchronous ingest of the
    // gathered images in id order.  The NCCL kernels do not fit next to the persistent kNN kernels, so a gather that is
    // queued while the pair loop runs has to wait for a gap between two of them.  Order (default): the HEAD -- the first
    // ~32 images -- is uploaded, gathered and ingested first, so that every rank starts matching (block-cyclic shares begin
    // with pairs of the first images) while the copy engine brings in the rest of its own images; ONE gather for all the
    // remaining slots follows and finds its gap when the kNN stream runs out of batches whose images are resident.
    // debug_flags bit 24 keeps the earlier order (all gathers in chunks of doubling size, then all ingests: the whole
    // upload of the rank's own images is exposed -- 6 ms of a 41 ms step at 2 GPUs).
    auto wire_chunk = [&](int s0, int s1) -> int {
      for (int sl = s0; sl < s1; ++sl) {
        uint8_t* w = ag_own + static_cast<size_t>(sl) * slot_w;
        if (sl < n_own) {
          const uint8_t* src = static_cast<const uint8_t*>(own_desc) + static_cast<size_t>(sl) * img_in;
          if (wire == dtype_) {
            PM_CUDA(cudaMemcpyAsync(w, src, img_in, up, ingest));
          } else {                                                 // f32 rows -> bytes on the device (checked integral)
            uint8_t* st = raw_stage + static_cast<size_t>(sl) * img_in;
            PM_CUDA(cudaMemcpyAsync(st, src, img_in, up, ingest));
            PM_CUDA(launch_f32_to_u8(reinterpret_cast<const float*>(st), w, static_cast<size_t>(n_kp) * dim_, ag_flag, ingest));
            ++stats.kernel_launches;
          }
          if (has_xy) PM_CUDA(cudaMemcpyAsync(w + img_w, own_xy + 2 * static_cast<size_t>(sl) * n_kp, img_xy, up, ingest));
          if (!own_on_device) stats.h2d_bytes += img_in + (has_xy ? img_xy : 0);
        } else {
          PM_CUDA(cudaMemsetAsync(w, 0, slot_w, ingest));          // ranks without an image in the last slot send zeros
        }
      }
      int rc = nccl_api().GroupStart();
      for (int sl = s0; sl < s1 && rc == kNcclSuccess; ++sl)
        rc = nccl_api().AllGather(ag_own + static_cast<size_t>(sl) * slot_w, ag_all + static_cast<size_t>(sl) * R * slot_w,
                                  slot_w, kNcclUint8, comm, ingest);
      const int rc_end = nccl_api().GroupEnd();
      if (rc == kNcclSuccess) rc = rc_end;
      if (rc != kNcclSuccess) return fail(PM_ERR_CUDA, "ncclAllGather: %s", nccl_api().GetErrorString(rc));
      return PM_OK;
    };
    auto ingest_images = [&](int img0, int img1) -> int {
      for (int img = img0; img < img1; ++img) {
        const uint8_t* g = ag_all + static_cast<size_t>(img) * slot_w;   // slot s, rank r sits at (s R + r) = image id
        const int rc2 = set_image(img, g, n_kp, dim_, wire, has_xy ? reinterpret_cast<const int32_t*>(g + img_w) : nullptr, true, true);
        if (rc2 != PM_OK) return rc2;
      }
      return PM_OK;
    };
    const int first = std::min(n_slots, std::max(1, (32 + R - 1) / R));
    int rcw = PM_OK;
    if ((prm.debug_flags >> 24) & 1) {
      for (int s0 = 0, s1 = first; s0 < n_slots && rcw == PM_OK; s0 = s1, s1 = std::min(n_slots, 2 * s1)) rcw = wire_chunk(s0, s1);
      if (rcw == PM_OK) rcw = ingest_images(0, n_total);
    } else if ((prm.debug_flags >> 25) & 1) {
      rcw = wire_chunk(0, first);
      if (rcw == PM_OK) rcw = ingest_images(0, std::min(n_total, first * R));
      if (rcw == PM_OK && first < n_slots) rcw = wire_chunk(first, n_slots);
      if (rcw == PM_OK && first * R < n_total) rcw = ingest_images(first * R, n_total);
    } else {
      // chunks of doubling size, each gathered and ingested before the next one is uploaded: the pairs among the first
      // k images take longer to match than the next k images take to arrive (quadratic against linear)
      for (int s0 = 0, s1 = first; s0 < n_slots && rcw == PM_OK; s0 = s1, s1 = std::min(n_slots, 2 * s1)) {
        rcw = wire_chunk(s0, s1);
        if (rcw == PM_OK) rcw = ingest_images(s0 * R, std::min(n_total, s1 * R));
      }
    }
    if (rcw != PM_OK) return rcw;
    if (wire != dtype_) {                                          // the caller promised integer-valued rows: verify
      PM_CUDA(cudaMemcpyAsync(h_agflag, ag_flag, sizeof(int), cudaMemcpyDeviceToHost, ingest));
      ag_check_pending = true;
    }
    return PM_OK;
  }
  bool ag_check_pending = false;

  int fill_job(Slot& s, int k, int i, int j) {
    auto a = images.find(i), b = images.find(j);
    if (a == images.end() || b == images.end())
      return fail(PM_ERR_STATE, "image id %d not set", a == images.end() ? i : j);
    if (a->second.pending || b->second.pending) {
      int rc = resolve(a->second);
      if (rc == PM_OK) rc = resolve(b->second);
      if (rc != PM_OK) return rc;
    }
    // Philox key of the pair: a function of (params.seed, image ids) only, so a pair's result does not depend on the
    // batch, the device or the rank that ran it; the CPU filter of the parity tests mixes the same way
    const uint64_t ps = pair_seed(prm.seed, i, j);
    s.h_jobs[k] = PairJob{a->second.row, b->second.row, a->second.n, b->second.n, a->second.maxn, b->second.maxn,
                          static_cast<uint32_t>(ps), static_cast<uint32_t>(ps >> 32)};
    if (static_cast<int>(job_integral.size()) <= k) { job_integral.resize(k + 1); job_unit.resize(k + 1); job_i8.resize(k + 1); job_s8.resize(k + 1); job_u1.resize(k + 1); }
    // an empty image is compatible with every path (its pairs have no rows to search)
    const Image &ia = a->second, &ib = b->second;
    job_integral[k] = (ia.integral || ia.n == 0) && (ib.integral || ib.n == 0);
    job_unit[k] = (ia.unit_ok || ia.n == 0) && (ib.unit_ok || ib.n == 0);
    job_i8[k] = (ia.i8_ok || ia.n == 0) && (ib.i8_ok || ib.n == 0);
    job_s8[k] = (ia.s8_ok || ia.n == 0) && (ib.s8_ok || ib.n == 0);
    job_u1[k] = ib.unit1_ok || ib.n == 0;              // the TRAIN image's norms are the ones the search would add
    return PM_OK;
  }

  int max_keypoints() const {
    int m = 0;
    for (auto& kv : images) m = std::max(m, kv.second.n);
    return m;
  }

  // Queues select -> ransac -> compact -> D2H(meta) for the batch in slot s.
  int enqueue_tail(Slot& s, int n, bool do_filter) {
    const bool mutual = prm.unique_mode == PM_MUTUAL_NN;
    PM_CUDA(launch_select(s.d_jobs, n, s.knn_idx, s.knn_dist, mutual ? s.rev_idx : nullptr, xy, s.stride,
                          prm.ratio, prm.unique_mode, s.owner, s.match_q, s.match_t, s.pts1, s.pts2,
                          s.count, s.stream));
    if (trace_on && s.ev_tr[1]) PM_CUDA(cudaEventRecord(s.ev_tr[1], s.stream));
    int rs_launches = 0;
    const bool staged = !((prm.debug_flags >> 21) & 1) && (staged_hint || ((prm.debug_flags >> 22) & 1));
    PM_CUDA(launch_ransac(s.pts1, s.pts2, s.count, n, s.stride, ransac_dev(do_filter), s.mask, s.F,
                          s.status, s.n_inl, s.iters, s.stream, s.d_jobs, staged ? s.rs_ws : nullptr, &rs_launches));
    stats.kernel_launches += rs_launches - 1;            // (the per-pair kernel is part of the 4 below)
    if (trace_on && s.ev_tr[2]) PM_CUDA(cudaEventRecord(s.ev_tr[2], s.stream));
    PM_CUDA(launch_compact(s.count, n, s.stride, s.match_q, s.match_t, s.mask, s.offsets, s.out_q,
                           s.out_t, s.out_mask, s.stream));
    stats.kernel_launches += 4;
    // by kernels writing pinned host memory, not the copy engine (see launch_copy_pinned)
    PM_CUDA(launch_copy_pinned(s.offsets, s.h_offsets, 8 * static_cast<size_t>(n + 1), s.stream));
    PM_CUDA(launch_copy_pinned(s.F, s.h_F, 72 * static_cast<size_t>(n), s.stream));
    PM_CUDA(launch_copy_pinned(s.status, s.h_status, 4 * static_cast<size_t>(n), s.stream));
    PM_CUDA(launch_copy_pinned(s.n_inl, s.h_ninl, 4 * static_cast<size_t>(n), s.stream));
    PM_CUDA(launch_copy_pinned(s.iters, s.h_iters, 4 * static_cast<size_t>(n), s.stream));
    stats.kernel_launches += 5;
    PM_CUDA(cudaEventRecord(s.ev_done, s.stream));
    stats.d2h_bytes += 8 * (n + 1) + 72 * n + 12 * n;
    return PM_OK;
  }

  // Waits for the batch's per-pair metadata, then queues the D2H of its compacted matches straight
  // into the (pinned) result arrays; nothing is copied on the host.
  int retrieve(Slot& s, Result& R) {
    if (!s.busy) return PM_OK;
    PM_CUDA(cudaEventSynchronize(s.ev_done));
    const int n = s.n_jobs;
    const int64_t total = s.h_offsets[n];
    const int64_t base = R.size;
    if (base + total > R.cap) {
      // growing moves the arrays: drain every copy that targets the old ones first
      for (auto& o : slots) if (o.stream) PM_CUDA(cudaStreamSynchronize(o.stream));
      if (single.stream) PM_CUDA(cudaStreamSynchronize(single.stream));
      if (!R.reserve(base + total)) return fail(PM_ERR_OOM, "pinned result buffers (%lld matches)", static_cast<long long>(base + total));
    }
    if (total > 0) {
      if (result_by_ce) {
        // the compacted matches (MBs per batch) stay on the copy engine: as kernels they cost 3 % of the resident step
        // (SM slots + PCIe writes) and measured no better end to end; PM_RESULT_COPY=kernel selects them (development)
        PM_CUDA(cudaMemcpyAsync(R.q + base, s.out_q, 4 * total, cudaMemcpyDeviceToHost, s.stream));
        PM_CUDA(cudaMemcpyAsync(R.t + base, s.out_t, 4 * total, cudaMemcpyDeviceToHost, s.stream));
        PM_CUDA(cudaMemcpyAsync(R.inlier + base, s.out_mask, total, cudaMemcpyDeviceToHost, s.stream));
      } else {
        PM_CUDA(launch_copy_pinned(s.out_q, R.q + base, 4 * static_cast<size_t>(total), s.stream));
        PM_CUDA(launch_copy_pinned(s.out_t, R.t + base, 4 * static_cast<size_t>(total), s.stream));
        PM_CUDA(launch_copy_pinned(s.out_mask, R.inlier + base, static_cast<size_t>(total), s.stream));
        stats.kernel_launches += 3;
      }
      stats.d2h_bytes += 9 * total;
    }
    R.size = base + total;
    if (s.timed) {
      float ms = 0;
      PM_CUDA(cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
      stats.knn_ms += ms;
      if (trace_on) {                     // development: timeline of the batch relative to the call's first event
        float t0 = 0, t1 = 0, t2 = 0;
        cudaEventElapsedTime(&t0, ev_a, s.ev_k0); cudaEventElapsedTime(&t1, ev_a, s.ev_k1);
        cudaEventElapsedTime(&t2, ev_a, s.ev_done);
        float f0 = -1, f1 = -1, f2 = -1;
        if (s.ev_tr[0]) { cudaEventElapsedTime(&f0, ev_a, s.ev_tr[0]); cudaEventElapsedTime(&f1, ev_a, s.ev_tr[1]); cudaEventElapsedTime(&f2, ev_a, s.ev_tr[2]); }
        std::fprintf(stderr, "[pm trace] batch pair0=%lld n=%d knn %.3f..%.3f ms fix-up done %.3f select %.3f ransac %.3f tail done %.3f ms\n",
                     static_cast<long long>(s.first_pair), s.n_jobs, t0, t1, f0, f1, f2, t2);
      }
      stats.knn_launches += 1;
      stats.knn_work += s.knn_work;
    }
    int long_runs = 0;
    for (int k = 0; k < n; ++k) {
      const int64_t p = s.first_pair + k;
      R.offsets[p + 1] = base + s.h_offsets[k + 1];
      R.status[p] = s.h_status[k];
      R.n_inliers[p] = s.h_ninl[k];
      R.iters[p] = s.h_iters[k];
      long_runs += s.h_iters[k] > ransac_stage_cut() ? 1 : 0;
      std::memcpy(&R.F[9 * p], &s.h_F[9 * k], 72);
      stats.putative_matches += s.h_offsets[k + 1] - s.h_offsets[k];
      stats.inlier_matches += s.h_status[k] == PM_PAIR_DROPPED ? 0 : s.h_ninl[k];
    }
    stats.pairs_matched += n;
    // Scheduling hint only (both forms of the filter give identical results): the staged continuation costs a dozen
    // nearly empty launches per batch, so it is queued while recent batches held pairs that ran past the first rounds.
    staged_hint = long_runs > 0;
    s.busy = false;
    return PM_OK;
  }

  // pairs: [n][2]; results appended at R positions [first, first+n)
  // Error exits of run_pairs: nothing of the failed call may stay in flight (its Result -- the target of queued D2H
  // copies -- is about to be destroyed, and a later call must not find slots that still look busy).
  void abandon_batches() {
    if (knn_stream) (void)cudaStreamSynchronize(knn_stream);
    for (auto& s : slots) {
      if (s.stream) (void)cudaStreamSynchronize(s.stream);
      s.busy = false;
    }
    (void)cudaGetLastError();
  }

  int run_pairs(const int32_t* pairs, int64_t n_pairs, int64_t first, Result& R, double* device_ms) {
    const int rc = run_pairs_impl(pairs, n_pairs, first, R, device_ms);
    if (rc != PM_OK) abandon_batches();
    return rc;
  }

  int run_pairs_impl(const int32_t* pairs, int64_t n_pairs, int64_t first, Result& R, double* device_ms) {
    PM_CUDA(cudaSetDevice(dev));
    if (n_pairs == 0) return PM_OK;
    if (dtype < 0) return fail(PM_ERR_STATE, "no images set");
    for (auto& s : slots) s.busy = false;              // (a failed earlier call drained them in abandon_batches)
    // every image id is checked before the first batch is queued: an unknown id at pair 10,000 must not fail the call
    // with batches in flight
    for (int64_t k = 0; k < 2 * n_pairs; ++k)
      if (images.find(pairs[k]) == images.end())
        return fail(PM_ERR_STATE, "image id %d not set (pair %lld)", pairs[k], static_cast<long long>(k / 2));
    const int maxn = std::max(max_keypoints(), 1);
    const int stride = (maxn + 255) / 256 * 256;
    int B = prm.batch_pairs > 0 ? prm.batch_pairs : static_cast<int>(std::clamp<int64_t>((2 << 20) / stride, 32, 2048));
    B = static_cast<int>(std::min<int64_t>(B, n_pairs));
    // batches in flight: the kNN kernels run back to back on their own stream, up to S batches ahead of the tails
    // (6: outlier-heavy pairs have a long chain of staged filter kernels per batch, +3 % there over 4; otherwise
    // measured: 8 or 16 instead of 4 changes nothing -- where the tails cannot be co-resident at a useful occupancy
    // they are throughput-bound, not latency-bound, once they get the machine).  PM_SLOTS overrides (development).
    const int S = n_pairs > B ? (opt_slots > 0 ? opt_slots : 6) : 1;
    const bool mutual = prm.unique_mode == PM_MUTUAL_NN;
    if (static_cast<int>(slots.size()) < S) slots.resize(S);
    for (int s = 0; s < S; ++s) {
      const int rc = ensure_slot(slots[s], B, stride, mutual);
      if (rc != PM_OK) return rc;
    }
    if (!R.reserve(R.size + std::max<int64_t>(1024, n_pairs * static_cast<int64_t>(stride) * 3 / 10)))
      return fail(PM_ERR_OOM, "pinned result buffers");
    PM_CUDA(cudaEventRecord(ev_a, knn_stream));
    const bool trace_host = trace_on;
    const auto t_host0 = std::chrono::steady_clock::now();
    trace_t0 = t_host0;
    int64_t done = 0;
    int b = 0;
    // results must be appended in pair order: retrieve slots in issue order
    while (done < n_pairs) {
      Slot& s = slots[b % S];
      int rc = retrieve(s, R);
      if (rc != PM_OK) return rc;
      // the first batches are short: with asynchronous ingest they need only the first few images (pair lists run
      // image by image), so the device starts after a fraction of the upload
      const int ramp = b < 3 ? std::max(16, B >> (3 - b)) : B;
      // ... and the last batches shrink again (halving, down to 32 pairs): what follows the last kNN kernel of the call
      // -- fix-up, selection, filter, compaction, D2H of its batch -- overlaps nothing, so that batch is kept small
      const int64_t rem = n_pairs - done;
      int64_t want = std::min(B, ramp);
      if (opt_rampdown && rem < 2 * static_cast<int64_t>(B)) want = std::min<int64_t>(want, std::max<int64_t>(32, (rem + 1) / 2));
      const int n = static_cast<int>(std::min<int64_t>(want, rem));
      for (int k = 0; k < n; ++k) {
        rc = fill_job(s, k, pairs[2 * (done + k)], pairs[2 * (done + k) + 1]);
        if (rc != PM_OK) return rc;
      }
      const auto t_fill = std::chrono::steady_clock::now();
      if ((rc = enqueue_knn(s, n, mutual, true, nullptr, true)) != PM_OK) return rc;
      if ((rc = enqueue_tail(s, n, prm.do_filter != 0)) != PM_OK) return rc;
      if (trace_host) {                      // development (PM_TRACE): when did the host get this batch out?
        const auto t_q = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[pm trace] host: batch %d (pair0=%lld) retrieved+filled at %.3f ms, queued at %.3f ms\n", b,
                     static_cast<long long>(first + done), std::chrono::duration<double, std::milli>(t_fill - t_host0).count(),
                     std::chrono::duration<double, std::milli>(t_q - t_host0).count());
      }
      s.busy = true; s.n_jobs = n; s.first_pair = first + done;
      done += n;
      ++b;
    }
    // drain in issue order
    for (int k = 0; k < S; ++k) {
      Slot& s = slots[(b + k) % S];
      const int rc = retrieve(s, R);
      if (rc != PM_OK) return rc;
    }
    // device clock: from ev_a to the completion of every stream (incl. the last D2H copies)
    for (int s = 1; s < S; ++s) {
      PM_CUDA(cudaEventRecord(slots[s].ev_done, slots[s].stream));
      PM_CUDA(cudaStreamWaitEvent(slots[0].stream, slots[s].ev_done, 0));
    }
    PM_CUDA(cudaEventRecord(ev_b, slots[0].stream));
    PM_CUDA(cudaEventSynchronize(ev_b));
    float ms = 0;
    PM_CUDA(cudaEventElapsedTime(&ms, ev_a, ev_b));
    if (device_ms) *device_ms = ms;
    return PM_OK;
  }

  // single-pair helpers ----------------------------------------------------------------------
  int single_prepare(int i, int j, bool need_rev) {
    PM_CUDA(cudaSetDevice(dev));
    auto a = images.find(i), b = images.find(j);
    if (a == images.end() || b == images.end())
      return fail(PM_ERR_STATE, "image id %d not set", a == images.end() ? i : j);
    const int stride = (std::max({a->second.n, b->second.n, 1}) + 255) / 256 * 256;
    int rc = ensure_slot(single, 1, stride, need_rev);
    if (rc != PM_OK) return rc;
    return fill_job(single, 0, i, j);
  }
};

}  // namespace

// ================================================================================================
struct pm_context {
  std::mutex mu;
  std::vector<std::unique_ptr<DeviceCtx>> devs;
  pm_params prm{};
  std::string err;

  int fail(int code, const std::string& m) { err = m; return code; }
  int from(DeviceCtx& d, int rc) { if (rc != PM_OK) err = d.err; return rc; }
};

extern "C" {

const char* pm_version(void) { return "pairmatch_b200 0.1 (sm_100a)"; }

void pm_default_params(pm_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof *p);
  p->ratio = 0.7f;                 // FeatureMatcher.h:45
  p->unique_mode = PM_UNIQUE_FIRST_WINS;
  p->min_matches = 7;              // SequentialReconstructor.cpp:237
  p->do_filter = 1;
  p->ransac_threshold = 3.0;       // cv::findFundamentalMat defaults (GeometricFilter.cpp:47)
  p->ransac_confidence = 0.99;
  p->ransac_max_iters = 1000;
  p->residual_mode = PM_RESID_SYMMETRIC_EPIPOLAR;
  p->sampler = PM_SAMPLER_OPENCV_MWC;
  p->essential_confidence = 0.999; // cv::findEssentialMat defaults (GeometricFilter.cpp:26-31)
  p->essential_threshold = 1.0;
}

int pm_create(const pm_params* p, const int* device_ids, int n_dev, pm_handle* out) {
  if (!out) { g_create_error = "pm_create: out is NULL"; return PM_ERR_INVALID; }
  *out = nullptr;
  pm_params prm;
  if (p) prm = *p; else pm_default_params(&prm);
  if (prm.sampler != PM_SAMPLER_OPENCV_MWC && prm.sampler != PM_SAMPLER_PHILOX) { g_create_error = "unknown sampler"; return PM_ERR_UNSUPPORTED; }
  if (prm.unique_mode < 0 || prm.unique_mode > 2 || prm.residual_mode < 0 || prm.residual_mode > 1 ||
      !(prm.ratio > 0.f) || prm.ransac_max_iters < 1 || prm.min_matches < 0) {
    g_create_error = "pm_create: invalid parameters";
    return PM_ERR_INVALID;
  }
  if (prm.min_matches < 7 && prm.do_filter) prm.min_matches = 7;   // the 7-point solver needs 7
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    (void)cudaGetLastError();
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                     " (this library has no CPU fallback)";
    return PM_ERR_NO_DEVICE;
  }
  std::vector<int> ids;
  if (!device_ids || n_dev <= 0) {
    int cur = 0;
    cudaGetDevice(&cur);
    ids.push_back(cur);
  } else {
    ids.assign(device_ids, device_ids + n_dev);
  }
  auto ctx = std::make_unique<pm_context>();
  ctx->prm = prm;
  for (int id : ids) {
    if (id < 0 || id >= count) {
      g_create_error = "pm_create: bad device id";
      for (auto& e2 : ctx->devs) e2->shutdown();
      return PM_ERR_INVALID;
    }
    auto d = std::make_unique<DeviceCtx>();
    const int rc = d->init(id, prm);
    if (rc != PM_OK) {
      g_create_error = d->err;
      d->shutdown();
      for (auto& e2 : ctx->devs) e2->shutdown();       // the devices initialised before this one
      return rc;
    }
    ctx->devs.push_back(std::move(d));
  }
  *out = ctx.release();
  return PM_OK;
}

int pm_destroy(pm_handle h) {
  if (!h) return PM_OK;
  for (auto& d : h->devs) d->shutdown();
  delete h;
  return PM_OK;
}

const char* pm_last_error(pm_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int pm_set_image(pm_handle h, int img_id, const void* desc, int n, int dim, int dtype, const int32_t* xy) {
  if (!h) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  if (img_id == kTmpA || img_id == kTmpB) return h->fail(PM_ERR_INVALID, "reserved image id");
  for (auto& d : h->devs) {
    const int rc = d->set_image(img_id, desc, n, dim, dtype, xy, false);
    if (rc != PM_OK) return h->from(*d, rc);
  }
  return PM_OK;
}

int pm_set_image_async(pm_handle h, int img_id, const void* desc, int n, int dim, int dtype, const int32_t* xy) {
  if (!h) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  if (img_id == kTmpA || img_id == kTmpB) return h->fail(PM_ERR_INVALID, "reserved image id");
  for (auto& d : h->devs) {
    const int rc = d->set_image(img_id, desc, n, dim, dtype, xy, false, true);
    if (rc != PM_OK) return h->from(*d, rc);
  }
  return PM_OK;
}

int pm_set_images_async(pm_handle h, int n_images, const int* img_ids, const void* const* descs, const int* ns,
                        int dim, int dtype, const int32_t* const* xys) {
  if (!h) return PM_ERR_INVALID;
  if (n_images < 0 || (n_images > 0 && (!img_ids || !descs || !ns))) return h->fail(PM_ERR_INVALID, "set_images: bad arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  for (int k = 0; k < n_images; ++k) {
    if (img_ids[k] == kTmpA || img_ids[k] == kTmpB) return h->fail(PM_ERR_INVALID, "reserved image id");
    for (auto& d : h->devs) {
      const int rc = d->set_image(img_ids[k], descs[k], ns[k], dim, dtype, xys ? xys[k] : nullptr, false, true);
      if (rc != PM_OK) return h->from(*d, rc);
    }
  }
  return PM_OK;
}

int pm_sync_images(pm_handle h) {
  if (!h) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  for (auto& d : h->devs) {
    if (cudaSetDevice(d->dev) != cudaSuccess) return h->fail(PM_ERR_CUDA, "cudaSetDevice failed");
    const int rc = d->resolve_all();
    if (rc != PM_OK) return h->from(*d, rc);
    if (cudaStreamSynchronize(d->ingest) != cudaSuccess) return h->fail(PM_ERR_CUDA, "ingest stream failed");
    if (d->ag_check_pending) {
      d->ag_check_pending = false;
      if (*d->h_agflag != 0) return h->from(*d, d->fail(PM_ERR_INVALID, "pm_ingest_allgather: wire dtype U8 was requested for rows that are not integer-valued in [0,255]"));
    }
  }
  return PM_OK;
}

int pm_set_image_device(pm_handle h, int img_id, const void* d_desc, int n, int dim, int dtype,
                        const int32_t* d_xy) {
  if (!h) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->devs.size() != 1) return h->fail(PM_ERR_UNSUPPORTED, "pm_set_image_device needs a single-device handle");
  return h->from(*h->devs[0], h->devs[0]->set_image(img_id, d_desc, n, dim, dtype, d_xy, true));
}

int pm_set_image_device_async(pm_handle h, int img_id, const void* d_desc, int n, int dim, int dtype,
                              const int32_t* d_xy) {
  if (!h) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->devs.size() != 1) return h->fail(PM_ERR_UNSUPPORTED, "pm_set_image_device needs a single-device handle");
  return h->from(*h->devs[0], h->devs[0]->set_image(img_id, d_desc, n, dim, dtype, d_xy, true, true));
}

int pm_num_keypoints(pm_handle h, int img_id) {
  if (!h) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  auto& im = h->devs[0]->images;
  auto it = im.find(img_id);
  return it == im.end() ? PM_ERR_STATE : it->second.n;
}

int pm_remove_image(pm_handle h, int img_id) {
  if (!h) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  if (img_id == kTmpA || img_id == kTmpB) return h->fail(PM_ERR_INVALID, "reserved image id");
  for (auto& d : h->devs) {
    const int rc = d->remove_image(img_id);
    if (rc != PM_OK) return h->from(*d, rc);
  }
  return PM_OK;
}

static int knn_single(pm_context* h, DeviceCtx& d, int i, int j, int32_t* idx, float* dist, float* dump) {
  int rc = d.single_prepare(i, j, false);
  if (rc != PM_OK) return h->from(d, rc);
  Slot& s = d.single;
  if ((rc = d.enqueue_knn(s, 1, false, false, dump)) != PM_OK) return h->from(d, rc);
  const int nq = s.h_jobs[0].nq;
  if (nq > 0) {
    if (cudaMemcpyAsync(idx, s.knn_idx, sizeof(int2) * nq, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess ||
        cudaMemcpyAsync(dist, s.knn_dist, sizeof(float2) * nq, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess)
      return h->from(d, d.fail(PM_ERR_CUDA, "D2H of kNN rows failed"));
  }
  cudaError_t e = cudaStreamSynchronize(s.stream);
  if (e != cudaSuccess) return h->from(d, d.fail_cuda(e, "pm_knn_pair sync", __LINE__));
  d.stats.d2h_bytes += 16ll * nq;
  return PM_OK;
}

int pm_knn_pair(pm_handle h, int img_i, int img_j, int32_t* idx, float* dist) {
  if (!h || !idx || !dist) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  return knn_single(h, *h->devs[0], img_i, img_j, idx, dist, nullptr);
}

// Debug aid for the tensor path: also returns the raw fp32 accumulators (nb - 2 a.b) of the
// first 256 x 128 block.  Not part of the drop-in surface.
int pm_debug_tc_dump(pm_handle h, int img_i, int img_j, int32_t* idx, float* dist, float* acc256x128) {
  if (!h || !idx || !dist || !acc256x128) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceCtx& d = *h->devs[0];
  cudaSetDevice(d.dev);
  if (!d.d_dump && cudaMalloc(&d.d_dump, 256 * 128 * 4) != cudaSuccess) return h->fail(PM_ERR_OOM, "dump alloc");
  cudaMemset(d.d_dump, 0, 256 * 128 * 4);
  const int rc = knn_single(h, d, img_i, img_j, idx, dist, d.d_dump);
  if (rc != PM_OK) return rc;
  if (cudaMemcpy(acc256x128, d.d_dump, 256 * 128 * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
    return h->fail(PM_ERR_CUDA, "dump D2H");
  return PM_OK;
}

static int pair_single(pm_context* h, DeviceCtx& d, int i, int j, bool do_filter, pm_pair_result* out) {
  const bool mutual = d.prm.unique_mode == PM_MUTUAL_NN;
  int rc = d.single_prepare(i, j, mutual);
  if (rc != PM_OK) return h->from(d, rc);
  Slot& s = d.single;
  if (out->capacity < s.h_jobs[0].nq)
    return h->from(d, d.fail(PM_ERR_INVALID, "pm_pair_result.capacity %d < query keypoints %d", out->capacity,
                             s.h_jobs[0].nq));
  if ((rc = d.enqueue_knn(s, 1, mutual, false)) != PM_OK) return h->from(d, rc);
  if ((rc = d.enqueue_tail(s, 1, do_filter)) != PM_OK) return h->from(d, rc);
  s.busy = true; s.n_jobs = 1; s.first_pair = 0; s.timed = false;
  Result R;
  R.pool = d.pool;
  R.offsets.assign(2, 0); R.status.assign(1, 0); R.n_inliers.assign(1, 0); R.iters.assign(1, 0);
  R.F.assign(9, 0.0);
  if ((rc = d.retrieve(s, R)) != PM_OK) return h->from(d, rc);
  if (cudaStreamSynchronize(s.stream) != cudaSuccess) return h->from(d, d.fail(PM_ERR_CUDA, "pair D2H failed"));
  const int m = static_cast<int>(R.size);
  out->n_matches = m;
  out->status = R.status[0];
  out->n_inliers = R.status[0] == PM_PAIR_DROPPED ? 0 : R.n_inliers[0];
  out->ransac_iters = R.iters[0];
  std::memcpy(out->F, R.F.data(), 72);
  if (m > 0) {
    if (out->q) std::memcpy(out->q, R.q, 4 * static_cast<size_t>(m));
    if (out->t) std::memcpy(out->t, R.t, 4 * static_cast<size_t>(m));
    if (out->inlier) std::memcpy(out->inlier, R.inlier, static_cast<size_t>(m));
  }
  return PM_OK;
}

int pm_match_pair(pm_handle h, int img_i, int img_j, pm_pair_result* out) {
  if (!h || !out) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  return pair_single(h, *h->devs[0], img_i, img_j, false, out);
}

int pm_match_filter_pair(pm_handle h, int img_i, int img_j, pm_pair_result* out) {
  if (!h || !out) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  return pair_single(h, *h->devs[0], img_i, img_j, h->prm.do_filter != 0, out);
}

int pm_match_descriptors(pm_handle h, const void* desc1, int n1, const void* desc2, int n2, int dim,
                         int dtype, pm_pair_result* out) {
  if (!h || !out) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceCtx& d = *h->devs[0];
  int rc = d.set_image(kTmpA, desc1, n1, dim, dtype, nullptr, false);
  if (rc != PM_OK) return h->from(d, rc);
  if ((rc = d.set_image(kTmpB, desc2, n2, dim, dtype, nullptr, false)) != PM_OK) return h->from(d, rc);
  return pair_single(h, d, kTmpA, kTmpB, false, out);
}

uint64_t pm_pair_seed(uint64_t seed, int32_t img_i, int32_t img_j) { return pair_seed(seed, img_i, img_j); }

static int filter_pair_F(pm_handle h, const float* xy1, const float* xy2, int M, uint64_t key, double F[9], uint8_t* mask,
                         int32_t* status, int32_t* iters);

int pm_filter_pair_F(pm_handle h, const float* xy1, const float* xy2, int M, double F[9], uint8_t* mask,
                     int32_t* status, int32_t* iters) {
  if (!h) return PM_ERR_INVALID;
  return filter_pair_F(h, xy1, xy2, M, h->prm.seed, F, mask, status, iters);
}

int pm_filter_pair_F_seeded(pm_handle h, const float* xy1, const float* xy2, int M, uint64_t pair_key, double F[9],
                            uint8_t* mask, int32_t* status, int32_t* iters) {
  return filter_pair_F(h, xy1, xy2, M, pair_key, F, mask, status, iters);
}

static int filter_pair_F(pm_handle h, const float* xy1, const float* xy2, int M, uint64_t key, double F[9], uint8_t* mask,
                         int32_t* status, int32_t* iters) {
  if (!h || M < 0 || (M > 0 && (!xy1 || !xy2 || !mask)) || !F) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceCtx& d = *h->devs[0];
  cudaSetDevice(d.dev);
  const int stride = (std::max(M, 1) + 255) / 256 * 256;
  int rc = d.ensure_slot(d.single, 1, stride, false);
  if (rc != PM_OK) return h->from(d, rc);
  Slot& s = d.single;
  auto ck = [&](cudaError_t e, const char* what) { return e == cudaSuccess ? PM_OK : h->from(d, d.fail_cuda(e, what, __LINE__)); };
  s.h_count[0] = M;
  if ((rc = ck(cudaMemcpyAsync(s.count, s.h_count, 4, cudaMemcpyHostToDevice, s.stream), "count H2D"))) return rc;
  if (M > 0) {
    if ((rc = ck(cudaMemcpyAsync(s.pts1, xy1, 8ull * M, cudaMemcpyHostToDevice, s.stream), "pts1 H2D"))) return rc;
    if ((rc = ck(cudaMemcpyAsync(s.pts2, xy2, 8ull * M, cudaMemcpyHostToDevice, s.stream), "pts2 H2D"))) return rc;
  }
  RansacDev rp = d.ransac_dev(true);
  rp.seed = key;
  rp.min_matches = 7;   // estimateFundamental itself has no gate; < 7 points cannot be solved
  int rs_launches = 0;
  if ((rc = ck(launch_ransac(s.pts1, s.pts2, s.count, 1, s.stride, rp, s.mask, s.F, s.status, s.n_inl, s.iters,
                             s.stream, nullptr, ((d.prm.debug_flags >> 21) & 1) ? nullptr : s.rs_ws, &rs_launches),
               "ransac launch"))) return rc;
  d.stats.kernel_launches += rs_launches;
  if (M > 0 && (rc = ck(cudaMemcpyAsync(mask, s.mask, M, cudaMemcpyDeviceToHost, s.stream), "mask D2H"))) return rc;
  if ((rc = ck(cudaMemcpyAsync(s.h_F, s.F, 72, cudaMemcpyDeviceToHost, s.stream), "F D2H"))) return rc;
  if ((rc = ck(cudaMemcpyAsync(s.h_status, s.status, 4, cudaMemcpyDeviceToHost, s.stream), "status D2H"))) return rc;
  if ((rc = ck(cudaMemcpyAsync(s.h_iters, s.iters, 4, cudaMemcpyDeviceToHost, s.stream), "iters D2H"))) return rc;
  if ((rc = ck(cudaStreamSynchronize(s.stream), "sync"))) return rc;
  d.stats.h2d_bytes += 16ll * M + 4;
  d.stats.d2h_bytes += M + 80;
  int st = s.h_status[0];
  if (M < 7) st = PM_PAIR_DROPPED;            // cv::findFundamentalMat returns an empty F for N < 7
  std::memcpy(F, s.h_F, 72);
  if (st != PM_PAIR_FILTERED) { std::memset(F, 0, 72); if (M > 0) std::memset(mask, 0, M); }
  if (status) *status = st;
  if (iters) *iters = s.h_iters[0];
  return PM_OK;
}

int pm_filter_pair_E(pm_handle h, const float* xy1, const float* xy2, int M, const pm_camera* cam1, const pm_camera* cam2,
                     double E[9], uint8_t* mask, int32_t* status, int32_t* iters) {
  if (!h || M < 0 || (M > 0 && (!xy1 || !xy2)) || !E || !cam1 || !cam2) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceCtx& d = *h->devs[0];
  if (!(cam1->fx > 0) || !(cam1->fy > 0) || !(cam2->fx > 0) || !(cam2->fy > 0))
    return h->from(d, d.fail(PM_ERR_INVALID, "pm_filter_pair_E: focal lengths must be positive"));
  cudaSetDevice(d.dev);
  const int stride = (std::max(M, 1) + 255) / 256 * 256;
  int rc = d.ensure_slot(d.single, 1, stride, false);
  if (rc != PM_OK) return h->from(d, rc);
  Slot& s = d.single;
  auto ck = [&](cudaError_t e, const char* what) { return e == cudaSuccess ? PM_OK : h->from(d, d.fail_cuda(e, what, __LINE__)); };
  const size_t need = emat_scratch_bytes(M);
  if (need > d.e_scratch_bytes) {
    if (d.e_scratch) cudaFree(d.e_scratch);
    d.e_scratch = nullptr; d.e_scratch_bytes = 0;
    if ((rc = ck(cudaMalloc(&d.e_scratch, need), "essential scratch"))) return rc;
    d.e_scratch_bytes = need;
  }
  if (M > 0) {
    if ((rc = ck(cudaMemcpyAsync(s.pts1, xy1, 8ull * M, cudaMemcpyHostToDevice, s.stream), "pts1 H2D"))) return rc;
    if ((rc = ck(cudaMemcpyAsync(s.pts2, xy2, 8ull * M, cudaMemcpyHostToDevice, s.stream), "pts2 H2D"))) return rc;
  }
  const double c1[6] = {cam1->fx, cam1->fy, cam1->cx, cam1->cy, cam1->k1, cam1->k2};
  const double c2[6] = {cam2->fx, cam2->fy, cam2->cx, cam2->cy, cam2->k1, cam2->k2};
  if ((rc = ck(launch_emat_ransac(s.pts1, s.pts2, M, c1, c2, d.prm.essential_confidence, d.prm.essential_threshold,
                                  d.prm.ransac_max_iters, d.prm.sampler, d.prm.seed, d.e_scratch, s.mask, s.F, s.status,
                                  s.n_inl, s.iters, s.stream), "essential launch"))) return rc;
  ++d.stats.kernel_launches;
  if (M > 0 && mask && (rc = ck(cudaMemcpyAsync(mask, s.mask, M, cudaMemcpyDeviceToHost, s.stream), "mask D2H"))) return rc;
  if ((rc = ck(cudaMemcpyAsync(s.h_F, s.F, 72, cudaMemcpyDeviceToHost, s.stream), "E D2H"))) return rc;
  if ((rc = ck(cudaMemcpyAsync(s.h_status, s.status, 4, cudaMemcpyDeviceToHost, s.stream), "status D2H"))) return rc;
  if ((rc = ck(cudaMemcpyAsync(s.h_iters, s.iters, 4, cudaMemcpyDeviceToHost, s.stream), "iters D2H"))) return rc;
  if ((rc = ck(cudaStreamSynchronize(s.stream), "sync"))) return rc;
  d.stats.h2d_bytes += 16ll * M;
  d.stats.d2h_bytes += (mask ? M : 0) + 80;
  const int st = s.h_status[0];
  std::memcpy(E, s.h_F, 72);
  if (st != PM_PAIR_FILTERED) { std::memset(E, 0, 72); if (M > 0 && mask) std::memset(mask, 0, M); }
  if (status) *status = st;
  if (iters) *iters = s.h_iters[0];
  return PM_OK;
}

int pm_match_all_pairs(pm_handle h, const int32_t* pairs, int64_t n_pairs, pm_csr_result** out) {
  if (!h || !out) return PM_ERR_INVALID;
  *out = nullptr;
  std::lock_guard<std::mutex> lk(h->mu);
  auto R = std::make_unique<Result>();
  R->pool = h->devs[0]->pool;
  if (n_pairs == 0 || (!pairs && n_pairs != PM_ALL_PAIRS)) {
    if (n_pairs != 0) return h->fail(PM_ERR_INVALID, "pairs is NULL: pass n_pairs = PM_ALL_PAIRS for the implicit all-pairs list");
    n_pairs = 0;                                          // an empty list is an empty result, not "all pairs"
  } else if (!pairs) {   // all i < j over the images set so far: the FakeImgMatcher pair list (ImageMatcher.cpp:6-24)
    std::vector<int> ids;
    for (auto& kv : h->devs[0]->images)
      if (kv.first != kTmpA && kv.first != kTmpB) ids.push_back(kv.first);
    std::sort(ids.begin(), ids.end());
    // every image against all EARLIER ones (query = lower id as at SequentialReconstructor.cpp:203-227): the first
    // batches touch only the first few images, so they overlap the upload of the rest (pm_set_image_async), and
    // consecutive pairs share their train image
    for (size_t b = 1; b < ids.size(); ++b)
      for (size_t a = 0; a < b; ++a) { R->pair_ij.push_back(ids[a]); R->pair_ij.push_back(ids[b]); }
    n_pairs = static_cast<int64_t>(R->pair_ij.size() / 2);
  } else {
    if (n_pairs < 0) return h->fail(PM_ERR_INVALID, "n_pairs < 0");
    R->pair_ij.assign(pairs, pairs + 2 * n_pairs);
  }
  R->offsets.assign(n_pairs + 1, 0);
  R->status.assign(n_pairs, 0); R->n_inliers.assign(n_pairs, 0); R->iters.assign(n_pairs, 0);
  R->F.assign(9 * n_pairs, 0.0);
  double ms = 0;
  const int nd = static_cast<int>(h->devs.size());
  if (nd == 1 || n_pairs < 2 * nd) {
    const int rc = h->devs[0]->run_pairs(R->pair_ij.data(), n_pairs, 0, *R, &ms);
    if (rc != PM_OK) return h->from(*h->devs[0], rc);
  } else {
    // contiguous shares of the pair list, one host thread per device; no data-path collective
    std::vector<std::unique_ptr<Result>> part(nd);
    std::vector<int> rcs(nd, PM_OK);
    std::vector<double> dms(nd, 0.0);
    std::vector<std::thread> th;
    std::vector<int64_t> lo(nd + 1);
    for (int k = 0; k <= nd; ++k) lo[k] = n_pairs * k / nd;
    for (int k = 0; k < nd; ++k) {
      part[k] = std::make_unique<Result>();
      part[k]->pool = h->devs[k]->pool;
      const int64_t n = lo[k + 1] - lo[k];
      part[k]->offsets.assign(n + 1, 0);
      part[k]->status.assign(n, 0); part[k]->n_inliers.assign(n, 0); part[k]->iters.assign(n, 0);
      part[k]->F.assign(9 * n, 0.0);
      th.emplace_back([&, k, n] {
        rcs[k] = h->devs[k]->run_pairs(R->pair_ij.data() + 2 * lo[k], n, 0, *part[k], &dms[k]);
      });
    }
    for (auto& t : th) t.join();
    for (int k = 0; k < nd; ++k)
      if (rcs[k] != PM_OK) return h->from(*h->devs[k], rcs[k]);
    for (int k = 0; k < nd; ++k) {
      Result& P = *part[k];
      const int64_t base = R->size;
      if (!R->reserve(base + P.size)) return h->fail(PM_ERR_OOM, "pinned result buffers");
      if (P.size > 0) {
        std::memcpy(R->q + base, P.q, 4 * static_cast<size_t>(P.size));
        std::memcpy(R->t + base, P.t, 4 * static_cast<size_t>(P.size));
        std::memcpy(R->inlier + base, P.inlier, static_cast<size_t>(P.size));
      }
      R->size = base + P.size;
      const int64_t n = lo[k + 1] - lo[k];
      for (int64_t p = 0; p < n; ++p) {
        R->offsets[lo[k] + p + 1] = base + P.offsets[p + 1];
        R->status[lo[k] + p] = P.status[p];
        R->n_inliers[lo[k] + p] = P.n_inliers[p];
        R->iters[lo[k] + p] = P.iters[p];
      }
      if (n > 0) std::memcpy(&R->F[9 * lo[k]], P.F.data(), 72 * n);
      ms = std::max(ms, dms[k]);
    }
  }
  R->pub.n_pairs = n_pairs;
  R->pub.pair_ij = R->pair_ij.data();
  R->pub.offsets = R->offsets.data();
  R->pub.q = R->q; R->pub.t = R->t; R->pub.inlier = R->inlier;
  R->pub.F = R->F.data(); R->pub.status = R->status.data();
  R->pub.n_inliers = R->n_inliers.data(); R->pub.ransac_iters = R->iters.data();
  R->pub.device_ms = ms;
  R->pub.owner_ = R.get();
  *out = &R.release()->pub;
  return PM_OK;
}

int pm_free_result(pm_csr_result* r) {
  if (!r) return PM_OK;
  delete static_cast<Result*>(r->owner_);
  return PM_OK;
}

int pm_get_stats(pm_handle h, pm_stats* out) {
  if (!h || !out) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  pm_stats s{};
  for (auto& d : h->devs) {
    s.pairs_matched += d->stats.pairs_matched; s.putative_matches += d->stats.putative_matches;
    s.inlier_matches += d->stats.inlier_matches; s.kernel_launches += d->stats.kernel_launches;
    s.h2d_bytes += d->stats.h2d_bytes; s.d2h_bytes += d->stats.d2h_bytes;
    s.knn_ms += d->stats.knn_ms; s.knn_launches += d->stats.knn_launches; s.knn_work += d->stats.knn_work;
    unsigned long long c[4] = {0, 0, 0, 0};
    if (cudaSetDevice(d->dev) == cudaSuccess && d->d_l2f &&
        cudaMemcpy(c, d->d_l2f, sizeof c, cudaMemcpyDeviceToHost) == cudaSuccess) {   // device-wide sync point
      s.rerank_rows += static_cast<int64_t>(c[0]); s.rerank_chunks += static_cast<int64_t>(c[1]);
      s.rerank_overflow += static_cast<int64_t>(c[2]);
      float w;
      const unsigned int wb = static_cast<unsigned int>(c[3]);
      std::memcpy(&w, &wb, sizeof w);
      s.rerank_worst_err = std::max(s.rerank_worst_err, static_cast<double>(w));
    } else {
      (void)cudaGetLastError();
    }
  }
  s.device_id = h->devs[0]->dev;
  s.n_images = h->devs[0]->stats.n_images;
  *out = s;
  return PM_OK;
}

int pm_reset_stats(pm_handle h) {
  if (!h) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  for (auto& d : h->devs) {
    const int dev = d->stats.device_id, ni = d->stats.n_images;
    d->stats = pm_stats{};
    d->stats.device_id = dev; d->stats.n_images = ni;
    if (cudaSetDevice(d->dev) == cudaSuccess && d->d_l2f) cudaMemset(d->d_l2f, 0, 4 * sizeof(unsigned long long));
  }
  return PM_OK;
}

int pm_measure_popc_peak(pm_handle h, double* popc32_per_s) {
  if (!h || !popc32_per_s) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceCtx& d = *h->devs[0];
  cudaSetDevice(d.dev);
  const int blocks = d.num_sms * 8, iters = 4096;
  uint32_t* buf = nullptr;
  if (cudaMalloc(&buf, 4ull * blocks * 256) != cudaSuccess) return h->fail(PM_ERR_OOM, "popc buffer");
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(d.ev_a, d.ingest);
    launch_popc_peak(buf, blocks, iters, d.ingest);
    cudaEventRecord(d.ev_b, d.ingest);
    if (cudaEventSynchronize(d.ev_b) != cudaSuccess) { cudaFree(buf); return h->from(d, d.fail(PM_ERR_CUDA, "popc kernel failed")); }
    float ms = 0;
    cudaEventElapsedTime(&ms, d.ev_a, d.ev_b);
    ++d.stats.kernel_launches;
    const double rate = static_cast<double>(blocks) * 256 * iters * 64 / (ms * 1e-3);
    if (rep > 0) best = std::max(best, rate);
  }
  cudaFree(buf);
  *popc32_per_s = best;
  return PM_OK;
}

int pm_comm_get_unique_id(uint8_t id[128]) {
  if (!id) return PM_ERR_INVALID;
  if (const char* why = nccl_api().load()) { g_create_error = std::string("NCCL unavailable: ") + why; return PM_ERR_UNSUPPORTED; }
  NcclUniqueId uid;
  const int rc = nccl_api().GetUniqueId(&uid);
  if (rc != kNcclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + nccl_api().GetErrorString(rc); return PM_ERR_CUDA; }
  std::memcpy(id, uid.internal, sizeof uid.internal);
  return PM_OK;
}

int pm_comm_init(pm_handle h, const uint8_t id[128], int rank, int n_ranks) {
  if (!h || !id) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->devs.size() != 1) return h->fail(PM_ERR_UNSUPPORTED, "pm_comm_init needs a single-device handle (one process per GPU)");
  return h->from(*h->devs[0], h->devs[0]->comm_init(id, rank, n_ranks));
}

int pm_ingest_allgather(pm_handle h, int n_images_total, int n_keypoints, int dim, int dtype, int wire_dtype,
                        const void* own_desc, const int32_t* own_xy, int own_on_device) {
  if (!h) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->devs.size() != 1) return h->fail(PM_ERR_UNSUPPORTED, "pm_ingest_allgather needs a single-device handle");
  return h->from(*h->devs[0], h->devs[0]->ingest_allgather(n_images_total, n_keypoints, dim, dtype, wire_dtype, own_desc,
                                                           own_xy, own_on_device != 0));
}

int pm_select_pairs(pm_handle h, int top_k, int32_t** pairs_out, int64_t* n_pairs_out, double* scores) {
  return pm_select_pairs_among(h, nullptr, 0, top_k, pairs_out, n_pairs_out, scores);
}

int pm_select_pairs_among(pm_handle h, const int32_t* img_ids, int n_ids, int top_k, int32_t** pairs_out,
                          int64_t* n_pairs_out, double* scores) {
  if (!h || !pairs_out || !n_pairs_out || n_ids < 0 || (n_ids > 0 && !img_ids)) return PM_ERR_INVALID;
  *pairs_out = nullptr; *n_pairs_out = 0;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceCtx& d = *h->devs[0];
  if (cudaSetDevice(d.dev) != cudaSuccess) return h->from(d, d.fail(PM_ERR_CUDA, "cudaSetDevice failed"));
  int rc = d.resolve_all();
  if (rc != PM_OK) return h->from(d, rc);
  if (cudaStreamSynchronize(d.ingest) != cudaSuccess) return h->from(d, d.fail(PM_ERR_CUDA, "ingest stream failed"));
  std::vector<int> ids;
  if (img_ids) {
    ids.assign(img_ids, img_ids + n_ids);
    std::sort(ids.begin(), ids.end());
    if (std::adjacent_find(ids.begin(), ids.end()) != ids.end())
      return h->from(d, d.fail(PM_ERR_INVALID, "pm_select_pairs_among: duplicate image id"));
    for (int id : ids)
      if (id == kTmpA || id == kTmpB || !d.images.count(id))
        return h->from(d, d.fail(PM_ERR_STATE, "pm_select_pairs_among: image %d not set", id));
  } else {
    for (auto& kv : d.images)
      if (kv.first != kTmpA && kv.first != kTmpB) ids.push_back(kv.first);
    std::sort(ids.begin(), ids.end());
  }
  const int n = static_cast<int>(ids.size());
  std::vector<int32_t> out;
  const bool all = top_k <= 0 || top_k >= n - 1;
  const int k = all ? 0 : top_k;
  if (n >= 2 && (!all || scores)) {
    const int gdim = d.dtype == PM_DESC_U8_BITS ? 32 * d.words : d.dim;
    if (gdim > 512) return h->from(d, d.fail(PM_ERR_UNSUPPORTED, "pm_select_pairs: descriptors of more than 512 elements"));
    std::vector<int2> himg(n);
    for (int a = 0; a < n; ++a) { const Image& im = d.images[ids[a]]; himg[a] = make_int2(im.row, im.n); }
    int2* dimg = nullptr; double *dg = nullptr, *dsim = nullptr; int32_t* dtop = nullptr;
    auto freeall = [&] { cudaFree(dimg); cudaFree(dg); cudaFree(dsim); cudaFree(dtop); };
    cudaError_t e = cudaMalloc(&dimg, sizeof(int2) * n);
    if (e == cudaSuccess) e = cudaMalloc(&dg, sizeof(double) * static_cast<size_t>(n) * gdim);
    if (e == cudaSuccess) e = cudaMalloc(&dsim, sizeof(double) * static_cast<size_t>(n) * n);
    if (e == cudaSuccess && k > 0) e = cudaMalloc(&dtop, sizeof(int32_t) * static_cast<size_t>(n) * k);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dimg, himg.data(), sizeof(int2) * n, cudaMemcpyHostToDevice, d.ingest);
    if (e == cudaSuccess) e = launch_retrieval(d.dtype == PM_DESC_U8_BITS ? nullptr : d.raw, d.dtype == PM_DESC_U8_BITS ? d.bits : nullptr,
                                               d.dim, d.words, dimg, n, 0, dg, dsim, nullptr, d.ingest);
    d.stats.kernel_launches += 2;
    if (e == cudaSuccess && scores)          // the clean matrix, before the selection marks the entries it takes
      e = cudaMemcpyAsync(scores, dsim, sizeof(double) * static_cast<size_t>(n) * n, cudaMemcpyDeviceToHost, d.ingest);
    std::vector<int32_t> top(static_cast<size_t>(n) * std::max(k, 1));
    if (e == cudaSuccess && k > 0) {
      e = launch_topk(dsim, n, k, dtop, d.ingest);
      ++d.stats.kernel_launches;
      if (e == cudaSuccess) e = cudaMemcpyAsync(top.data(), dtop, sizeof(int32_t) * static_cast<size_t>(n) * k, cudaMemcpyDeviceToHost, d.ingest);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.ingest);
    freeall();
    if (e != cudaSuccess) return h->from(d, d.fail_cuda(e, "pm_select_pairs", __LINE__));
    if (!all) {
      std::vector<std::pair<int32_t, int32_t>> pr;          // (j, i) with i < j: the order of the implicit all-pairs list
      for (int a = 0; a < n; ++a)
        for (int t = 0; t < k; ++t) {
          const int b = top[static_cast<size_t>(a) * k + t];
          if (b < 0 || b == a) continue;
          pr.emplace_back(std::max(a, b), std::min(a, b));
        }
      std::sort(pr.begin(), pr.end());
      pr.erase(std::unique(pr.begin(), pr.end()), pr.end());
      for (auto& p : pr) { out.push_back(ids[p.second]); out.push_back(ids[p.first]); }
    }
  }
  if (all)
    for (int b = 1; b < n; ++b)
      for (int a = 0; a < b; ++a) { out.push_back(ids[a]); out.push_back(ids[b]); }
  int32_t* buf = static_cast<int32_t*>(std::malloc(std::max<size_t>(out.size(), 1) * sizeof(int32_t)));
  if (!buf) return h->fail(PM_ERR_OOM, "pm_select_pairs: pair list");
  if (!out.empty()) std::memcpy(buf, out.data(), out.size() * sizeof(int32_t));
  *pairs_out = buf;
  *n_pairs_out = static_cast<int64_t>(out.size() / 2);
  return PM_OK;
}

int pm_free_pairs(int32_t* pairs) {
  std::free(pairs);
  return PM_OK;
}

int pm_measure_tensor_peak(pm_handle h, int kind, double* flop_per_s) {
  if (!h || !flop_per_s || kind < 0 || kind > 2) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceCtx& d = *h->devs[0];
  cudaSetDevice(d.dev);
  const int iters = 20000;        // x 4 instructions of 128 tensor-pipe cycles each: a few ms per launch
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    double flop = 0;
    cudaEventRecord(d.ev_a, d.ingest);
    const cudaError_t e = launch_tensor_peak(kind, iters, d.num_sms, &flop, d.ingest);
    cudaEventRecord(d.ev_b, d.ingest);
    if (e != cudaSuccess || cudaEventSynchronize(d.ev_b) != cudaSuccess)
      return h->from(d, d.fail(PM_ERR_CUDA, "tensor peak kernel failed: %s", cudaGetErrorString(cudaGetLastError())));
    float ms = 0;
    cudaEventElapsedTime(&ms, d.ev_a, d.ev_b);
    ++d.stats.kernel_launches;
    if (rep > 0) best = std::max(best, flop / (ms * 1e-3));
  }
  *flop_per_s = best;
  return PM_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// On-disk cache (the reference's README lists "save intermediate steps" as a todo):
// one little-endian container for (a) the ingested images of a handle and (b) a CSR match result.
//   header  : magic "PMB200\0\1", u32 version, u32 kind (1 images, 2 result), u64 count, u64 payload bytes,
//             u64 checksum of the payload (s1 ^ rotl(s2, 32), s1 = sum of u64 words, s2 = sum of (i+1)*word)
//   payload : 8-byte aligned sections (layout in reconstructor_b200/cache.py, the host-side mirror)
// ------------------------------------------------------------------------------------------------
namespace {

struct CacheHeader {
  char magic[8];
  uint32_t version, kind;
  uint64_t count, payload_bytes, checksum;
};
static_assert(sizeof(CacheHeader) == 40, "CacheHeader layout");
const char kCacheMagic[8] = {'P', 'M', 'B', '2', '0', '0', '\0', '\1'};

struct Checksum {
  uint64_t s1 = 0, s2 = 0, i = 0;
  void add(const void* p, size_t bytes) {          // bytes % 8 == 0
    const uint64_t* w = static_cast<const uint64_t*>(p);
    for (size_t k = 0; k < bytes / 8; ++k) { uint64_t v; std::memcpy(&v, w + k, 8); s1 += v; s2 += (++i) * v; }
  }
  uint64_t value() const { return s1 ^ ((s2 << 32) | (s2 >> 32)); }
};

struct CacheWriter {
  FILE* f = nullptr;
  Checksum ck;
  uint64_t bytes = 0;
  bool ok = true;
  bool open(const char* path) {
    f = std::fopen(path, "wb");
    if (!f) return false;
    CacheHeader h{};
    ok = std::fwrite(&h, sizeof h, 1, f) == 1;    // placeholder, rewritten by finish()
    return ok;
  }
  void section(const void* p, size_t n) {          // pads to 8 bytes
    if (!ok) return;
    const size_t whole = n / 8 * 8;
    if (whole) { ok = std::fwrite(p, 1, whole, f) == whole; ck.add(p, whole); }
    if (ok && n > whole) {
      unsigned char tail[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      std::memcpy(tail, static_cast<const unsigned char*>(p) + whole, n - whole);
      ok = std::fwrite(tail, 1, 8, f) == 8;
      ck.add(tail, 8);
    }
    bytes += (n + 7) / 8 * 8;
  }
  bool finish(uint32_t kind, uint64_t count) {
    if (f && ok) {
      CacheHeader h{};
      std::memcpy(h.magic, kCacheMagic, 8);
      h.version = 1; h.kind = kind; h.count = count; h.payload_bytes = bytes; h.checksum = ck.value();
      ok = std::fseek(f, 0, SEEK_SET) == 0 && std::fwrite(&h, sizeof h, 1, f) == 1;
    }
    if (f) ok = (std::fclose(f) == 0) && ok;
    f = nullptr;
    return ok;
  }
};

// Reads and verifies a whole cache file.  Returns PM_OK or an error code with g_create_error set.
int read_cache(const char* path, uint32_t kind, CacheHeader* hdr, std::vector<unsigned char>* payload) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { g_create_error = std::string("cannot open ") + path; return PM_ERR_INVALID; }
  CacheHeader h{};
  bool ok = std::fread(&h, sizeof h, 1, f) == 1 && std::memcmp(h.magic, kCacheMagic, 8) == 0 && h.version == 1 &&
            h.kind == kind && h.payload_bytes % 8 == 0;
  if (ok) {
    try { payload->resize(h.payload_bytes); } catch (...) { std::fclose(f); g_create_error = "cache file too large"; return PM_ERR_OOM; }
    ok = h.payload_bytes == 0 || std::fread(payload->data(), 1, h.payload_bytes, f) == h.payload_bytes;
    unsigned char extra;
    ok = ok && std::fread(&extra, 1, 1, f) == 0;   // nothing may follow the payload
  }
  std::fclose(f);
  if (ok) {
    Checksum ck;
    ck.add(payload->data(), payload->size());
    ok = ck.value() == h.checksum;
  }
  if (!ok) { g_create_error = std::string(path) + ": not a valid pairmatch_b200 cache file of the requested kind (magic / version / size / checksum)"; return PM_ERR_INVALID; }
  *hdr = h;
  return PM_OK;
}

struct Cursor {
  const unsigned char* p;
  size_t left;
  bool take(void* dst, size_t n) {
    const size_t padded = (n + 7) / 8 * 8;
    if (padded > left) return false;
    if (n) std::memcpy(dst, p, n);
    p += padded; left -= padded;
    return true;
  }
  const unsigned char* view(size_t n) {
    const size_t padded = (n + 7) / 8 * 8;
    if (padded > left) return nullptr;
    const unsigned char* r = p;
    p += padded; left -= padded;
    return r;
  }
};

}  // namespace

extern "C" {

int pm_save_result(const pm_csr_result* r, const char* path) {
  if (!r || !path || r->n_pairs < 0) { g_create_error = "pm_save_result: bad arguments"; return PM_ERR_INVALID; }
  CacheWriter w;
  if (!w.open(path)) { g_create_error = std::string("cannot create ") + path; return PM_ERR_INVALID; }
  const int64_t np = r->n_pairs, nm = np > 0 ? r->offsets[np] : 0;
  const int64_t head[4] = {np, nm, 0, 0};
  w.section(head, sizeof head);
  w.section(&r->device_ms, 8);
  w.section(r->pair_ij, 8 * static_cast<size_t>(np));
  const int64_t zero = 0;
  w.section(np > 0 ? r->offsets : &zero, 8 * static_cast<size_t>(np + 1));
  w.section(r->q, 4 * static_cast<size_t>(nm));
  w.section(r->t, 4 * static_cast<size_t>(nm));
  w.section(r->inlier, static_cast<size_t>(nm));
  w.section(r->F, 72 * static_cast<size_t>(np));
  w.section(r->status, 4 * static_cast<size_t>(np));
  w.section(r->n_inliers, 4 * static_cast<size_t>(np));
  w.section(r->ransac_iters, 4 * static_cast<size_t>(np));
  if (!w.finish(2, static_cast<uint64_t>(np))) { g_create_error = std::string("write error on ") + path; return PM_ERR_INVALID; }
  return PM_OK;
}

int pm_load_result(const char* path, pm_csr_result** out) {
  if (!path || !out) { g_create_error = "pm_load_result: bad arguments"; return PM_ERR_INVALID; }
  *out = nullptr;
  CacheHeader h{};
  std::vector<unsigned char> buf;
  const int rc = read_cache(path, 2, &h, &buf);
  if (rc != PM_OK) return rc;
  Cursor c{buf.data(), buf.size()};
  int64_t head[4];
  double ms = 0;
  bool ok = c.take(head, sizeof head) && c.take(&ms, 8);
  const int64_t np = ok ? head[0] : 0, nm = ok ? head[1] : 0;
  ok = ok && np >= 0 && nm >= 0 && static_cast<uint64_t>(np) == h.count;
  auto R = std::make_unique<Result>();
  if (ok) {
    try {
      R->pair_ij.resize(2 * np); R->offsets.resize(np + 1); R->vq.resize(nm); R->vt.resize(nm); R->vin.resize(nm);
      R->F.resize(9 * np); R->status.resize(np); R->n_inliers.resize(np); R->iters.resize(np);
    } catch (...) { g_create_error = "pm_load_result: out of memory"; return PM_ERR_OOM; }
    ok = c.take(R->pair_ij.data(), 8 * static_cast<size_t>(np)) && c.take(R->offsets.data(), 8 * static_cast<size_t>(np + 1)) &&
         c.take(R->vq.data(), 4 * static_cast<size_t>(nm)) && c.take(R->vt.data(), 4 * static_cast<size_t>(nm)) &&
         c.take(R->vin.data(), static_cast<size_t>(nm)) && c.take(R->F.data(), 72 * static_cast<size_t>(np)) &&
         c.take(R->status.data(), 4 * static_cast<size_t>(np)) && c.take(R->n_inliers.data(), 4 * static_cast<size_t>(np)) &&
         c.take(R->iters.data(), 4 * static_cast<size_t>(np)) && c.left == 0;
    ok = ok && R->offsets[0] == 0 && R->offsets[np] == nm;
    for (int64_t p = 0; ok && p < np; ++p) ok = R->offsets[p] <= R->offsets[p + 1];
  }
  if (!ok) { g_create_error = std::string(path) + ": malformed result sections"; return PM_ERR_INVALID; }
  R->size = nm;
  R->pub.n_pairs = np;
  R->pub.pair_ij = R->pair_ij.data(); R->pub.offsets = R->offsets.data();
  R->pub.q = R->vq.data(); R->pub.t = R->vt.data(); R->pub.inlier = R->vin.data();
  R->pub.F = R->F.data(); R->pub.status = R->status.data();
  R->pub.n_inliers = R->n_inliers.data(); R->pub.ransac_iters = R->iters.data();
  R->pub.device_ms = ms;
  R->pub.owner_ = R.get();
  *out = &R.release()->pub;
  return PM_OK;
}

int pm_save_images(pm_handle h, const char* path) {
  if (!h || !path) return PM_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceCtx& d = *h->devs[0];
  if (cudaSetDevice(d.dev) != cudaSuccess) return h->from(d, d.fail(PM_ERR_CUDA, "cudaSetDevice failed"));
  // asynchronously ingested images must be resident before the arenas are read (the blocking copies below run on the
  // legacy stream, which does not order against the non-blocking ingest stream)
  {
    const int rc = d.resolve_all();
    if (rc != PM_OK) return h->from(d, rc);
    if (cudaStreamSynchronize(d.ingest) != cudaSuccess) return h->from(d, d.fail(PM_ERR_CUDA, "ingest stream failed"));
  }
  std::vector<int> ids;
  for (auto& kv : d.images)
    if (kv.first != kTmpA && kv.first != kTmpB) ids.push_back(kv.first);
  std::sort(ids.begin(), ids.end());
  CacheWriter w;
  if (!w.open(path)) return h->fail(PM_ERR_INVALID, std::string("cannot create ") + path);
  std::vector<unsigned char> host;
  std::vector<float> frow;
  for (int id : ids) {
    const Image& im = d.images[id];
    const int32_t rec[6] = {id, im.n, d.dim, d.dtype, im.has_xy ? 1 : 0, 0};
    w.section(rec, sizeof rec);
    size_t bytes = 0;
    cudaError_t e = cudaSuccess;
    if (d.dtype == PM_DESC_U8_BITS) {
      bytes = static_cast<size_t>(im.n) * d.words * 4;
      host.resize(bytes);
      if (bytes) e = cudaMemcpy(host.data(), d.bits + static_cast<size_t>(im.row) * d.words, bytes, cudaMemcpyDeviceToHost);
    } else {
      const size_t elems = static_cast<size_t>(im.n) * d.dim;
      frow.resize(elems);
      if (elems) e = cudaMemcpy(frow.data(), d.raw + static_cast<size_t>(im.row) * d.dim, elems * 4, cudaMemcpyDeviceToHost);
      if (d.dtype == PM_DESC_U8) {                   // the arena keeps the byte values as floats
        bytes = elems;
        host.resize(bytes);
        for (size_t k = 0; k < elems; ++k) host[k] = static_cast<unsigned char>(frow[k]);
      } else {
        bytes = elems * 4;
        host.resize(bytes);
        if (bytes) std::memcpy(host.data(), frow.data(), bytes);
      }
    }
    if (e != cudaSuccess) { w.finish(1, 0); return h->from(d, d.fail_cuda(e, "pm_save_images D2H", __LINE__)); }
    w.section(host.data(), bytes);
    if (im.has_xy) {
      host.resize(static_cast<size_t>(im.n) * 8);
      if (im.n && (e = cudaMemcpy(host.data(), d.xy + 2 * static_cast<size_t>(im.row), host.size(), cudaMemcpyDeviceToHost)) != cudaSuccess) {
        w.finish(1, 0);
        return h->from(d, d.fail_cuda(e, "pm_save_images D2H", __LINE__));
      }
      w.section(host.data(), host.size());
    }
  }
  if (!w.finish(1, ids.size())) return h->fail(PM_ERR_INVALID, std::string("write error on ") + path);
  return PM_OK;
}

int pm_load_images(pm_handle h, const char* path) {
  if (!h || !path) return PM_ERR_INVALID;
  CacheHeader hdr{};
  std::vector<unsigned char> buf;
  int rc = read_cache(path, 1, &hdr, &buf);
  if (rc != PM_OK) return h->fail(rc, g_create_error);
  std::lock_guard<std::mutex> lk(h->mu);
  Cursor c{buf.data(), buf.size()};
  for (uint64_t k = 0; k < hdr.count; ++k) {
    int32_t rec[6];
    if (!c.take(rec, sizeof rec) || rec[1] < 0 || rec[2] <= 0) return h->fail(PM_ERR_INVALID, std::string(path) + ": malformed image record");
    const size_t n = static_cast<size_t>(rec[1]);
    const size_t bytes = rec[3] == PM_DESC_U8_BITS ? n * (rec[2] / 8) : (rec[3] == PM_DESC_U8 ? n * rec[2] : n * rec[2] * 4);
    const unsigned char* desc = c.view(bytes);
    const unsigned char* xy = rec[4] ? c.view(n * 8) : nullptr;
    if (!desc || (rec[4] && !xy)) return h->fail(PM_ERR_INVALID, std::string(path) + ": truncated image record");
    if (rec[0] == kTmpA || rec[0] == kTmpB) return h->fail(PM_ERR_INVALID, "reserved image id in cache file");
    for (auto& d : h->devs) {
      rc = d->set_image(rec[0], desc, rec[1], rec[2], rec[3], reinterpret_cast<const int32_t*>(xy), false);
      if (rc != PM_OK) return h->from(*d, rc);
    }
  }
  if (c.left != 0) return h->fail(PM_ERR_INVALID, std::string(path) + ": trailing bytes after the last image");
  return PM_OK;
}

}  // extern "C"

