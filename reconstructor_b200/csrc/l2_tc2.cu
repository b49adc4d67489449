// l2_tc2.cu -- K2 (CTA-pair version): exact L2 2-NN for integer-valued 128-d descriptors on tcgen05
// with cta_group::2 -- two SMs of one TPC cooperate on a 256 x 256 output tile per MMA.
//
// Why pairs: with one CTA per tile an M=128 x N=128 x K=16 MMA needs 8 KB of shared-memory operands
// every 64 cycles, i.e. the full 128 B/clk of the SM; measured, the single-CTA kernel (l2_tc.cu) tops
// out at ~77 % of the tensor pipe even with an empty epilogue.  In pair mode each SM supplies its own
// 128 query rows (A) and HALF of the train tile (B, 128 of 256 columns): 8 KB per 128 cycles.
//
// Same arithmetic and the same exactness argument as l2_tc.cu (fp16 operands, integer partial sums
// < 2^24 in the fp32 accumulator, train-row norm folded in as a 9th K=16 step).  Replaces knnMatch at
// Mapper/libMapper/FeatureMatcher.cpp:48-49; parity: cv::BFMatcher(NORM_L2), bit-exact.
//
// Per CTA: 20 warps.
//   warp 0      TMA producer: its 128 query rows (resident per work item) and a ring of its halves of
//               the train tiles; completion is signalled on the LEADER CTA's mbarriers
//   warp 1      MMA issuer (leader CTA only): 9 x tcgen05.mma.cta_group::2 M=256 N=256 K=16 per tile,
//               accumulator = 128 lanes x 256 columns in each CTA's TMEM, 2 stages
//   warp 2      TMEM allocator (cta_group::2, 512 columns)
//   warps 4-19  epilogue: warp = (lane quarter, 64-column slice); tcgen05.ld 32x32b.x32, min tree,
//               running top-2 per row; slices merged through shared memory once per work item
// Work item = (pair, 256-row query super-tile), dealt round-robin to the clusters.
#include "common.cuh"
#include "kernels.h"

// timing-probe dissection (MODE 1 only, A/B builds): skip the accumulator hand-shake / the train-tile TMA
#ifndef PM_PROBE_NOACC
#define PM_PROBE_NOACC 0
#endif
#ifndef PM_PROBE_NOTMA
#define PM_PROBE_NOTMA 0
#endif


namespace pm {

static constexpr int T2_BM = 128;               // query rows per CTA
static constexpr int T2_ROWS = 2 * T2_BM;       // per cluster work item
static constexpr int T2_ATOM = T2_BM * 128;     // 16 KB: 128 rows x 128 B (SWIZZLE_128B)
static constexpr int T2_EXT = T2_BM * 32;       // 4 KB: 128 rows x 32 B (SWIZZLE_32B)
static constexpr uint32_t T2_TMEM_COLS = 512;
static constexpr uint32_t T2_PEER_MASK = 0xFEFFFFFFu;     // clears the CTA-rank bit of a shared::cluster address

// Geometry of a variant.  BN = train rows per tile (UMMA N); every epilogue warp owns CPW chunks of
// 32 columns of its lane quarter, so there are BN / (32 * CPW) column slices and 4x that many
// epilogue warps.
//   BN = 256, CPW = 2: 16 epilogue warps (640 threads)
//   BN = 192, CPW = 1: 24 epilogue warps (896 threads) -- more warps in flight per scheduler
//   KA = number of 64-wide K atoms of a row: 2 (128-d) or 4 (256-d); a row is 64*KA + 16 fp16.
template <int BN, int CPW, int KA = 2>
struct T2Cfg {
  static constexpr int kBN = BN, kBNH = BN / 2, kCPW = CPW, kKA = KA;
  static constexpr int kATile = KA * T2_ATOM + T2_EXT;          // query tile of one CTA (20 / 36 / 68 KB)
  static constexpr int kAStages = KA == 1 ? 2 : 1;              // the short byte form prefetches the next item's tile
  static constexpr int kSlices = BN / (32 * CPW);
  static constexpr int kEpiWarps = 4 * kSlices;
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
  // the train-tile ring is staged in K GROUPS of two 64-wide atoms (the norm block rides with the last
  // group): one group per tile for 128-d rows, two for 256-d rows -- finer stages hide the TMA latency
  // that two whole-tile stages of 68 KB could not
  static constexpr int kAG = KA >= 2 ? 2 : 1;                    // 128-byte K atoms per stage
  static constexpr int kBAtom = kBNH * 128, kBExt = kBNH * 32, kBTile = kAG * kBAtom + kBExt;   // one stage
  static constexpr int kGroups = KA / kAG;
  static constexpr int kStages = KA == 4 ? 3 : (KA == 1 ? 6 : (BN == 256 ? 3 : 4));
  static constexpr int kSmemB = kStages * kBTile;
  static constexpr int kXchg = 4 * (kSlices - 1) * 32 * 32;   // two float4 per (quarter, slice, lane)
  static constexpr int kSmemBytes = kAStages * kATile + kSmemB + 1024 + 256 + kXchg;
  // kind::f16: D=f32, A=B=f16, K-major, N=BN, M=256 (pair)
  static constexpr uint32_t kShape = ((BN >> 3) << 17) | ((T2_ROWS >> 4) << 24);
  static constexpr uint32_t kIdesc = (1u << 4) | kShape;
  // kind::i8: D=s32 (c_format 2); main K-steps A=u8 (format 0), B=s8 (format 1); norm block A=B=u8
  static constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 10) | kShape;
  static constexpr uint32_t kIdescI8Ext = (2u << 4) | kShape;
  // kind::i8 with both operands signed (quantised real-valued rows, KIND 3)
  static constexpr uint32_t kIdescS8 = (2u << 4) | (1u << 7) | (1u << 10) | kShape;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Waiting on an mbarrier WITHOUT burning issue slots.  The persistent kernels below share their SMs with the small tail
// kernels of earlier batches (api.cu); a warp that polls in a tight loop is always eligible, and measured next to the
// candidate-key kernels (whose epilogue warps wait for the tensor pipe ~45 % of the time) every co-resident kernel ran
// 6-40x slower than alone.  So a failed poll puts the warp to sleep for NS nanoseconds (role-specific: the MMA issuer
// does not sleep, it is on the critical path).  A wait that lasts longer than PM_WAIT_TIMEOUT_S seconds of %globaltimer
// (a lost arrival: a bug, not load -- profilers and sanitizers stretch a wait by orders of magnitude less) traps
// instead of hanging the device.
#ifndef PM_WAIT_NS_EPI
#define PM_WAIT_NS_EPI 96
#endif
#ifndef PM_WAIT_NS_MMA
#define PM_WAIT_NS_MMA 0     // one warp per CTA pair, on the critical hand-over path: measured +1 % (ORB) .. +3 % (SuperPoint) over 24 ns
#endif
#ifndef PM_WAIT_NS_TMA
#define PM_WAIT_NS_TMA 96
#endif
#ifndef PM_WAIT_TIMEOUT_S
#define PM_WAIT_TIMEOUT_S 20
#endif
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
template <int NS>
__device__ __forceinline__ void wait_sleep(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0 = 0;
  for (uint32_t spin = 1; !mbar_try_wait(bar, parity); ++spin) {
    if (NS > 0) __nanosleep(NS);
    if ((spin & 0xFFFu) == 0) {
      const unsigned long long t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > static_cast<unsigned long long>(PM_WAIT_TIMEOUT_S) * 1000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void wait_epi(uint64_t* bar, uint32_t parity) { wait_sleep<PM_WAIT_NS_EPI>(bar, parity); }
__device__ __forceinline__ void wait_mma(uint64_t* bar, uint32_t parity) { wait_sleep<PM_WAIT_NS_MMA>(bar, parity); }
__device__ __forceinline__ void wait_tma(uint64_t* bar, uint32_t parity) { wait_sleep<PM_WAIT_NS_TMA>(bar, parity); }
// TMA load whose completion bytes land on the leader CTA's mbarrier.
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar) & T2_PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// KIND 0: kind::f16 (fp16 operands, K = 16 per instruction); KIND 1: kind::f8f6f4 with E4M3 operands (K = 32 per
// instruction, twice the rate).  Both consume 32 bytes of a K-major row per instruction and share the
// instruction-descriptor encoding used here (format field 0 = F16 resp. E4M3, D = F32).
// KIND 2: kind::i8 (u8 x s8 -> s32, K = 32 per instruction, the rate of kind::f8f6f4): integer-valued 128-d rows,
// see the I8 FORM note above t2i_fast16.
template <int KIND>
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  if (KIND >= 2)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else if (KIND == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::mxf4 (E2M1 operands packed two per byte, 64 values of K per instruction, UE8M0 scale factors per 32 values
// read from TMEM at sfa / sfb).
__device__ __forceinline__ void umma_mxf4_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t sfa, uint32_t sfb, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
      : "memory");
}
// 32 columns x this warp's 32 lanes of TMEM <- one value
__device__ __forceinline__ void tmem_st_32x32b_x32_fill(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(v)
      : "memory");
}
// Arrives (once all prior MMAs of this thread completed) on the barrier at the same offset in both CTAs.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// Arrive on the barrier at the same offset in the leader CTA (rank 0) of the pair.
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

struct Top2p {
  float m1, m2;
  int i1, i2;
};
__device__ __forceinline__ void t2_push(Top2p& s, float v, int col) {
  if (v < s.m2) {
    if (v < s.m1) { s.m2 = s.m1; s.i2 = s.i1; s.m1 = v; s.i1 = col; }
    else { s.m2 = v; s.i2 = col; }
  }
}
__device__ __noinline__ void t2_group_generic(Top2p* sp, float v0, float v1, float v2, float v3, int col) {
  Top2p s = *sp;
  t2_push(s, v0, col); t2_push(s, v1, col + 1); t2_push(s, v2, col + 2); t2_push(s, v3, col + 3);
  *sp = s;
}
// 32 consecutive columns in registers -> running top-2 (strict <, ascending columns: lowest index
// wins ties).  Minima of 8 groups of 4, one compare against the current 2nd best; on a hit a
// per-lane mask selects the groups to visit, each handled by one shared single-insert body.
__device__ __forceinline__ void t2_scan32(Top2p& s, const uint32_t* r, int col0) {
  float g[8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    g[k] = fminf(fminf(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1])),
                 fminf(__uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3])));
  const float cm = fminf(fminf(fminf(g[0], g[1]), fminf(g[2], g[3])), fminf(fminf(g[4], g[5]), fminf(g[6], g[7])));
  if (cm < s.m2) {
    uint32_t mask = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) mask |= (g[k] < s.m2 ? 1u : 0u) << k;
    do {
      const int k = __ffs(mask) - 1;
      mask &= mask - 1;
      float v0, v1, v2, v3, gk;
      switch (k) {
#define PM_CASE(K)                                                                                 \
  case K:                                                                                          \
    v0 = __uint_as_float(r[4 * K]); v1 = __uint_as_float(r[4 * K + 1]);                             \
    v2 = __uint_as_float(r[4 * K + 2]); v3 = __uint_as_float(r[4 * K + 3]); gk = g[K];              \
    break;
        PM_CASE(0) PM_CASE(1) PM_CASE(2) PM_CASE(3) PM_CASE(4) PM_CASE(5) PM_CASE(6)
        default:
          v0 = __uint_as_float(r[28]); v1 = __uint_as_float(r[29]);
          v2 = __uint_as_float(r[30]); v3 = __uint_as_float(r[31]); gk = g[7];
          break;
#undef PM_CASE
      }
      if (gk < s.m2) {
        const int col = col0 + 4 * k;
        int pos = 3;
        pos = v2 == gk ? 2 : pos;
        pos = v1 == gk ? 1 : pos;
        pos = v0 == gk ? 0 : pos;
        const float second = fminf(fmaxf(fminf(v0, v1), fminf(v2, v3)), fminf(fmaxf(v0, v1), fmaxf(v2, v3)));
        const float nm2 = fmaxf(gk, s.m1);
        if (second < nm2) {
          t2_group_generic(&s, v0, v1, v2, v3, col);
        } else {
          const bool first = gk < s.m1;
          const int c = col + pos;
          s.i2 = first ? s.i1 : c;
          s.i1 = first ? c : s.i1;
          s.m2 = nm2;
          s.m1 = fminf(gk, s.m1);
        }
      }
    } while (mask);
  }
}
// Values-only update (MODE 2): branch-free; s.i1 carries the base column of the earliest 16-column
// chunk that attained the minimum, s.m2 the second smallest chunk minimum (see l2_fixup.cu).
__device__ __forceinline__ void t2_fast16(Top2p& s, const uint32_t* r, int cbase) {
  float g[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    g[k] = fminf(fminf(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1])),
                 fminf(__uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3])));
  const float cm = fminf(fminf(g[0], g[1]), fminf(g[2], g[3]));
  const float t = fmaxf(cm, s.m1);
  s.i1 = cm < s.m1 ? cbase : s.i1;
  s.m1 = fminf(cm, s.m1);
  s.m2 = fminf(s.m2, t);
}
__device__ __forceinline__ void t2_fast(Top2p& s, const uint32_t* r, int cbase) {
  t2_fast16(s, r, cbase);
  t2_fast16(s, r + 16, cbase + 16);
}
// ---- I8 FORM (KIND 2): integer-valued 128-d rows as bytes --------------------------------------------
// query row  [ a_0 .. a_127 (u8) | 1, 255 x31 (u8) ]      train row [ 127 - b_k (s8) | r, e_1 .. e_31 (u8) ]
// with h = floor(|b|^2 / 2) = r + 255 * (e_1 + .. + e_31).  The s32 accumulator is
//     D' = a.(127 - b) + h = 127 * sum(a) - a.b + floor(|b|^2 / 2),
// so |a - b|^2 = 2 D' + (|b|^2 & 1) + (|a|^2 - 254 * sum(a)): the order of the distances is the order of
// (D', parity of |b|^2).  The values-only epilogue tracks D'; l2_fixup_i8_kernel resolves the parity exactly
// (and rescans the rare rows where it could matter).  5 K-steps per tile instead of the fp16 form's 9.
struct Top2i {
  int m1, m2, i1;
};
static constexpr int T2I_INF = 0x7f800000;      // also what the column mask writes (bit pattern of +inf)
__device__ __forceinline__ void t2i_fast16(Top2i& s, const uint32_t* r, int cbase) {
  int g[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    g[k] = min(min(static_cast<int>(r[4 * k]), static_cast<int>(r[4 * k + 1])),
               min(static_cast<int>(r[4 * k + 2]), static_cast<int>(r[4 * k + 3])));
  const int cm = min(min(g[0], g[1]), min(g[2], g[3]));
  const int t = max(cm, s.m1);
  s.i1 = cm < s.m1 ? cbase : s.i1;
  s.m1 = min(cm, s.m1);
  s.m2 = min(s.m2, t);
}
// 32-column chunks: a tree of 3-input minima (16 instructions for 32 values) and a 4-instruction update
__device__ __forceinline__ void t2i_fast32(Top2i& s, const uint32_t* r, int cbase) {
  int a[10];
#pragma unroll
  for (int k = 0; k < 10; ++k)
    a[k] = __vimin3_s32(static_cast<int>(r[3 * k]), static_cast<int>(r[3 * k + 1]), static_cast<int>(r[3 * k + 2]));
  const int b0 = __vimin3_s32(a[0], a[1], a[2]), b1 = __vimin3_s32(a[3], a[4], a[5]);
  const int b2 = __vimin3_s32(a[6], a[7], a[8]);
  const int b3 = __vimin3_s32(a[9], static_cast<int>(r[30]), static_cast<int>(r[31]));
  const int cm = min(__vimin3_s32(b0, b1, b2), b3);
  bool keep;                                         // m1 <= cm: the earlier chunk stays the winner
  const int nm1 = __vibmin_s32(s.m1, cm, &keep);
  const int t = max(cm, s.m1);
  s.i1 = keep ? s.i1 : cbase;
  s.m1 = nm1;
  s.m2 = min(s.m2, t);
}
// 32 columns starting at column c of the train image, of which the first `lim` exist
__device__ __forceinline__ void t2i_chunk32(Top2i& s, uint32_t (&v)[32], int c, int lim) {
  if (lim < 32) {
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (e >= lim) v[e] = static_cast<uint32_t>(T2I_INF);
  }
  if (TC_I8_CHUNK == 32) {
    if (lim > 0) t2i_fast32(s, v, c);
  } else {
    if (lim > 0) t2i_fast16(s, v, c);
    if (lim > 16) t2i_fast16(s, v + 16, c + 16);
  }
}
// ---- packed pairs (l2_i8x2_kernel PK): two Hamming distances per fp32 accumulator ------------------------
// v = 2^19 + 1024 H(c) + (512 + H(c + 192)) (see the kernel): bit pattern 0x49000000 | n << 4 with the high field
// n[10..18] = H(c) and the low field n[0..9] = 512 + H(c + 192).  One multiplication by 4 (fma pipe; the minima keep the
// alu pipe busy, and each pipe issues one warp instruction per two cycles) drops the top exponent bits and puts the two
// fields into the two 16-bit HALVES of the word -- high half 0x2400 | H(c), low half (512 + H(c + 192)) << 6 -- so that
// ONE tree of three-input unsigned 16x2 minima (VIMNMX3.U16x2) reduces both at once: 32 fma + 16 alu instructions for
// the 64 distances of 32 columns (the one-row form: 16 alu instructions for 32 distances).
#ifndef PM_PK_PROBE
#define PM_PK_PROBE 0        // development: 1 = no field shift, 2 = XOR tree instead of the minima (timing only, wrong results)
#endif
static constexpr uint32_t T2P_SF_BIG = 0x89898989u;             // UE8M0 2^10, four scale factors per TMEM column
// scale_vec::2X reads the scales of a row's two K blocks from bytes 0 | 1 of its scale column (sf_id 0): 2^10 | 1.  (Bytes 1 | 0
// and scales split over even | odd columns were tried on the device: wrong results.)
static constexpr uint32_t T2P_SF_MIX = 0x7F897F89u;
static constexpr float T2P_BIAS = 512.f;                        // what the bias slots of the norm block add (pack.cu)
__device__ __forceinline__ uint32_t t2_umin32x2(const uint32_t (&r)[32]) {
  uint32_t a[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) a[k] = __vimin3_u16x2(r[3 * k], r[3 * k + 1], r[3 * k + 2]);
  const uint32_t b0 = __vimin3_u16x2(a[0], a[1], a[2]), b1 = __vimin3_u16x2(a[3], a[4], a[5]);
  const uint32_t b2 = __vimin3_u16x2(a[6], a[7], a[8]), b3 = __vimin3_u16x2(a[9], r[30], r[31]);
  return __vminu2(__vimin3_u16x2(b0, b1, b2), b3);
}
// v: 32 columns, of which the first lim_hi / lim_lo carry a row behind the high / low field; m_hi / m_lo: the smallest H
// of those rows (garbage when lim <= 0: not used)
template <bool FULL>
__device__ __forceinline__ void t2p_chunk32(const uint32_t* v, int lim_hi, int lim_lo, int& m_hi, int& m_lo) {
  uint32_t four;
  asm volatile("mov.u32 %0, 4;" : "=r"(four));
  uint32_t w[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) w[e] = PM_PK_PROBE == 1 ? v[e] : v[e] * four;
  if (!FULL && lim_lo < 32) {                          // ragged / absent second sub-tile
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (e >= lim_lo) w[e] |= 0x0000FFFFu;
  }
  if (!FULL && lim_hi < 32) {
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (e >= lim_hi) w[e] = 0xFFFFFFFFu;
  }
  uint32_t m = t2_umin32x2(w);
  if (PM_PK_PROBE == 2) {                                // timing probe: a tree of three-input XORs instead of the minima
    uint32_t a[11];
#pragma unroll
    for (int k = 0; k < 10; ++k) a[k] = w[3 * k] ^ w[3 * k + 1] ^ w[3 * k + 2];
    a[10] = w[30] ^ w[31];
    m = (a[0] ^ a[1] ^ a[2]) ^ (a[3] ^ a[4] ^ a[5]) ^ (a[6] ^ a[7] ^ a[8]) ^ (a[9] ^ a[10]);
  }
  m_hi = static_cast<int>((m >> 16) & 0x1FFu);
  m_lo = static_cast<int>((m & 0xFFFFu) >> 6) - 512;
}
__device__ __forceinline__ void t2i_update(Top2i& s, int cm, int cbase) {
  bool keep;                                         // m1 <= cm: the earlier chunk stays the winner
  const int nm1 = __vibmin_s32(s.m1, cm, &keep);
  const int t = max(cm, s.m1);
  s.i1 = keep ? s.i1 : cbase;
  s.m1 = nm1;
  s.m2 = min(s.m2, t);
}
// ---- MODE 3: six smallest chunk minima as keys ---------------------------------------------------
// key = positive fp32 score with the low 10 mantissa bits replaced by the 16-column chunk id (the
// scores are approximate anyway; l2f_fixup.cu widens its error bound by the 2^-13 relative truncation).
// For positive floats, float order == integer order, so a min/max chain sorts (score, id) at once.
struct Keys4 {                       // the six smallest keys, ascending
  float k1, k2, k3, k4, k5, k6;
};
__device__ __forceinline__ void keys_insert(Keys4& s, float key) {
  float lo = fminf(s.k1, key), hi = fmaxf(s.k1, key);
  s.k1 = lo; key = hi;
  lo = fminf(s.k2, key); hi = fmaxf(s.k2, key);
  s.k2 = lo; key = hi;
  lo = fminf(s.k3, key); hi = fmaxf(s.k3, key);
  s.k3 = lo; key = hi;
  lo = fminf(s.k4, key); hi = fmaxf(s.k4, key);
  s.k4 = lo; key = hi;
  lo = fminf(s.k5, key); hi = fmaxf(s.k5, key);
  s.k5 = lo; key = hi;
  s.k6 = fminf(s.k6, key);
}
__device__ __forceinline__ void keys_chunk16(Keys4& s, const uint32_t* r, int col) {
  float g[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    g[k] = fminf(fminf(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1])),
                 fminf(__uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3])));
  const float cm = fminf(fminf(g[0], g[1]), fminf(g[2], g[3]));
  keys_insert(s, __uint_as_float((__float_as_uint(cm) & 0xFFFFFC00u) | static_cast<uint32_t>(col >> 4)));
}
// ---- MODE 3 on kind::i8 (KIND 3): quantised real-valued rows -------------------------------------------
// rows x with |x_k| <= 0.5 are quantised to q = rint(254 x) (s8); query form [ q_a | 1, 255 x31 (u8) ], train form
// [ -q_b | digits of h = rint(254^2 (|b|^2 / 2 + 1)) in base 255 ].  The s32 accumulator h - q_a.q_b approximates
// 254^2 * (|b|^2 / 2 - a.b + 1) = 254^2 / 2 * score, score = |b|^2 - 2 a.b + 2 as in the fp16 form, with
// |error| <= sqrt(D) (|a| + |b|) / 254 + D / (2 * 254^2) + 1 / 254^2 on the score (l2f_fixup.cu widens its bound).
// Keys are integers here: (accumulator << 10) | chunk id, 0 < accumulator < 2^18.
static constexpr float T2S_SCALE = 254.f;
struct Keys6i {
  int k1, k2, k3, k4, k5, k6;
};
__device__ __forceinline__ void keysi_insert(Keys6i& s, int key) {
  int lo = min(s.k1, key), hi = max(s.k1, key);
  s.k1 = lo; key = hi;
  lo = min(s.k2, key); hi = max(s.k2, key);
  s.k2 = lo; key = hi;
  lo = min(s.k3, key); hi = max(s.k3, key);
  s.k3 = lo; key = hi;
  lo = min(s.k4, key); hi = max(s.k4, key);
  s.k4 = lo; key = hi;
  lo = min(s.k5, key); hi = max(s.k5, key);
  s.k5 = lo; key = hi;
  s.k6 = min(s.k6, key);
}
__device__ __forceinline__ void keysi_chunk16(Keys6i& s, const uint32_t* r, int col) {
  const int a0 = __vimin3_s32(static_cast<int>(r[0]), static_cast<int>(r[1]), static_cast<int>(r[2]));
  const int a1 = __vimin3_s32(static_cast<int>(r[3]), static_cast<int>(r[4]), static_cast<int>(r[5]));
  const int a2 = __vimin3_s32(static_cast<int>(r[6]), static_cast<int>(r[7]), static_cast<int>(r[8]));
  const int a3 = __vimin3_s32(static_cast<int>(r[9]), static_cast<int>(r[10]), static_cast<int>(r[11]));
  const int a4 = __vimin3_s32(static_cast<int>(r[12]), static_cast<int>(r[13]), static_cast<int>(r[14]));
  const int cm = min(__vimin3_s32(a0, a1, a2), __vimin3_s32(a3, a4, static_cast<int>(r[15])));
  keysi_insert(s, (cm << 10) | (col >> 4));
}
// 32 columns starting at column c of the train image, of which the first `lim` exist (two 16-column chunks)
__device__ __forceinline__ void keysi_chunk32(Keys6i& s, uint32_t (&v)[32], int c, int lim) {
  if (lim < 32) {
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (e >= lim) v[e] = static_cast<uint32_t>(T2I_INF);
  }
  if (lim > 0) keysi_chunk16(s, v, c);
  if (lim > 16) keysi_chunk16(s, v + 16, c + 16);
}
// ---- MODE 4 (quantised real-valued rows, the batched loop's default): exact argmin + second-smallest score -------
// MODE 3 hands the re-rank a 16-column CHUNK per candidate row, and the re-rank has to evaluate all of it in fp32: a
// latency-bound kernel that does not fit next to this persistent kernel at a useful occupancy (measured: the step cost
// tensor kernel + tails).  Here the epilogue does the bookkeeping instead:
//   * keys are (acc << 13) | column; the three smallest CHUNK keys k1 < k2 < k3 are kept branch-free (VIMNMX3 tree over
//     the 16 accumulators + 5 min/max), with the chunk's base column in the low bits -- cheaper than MODE 3's six keys;
//   * only when a chunk becomes the new row minimum AND its score is match-like (key < thresh: d^2 below ~0.9 for a
//     unit-norm query -- a few times per row, a few percent of the chunks per warp) the chunk is REFINED on the spot:
//     the column of its minimum goes into the key and the second smallest accumulator of the chunk into w2.  At the end
//     min(w2, k2 >> 13) is the second smallest approximate score of the whole row, so every column but the argmin is
//     bounded from below without looking at it again.  w2 == T2K_NONE marks a row whose minimum was never refined.
// l2f_rerank1_kernel (l2f_fixup.cu) then evaluates ONE column per candidate row in fp32 and certifies the nearest
// index and the outcome of Lowe's test; rows it cannot close (not refined but still a candidate, or a test that hinges
// on the uncertainty of the second score; rare) go to l2f_fixup_kernel with the three chunk keys.  The threshold only
// steers work between the two kernels, never a result.
// Needs acc < 2^18 - 1 (holds: h <= 1.505 * 254^2, -q_a.q_b <= 1.01 * 254^2) and nt <= 8192 (13 column bits).
struct Keys3w {
  int k1, k2, k3, w2;             // keys = (acc << 13) | column; w2 = second smallest acc inside k1's chunk
};
static constexpr int T2K_COLBITS = 13;
static constexpr int T2K_NONE = 0x7fffffff;
static constexpr int T2K_ACC_NONE = (1 << 18) - 1;           // masks absent columns; == T2K_NONE >> T2K_COLBITS
#ifndef PM_KEYS3_THRESH_SCORE
#define PM_KEYS3_THRESH_SCORE 1.9f                            // score = d^2 - |a|^2 + 2
#endif
static constexpr int T2K_THRESH_KEY = static_cast<int>(PM_KEYS3_THRESH_SCORE * (T2S_SCALE * T2S_SCALE / 2.f)) << T2K_COLBITS;
__device__ __forceinline__ void keys3_insert(Keys3w& s, int key) {
  int lo = min(s.k1, key), hi = max(s.k1, key);
  s.k1 = lo; key = hi;
  lo = min(s.k2, key); hi = max(s.k2, key);
  s.k2 = lo;
  s.k3 = min(s.k3, hi);
}
// column of the minimum + second smallest value of a chunk (the rare path): values packed with their position,
// p_j = v_j * 16 + j (distinct), tree -> minimum with its column; u_j = p_j - pm - 1 wraps to 0xFFFFFFFF for the minimum
// itself and keeps the order of the others, so an unsigned tree yields the second smallest.
__device__ __forceinline__ int2 keys3_refine(const int (&v)[16]) {
  int p[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) p[j] = v[j] * 16 + j;
  const int a0 = __vimin3_s32(p[0], p[1], p[2]), a1 = __vimin3_s32(p[3], p[4], p[5]);
  const int a2 = __vimin3_s32(p[6], p[7], p[8]), a3 = __vimin3_s32(p[9], p[10], p[11]);
  const int a4 = __vimin3_s32(p[12], p[13], p[14]);
  const int pm = min(__vimin3_s32(a0, a1, a2), __vimin3_s32(a3, a4, p[15]));
  const unsigned int d = static_cast<unsigned int>(-pm - 1);
  unsigned int u[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) u[j] = static_cast<unsigned int>(p[j]) + d;
  const unsigned int b0 = __vimin3_u32(u[0], u[1], u[2]), b1 = __vimin3_u32(u[3], u[4], u[5]);
  const unsigned int b2 = __vimin3_u32(u[6], u[7], u[8]), b3 = __vimin3_u32(u[9], u[10], u[11]);
  const unsigned int b4 = __vimin3_u32(u[12], u[13], u[14]);
  const unsigned int um = min(__vimin3_u32(b0, b1, b2), __vimin3_u32(b3, b4, u[15]));
  return make_int2(pm & 15, static_cast<int>(um - d) >> 4);        // (column inside the chunk, second smallest acc;
}                                                                   //  arithmetic shift: accumulators may be negative)
// 16 columns starting at train column c (multiple of 16); RAGGED: only the first `lim` (>= 1) exist
// OFF: constant added to every accumulator of the tile (the norm term of the unit-norm form, which has no norm K-step:
// see l2_i8x2_kernel NX); it enters the chunk key and the refined second score, not the 16 values
template <bool RAGGED, int OFF = 0>
__device__ __forceinline__ void keys3_chunk16(Keys3w& s, const uint32_t* r, int c, int lim) {
  int v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = (RAGGED && j >= lim) ? T2K_ACC_NONE - OFF : static_cast<int>(r[j]);
  const int a0 = __vimin3_s32(v[0], v[1], v[2]), a1 = __vimin3_s32(v[3], v[4], v[5]);
  const int a2 = __vimin3_s32(v[6], v[7], v[8]), a3 = __vimin3_s32(v[9], v[10], v[11]);
  const int a4 = __vimin3_s32(v[12], v[13], v[14]);
  const int cm = min(__vimin3_s32(a0, a1, a2), __vimin3_s32(a3, a4, v[15]));
  int key = cm * (1 << T2K_COLBITS) + (c + OFF * (1 << T2K_COLBITS));
  if (key < min(s.k1, T2K_THRESH_KEY)) {             // new, match-like row minimum: refine (rare)
    const int2 js = keys3_refine(v);
    key += js.x;
    s.w2 = js.y + OFF;
  }
  keys3_insert(s, key);
}
template <int OFF = 0>
__device__ __forceinline__ void keys3_chunk32(Keys3w& s, const uint32_t (&v)[32], int c, int lim) {
  if (lim >= 32) {                                   // warp-uniform: the whole 32-column piece exists
    keys3_chunk16<false, OFF>(s, v, c, 16);
    keys3_chunk16<false, OFF>(s, v + 16, c + 16, 16);
  } else {
    if (lim > 0) keys3_chunk16<true, OFF>(s, v, c, lim);
    if (lim > 16) keys3_chunk16<true, OFF>(s, v + 16, c + 16, lim - 16);
  }
}
// integer key -> the float key l2f_fixup.cu expects: score with the chunk id in the low 10 mantissa bits
__device__ __forceinline__ float keysi_to_float(int key) {
  if (key == T2I_INF) return __int_as_float(0x7f800000);
  const float sc = static_cast<float>(key >> 10) * (2.f / (T2S_SCALE * T2S_SCALE));
  return __uint_as_float((__float_as_uint(sc) & 0xFFFFFC00u) | static_cast<uint32_t>(key & 0x3FF));
}
// ordered by (value, index)
__device__ __forceinline__ bool t2_less(float va, int ia, float vb, int ib) { return va < vb || (va == vb && ia < ib); }
__device__ __forceinline__ void t2_merge(Top2p& s, float om1, int oi1, float om2, int oi2) {
  Top2p t;
  if (oi1 >= 0 && (s.i1 < 0 || t2_less(om1, oi1, s.m1, s.i1))) {
    t.m1 = om1; t.i1 = oi1;
    if (oi2 >= 0 && (s.i1 < 0 || t2_less(om2, oi2, s.m1, s.i1))) { t.m2 = om2; t.i2 = oi2; }
    else { t.m2 = s.m1; t.i2 = s.i1; }
  } else {
    t.m1 = s.m1; t.i1 = s.i1;
    if (oi1 >= 0 && (s.i2 < 0 || t2_less(om1, oi1, s.m2, s.i2))) { t.m2 = om1; t.i2 = oi1; }
    else { t.m2 = s.m2; t.i2 = s.i2; }
  }
  s = t;
}

// MODE 0: exact top-2 with indices, 1: timing probe, 2: values-only (l2_fixup follows), 3: real-valued rows.
// LEAN variants are compiled for <= 64 registers per thread (launch bound of 1024 threads; no spills):
// 20 warps x 64 = 40,960 registers and <= 190 KB of shared memory per SM leave room for the small tail
// kernels of the previous batch (re-rank / fix-up, select, RANSAC) to be co-resident with this persistent
// kernel, which runs on a higher-priority stream.
template <class Cfg, int MODE, bool LEAN = (MODE == 3), int KIND = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LEAN ? 1024 : Cfg::kThreads, 1)
l2_top2_tc2_kernel(const __grid_constant__ CUtensorMap q_main, const __grid_constant__ CUtensorMap q_ext,
                   const __grid_constant__ CUtensorMap t_main, const __grid_constant__ CUtensorMap t_ext,
                   const int32_t* __restrict__ qnorm, const PairJob* __restrict__ jobs, int n_jobs,
                   int tiles_per_job, int2* __restrict__ knn_idx, float2* __restrict__ knn_dist, int stride,
                   float2* __restrict__ extra_keys) {
  // MODE 3 (real-valued rows): the accumulator is an APPROXIMATE score |b|^2 - 2 a.b + 2 (fp16 operands);
  // the epilogue keeps the 6 smallest 16-column chunk minima as keys (score with the chunk id in the low
  // 10 mantissa bits) and l2f_fixup.cu re-ranks exactly in fp32.  qnorm is unused in that mode.
  constexpr int T2_BN = Cfg::kBN, T2_BNH = Cfg::kBNH, T2_STAGES = Cfg::kStages, T2_SMEM_B = Cfg::kSmemB;
  constexpr int KA = Cfg::kKA, T2_TILE = Cfg::kATile, T2_SMEM_A = Cfg::kAStages * Cfg::kATile, AST = Cfg::kAStages;
  constexpr int KEL = KIND >= 1 ? 128 : 64;        // tensor-map elements per 128-byte K atom (bytes / fp16)
  constexpr int AG = Cfg::kAG;
  static_assert(KIND != 2 || MODE == 2 || MODE == 1, "the i8 form has a values-only epilogue");
    constexpr int KDIM = KEL * KA;
  constexpr int T2_BATOM = Cfg::kBAtom, T2_BTILE = Cfg::kBTile, CPW = Cfg::kCPW, NSL = Cfg::kSlices;
  constexpr int NGRP = Cfg::kGroups;
  constexpr uint32_t T2_IDESC = KIND == 2 ? Cfg::kIdescI8 : KIND == 3 ? Cfg::kIdescS8 : Cfg::kIdesc;
  constexpr uint32_t T2_IDESC_EXT = KIND >= 2 ? Cfg::kIdescI8Ext : Cfg::kIdesc;
  static_assert(KIND != 3 || MODE == 3, "quantised real-valued rows use the candidate-key epilogue");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + T2_SMEM_A;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T2_SMEM_A + T2_SMEM_B);
  uint64_t* a_full = bars + 0;                       // [2] leader: 1 arrival + 2 x tile of tx
  uint64_t* a_empty = bars + 2;                      // [2] both: multicast commit
  uint64_t* b_full = bars + 4;                       // [STAGES] leader
  uint64_t* b_empty = bars + 4 + T2_STAGES;          // [STAGES] both
  uint64_t* acc_full = bars + 4 + 2 * T2_STAGES;     // [2] both: multicast commit
  uint64_t* acc_empty = acc_full + 2;                // [2] leader: 32 epilogue warps of the pair
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float4* xchg = reinterpret_cast<float4*>(smem + T2_SMEM_A + T2_SMEM_B + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&q_main); tma_prefetch_desc(&q_ext);
    tma_prefetch_desc(&t_main); tma_prefetch_desc(&t_ext);
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < T2_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 2 * Cfg::kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(T2_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // barriers of both CTAs initialised and visible
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_items = n_jobs * tiles_per_job;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ==========================
    if (lane == 0) {
      uint32_t ai = 0, bi = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int jb = item / tiles_per_job, r = item - jb * tiles_per_job;
        const PairJob job = jobs[jb];
        if (r * T2_ROWS >= job.nq) continue;
        {
          const uint32_t as = ai % AST, use = ai / AST;
          wait_tma(&a_empty[as], (use & 1) ^ 1);
          if (leader) mbar_expect_tx(&a_full[as], 2 * T2_TILE);
          uint8_t* dstA = sA + as * T2_TILE;
          const int row = job.q_row + r * T2_ROWS + rank * T2_BM;
#pragma unroll
          for (int a = 0; a < KA; ++a) tma_load_2d_pair(dstA + a * T2_ATOM, &q_main, KEL * a, row, &a_full[as]);
          tma_load_2d_pair(dstA + KA * T2_ATOM, &q_ext, KDIM, row, &a_full[as]);
        }
        ++ai;
        const int n_tiles = (MODE == 1 && PM_PROBE_NOTMA) ? 0 : (job.nt + T2_BN - 1) / T2_BN;
        for (int n = 0; n < n_tiles; ++n) {
          const int row = job.t_row + n * T2_BN + rank * T2_BNH;      // this CTA's half of the train tile
#pragma unroll
          for (int g = 0; g < NGRP; ++g, ++bi) {
            const uint32_t st = bi % T2_STAGES;
            wait_tma(&b_empty[st], ((bi / T2_STAGES) & 1) ^ 1);
            const bool last = g == NGRP - 1;
            if (leader) mbar_expect_tx(&b_full[st], 2 * (AG * T2_BATOM + (last ? Cfg::kBExt : 0)));
            uint8_t* dst = sB + st * T2_BTILE;
#pragma unroll
            for (int a = 0; a < AG; ++a)
              tma_load_2d_pair(dst + a * T2_BATOM, &t_main, (AG * g + a) * KEL, row, &b_full[st]);
            if (last) tma_load_2d_pair(dst + AG * T2_BATOM, &t_ext, KDIM, row, &b_full[st]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer (leader CTA) ===========================
    if (leader) {
      uint32_t ai = 0, bi = 0, ti = 0;
      constexpr uint32_t HI128 = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t HI32 = (256u >> 4) | (1u << 14) | (6u << 29);
      const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int jb = item / tiles_per_job, r = item - jb * tiles_per_job;
        const int job_nq = jobs[jb].nq, job_nt = jobs[jb].nt;
        if (r * T2_ROWS >= job_nq) continue;
        const uint32_t a_st = ai % AST;
        wait_mma(&a_full[a_st], (ai / AST) & 1);
        const uint32_t a_lo = a_lo0 + a_st * (T2_TILE >> 4);
        ++ai;
        const int n_tiles = (job_nt + T2_BN - 1) / T2_BN;
        for (int n = 0; n < n_tiles; ++n, ++ti) {
          const uint32_t as = ti & 1, use = ti >> 1;
          if (!(MODE == 1 && PM_PROBE_NOACC)) wait_mma(&acc_empty[as], (use & 1) ^ 1);
          const uint32_t d_tmem = tmem_base + as * T2_BN;
#pragma unroll
          for (int g = 0; g < NGRP; ++g, ++bi) {
            const uint32_t st = bi % T2_STAGES;
            if (!(MODE == 1 && PM_PROBE_NOTMA)) wait_mma(&b_full[st], (bi / T2_STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t b_lo = b_lo0 + st * (T2_BTILE >> 4);
#pragma unroll
              for (int k = 0; k < 4 * AG; ++k) {
                const uint32_t aoff = ((AG * g + (k >> 2)) * T2_ATOM + (k & 3) * 32) >> 4;
                const uint32_t boff = ((k >> 2) * T2_BATOM + (k & 3) * 32) >> 4;
                umma_f16_pair<KIND>(d_tmem, (static_cast<uint64_t>(HI128) << 32) | (a_lo + aoff),
                              (static_cast<uint64_t>(HI128) << 32) | (b_lo + boff), T2_IDESC, (g > 0 || k > 0) ? 1u : 0u);
              }
              if (g == NGRP - 1) {
                umma_f16_pair<KIND>(d_tmem, (static_cast<uint64_t>(HI32) << 32) | (a_lo + ((KA * T2_ATOM) >> 4)),
                              (static_cast<uint64_t>(HI32) << 32) | (b_lo + ((AG * T2_BATOM) >> 4)), T2_IDESC_EXT, 1u);
                umma_commit_pair(&acc_full[as]);
              }
              if (!(MODE == 1 && PM_PROBE_NOTMA)) umma_commit_pair(&b_empty[st]);
            }
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit_pair(&a_empty[a_st]);
        __syncwarp();
      }
      if (MODE == 1 && PM_PROBE_NOACC && ai > 0)     // drain before the dealloc
        wait_mma(&a_empty[(ai - 1) % AST], ((ai - 1) / AST) & 1);
    }
  } else if (warp >= 4) {
    // ======================================= epilogue (both CTAs) =============================
    const int quarter = warp & 3, slice = (warp - 4) >> 2;       // which 32*CPW-column slice of the tile
    uint32_t ti = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int jb = item / tiles_per_job, r = item - jb * tiles_per_job;
      const PairJob job = jobs[jb];
      if (r * T2_ROWS >= job.nq) continue;
      const int row = r * T2_ROWS + rank * T2_BM + quarter * 32 + lane;
      Top2p s;
      s.m1 = s.m2 = __int_as_float(0x7f800000);
      s.i1 = s.i2 = -1;                                // MODE 2: i1 = base column of the winning 16-column chunk
      Keys4 ks;
      ks.k1 = ks.k2 = ks.k3 = ks.k4 = ks.k5 = ks.k6 = __int_as_float(0x7f800000);
      Keys6i ki;                                       // KIND 3
      ki.k1 = ki.k2 = ki.k3 = ki.k4 = ki.k5 = ki.k6 = T2I_INF;
      Top2i si;                                        // KIND 2
      si.m1 = si.m2 = T2I_INF;
      si.i1 = -1;
      const int n_tiles = (job.nt + T2_BN - 1) / T2_BN;
      for (int n = 0; n < ((MODE == 1 && PM_PROBE_NOACC) ? 0 : n_tiles); ++n, ++ti) {
        const uint32_t as = ti & 1, use = ti >> 1;
        wait_epi(&acc_full[as], use & 1);
        tc_fence_after();
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * T2_BN + slice * (32 * CPW);
        const int c0 = n * T2_BN + slice * (32 * CPW);
        const int lim = job.nt - c0;                   // columns of this slice that exist
        uint32_t v[32];
        if (MODE == 1) {
          tc_fence_before();
          if (lane == 0) mbar_arrive_leader(&acc_empty[as]);
        } else if (lim >= 32 * CPW) {
#pragma unroll
          for (int c = 0; c < CPW; ++c) {
            tmem_ld_32x32b_x32(t0 + 32 * c, v);
            if (c == CPW - 1) {                         // accumulator slice drained
              tc_fence_before();
              if (lane == 0) mbar_arrive_leader(&acc_empty[as]);
            }
            if (MODE == 3 && KIND == 3) { keysi_chunk16(ki, v, c0 + 32 * c); keysi_chunk16(ki, v + 16, c0 + 32 * c + 16); }
            else if (MODE == 3) { keys_chunk16(ks, v, c0 + 32 * c); keys_chunk16(ks, v + 16, c0 + 32 * c + 16); }
            else if (MODE == 2 && KIND == 2) t2i_chunk32(si, v, c0 + 32 * c, 32);
            else if (MODE == 2) t2_fast(s, v, c0 + 32 * c);
            else t2_scan32(s, v, c0 + 32 * c);
          }
        } else {
#pragma unroll
          for (int c = 0; c < CPW; ++c) {
            tmem_ld_32x32b_x32(t0 + 32 * c, v);
            if (c == CPW - 1) {
              tc_fence_before();
              if (lane == 0) mbar_arrive_leader(&acc_empty[as]);
            }
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (32 * c + e >= lim) v[e] = 0x7f800000u;
            if (MODE == 3 && KIND == 3) {
              if (lim > 32 * c) keysi_chunk16(ki, v, c0 + 32 * c);
              if (lim > 32 * c + 16) keysi_chunk16(ki, v + 16, c0 + 32 * c + 16);
            } else if (MODE == 3) {
              if (lim > 32 * c) keys_chunk16(ks, v, c0 + 32 * c);
              if (lim > 32 * c + 16) keys_chunk16(ks, v + 16, c0 + 32 * c + 16);
            } else if (MODE == 2 && KIND == 2) {
              t2i_chunk32(si, v, c0 + 32 * c, lim - 32 * c);
            } else if (MODE == 2) {
              if (lim > 32 * c) t2_fast16(s, v, c0 + 32 * c);
              if (lim > 32 * c + 16) t2_fast16(s, v + 16, c0 + 32 * c + 16);
            }
            else t2_scan32(s, v, c0 + 32 * c);
          }
        }
      }
      // merge the column slices of this lane quarter: slices 1.. publish, slice 0 merges
      {
        float4* slot = xchg + ((quarter * (NSL - 1)) * 64 + lane);
        const int bar_id = 1 + quarter;
        if (slice > 0) {
          slot[(slice - 1) * 64] =
              (MODE == 3 && KIND == 3) ? make_float4(__int_as_float(ki.k1), __int_as_float(ki.k2), __int_as_float(ki.k3), __int_as_float(ki.k4))
              : MODE == 3 ? make_float4(ks.k1, ks.k2, ks.k3, ks.k4)
              : KIND == 2 ? make_float4(__int_as_float(si.m1), __int_as_float(si.i1), __int_as_float(si.m2), 0.f)
                          : make_float4(s.m1, __int_as_float(s.i1), s.m2, __int_as_float(s.i2));
          if (MODE == 3 && KIND == 3) slot[(slice - 1) * 64 + 32] = make_float4(__int_as_float(ki.k5), __int_as_float(ki.k6), 0.f, 0.f);
          else if (MODE == 3) slot[(slice - 1) * 64 + 32] = make_float4(ks.k5, ks.k6, 0.f, 0.f);
        }
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NSL) : "memory");
        if (slice == 0) {
#pragma unroll
          for (int o = 0; o < NSL - 1; ++o) {
            const float4 x = slot[o * 64];
            if (MODE == 3 && KIND == 3) {
              const float4 x2 = slot[o * 64 + 32];
              keysi_insert(ki, __float_as_int(x.x)); keysi_insert(ki, __float_as_int(x.y));
              keysi_insert(ki, __float_as_int(x.z)); keysi_insert(ki, __float_as_int(x.w));
              keysi_insert(ki, __float_as_int(x2.x)); keysi_insert(ki, __float_as_int(x2.y));
            } else if (MODE == 3) {
              const float4 x2 = slot[o * 64 + 32];
              keys_insert(ks, x.x); keys_insert(ks, x.y); keys_insert(ks, x.z); keys_insert(ks, x.w);
              keys_insert(ks, x2.x); keys_insert(ks, x2.y);
            } else if (MODE == 2 && KIND == 2) {
              const int ob = __float_as_int(x.y), om1 = __float_as_int(x.x), om2 = __float_as_int(x.z);
              const bool take = ob >= 0 && (si.i1 < 0 || om1 < si.m1 || (om1 == si.m1 && ob < si.i1));
              const int hi = max(si.m1, om1);
              si.m2 = min(min(si.m2, om2), hi);
              si.m1 = min(si.m1, om1);
              si.i1 = take ? ob : si.i1;
            } else if (MODE == 2) {
              const int ob = __float_as_int(x.y);
              const bool take = ob >= 0 && (s.i1 < 0 || x.x < s.m1 || (x.x == s.m1 && ob < s.i1));
              const float hi = fmaxf(s.m1, x.x);
              s.m2 = fminf(fminf(s.m2, x.z), hi);
              s.m1 = fminf(s.m1, x.x);
              s.i1 = take ? ob : s.i1;
            } else {
              t2_merge(s, x.x, __float_as_int(x.y), x.z, __float_as_int(x.w));
            }
          }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NSL) : "memory");
      }
      if (MODE == 3 && KIND == 3 && slice == 0 && row < job.nq) {
        const size_t o = static_cast<size_t>(jb) * stride + row;
        knn_dist[o] = make_float2(keysi_to_float(ki.k1), keysi_to_float(ki.k2));
        knn_idx[o] = make_int2(__float_as_int(keysi_to_float(ki.k3)), __float_as_int(keysi_to_float(ki.k4)));
        extra_keys[o] = make_float2(keysi_to_float(ki.k5), keysi_to_float(ki.k6));
      } else if (MODE == 3 && slice == 0 && row < job.nq) {
        const size_t o = static_cast<size_t>(jb) * stride + row;
        knn_dist[o] = make_float2(ks.k1, ks.k2);
        knn_idx[o] = make_int2(__float_as_int(ks.k3), __float_as_int(ks.k4));
        extra_keys[o] = make_float2(ks.k5, ks.k6);
      } else if (MODE == 2 && KIND == 2 && slice == 0 && row < job.nq) {
        const size_t o = static_cast<size_t>(jb) * stride + row;
        knn_idx[o] = make_int2(si.i1, -3);             // -3: D' values of the i8 form (l2_fixup_i8_kernel follows)
        knn_dist[o] = make_float2(__int_as_float(si.m1), __int_as_float(si.m2));
      } else if (MODE == 2 && slice == 0 && row < job.nq) {
        const float na = static_cast<float>(qnorm[job.q_row + row]);
        const size_t o = static_cast<size_t>(jb) * stride + row;
        knn_idx[o] = make_int2(s.i1, -2);
        knn_dist[o] = make_float2(__fadd_rn(s.m1, na), __fadd_rn(s.m2, na));
      } else if (KIND != 2 && slice == 0 && row < job.nq) {
        const float na = static_cast<float>(qnorm[job.q_row + row]);
        int2 oi;
        float2 od;
        oi.x = s.i1; oi.y = s.i2;
        od.x = s.i1 < 0 ? s.m1 : __fsqrt_rn(fmaxf(__fadd_rn(s.m1, na), 0.f));
        od.y = s.i2 < 0 ? s.m2 : __fsqrt_rn(fmaxf(__fadd_rn(s.m2, na), 0.f));
        const size_t o = static_cast<size_t>(jb) * stride + row;
        knn_idx[o] = oi;
        knn_dist[o] = od;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // no remote arrive / multicast may still be in flight
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T2_TMEM_COLS) : "memory");
  }
}

using T2Wide = T2Cfg<256, 2>;     // 16 epilogue warps
using T2Deep = T2Cfg<192, 1>;     // 24 epilogue warps
using T2F128 = T2Cfg<256, 2, 2>;  // real-valued rows, 128-d
using T2F256 = T2Cfg<256, 2, 4>;  // real-valued rows, 256-d (SuperPoint)
using T2I8 = T2Cfg<256, 2, 1>;    // integer-valued 128-d rows as bytes (kind::i8)

cudaError_t tc2_configure() {
  cudaError_t e;
#define PM_T2_ATTR(CFG, M)                                                                                       \
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<CFG, M>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                CFG::kSmemBytes)) != cudaSuccess) return e;                                        \
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<CFG, M>, cudaFuncAttributePreferredSharedMemoryCarveout,           \
                                cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e
  PM_T2_ATTR(T2Wide, 0); PM_T2_ATTR(T2Wide, 1); PM_T2_ATTR(T2Wide, 2);
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2Wide, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                T2Wide::kSmemBytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2Wide, 2, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  PM_T2_ATTR(T2Deep, 0); PM_T2_ATTR(T2Deep, 1); PM_T2_ATTR(T2Deep, 2);
  PM_T2_ATTR(T2F128, 3); PM_T2_ATTR(T2F256, 3);
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2Wide, 2, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                T2Wide::kSmemBytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2Wide, 2, false, 1>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2F256, 2, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                T2F256::kSmemBytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2F256, 2, false, 1>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2I8, 2, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                T2I8::kSmemBytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2I8, 2, false, 2>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2I8, 1, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                T2I8::kSmemBytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2I8, 2, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                T2I8::kSmemBytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2I8, 2, true, 2>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
#undef PM_T2_ATTR
  return cudaSuccess;
}

// variant 0: 256-column tiles / 16 epilogue warps; variant 1: 192-column tiles / 24 epilogue warps.
// mode 0: exact top-2 with indices; 1: timing probe (garbage results); 2: values only (run l2_fixup next);
// 4: as 2, 64-register build (256-column variant only).
cudaError_t launch_l2_tc2(const TcMaps& maps, const int32_t* qnorm, const PairJob* jobs, int n_jobs, int max_nq,
                          int2* idx, float2* dist, int stride, int num_sms, int variant, int mode, cudaStream_t st) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  const int tiles_per_job = (max_nq + T2_ROWS - 1) / T2_ROWS;
  const int n_items = n_jobs * tiles_per_job;
  int clusters = num_sms / 2;
  if (n_items < clusters) clusters = n_items;
  const int grid = clusters * 2;
  const CUtensorMap& tm = variant == 1 ? maps.t_main96 : maps.t_main;
  const CUtensorMap& te = variant == 1 ? maps.t_ext96 : maps.t_ext;
#define PM_T2_LAUNCH(CFG, M)                                                                         \
  l2_top2_tc2_kernel<CFG, M><<<grid, CFG::kThreads, CFG::kSmemBytes, st>>>(                          \
      maps.q_main, maps.q_ext, tm, te, qnorm, jobs, n_jobs, tiles_per_job, idx, dist, stride, nullptr)
  if (variant == 1) {
    if (mode == 1) PM_T2_LAUNCH(T2Deep, 1); else if (mode == 2) PM_T2_LAUNCH(T2Deep, 2); else PM_T2_LAUNCH(T2Deep, 0);
  } else {
    if (mode == 1) PM_T2_LAUNCH(T2Wide, 1);
    else if (mode == 2) PM_T2_LAUNCH(T2Wide, 2);
    else if (mode == 4)     // values-only, 64-register build (co-resident tail kernels)
      l2_top2_tc2_kernel<T2Wide, 2, true><<<grid, T2Wide::kThreads, T2Wide::kSmemBytes, st>>>(
          maps.q_main, maps.q_ext, tm, te, qnorm, jobs, n_jobs, tiles_per_job, idx, dist, stride, nullptr);
    else PM_T2_LAUNCH(T2Wide, 0);
  }
#undef PM_T2_LAUNCH
  return cudaGetLastError();
}

// Binary descriptors (256 / 512 bit) as E4M3 rows of 32*words + 32 bytes (pack_bits_kernel): the accumulator is
// |b| - 2 a.b, an exact integer; values-only epilogue (MODE 2), hamming_fixup follows.  qnorm = popcounts.
cudaError_t launch_ham_tc2(const TcMaps& maps, int words, const int32_t* qnorm, const PairJob* jobs, int n_jobs,
                           int max_nq, int2* idx, float2* dist, int stride, int num_sms, cudaStream_t st) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  const int tiles_per_job = (max_nq + T2_ROWS - 1) / T2_ROWS;
  const int n_items = n_jobs * tiles_per_job;
  int clusters = num_sms / 2;
  if (n_items < clusters) clusters = n_items;
  const int grid = clusters * 2;
  if (words == 8)
    l2_top2_tc2_kernel<T2Wide, 2, false, 1><<<grid, T2Wide::kThreads, T2Wide::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, qnorm, jobs, n_jobs, tiles_per_job, idx, dist, stride, nullptr);
  else if (words == 16)
    l2_top2_tc2_kernel<T2F256, 2, false, 1><<<grid, T2F256::kThreads, T2F256::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, qnorm, jobs, n_jobs, tiles_per_job, idx, dist, stride, nullptr);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}


// =========================================================================================================
// K2-i8x2: the byte form with TWO query row sets per cluster.  The one-set kernel above streams the whole
// train image (8192 x 160 B = 1.3 MB) through L2 once per 256 query rows: 42 MB per pair, 7-8 TB/s at the
// rate kind::i8 consumes it -- measured, that is the L2 ceiling of this GPU, and the tensor pipe waits.
// Here a work item is 512 query rows: each CTA keeps TWO resident 128-row query tiles and every train tile
// is multiplied with both (set 0 -> accumulator stage 0, set 1 -> stage 1), so L2 traffic per pair halves.
// The two accumulator stages are double-buffered by construction: the epilogue reduces stage 0 while the
// MMAs of set 1 run, and vice versa.  Everything else as in l2_top2_tc2_kernel<T2I8, 2, *, 2>.
//   smem per CTA: query tiles 2 items x 2 sets x 20 KB, train ring 4 x 20 KB, barriers, slice exchange.
// MODE 2: product; 1: TMA + MMA only; 5: + tcgen05.ld without the reduction (timing probes, no results).
template <int KA, int BN_ = 256, int STG = 0, int PK = 0>
struct I8X2Cfg {                                                 // KA = 128-byte K atoms per row: 1 (SIFT bytes), 2 (256-bit rows)
  static constexpr int kKA = KA;                                 // BN_ = 192: the fp4 form (TMEM columns 384.. hold the scale factors)
  static constexpr int kBN = BN_, kBNH = BN_ / 2, kSets = 2;
  static constexpr int kSub = PK ? 2 : 1;                        // PK: a stage holds TWO train sub-tiles (rows c and c + BN share a column)
  static constexpr int kAStages = KA == 1 ? 2 : 1, kStages = STG ? STG : (PK ? 3 : (KA == 1 ? (BN_ == 256 ? 4 : 5) : 3));
  static constexpr int kTile = KA * T2_ATOM + T2_EXT;            // 20 / 36 KB: 128 rows x (128 KA + 32) B
  static constexpr int kSmemA = kAStages * kSets * kTile;        // 80 / 72 KB
  static constexpr int kBAtom = kBNH * 128;
  static constexpr int kBSub = KA * kBAtom + kBNH * 32;          // 20 / 36 KB (15 KB for 96-row halves)
  static constexpr int kBMain = KA * kBAtom;                     // PK: the sub-tiles' main atoms, then ONE shared norm tile
  static constexpr int kBTile = PK ? 2 * kBMain + kBNH * 32 : kBSub;
  static constexpr int kTileRows = kSub * BN_;                   // train rows per stage
  static constexpr int kSmemB = kStages * kBTile;                // 80 / 108 KB
  static constexpr int kSlices = BN_ / 64, kEpiWarps = 4 * kSlices, kThreads = 128 + 32 * kEpiWarps;
  static constexpr int kXchg = 4 * (kSlices - 1) * 32 * 32;
  static constexpr int kSmemBytes = kSmemA + kSmemB + 1024 + 256 + kXchg;
  static constexpr int kRows = kSets * T2_ROWS;                  // 512 query rows per work item
  static constexpr uint32_t kSfCol = 2 * BN_;                    // fp4 form: first TMEM column of the scale factors
  static constexpr uint32_t kShape = ((kBN >> 3) << 17) | ((T2_ROWS >> 4) << 24);
  static constexpr uint32_t kIdesc = (2u << 4) | (1u << 10) | kShape;      // D = s32, A = u8, B = s8
  static constexpr uint32_t kIdescExt = (2u << 4) | kShape;                // norm block: A = B = u8
  static constexpr uint32_t kIdescS8 = (2u << 4) | (1u << 7) | (1u << 10) | kShape;   // MODE 3: A = B = s8 (quantised rows)
  // kind::mxf4 (block-scaled instruction descriptor): A = B = E2M1 (format 1), scale factors UE8M0, K = 64, D = f32
  static constexpr uint32_t kIdescFp4 = (1u << 7) | (1u << 10) | (1u << 23) | kShape;
};
using I8X2 = I8X2Cfg<1>;

// LEAN: 0 = no register cap, 1 = 64 registers (launch bound of 1024 threads), 3 = 96 registers (bound of 640 threads: the
// packed form's 512 threads then leave the 16 K registers of one RANSAC block), 2 = 72 registers (bound of 896 threads):
// 640 threads x 72 leave 19 K registers for the tail kernels of earlier batches (the RANSAC block needs 16 K)
// NX = 1 (MODE 4 only, real-valued rows whose squared norms are all within L2S8_UNIT_TOL of 1 -- SuperPoint rows are
// L2-normalised, FeatureSuperPoint.cpp:195-205): the norm term of the score is the same constant for every train row, so
// the norm block is neither loaded nor multiplied -- 4 KA K-steps per tile instead of 4 KA + 1 -- and the constant
// T2K_UNIT_OFF = 254^2 (1/2 + 1) enters the chunk keys in the epilogue.  The scores are those of rows of norm exactly 1;
// the re-rank's certified bound carries the tolerance (ff_bound e_mode 2).
static constexpr int T2K_UNIT_OFF = 96774;                      // 254^2 * 1.5, exact
// PK = 1 (256-bit rows on kind::mxf4, MODE 1 / 2): TWO train rows per accumulator column.  Draining a 192-column
// accumulator through tcgen05.ld (64 B/clk per scheduler) takes 80 % of the time the tensor pipe needs for the other row
// set's 5 K-steps, so the hand-over commit -> wake -> ld -> arrive -> issue never hid.  The
// block scale factors of kind::mxf4 halve the drain: a stage holds two train sub-tiles (rows c.. and c + 192..), the
// K-steps of the FIRST one run with UE8M0 scale 2^10 on the B side, those of the second with scale 1, all into one
// accumulator.  Every packed row carries bias slots in its norm block worth 512 (pack.cu), so a column ends as
//     1024 (512 + H(c)) + (512 + H(c + 192)) = 2^19 + 1024 H(c) + (512 + H(c + 192)):
// every value lies in the binade of 2^19 (the first sub-tile's row exists whenever the second one's does; an absent or
// zero second row adds 0), so the fp32 bit pattern is 0x49000000 | n << 4 and both Hamming distances (<= 256; exact:
// every partial sum is an integer below 2^20) are read with shifts and masks.  Ten K-steps between hand-overs instead of
// five, for half as many accumulator reads; both 32-column loads of a warp are issued before anything is reduced.
template <int MODE, int LEAN, int KA = 1, int FP4 = 0, int STG = 0, int NX = 0, int PK = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LEAN == 1 ? 1024 : LEAN == 2 ? 896 : LEAN == 3 ? 640 : I8X2Cfg<KA, FP4 ? 192 : 256>::kThreads, 1)
l2_i8x2_kernel(const __grid_constant__ CUtensorMap q_main, const __grid_constant__ CUtensorMap q_ext,
               const __grid_constant__ CUtensorMap t_main, const __grid_constant__ CUtensorMap t_ext,
               const PairJob* __restrict__ jobs, int n_jobs, int blocks_per_job, int2* __restrict__ knn_idx,
               float2* __restrict__ knn_dist, int stride, float2* __restrict__ extra_keys = nullptr) {
  using C = I8X2Cfg<KA, FP4 ? 192 : 256, STG, PK>;
  static_assert(MODE < 3 || MODE == 5 || !FP4, "candidate keys are integer keys of the kind::i8 accumulators");
  static_assert(!PK || (FP4 && KA == 1 && (MODE == 1 || MODE == 2)), "packed pairs: 256-bit rows on kind::mxf4");
  constexpr int BN = C::kBN, BNH = C::kBNH, ST = C::kStages, TILE = C::kTile, BTILE = C::kBTile, NSL = C::kSlices;
  constexpr int AST = C::kAStages, KDIM = 128 * KA, BATOM = C::kBAtom;
  static_assert(!FP4 || KA <= 2, "fp4 forms: one 128-byte K atom per 256 bits (256- and 512-bit rows)");
  static_assert(!NX || MODE == 4, "the unit-norm form exists for the argmin epilogue only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::kSmemA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kSmemA + C::kSmemB);
  uint64_t* a_full = bars + 0;                       // [2] leader
  uint64_t* a_empty = bars + 2;                      // [2] both (multicast commit)
  uint64_t* b_full = bars + 4;                       // [ST] leader
  uint64_t* b_empty = bars + 4 + ST;                 // [ST] both
  uint64_t* acc_full = bars + 4 + 2 * ST;            // [2] both; stage = row set
  uint64_t* acc_empty = acc_full + 2;                // [2] leader: 32 epilogue warps of the pair
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float4* xchg = reinterpret_cast<float4*>(smem + C::kSmemA + C::kSmemB + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&q_main); tma_prefetch_desc(&q_ext);
    tma_prefetch_desc(&t_main); tma_prefetch_desc(&t_ext);
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < ST; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 2 * C::kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(T2_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (FP4) {
    // kind::mxf4 is block-scaled: every scale factor is 1.0 (UE8M0 0x7F).  64 TMEM columns of all 128 lanes of both
    // CTAs are filled with it once, so the scale-factor layout the instruction expects does not matter.
    if (warp >= 4 && warp < 8) {
      const uint32_t a = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + C::kSfCol;
      tmem_st_32x32b_x32_fill(a, 0x7F7F7F7Fu);
      tmem_st_32x32b_x32_fill(a + 32, 0x7F7F7F7Fu);
      if (PK) {                                          // columns 448..479: scale 2^10 for the first sub-tile's B rows;
        tmem_st_32x32b_x32_fill(a + 64, T2P_SF_BIG);     // 480..511: 2^10 for the first K block, 1 for the second (norm step)
        tmem_st_32x32b_x32_fill(a + 96, T2P_SF_MIX);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
  }

  const int n_items = n_jobs * blocks_per_job;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ==========================
    if (lane == 0) {
      uint32_t ai = 0, bi = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int jb = item / blocks_per_job, blk = item - jb * blocks_per_job;
        const PairJob job = jobs[jb];
        if (blk * C::kRows >= job.nq) continue;
        const int nset = blk * C::kRows + T2_ROWS < job.nq ? 2 : 1;
        {
          const uint32_t as = ai % AST, use = ai / AST;
          wait_tma(&a_empty[as], (use & 1) ^ 1);
          if (leader) mbar_expect_tx(&a_full[as], nset * 2 * (NX ? TILE - T2_EXT : TILE));
          for (int set = 0; set < nset; ++set) {
            uint8_t* dst = sA + (as * 2 + set) * TILE;
            const int row = job.q_row + blk * C::kRows + set * T2_ROWS + rank * T2_BM;
#pragma unroll
            for (int a = 0; a < KA; ++a) tma_load_2d_pair(dst + a * T2_ATOM, &q_main, 128 * a, row, &a_full[as]);
            if (!NX) tma_load_2d_pair(dst + KA * T2_ATOM, &q_ext, KDIM, row, &a_full[as]);
          }
        }
        ++ai;
        const int n_tiles = (job.nt + C::kTileRows - 1) / C::kTileRows;
        for (int n = 0; n < n_tiles; ++n, ++bi) {
          const uint32_t st = bi % ST;
          wait_tma(&b_empty[st], ((bi / ST) & 1) ^ 1);
          // PK: the second sub-tile (rows + BN) is neither loaded nor multiplied when it lies past the image
          const int nsub = PK && n * C::kTileRows + BN < job.nt ? 2 : 1;
          if (leader)
            mbar_expect_tx(&b_full[st], 2 * (PK ? nsub * C::kBMain + BNH * 32 : (NX ? BTILE - BNH * 32 : BTILE)));
          for (int sub = 0; sub < nsub; ++sub) {
            const int row = job.t_row + n * C::kTileRows + sub * BN + rank * BNH;   // this CTA's half of the train (sub-)tile
            uint8_t* dst = sB + st * BTILE + sub * (PK ? C::kBMain : C::kBSub);
#pragma unroll
            for (int a = 0; a < KA; ++a) tma_load_2d_pair(dst + a * BATOM, &t_main, 128 * a, row, &b_full[st]);
            if (!NX && !PK) tma_load_2d_pair(dst + KA * BATOM, &t_ext, KDIM, row, &b_full[st]);
          }
          // PK: t_ext is the map of the pair norm blocks -- rows r and r + BN side by side in one 32-byte operand row
          if (PK) tma_load_2d_pair(sB + st * BTILE + 2 * C::kBMain, &t_ext, 0, job.t_row + n * C::kTileRows + rank * BNH, &b_full[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer (leader CTA) ===========================
    if (leader) {
      uint32_t ai = 0, bi = 0, use[2] = {0, 0};
      constexpr uint32_t HI128 = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t HI32 = (256u >> 4) | (1u << 14) | (6u << 29);
      const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int jb = item / blocks_per_job, blk = item - jb * blocks_per_job;
        const int job_nq = jobs[jb].nq, job_nt = jobs[jb].nt;
        if (blk * C::kRows >= job_nq) continue;
        const int nset = blk * C::kRows + T2_ROWS < job_nq ? 2 : 1;
        const uint32_t a_st = ai % AST;
        wait_mma(&a_full[a_st], (ai / AST) & 1);
        ++ai;
        const int n_tiles = (job_nt + C::kTileRows - 1) / C::kTileRows;
        for (int n = 0; n < n_tiles; ++n, ++bi) {
          const uint32_t st = bi % ST;
          wait_mma(&b_full[st], (bi / ST) & 1);
          const uint32_t b_lo = b_lo0 + st * (BTILE >> 4);
          const int nsub = PK && n * C::kTileRows + BN < job_nt ? 2 : 1;
          for (int set = 0; set < nset; ++set) {
            wait_mma(&acc_empty[set], (use[set] & 1) ^ 1);
            ++use[set];
            tc_fence_after();
            if (elect_one()) {
              const uint32_t d_tmem = tmem_base + set * BN;
              const uint32_t a_lo = a_lo0 + (a_st * 2 + set) * (TILE >> 4);
              if (PK) {
                // two sub-tiles into ONE accumulator: 4 K-steps with B scale 2^10 (TMEM columns kSfCol + 64..), 4 with B
                // scale 1, and ONE norm K-step whose two K blocks are the norm blocks of the two rows (scales 2^10 | 1)
                const uint32_t sfa = tmem_base + C::kSfCol;
                for (int sub = 0; sub < nsub; ++sub) {
                  const uint32_t sfb = tmem_base + C::kSfCol + (sub ? 16 : 64);
                  const uint32_t bl = b_lo + sub * (C::kBMain >> 4);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_mxf4_pair(d_tmem, (static_cast<uint64_t>(HI128) << 32) | (a_lo + ((k * 32) >> 4)),
                                   (static_cast<uint64_t>(HI128) << 32) | (bl + ((k * 32) >> 4)), C::kIdescFp4, sfa, sfb,
                                   (sub | k) > 0 ? 1u : 0u);
                }
                umma_mxf4_pair(d_tmem, (static_cast<uint64_t>(HI32) << 32) | (a_lo + (T2_ATOM >> 4)),
                               (static_cast<uint64_t>(HI32) << 32) | (b_lo + ((2 * C::kBMain) >> 4)), C::kIdescFp4, sfa,
                               tmem_base + C::kSfCol + 96, 1u);
              } else if (FP4) {
                // 4 KA x 64 bit values of the row + the 64-value norm block: 5 / 9 K-steps of kind::mxf4 (K = 64)
                const uint32_t sfa = tmem_base + C::kSfCol, sfb = tmem_base + C::kSfCol + 16;
#pragma unroll
                for (int k = 0; k < 4 * KA; ++k)
                  umma_mxf4_pair(d_tmem,
                                 (static_cast<uint64_t>(HI128) << 32) | (a_lo + (((k >> 2) * T2_ATOM + (k & 3) * 32) >> 4)),
                                 (static_cast<uint64_t>(HI128) << 32) | (b_lo + (((k >> 2) * BATOM + (k & 3) * 32) >> 4)),
                                 C::kIdescFp4, sfa, sfb, k > 0 ? 1u : 0u);
                umma_mxf4_pair(d_tmem, (static_cast<uint64_t>(HI32) << 32) | (a_lo + ((KA * T2_ATOM) >> 4)),
                               (static_cast<uint64_t>(HI32) << 32) | (b_lo + ((KA * BATOM) >> 4)), C::kIdescFp4, sfa, sfb, 1u);
              } else {
#pragma unroll
                for (int k = 0; k < 4 * KA; ++k)
                  umma_f16_pair<2>(d_tmem,
                                   (static_cast<uint64_t>(HI128) << 32) | (a_lo + (((k >> 2) * T2_ATOM + (k & 3) * 32) >> 4)),
                                   (static_cast<uint64_t>(HI128) << 32) | (b_lo + (((k >> 2) * BATOM + (k & 3) * 32) >> 4)),
                                   MODE >= 3 ? C::kIdescS8 : C::kIdesc, k > 0 ? 1u : 0u);
                if (!NX)
                  umma_f16_pair<2>(d_tmem, (static_cast<uint64_t>(HI32) << 32) | (a_lo + ((KA * T2_ATOM) >> 4)),
                                   (static_cast<uint64_t>(HI32) << 32) | (b_lo + ((KA * BATOM) >> 4)), C::kIdescExt, 1u);
              }
              umma_commit_pair(&acc_full[set]);
              if (set == nset - 1) umma_commit_pair(&b_empty[st]);
            }
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit_pair(&a_empty[a_st]);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ======================================= epilogue (both CTAs) =============================
    const int quarter = warp & 3, slice = (warp - 4) >> 2;       // 64-column slice of the tile
    uint32_t use[2] = {0, 0};
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + slice * 64;
    uint32_t rem_empty0, rem_empty1;                             // the leader's acc_empty barriers (PK: mapped once)
    asm("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(rem_empty0) : "r"(smem_u32(&acc_empty[0])));
    asm("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(rem_empty1) : "r"(smem_u32(&acc_empty[1])));
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int jb = item / blocks_per_job, blk = item - jb * blocks_per_job;
      const PairJob job = jobs[jb];
      if (blk * C::kRows >= job.nq) continue;
      const int nset = blk * C::kRows + T2_ROWS < job.nq ? 2 : 1;
      Top2i s0, s1;
      s0.m1 = s0.m2 = s1.m1 = s1.m2 = T2I_INF;
      s0.i1 = s1.i1 = -1;
      Keys6i g0, g1;                                    // MODE 3: six candidate keys per row and row set
      g0.k1 = g0.k2 = g0.k3 = g0.k4 = g0.k5 = g0.k6 = T2I_INF;
      g1 = g0;
      Keys3w h0, h1;                                    // MODE 4
      h0.k1 = h0.k2 = h0.k3 = h0.w2 = T2K_NONE;
      h1 = h0;
      const int n_tiles = (job.nt + C::kTileRows - 1) / C::kTileRows;
      for (int n = 0; n < n_tiles; ++n) {
        const int c0 = n * C::kTileRows + slice * 64;
        const int lim = job.nt - c0;
#pragma unroll
        for (int set = 0; set < 2; ++set) {
          if (set < nset) {
            wait_epi(&acc_full[set], use[set] & 1);
            ++use[set];
            tc_fence_after();
            if (MODE == 1) {
              tc_fence_before();
              if (lane == 0) mbar_arrive_leader(&acc_empty[set]);
            } else if (PK) {
              // column e of this slice: high field = train row c0 + e, low field = train row c0 + BN + e.  Both loads,
              // the hand-back, then the chunk minima of both fields and the four updates in ascending order of the
              // chunk base (ties keep the earlier chunk).
              Top2i& s = set ? s1 : s0;
              uint32_t v[64];
              int hi0, lo0, hi1, lo1;
              tmem_ld_32x32b_x64_async(tq + set * BN, v);
              tmem_wait_pin(v);
              tc_fence_before();
              if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(set ? rem_empty1 : rem_empty0) : "memory");
              if (lim >= BN + 64) {                      // every column of both fields carries a row: one straight block
                t2p_chunk32<true>(v, 32, 32, hi0, lo0);
                t2p_chunk32<true>(v + 32, 32, 32, hi1, lo1);
                t2i_update(s, hi0, c0);
                t2i_update(s, hi1, c0 + 32);
                t2i_update(s, lo0, c0 + BN);
                t2i_update(s, lo1, c0 + BN + 32);
              } else {
                t2p_chunk32<false>(v, lim, lim - BN, hi0, lo0);
                t2p_chunk32<false>(v + 32, lim - 32, lim - BN - 32, hi1, lo1);
                if (lim > 0) t2i_update(s, hi0, c0);
                if (lim > 32) t2i_update(s, hi1, c0 + 32);
                if (lim > BN) t2i_update(s, lo0, c0 + BN);
                if (lim > BN + 32) t2i_update(s, lo1, c0 + BN + 32);
              }
            } else {
              uint32_t v[32];
              tmem_ld_32x32b_x32(tq + set * BN, v);
              if (MODE == 2) t2i_chunk32(set ? s1 : s0, v, c0, lim);
              if (MODE == 3) keysi_chunk32(set ? g1 : g0, v, c0, lim);
              if (MODE == 4) keys3_chunk32<NX ? T2K_UNIT_OFF : 0>(set ? h1 : h0, v, c0, lim);
              tmem_ld_32x32b_x32(tq + set * BN + 32, v);
              tc_fence_before();
              if (lane == 0) mbar_arrive_leader(&acc_empty[set]);
              if (MODE == 2) t2i_chunk32(set ? s1 : s0, v, c0 + 32, lim - 32);
              if (MODE == 3) keysi_chunk32(set ? g1 : g0, v, c0 + 32, lim - 32);
              if (MODE == 4) keys3_chunk32<NX ? T2K_UNIT_OFF : 0>(set ? h1 : h0, v, c0 + 32, lim - 32);
            }
          }
        }
      }
      // MODE 3: merge the six keys of the column slices (slices 1.. publish, slice 0 merges and writes the keys in
      // the layout of l2_top2_tc2_kernel<.., 3, .., 3>: knn_dist = k1,k2 | knn_idx = bits of k3,k4 | extra = k5,k6)
      if (MODE == 3) {
        float4* slot = xchg + ((quarter * (NSL - 1)) * 64 + lane);
        const int bar_id = 1 + quarter;
#pragma unroll
        for (int set = 0; set < 2; ++set) {
          if (set < nset) {
            Keys6i& g = set ? g1 : g0;
            if (slice > 0) {
              slot[(slice - 1) * 64] = make_float4(__int_as_float(g.k1), __int_as_float(g.k2), __int_as_float(g.k3), __int_as_float(g.k4));
              slot[(slice - 1) * 64 + 32] = make_float4(__int_as_float(g.k5), __int_as_float(g.k6), 0.f, 0.f);
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NSL) : "memory");
            if (slice == 0) {
#pragma unroll
              for (int o = 0; o < NSL - 1; ++o) {
                const float4 x = slot[o * 64], x2 = slot[o * 64 + 32];
                keysi_insert(g, __float_as_int(x.x)); keysi_insert(g, __float_as_int(x.y));
                keysi_insert(g, __float_as_int(x.z)); keysi_insert(g, __float_as_int(x.w));
                keysi_insert(g, __float_as_int(x2.x)); keysi_insert(g, __float_as_int(x2.y));
              }
              const int row = blk * C::kRows + set * T2_ROWS + rank * T2_BM + quarter * 32 + lane;
              if (row < job.nq) {
                const size_t o = static_cast<size_t>(jb) * stride + row;
                knn_dist[o] = make_float2(keysi_to_float(g.k1), keysi_to_float(g.k2));
                knn_idx[o] = make_int2(__float_as_int(keysi_to_float(g.k3)), __float_as_int(keysi_to_float(g.k4)));
                extra_keys[o] = make_float2(keysi_to_float(g.k5), keysi_to_float(g.k6));
              }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NSL) : "memory");
          }
        }
      }
      // MODE 4: merge (k1, k2, k3, w2) of the column slices; w2 follows the slice that holds the row minimum.  Per row:
      // knn_idx = (k1, k2), knn_dist = bits of (k3, w2); l2f_rerank1_kernel follows.
      if (MODE == 4) {
        int4* slot = reinterpret_cast<int4*>(xchg) + ((quarter * (NSL - 1)) * 32 + lane);
        const int bar_id = 1 + quarter;
#pragma unroll
        for (int set = 0; set < 2; ++set) {
          if (set < nset) {
            Keys3w& g = set ? h1 : h0;
            if (slice > 0) slot[(slice - 1) * 32] = make_int4(g.k1, g.k2, g.k3, g.w2);
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NSL) : "memory");
            if (slice == 0) {
#pragma unroll
              for (int o = 0; o < NSL - 1; ++o) {
                const int4 x = slot[o * 32];
                if (x.x < g.k1) g.w2 = x.w;
                keys3_insert(g, x.x); keys3_insert(g, x.y); keys3_insert(g, x.z);
              }
              const int row = blk * C::kRows + set * T2_ROWS + rank * T2_BM + quarter * 32 + lane;
              if (row < job.nq) {
                const size_t o = static_cast<size_t>(jb) * stride + row;
                knn_idx[o] = make_int2(g.k1, g.k2);
                knn_dist[o] = make_float2(__int_as_float(g.k3), __int_as_float(g.w2));
              }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NSL) : "memory");
          }
        }
      }
      // merge the column slices of this lane quarter (slices 1.. publish, slice 0 merges), one row set at a time
      if (MODE == 2) {
        float4* slot = xchg + ((quarter * (NSL - 1)) * 64 + lane);
        const int bar_id = 1 + quarter;
#pragma unroll
        for (int set = 0; set < 2; ++set) {
          if (set < nset) {
            Top2i& s = set ? s1 : s0;
            if (slice > 0)
              slot[(slice - 1) * 64] = make_float4(__int_as_float(s.m1), __int_as_float(s.i1), __int_as_float(s.m2), 0.f);
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NSL) : "memory");
            if (slice == 0) {
#pragma unroll
              for (int o = 0; o < NSL - 1; ++o) {
                const float4 x = slot[o * 64];
                const int ob = __float_as_int(x.y), om1 = __float_as_int(x.x), om2 = __float_as_int(x.z);
                const bool take = ob >= 0 && (s.i1 < 0 || om1 < s.m1 || (om1 == s.m1 && ob < s.i1));
                const int hi = max(s.m1, om1);
                s.m2 = min(min(s.m2, om2), hi);
                s.m1 = min(s.m1, om1);
                s.i1 = take ? ob : s.i1;
              }
              const int row = blk * C::kRows + set * T2_ROWS + rank * T2_BM + quarter * 32 + lane;
              if (row < job.nq) {
                const size_t o = static_cast<size_t>(jb) * stride + row;
                // -3: integer values of the byte forms; -2: the fp4 form's accumulators are non-negative floats
                // (integer order == float order), hamming_fixup reads them as floats
                knn_idx[o] = make_int2(s.i1, FP4 ? -2 : -3);
                if (PK)                             // integer distances -> the floats hamming_fixup reads
                  knn_dist[o] = make_float2(s.m1 == T2I_INF ? __int_as_float(T2I_INF) : static_cast<float>(s.m1),
                                            s.m2 == T2I_INF ? __int_as_float(T2I_INF) : static_cast<float>(s.m2));
                else if (FP4 && KA == 1)            // 256-bit E2M1 rows: the bias slots of the norm block added 512
                  knn_dist[o] = make_float2(s.m1 == T2I_INF ? __int_as_float(T2I_INF) : __int_as_float(s.m1) - T2P_BIAS,
                                            s.m2 == T2I_INF ? __int_as_float(T2I_INF) : __int_as_float(s.m2) - T2P_BIAS);
                else
                  knn_dist[o] = make_float2(__int_as_float(s.m1), __int_as_float(s.m2));
              }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NSL) : "memory");
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // no remote arrive / multicast may still be in flight
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T2_TMEM_COLS) : "memory");
  }
}

template <int MODE, int LEAN, int KA = 1, int FP4 = 0, int STG = 0, int NX = 0, int PK = 0>
static cudaError_t i8x2_attr() {
  cudaError_t e = cudaFuncSetAttribute(l2_i8x2_kernel<MODE, LEAN, KA, FP4, STG, NX, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       I8X2Cfg<KA, FP4 ? 192 : 256, STG, PK>::kSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(l2_i8x2_kernel<MODE, LEAN, KA, FP4, STG, NX, PK>, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}

using T2S256 = T2Cfg<256, 2, 2>;  // quantised real-valued rows, 256-d: 256 bytes of K + norm block
using T2S128 = T2Cfg<256, 2, 1>;  // 128-d
cudaError_t s8_configure() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2S256, 3, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                T2S256::kSmemBytes)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2S256, 3, true, 3>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(l2_top2_tc2_kernel<T2S128, 3, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                T2S128::kSmemBytes)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(l2_top2_tc2_kernel<T2S128, 3, true, 3>, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared);
}
// Quantised real-valued rows (s8 operand forms of dim + 32 bytes per row, pack_float_kernel): approximate scores on
// kind::i8, six candidate keys per query row in the layout of launch_l2f_tc2; l2f_fixup (e_mode 1) follows.
cudaError_t launch_l2s8_tc2(const TcMaps& maps, int dim, const PairJob* jobs, int n_jobs, int max_nq, int2* idx,
                            float2* dist, float2* extra, int stride, int num_sms, cudaStream_t st) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  const int tiles_per_job = (max_nq + T2_ROWS - 1) / T2_ROWS;
  const int n_items = n_jobs * tiles_per_job;
  int clusters = num_sms / 2;
  if (n_items < clusters) clusters = n_items;
  const int grid = clusters * 2;
  if (dim == 256)
    l2_top2_tc2_kernel<T2S256, 3, true, 3><<<grid, T2S256::kThreads, T2S256::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, nullptr, jobs, n_jobs, tiles_per_job, idx, dist, stride, extra);
  else if (dim == 128)
    l2_top2_tc2_kernel<T2S128, 3, true, 3><<<grid, T2S128::kThreads, T2S128::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, nullptr, jobs, n_jobs, tiles_per_job, idx, dist, stride, extra);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// Quantised real-valued rows with TWO query row sets per cluster (l2_i8x2_kernel MODE 3): the one-set kernel above
// streams the train image (8192 x 288 B) through L2 once per 256 query rows -- 75 MB per pair, 7.2 TB/s at the rate
// kind::i8 consumes 9 K-steps, the measured L2 ceiling -- so the tensor pipe waits AND every co-resident tail kernel of
// the previous batches starves on L2.  Two row sets halve that traffic.  Same candidate keys, same layout.
#ifndef PM_S8X2_STAGES
#define PM_S8X2_STAGES 3
#endif
#ifndef PM_S8X2_LEAN
#define PM_S8X2_LEAN 2
#endif
cudaError_t launch_l2s8x2(const TcMaps& maps, int dim, const PairJob* jobs, int n_jobs, int max_nq, int2* idx,
                          float2* dist, float2* extra, int stride, int num_sms, cudaStream_t st, int keys3) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  using C2 = I8X2Cfg<2, 256, PM_S8X2_STAGES>;
  using C1 = I8X2Cfg<1>;
  const int blocks_per_job = (max_nq + C2::kRows - 1) / C2::kRows;
  const int n_items = n_jobs * blocks_per_job;
  int clusters = num_sms / 2;
  if (n_items < clusters) clusters = n_items;
  const int grid = clusters * 2;
  if (dim == 256 && keys3 == 2)       // every train row of unit norm: no norm K-step (NX)
    l2_i8x2_kernel<4, PM_S8X2_LEAN, 2, 0, PM_S8X2_STAGES, 1><<<grid, C2::kThreads, C2::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, jobs, n_jobs, blocks_per_job, idx, dist, stride, extra);
  else if (dim == 128 && keys3 == 2)
    l2_i8x2_kernel<4, PM_S8X2_LEAN, 1, 0, 0, 1><<<grid, C1::kThreads, C1::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, jobs, n_jobs, blocks_per_job, idx, dist, stride, extra);
  else if (dim == 256 && keys3)
    l2_i8x2_kernel<4, PM_S8X2_LEAN, 2, 0, PM_S8X2_STAGES><<<grid, C2::kThreads, C2::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, jobs, n_jobs, blocks_per_job, idx, dist, stride, extra);
  else if (dim == 128 && keys3)
    l2_i8x2_kernel<4, PM_S8X2_LEAN, 1><<<grid, C1::kThreads, C1::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, jobs, n_jobs, blocks_per_job, idx, dist, stride, extra);
  else if (dim == 256)
    l2_i8x2_kernel<3, PM_S8X2_LEAN, 2, 0, PM_S8X2_STAGES><<<grid, C2::kThreads, C2::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, jobs, n_jobs, blocks_per_job, idx, dist, stride, extra);
  else if (dim == 128)
    l2_i8x2_kernel<3, PM_S8X2_LEAN, 1><<<grid, C1::kThreads, C1::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, jobs, n_jobs, blocks_per_job, idx, dist, stride, extra);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t i8x2_configure() {
  cudaError_t e;
  if ((e = i8x2_attr<4, PM_S8X2_LEAN, 2, 0, PM_S8X2_STAGES>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<4, PM_S8X2_LEAN, 1>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<4, PM_S8X2_LEAN, 2, 0, PM_S8X2_STAGES, 1>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<4, PM_S8X2_LEAN, 1, 0, 0, 1>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<3, PM_S8X2_LEAN, 2, 0, PM_S8X2_STAGES>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<3, PM_S8X2_LEAN, 1>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<2, false>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<2, true>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<1, false>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<2, false, 2>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<1, false, 2>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<2, false, 1, 1>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<1, false, 1, 1>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<2, false, 2, 1>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<1, false, 2, 1>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<2, 3, 1, 1, 0, 0, 1>()) != cudaSuccess) return e;
  if ((e = i8x2_attr<1, false, 1, 1, 0, 0, 1>()) != cudaSuccess) return e;
  return i8x2_attr<5, false>();
}

// 256-bit binary rows as bytes (pack_bits_kernel: query bit -> 0 / 1 (u8), train bit -> +1 / -1 (s8), norm block =
// popcount of the train row): the s32 accumulator IS the Hamming distance.  Two query row sets per cluster like the
// SIFT kernel; hamming_fixup (32-column chunks, integer inputs) follows.  probe: TMA + MMA timing probe.
cudaError_t launch_ham_i8x2(const TcMaps& maps, const PairJob* jobs, int n_jobs, int max_nq, int2* idx, float2* dist,
                            int stride, int num_sms, int probe, cudaStream_t st) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  using C = I8X2Cfg<2>;
  const int blocks_per_job = (max_nq + C::kRows - 1) / C::kRows;
  const int n_items = n_jobs * blocks_per_job;
  int clusters = num_sms / 2;
  if (n_items < clusters) clusters = n_items;
  const int grid = clusters * 2;
  if (probe)
    l2_i8x2_kernel<1, false, 2><<<grid, C::kThreads, C::kSmemBytes, st>>>(maps.q_main, maps.q_ext, maps.t_main, maps.t_ext,
                                                                          jobs, n_jobs, blocks_per_job, idx, dist, stride);
  else
    l2_i8x2_kernel<2, false, 2><<<grid, C::kThreads, C::kSmemBytes, st>>>(maps.q_main, maps.q_ext, maps.t_main, maps.t_ext,
                                                                          jobs, n_jobs, blocks_per_job, idx, dist, stride);
  return cudaGetLastError();
}

// 256-bit binary rows as E2M1 values on kind::mxf4 (pack_bits_kernel: query bit -> 0 / 1, train bit -> +1 / -1, 64-value
// norm block = popcount of the train row as products of representable digits): 64 values of K per instruction at the
// instruction rate of kind::i8, so a train tile costs 5 K-steps instead of 9 and the fp32 accumulator IS the Hamming
// distance (exact: every partial sum is an integer of magnitude <= 512).  192-column train tiles leave TMEM columns for
// the (all-ones) scale factors.  hamming_fixup (32-column chunks, float inputs) follows.  probe: TMA + MMA timing probe.
// packed (256-bit rows): two train rows per accumulator column (l2_i8x2_kernel PK), same results.
cudaError_t launch_ham_fp4x2(const TcMaps& maps, const PairJob* jobs, int n_jobs, int max_nq, int2* idx, float2* dist,
                             int stride, int num_sms, int probe, cudaStream_t st, int words, int packed) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  using C = I8X2Cfg<1, 192>;
  using C2 = I8X2Cfg<2, 192>;
  const int blocks_per_job = (max_nq + C::kRows - 1) / C::kRows;
  const int n_items = n_jobs * blocks_per_job;
  int clusters = num_sms / 2;
  if (n_items < clusters) clusters = n_items;
  const int grid = clusters * 2;
  if (words == 16) {                  // 512-bit rows: two K atoms, 9 K-steps (the E4M3 form needs 17)
    if (probe)
      l2_i8x2_kernel<1, false, 2, 1><<<grid, C2::kThreads, C2::kSmemBytes, st>>>(
          maps.q_main, maps.q_ext, maps.t_main96, maps.t_ext96, jobs, n_jobs, blocks_per_job, idx, dist, stride);
    else
      l2_i8x2_kernel<2, false, 2, 1><<<grid, C2::kThreads, C2::kSmemBytes, st>>>(
          maps.q_main, maps.q_ext, maps.t_main96, maps.t_ext96, jobs, n_jobs, blocks_per_job, idx, dist, stride);
    return cudaGetLastError();
  }
  if (words != 8) return cudaErrorInvalidValue;
  if (packed) {
    using CP = I8X2Cfg<1, 192, 0, 1>;
    if (probe)
      l2_i8x2_kernel<1, false, 1, 1, 0, 0, 1><<<grid, CP::kThreads, CP::kSmemBytes, st>>>(
          maps.q_main, maps.q_ext, maps.t_main96, maps.t_ext2x96, jobs, n_jobs, blocks_per_job, idx, dist, stride);
    else
      l2_i8x2_kernel<2, 3, 1, 1, 0, 0, 1><<<grid, CP::kThreads, CP::kSmemBytes, st>>>(
          maps.q_main, maps.q_ext, maps.t_main96, maps.t_ext2x96, jobs, n_jobs, blocks_per_job, idx, dist, stride);
    return cudaGetLastError();
  }
  if (probe)
    l2_i8x2_kernel<1, false, 1, 1><<<grid, C::kThreads, C::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main96, maps.t_ext96, jobs, n_jobs, blocks_per_job, idx, dist, stride);
  else
    l2_i8x2_kernel<2, false, 1, 1><<<grid, C::kThreads, C::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main96, maps.t_ext96, jobs, n_jobs, blocks_per_job, idx, dist, stride);
  return cudaGetLastError();
}

// variant: 0 = product (72 registers: the tail kernels of earlier batches stay co-resident), 1 = TMA + MMA probe,
// 2 = + accumulator loads probe, 3 = 64-register build
cudaError_t launch_l2i8x2(const TcMaps& maps, const PairJob* jobs, int n_jobs, int max_nq, int2* idx, float2* dist,
                          int stride, int num_sms, int variant, cudaStream_t st) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  const int blocks_per_job = (max_nq + I8X2::kRows - 1) / I8X2::kRows;
  const int n_items = n_jobs * blocks_per_job;
  int clusters = num_sms / 2;
  if (n_items < clusters) clusters = n_items;
  const int grid = clusters * 2;
#define PM_I8X2_LAUNCH(M, L)                                                                                     \
  l2_i8x2_kernel<M, L><<<grid, I8X2::kThreads, I8X2::kSmemBytes, st>>>(maps.q_main, maps.q_ext, maps.t_main,    \
                                                                       maps.t_ext, jobs, n_jobs, blocks_per_job, \
                                                                       idx, dist, stride)
  if (variant == 1) PM_I8X2_LAUNCH(1, false);
  else if (variant == 2) PM_I8X2_LAUNCH(5, false);
  else if (variant == 3) PM_I8X2_LAUNCH(2, true);
  else PM_I8X2_LAUNCH(2, false);
#undef PM_I8X2_LAUNCH
  return cudaGetLastError();
}

// Integer-valued 128-d rows in the i8 form (pack_sift_kernel): D' per row in (knn_idx = chunk base, -3 |
// knn_dist = int bits of m1, m2'); l2_fixup_i8 follows.  One query row set per cluster (the two-set kernel
// l2_i8x2_kernel above is the default; this one is kept for A/B and parity).  probe: 1 / 2 = TMA + MMA timing probe
// (no results), 3 = 64-register build.
cudaError_t launch_l2i8_tc2(const TcMaps& maps, const PairJob* jobs, int n_jobs, int max_nq, int2* idx, float2* dist,
                            int stride, int num_sms, int probe, cudaStream_t st) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  const int tiles_per_job = (max_nq + T2_ROWS - 1) / T2_ROWS;
  const int n_items = n_jobs * tiles_per_job;
  int clusters = num_sms / 2;
  if (n_items < clusters) clusters = n_items;
  const int grid = clusters * 2;
  if (probe == 3)     // 64-register build: the tail kernels of the previous batch can be co-resident
    l2_top2_tc2_kernel<T2I8, 2, true, 2><<<grid, T2I8::kThreads, T2I8::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, nullptr, jobs, n_jobs, tiles_per_job, idx, dist, stride, nullptr);
  else if (probe)
    l2_top2_tc2_kernel<T2I8, 1, false, 2><<<grid, T2I8::kThreads, T2I8::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, nullptr, jobs, n_jobs, tiles_per_job, idx, dist, stride, nullptr);
  else
    l2_top2_tc2_kernel<T2I8, 2, false, 2><<<grid, T2I8::kThreads, T2I8::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, nullptr, jobs, n_jobs, tiles_per_job, idx, dist, stride, nullptr);
  return cudaGetLastError();
}

// =========================================================================================================
// Tensor-pipe peak of one MMA kind, measured live (pm_measure_tensor_peak): the issue loop of the kernels above with
// nothing around it.  One CTA pair per two SMs, operands resident in shared memory (one 128-byte K atom of 128 rows per
// CTA and operand, pseudo-random bytes so that the data path toggles like real operands), cta_group::2 M = 256, N = 256,
// four K-steps per iteration alternating between two accumulators, no TMA, no epilogue, one commit at the end.
//   KIND 0: kind::f16 (K = 16 per instruction), 2: kind::i8 (K = 32), 4: kind::mxf4 block-scaled (K = 64)
// FLOP per instruction and cluster: 2 * 256 * 256 * K.
template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
tensor_peak_kernel(int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + T2_ATOM;
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 2 * T2_ATOM);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = cluster_ctarank() == 0;
  for (int i = threadIdx.x; i < 2 * T2_ATOM / 4; i += blockDim.x) {
    uint32_t x = static_cast<uint32_t>(i) * 2654435761u + blockIdx.x * 40503u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    if (KIND == 0) x &= 0x3BFF3BFFu;                   // fp16 operands: finite values of magnitude < 1.5
    reinterpret_cast<uint32_t*>(smem)[i] = x;
  }
  fence_proxy_async();
  if (warp == 0 && lane == 0) { mbar_init(done, 1); fence_mbar_init(); }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(T2_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (KIND == 4) {                                       // all scale factors 1.0 (UE8M0 0x7F), as in l2_i8x2_kernel
    const uint32_t a = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 384;
    tmem_st_32x32b_x32_fill(a, 0x7F7F7F7Fu);
    tmem_st_32x32b_x32_fill(a + 32, 0x7F7F7F7Fu);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
  }
  if (warp == 1 && leader) {
    constexpr uint32_t HI128 = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr int BN = KIND == 4 ? 192 : 256;          // the fp4 kernels use 192-column tiles (TMEM holds the scale factors)
    constexpr uint32_t shape = ((BN >> 3) << 17) | ((T2_ROWS >> 4) << 24);
    constexpr uint32_t idesc = KIND == 0 ? ((1u << 4) | shape) : KIND == 2 ? ((2u << 4) | (1u << 10) | shape)
                                                                         : ((1u << 7) | (1u << 10) | (1u << 23) | shape);
    const uint32_t a_lo = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const uint32_t d_tmem = tmem_base + (it & 1) * BN;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = (static_cast<uint64_t>(HI128) << 32) | (a_lo + ((k * 32) >> 4));
          const uint64_t bd = (static_cast<uint64_t>(HI128) << 32) | (b_lo + ((k * 32) >> 4));
          if (KIND == 4) umma_mxf4_pair(d_tmem, ad, bd, idesc, tmem_base + 384, tmem_base + 384 + 16, (it > 1 || k > 0) ? 1u : 0u);
          else umma_f16_pair<KIND>(d_tmem, ad, bd, idesc, (it > 1 || k > 0) ? 1u : 0u);
        }
      }
      umma_commit_pair(done);
    }
    __syncwarp();
  }
  wait_sleep<64>(done, 0);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(T2_TMEM_COLS) : "memory");
  }
}
// kind: 0 f16, 1 i8, 2 mxf4 (PM_PEAK_KIND_*).  flop_per_launch receives the FLOP of one launch.
cudaError_t launch_tensor_peak(int kind, int iters, int num_sms, double* flop_per_launch, cudaStream_t st) {
  const int grid = (num_sms / 2) * 2;
  const int smem = 2 * T2_ATOM + 1024 + 64;
  const int kvals = kind == 0 ? 16 : (kind == 1 ? 32 : 64);
  const int bn = kind == 2 ? 192 : 256;
  *flop_per_launch = static_cast<double>(grid / 2) * iters * 4.0 * 2.0 * 256.0 * bn * kvals;
  if (kind == 0) tensor_peak_kernel<0><<<grid, 128, smem, st>>>(iters);
  else if (kind == 1) tensor_peak_kernel<2><<<grid, 128, smem, st>>>(iters);
  else if (kind == 2) tensor_peak_kernel<4><<<grid, 128, smem, st>>>(iters);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// Real-valued rows (fp16 operand forms of 64*KA + 16 halfs per row): approximate scores, 4 best chunk
// keys per query row in (knn_dist = k1,k2 | knn_idx = bits of k3,k4 | extra = k5,k6).  dim: 128 or 256.
cudaError_t launch_l2f_tc2(const TcMaps& maps, int dim, const PairJob* jobs, int n_jobs, int max_nq, int2* idx,
                           float2* dist, float2* extra, int stride, int num_sms, cudaStream_t st) {
  if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
  const int tiles_per_job = (max_nq + T2_ROWS - 1) / T2_ROWS;
  const int n_items = n_jobs * tiles_per_job;
  int clusters = num_sms / 2;
  if (n_items < clusters) clusters = n_items;
  const int grid = clusters * 2;
  if (dim == 128)
    l2_top2_tc2_kernel<T2F128, 3><<<grid, T2F128::kThreads, T2F128::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, nullptr, jobs, n_jobs, tiles_per_job, idx, dist, stride, extra);
  else if (dim == 256)
    l2_top2_tc2_kernel<T2F256, 3><<<grid, T2F256::kThreads, T2F256::kSmemBytes, st>>>(
        maps.q_main, maps.q_ext, maps.t_main, maps.t_ext, nullptr, jobs, n_jobs, tiles_per_job, idx, dist, stride, extra);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace pm
