// tcgen05 / TMA / TMEM path of the GE2E loss (TF32 operands, fp32 accumulators in tensor memory).
//
// One warp-specialised kernel, three modes (the tensor-core twin of ge2e_simt.cu's strip kernel).
// An "owner" operand tile X[128, D] stays resident in shared memory, a "stream" operand Y is
// pulled through a TMA ring 128 rows at a time:
//
//   MMA1   T[128 x 128] = X . Y_tile^T          (SS, both K-major, K = D)          -> TMEM
//   FWD    epilogue: online log-sum-exp (softmax) / running arg-max (contrast) over T's columns;
//          S = w (cos + eps) + b is never written anywhere (reference s3:64-79, s3:27, s3:114-127)
//   BWD    epilogue: G = w g (softmax(S) - onehot) with the leave-one-out diagonal masked,
//          rounded to TF32 and written back over T in TMEM;
//   MMA2   Acc[128 x D] += G . Y_tile            (A from TMEM, B MN-major from smem) -> TMEM
//            BWD_DE: X = E_hat rows, Y = C_hat     -> dE_hat = (wG) C_hat
//            BWD_DC: X = C_hat rows, Y = E_hat     -> dC_hat = (wG)^T E_hat
//
// Thread-block clusters: the C CTAs of a cluster own C consecutive owner tiles and walk the SAME
// stream tiles in lock step; every stream slab is fetched from L2 once per cluster (each CTA
// issues 1/C of it with TMA multicast), which is what the measured ~6.3 TB/s L2->SM ceiling
// demands (one CTA alone needs 128 KB of stream operand per 128x128x256 tile).
//
// Work is a flat list of (owner group, stream tile) pairs cut into equal contiguous ranges over
// the clusters (stream-K): a cluster whose range covers only part of an owner group publishes a
// partial result (FWD: (max, sum) per row, merged by the last CTA to finish that tile; BWD: fp32
// atomic adds into a zeroed output).
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one thread),
// warp 2 = TMEM allocator, warps 4-11 = epilogue (TMEM lane quarter = warp % 4, column half =
// (warp - 4) / 4: two threads share a tile row).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <limits.h>
#include <stdlib.h>

#include <algorithm>

#include "ge2e_common.cuh"
#include "ge2e_tc_ptx.cuh"

namespace ge2e {

namespace {

using namespace ptx;

constexpr int kTile = 128;            // owner / stream tile rows (= UMMA M = MMA1 N)
constexpr int kSlabCols = 32;         // fp32 columns per 128-byte swizzled row
constexpr int kSlabBytes = kTile * 128;       // one [128 x 32] K-major slab = 16 KB
constexpr int kStages = 6;            // TMA ring depth (16 KB each)
constexpr int kStageBytes = 16384;
constexpr int kMma2Rows = 16;         // stream rows per MMA2 ring stage ([D/32][16][32] MN-major)
constexpr int kMaxSlabs = 8;          // D <= 256
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreadsTc = (kEpiWarp0 + kEpiWarps) * 32;
constexpr int kMaxCluster = 4;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

enum { TC_FWD = 0, TC_BWD_DE = 1, TC_BWD_DC = 2 };

// barrier indices inside the shared barrier array
enum {
  BAR_FULL = 0,                       // [kStages]  TMA -> MMA
  BAR_EMPTY = BAR_FULL + kStages,     // [kStages]  MMA (of every CTA in the cluster) -> TMA
  BAR_A_FULL = BAR_EMPTY + kStages,   // owner tile landed
  BAR_A_EMPTY,                        // owner tile no longer read by MMA1
  BAR_S_FULL,                         // [2] MMA1 result in TMEM
  BAR_S_EMPTY = BAR_S_FULL + 2,       // [2] FWD: epilogue drained T
  BAR_G_FULL = BAR_S_EMPTY + 2,       // [2] BWD: G written back to TMEM
  BAR_ACC_FULL = BAR_G_FULL + 2,      // BWD: accumulator complete for this segment
  BAR_ACC_EMPTY,                      // BWD: accumulator drained
  BAR_COUNT
};

struct TcParams {
  int n_own, n_str, D, kslabs;
  int M, spk_offset;
  int OT, ST;                 // owner tiles, stream tiles
  int C, OG;                  // cluster size, owner groups = ceil(OT / C)
  long long GP;               // OG * ST (group, stream tile) pairs
  const float* cos_diag;      // [U_local]
  const float* row_stat;      // BWD: lse per local utterance row
  const float* row_aux;       // BWD_DE: q = 1 - p_jj per local utterance row
  float* row_aux_out;         // FWD softmax
  const float* w;
  const float* b;
  const float* grad_out;
  float eps;
  // FWD outputs
  float* row_stat_out;
  int32_t* kstar_out;
  float* loss_accum;
  float* per_row_out;
  int* seg_done;              // [OT] stream tiles finished per owner tile (zeroed by the host)
  float2* seg_part;           // [OT][maxseg][128] partial row state
  int maxseg;
  // BWD outputs
  float* acc_out;             // dE_hat [n_own, D] or dC_hat_partial [n_own, D]
  float* dwdb;                // BWD_DE
  unsigned long long* trace;  // debug: [CTA][3 roles][kTraceEvents] globaltimer stamps, or nullptr
  int dbg;                    // debug (GE2E_TC_DEBUG): 1 = no TMA for stream stages, 2 = no MMA issue,
                              //                        4 = epilogue skips the math (results are garbage)
};

constexpr int kTraceEvents = 64;
struct Tracer {
  unsigned long long* buf;
  int n;
  __device__ __forceinline__ Tracer(unsigned long long* base, int role)
      : buf(base ? base + (static_cast<size_t>(blockIdx.x) * 3 + role) * kTraceEvents : nullptr), n(0) {}
  __device__ __forceinline__ void mark() {
    if (buf != nullptr && n < kTraceEvents) buf[n++] = globaltimer_ns();
  }
};

struct SharedTail {
  uint64_t bars[BAR_COUNT];
  uint32_t tmem_base;
  int flag;
  float red[2 * kEpiWarps];
  union {                                 // 1 KB either way: the budget above the ring is ~2 KB
    alignas(16) float lse_s[2][kTile];    // BWD_DC: lse (log2 domain) of the current stream rows
    alignas(16) float2 xch[kTile];        // FWD: row state of the upper column half
  };
};
static_assert(1024 + kMaxSlabs * kSlabBytes + kStages * kStageBytes + sizeof(SharedTail) <= 232448,
              "dynamic shared memory over the 227 KB per-CTA limit");

__device__ __forceinline__ int cluster_of_pair(long long gp, long long GP, int NC) {
  // largest c with floor(c * GP / NC) <= gp
  return static_cast<int>(((gp + 1) * NC + GP - 1) / GP) - 1;
}

template <int MODE, int VARIANT>
__global__ void __launch_bounds__(kThreadsTc, 1)
tc_strip_kernel(const __grid_constant__ CUtensorMap tm_own, const __grid_constant__ CUtensorMap tm_str2,
                const __grid_constant__ CUtensorMap tm_str3, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr bool kBwd = (MODE != TC_FWD);
  constexpr int kTmemCols = kBwd ? 512 : 256;

  // carve: [owner slabs 8 x 16 KB][ring kStages x 16 KB][tail]; the dynamic window starts at the
  // same CTA-relative offset in every CTA of the cluster, so multicast offsets line up
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_smem = smem_base;
  const uint32_t ring_smem = smem_base + kMaxSlabs * kSlabBytes;
  SharedTail* tail = reinterpret_cast<SharedTail*>(smem_al + kMaxSlabs * kSlabBytes + kStages * kStageBytes);
  const uint32_t bars = smem_u32(&tail->bars[0]);
  auto bar = [&](int i) { return bars + 8u * i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.C;
  const int NC = gridDim.x / C, cl = blockIdx.x / C, cr = blockIdx.x % C;   // cluster id / rank in cluster
  const uint16_t cmask = static_cast<uint16_t>((1u << C) - 1u);
  const long long gp_begin = (static_cast<long long>(cl) * p.GP) / NC;
  const long long gp_end = (static_cast<long long>(cl + 1) * p.GP) / NC;
  const int kslabs = p.kslabs;
  const uint32_t mma2_tx = kslabs * kMma2Rows * 128;         // MMA2 stage bytes ([D/32][16][32])

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_own);
    prefetch_tmap(&tm_str2);
    if (kBwd) prefetch_tmap(&tm_str3);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(bar(BAR_FULL + i), 1); mbar_init(bar(BAR_EMPTY + i), C); }
    mbar_init(bar(BAR_A_FULL), 1);
    mbar_init(bar(BAR_A_EMPTY), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(BAR_S_FULL + i), 1);
      mbar_init(bar(BAR_S_EMPTY + i), kEpiWarps);
      mbar_init(bar(BAR_G_FULL + i), kEpiWarps);
    }
    mbar_init(bar(BAR_ACC_FULL), 1);
    mbar_init(bar(BAR_ACC_EMPTY), kEpiWarps);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(smem_u32(&tail->tmem_base));
  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();     // peers' barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      Tracer tr(p.trace, 0);
      tr.mark();
      int stage = 0, phase = 0, sg = 0;
      const int rows_c = kTile / C;                 // this CTA's share of an MMA1 slab (rows)
      const int slabs_c = kslabs / C;               // ... and of an MMA2 stage (32-column chunks)
      auto load_mma1 = [&](int st) {
        for (int ks = 0; ks < kslabs; ++ks) {
          mbar_wait(bar(BAR_EMPTY + stage), phase ^ 1);
          if (p.dbg & 1) { mbar_arrive(bar(BAR_FULL + stage)); if (++stage == kStages) { stage = 0; phase ^= 1; } continue; }
          mbar_expect_tx(bar(BAR_FULL + stage), kSlabBytes);
          const uint32_t dst = ring_smem + stage * kStageBytes;
          if (C == 1) tma_load_2d(dst, &tm_str2, ks * kSlabCols, st * kTile, bar(BAR_FULL + stage));
          else tma_load_2d_mc(dst + cr * rows_c * 128, &tm_str2, ks * kSlabCols, st * kTile + cr * rows_c,
                              bar(BAR_FULL + stage), cmask);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      };
      auto load_mma2 = [&](int st) {
        for (int kc = 0; kc < kTile / kMma2Rows; ++kc) {
          mbar_wait(bar(BAR_EMPTY + stage), phase ^ 1);
          if (p.dbg & 1) { mbar_arrive(bar(BAR_FULL + stage)); if (++stage == kStages) { stage = 0; phase ^= 1; } continue; }
          mbar_expect_tx(bar(BAR_FULL + stage), mma2_tx);
          const uint32_t dst = ring_smem + stage * kStageBytes;
          const int row0 = st * kTile + kc * kMma2Rows;
          if (C == 1) tma_load_3d(dst, &tm_str3, 0, row0, 0, bar(BAR_FULL + stage));
          else tma_load_3d_mc(dst + cr * slabs_c * kMma2Rows * 128, &tm_str3, 0, row0, cr * slabs_c,
                              bar(BAR_FULL + stage), cmask);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      };
      for (long long gp = gp_begin; gp < gp_end; ++sg) {
        const int og = static_cast<int>(gp / p.ST), s0 = static_cast<int>(gp % p.ST);
        const int s1 = static_cast<int>(min(static_cast<long long>(p.ST), s0 + (gp_end - gp)));
        const int ot = og * C + cr;       // may be >= OT in the last group: TMA zero-fills, nothing is stored
        if (sg > 0) mbar_wait(bar(BAR_A_EMPTY), (sg - 1) & 1);
        tr.mark();   // owner tile issue
        mbar_expect_tx(bar(BAR_A_FULL), kslabs * kSlabBytes);
        for (int ks = 0; ks < kslabs; ++ks)
          tma_load_2d(a_smem + ks * kSlabBytes, &tm_own, ks * kSlabCols, ot * kTile, bar(BAR_A_FULL));
        if (!kBwd) {
          for (int st = s0; st < s1; ++st) load_mma1(st);
        } else {
          load_mma1(s0);
          for (int st = s0; st < s1; ++st) {
            if (st + 1 < s1) load_mma1(st + 1);
            load_mma2(st);
          }
        }
        gp += s1 - s0;
        tr.mark();   // all loads of the segment issued
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc1 = idesc_tf32(kTile, kTile, 0, 0);
      const uint32_t idesc2 = idesc_tf32(kTile, p.D, 0, 1);
      // descriptor templates: only the 14-bit start-address field changes per MMA
      const uint64_t dk = smem_desc(0, 16, 1024, kLayoutSw128);                        // K-major
      const uint64_t dmn = smem_desc(0, kMma2Rows * 128, 512, kLayoutSw128Base32);     // MN-major TF32
      int stage = 0, phase = 0, sg = 0, it = 0;
      Tracer tr(p.trace, 1);
      tr.mark();
      auto release = [&](int s) {
        if (C == 1) umma_commit(bar(BAR_EMPTY + s)); else umma_commit_mc(bar(BAR_EMPTY + s), cmask);
      };
      auto mma1 = [&](int iter) {
        const uint32_t d_tmem = tmem + (iter & 1) * kTile;
        for (int ks = 0; ks < kslabs; ++ks) {
          mbar_wait(bar(BAR_FULL + stage), phase);
          tc_fence_after();
          const uint64_t da = dk | ((a_smem + ks * kSlabBytes) >> 4);
          const uint64_t db = dk | ((ring_smem + stage * kStageBytes) >> 4);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)     // +32 B per K step of 8 inside the 128 B swizzled row
            if (!(p.dbg & 2)) umma_tf32_ss(d_tmem, da + 2 * k4, db + 2 * k4, idesc1, (ks | k4) != 0);
          release(stage);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar(BAR_S_FULL + (iter & 1)));
      };
      auto mma2 = [&](int iter, bool first) {
        const uint32_t a_tmem = tmem + (iter & 1) * kTile;
        const uint32_t d_tmem = tmem + 2 * kTile;
        for (int kc = 0; kc < kTile / kMma2Rows; ++kc) {
          mbar_wait(bar(BAR_FULL + stage), phase);
          tc_fence_after();
          // MN-major TF32 operand: 32-byte-atom 128B swizzle, chunks of 32 columns kMma2Rows*128 B
          // apart (LBO), groups of 4 k-rows 512 B apart (SBO); one MMA consumes 8 k-rows = 1024 B
          const uint64_t db = dmn | ((ring_smem + stage * kStageBytes) >> 4);
#pragma unroll
          for (int k2 = 0; k2 < kMma2Rows / 8; ++k2)
            if (!(p.dbg & 2)) umma_tf32_ts(d_tmem, a_tmem + kc * kMma2Rows + k2 * 8, db + 64 * k2, idesc2,
                         !(first && kc == 0 && k2 == 0));
          release(stage);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      };
      for (long long gp = gp_begin; gp < gp_end; ++sg) {
        const int s0 = static_cast<int>(gp % p.ST);
        const int s1 = static_cast<int>(min(static_cast<long long>(p.ST), s0 + (gp_end - gp)));
        mbar_wait(bar(BAR_A_FULL), sg & 1);
        tc_fence_after();
        tr.mark();   // owner tile landed
        if (!kBwd) {
          for (int st = s0; st < s1; ++st, ++it) {
            mbar_wait(bar(BAR_S_EMPTY + (it & 1)), ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            mma1(it);
            tr.mark();   // pair issued
          }
          umma_commit(bar(BAR_A_EMPTY));
        } else {
          if (sg > 0) { mbar_wait(bar(BAR_ACC_EMPTY), (sg - 1) & 1); tc_fence_after(); }
          mma1(it);
          for (int st = s0; st < s1; ++st, ++it) {
            if (st + 1 < s1) mma1(it + 1); else umma_commit(bar(BAR_A_EMPTY));
            tr.mark();   // MMA1(next) issued
            mbar_wait(bar(BAR_G_FULL + (it & 1)), (it >> 1) & 1);
            tc_fence_after();
            tr.mark();   // G landed
            mma2(it, st == s0);
            tr.mark();   // MMA2 issued
          }
          umma_commit(bar(BAR_ACC_FULL));
        }
        gp += s1 - s0;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================================== epilogue
    const int ew = warp - kEpiWarp0;
    const int quarter = ew & 3, half = ew >> 2;      // TMEM lanes [32 q, 32 q + 32); columns [64 h, 64 h + 64)
    const int trow = quarter * 32 + lane;            // row inside the owner tile
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const float w = __ldg(p.w), b = __ldg(p.b), eps = p.eps;
    const float w2 = w * kLog2e, b2 = fmaf(w, eps, b) * kLog2e;   // log2-domain affine: S*log2e
    const float bb = fmaf(w, eps, b);
    const float g = kBwd ? __ldg(p.grad_out) : 1.f;
    const float wg = w * g;
    float dw_acc = 0.f, db_acc = 0.f, loss_acc = 0.f;
    int sg = 0, it = 0;
    Tracer tr((trow == 0 && half == 0) ? p.trace : nullptr, 2);
    tr.mark();

    for (long long gp = gp_begin; gp < gp_end; ++sg) {
      const int og = static_cast<int>(gp / p.ST), s0 = static_cast<int>(gp % p.ST);
      const int s1 = static_cast<int>(min(static_cast<long long>(p.ST), s0 + (gp_end - gp)));
      const int ot = og * C + cr;
      const bool tile_valid = ot < p.OT;             // CTA-uniform
      const int orow = ot * kTile + trow;            // owner row (utterance, or centroid for DC)
      const bool ovalid = orow < p.n_own;
      // per-owner-row metadata
      int jg = -1;            // FWD / DE: global speaker of this utterance row
      float cd = 0.f, lse2 = INFINITY, qd = 0.f;
      int dlo = 0, dhi = 0;   // DC: local utterance rows [dlo, dhi) belong to this centroid
      if (MODE != TC_BWD_DC) {
        if (ovalid) {
          jg = p.spk_offset + orow / p.M;
          cd = __ldg(p.cos_diag + orow);
          if (MODE == TC_BWD_DE) { lse2 = __ldg(p.row_stat + orow) * kLog2e; qd = __ldg(p.row_aux + orow); }
        }
      } else {
        const int jl = orow - p.spk_offset;
        if (ovalid && jl >= 0 && (long long)jl * p.M < p.n_str) { dlo = jl * p.M; dhi = dlo + p.M; }
      }
      // FWD softmax running state, log2 domain, OFF-diagonal columns only: the running max starts at
      // the diagonal logit (known from cos_diag) and the diagonal term joins when the row is closed
      const float xd2 = fmaf(cd, w2, b2);
      float m2 = xd2, lsum = 0.f;
      float best = -INFINITY; int bestk = INT_MAX;   // FWD contrast

      for (int st = s0; st < s1; ++st, ++it) {
        const int buf = it & 1;
        if (MODE == TC_BWD_DC) {
          // stage the lse of the 128 stream rows (utterances) of this tile
          if (half == 0) {
            const int u = st * kTile + trow;
            tail->lse_s[buf][trow] = (u < p.n_str) ? __ldg(p.row_stat + u) * kLog2e : INFINITY;
          }
          named_bar_sync(1, kEpiThreads);
        }
        mbar_wait(bar(BAR_S_FULL + buf), (it >> 1) & 1);
        tc_fence_after();
        tr.mark();   // T tile ready
        const uint32_t t_addr = tmem + lane_addr + buf * kTile;
#pragma unroll 1
        for (int ch = 2 * half; ch < 2 * half + 2; ++ch) {
          const int c0 = st * kTile + ch * 32;       // first stream row (column of T) of this chunk
          uint32_t v[32];
          tmem_ld32(t_addr + ch * 32, v);
          tmem_ld_wait();
          if (p.dbg & 4) {
            if (MODE != TC_FWD) tmem_st32(t_addr + ch * 32, v);
            continue;
          }
          if (MODE == TC_FWD) {
            if (c0 < p.n_str) {
              const bool tailc = c0 + 32 > p.n_str;
              const bool diagc = static_cast<unsigned>(jg - c0) < 32u;
              if (VARIANT == GE2E_SOFTMAX) {
                float x[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] = fmaf(__uint_as_float(v[i]), w2, b2);
                if (__any_sync(0xffffffffu, tailc || diagc)) {
#pragma unroll
                  for (int i = 0; i < 32; ++i)   // own-speaker column (s3:78) and padding leave the sum
                    if (c0 + i == jg || c0 + i >= p.n_str) x[i] = -INFINITY;
                }
                float cm = x[0];
#pragma unroll
                for (int i = 1; i < 32; ++i) cm = fmaxf(cm, x[i]);
                const float mn = fmaxf(m2, cm);
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) s += ex2(x[i] - mn);
                lsum = fmaf(lsum, ex2(m2 - mn), s);
                m2 = mn;
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  float s = fmaf(__uint_as_float(v[i]), w, bb);
                  if (c0 + i == jg || c0 + i >= p.n_str) s = -INFINITY;
                  if (s > best) { best = s; bestk = c0 + i; }
                }
              }
            }
          } else {
            // ---- backward: G tile, written back over T
            uint32_t gq[32];
            if (MODE == TC_BWD_DE) {
              const bool special = (c0 + 32 > p.n_str) || (static_cast<unsigned>(jg - c0) < 32u);
              const bool any_special = __any_sync(0xffffffffu, special);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float dot = __uint_as_float(v[i]);
                float pr = ex2(fmaf(dot, w2, b2) - lse2);        // softmax prob (0 for padded rows)
                if (any_special) {
                  if (c0 + i >= p.n_str) pr = 0.f;
                  if (c0 + i == jg) {
                    // diagonal: p_jj - 1 = -q (saved by the forward), uses cos_diag, contributes to
                    // dw but not to the contraction
                    dw_acc = fmaf(-qd, cd + eps, dw_acc);
                    pr = 0.f;
                  }
                }
                dw_acc = fmaf(pr, dot + eps, dw_acc);
                gq[i] = __float_as_uint(round_tf32(wg * pr));
              }
            } else {
              const float4* ls4 = reinterpret_cast<const float4*>(&tail->lse_s[buf][ch * 32]);
              const bool special = (c0 < dhi) && (c0 + 32 > dlo);
              const bool any_special = __any_sync(0xffffffffu, special);
#pragma unroll
              for (int i4 = 0; i4 < 8; ++i4) {
                const float4 l4 = ls4[i4];
                const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const int i = i4 * 4 + q;
                  float pr = ex2(fmaf(__uint_as_float(v[i]), w2, b2) - ls[q]);   // lse = +inf past the end
                  if (any_special && c0 + i >= dlo && c0 + i < dhi) pr = 0.f;    // own speaker's rows
                  gq[i] = __float_as_uint(round_tf32(wg * pr));
                }
              }
            }
            tmem_st32(t_addr + ch * 32, gq);
          }
        }
        tr.mark();   // tile consumed
        if (MODE == TC_FWD) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(BAR_S_EMPTY + buf));
        } else {
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(BAR_G_FULL + buf));
        }
      }

      // ------------------------------------------------------------ segment flush
      const bool full = (s0 == 0 && s1 == p.ST);
      if (MODE == TC_FWD) {
        // fold the two column halves of every row (upper half hands its state over through smem)
        float2 mine;
        if (VARIANT == GE2E_SOFTMAX) mine = make_float2(m2, lsum);
        else mine = make_float2(best, __int_as_float(bestk));
        if (half == 1) tail->xch[trow] = mine;
        named_bar_sync(1, kEpiThreads);
        const float2 theirs = tail->xch[trow];
        named_bar_sync(1, kEpiThreads);    // xch may be rewritten by the next segment from here on
        if (half == 0 && tile_valid) {     // tile_valid is CTA-uniform: barrier 2 below stays consistent
          auto fold = [&](float2 q) {
            if (VARIANT == GE2E_SOFTMAX) {
              const float mn = fmaxf(m2, q.x);
              lsum = lsum * ex2(m2 - mn) + q.y * ex2(q.x - mn);
              m2 = mn;
            } else {
              const int qk = __float_as_int(q.y);
              if (q.x > best || (q.x == best && qk < bestk)) { best = q.x; bestk = qk; }
            }
          };
          fold(theirs);
          // publish the partial row state, last finisher of this owner tile merges
          const int first_cl = cluster_of_pair(static_cast<long long>(og) * p.ST, p.GP, NC);
          bool last = full;
          if (!full) {
            float2 part;
            if (VARIANT == GE2E_SOFTMAX) part = make_float2(m2, lsum);
            else part = make_float2(best, __int_as_float(bestk));
            p.seg_part[(static_cast<size_t>(ot) * p.maxseg + (cl - first_cl)) * kTile + trow] = part;
            __threadfence();
            named_bar_sync(2, kTile);
            if (trow == 0) {
              const int done = atomicAdd(p.seg_done + ot, s1 - s0) + (s1 - s0);
              tail->flag = (done == p.ST);
            }
            named_bar_sync(2, kTile);
            last = tail->flag != 0;
            if (last) {
              __threadfence();
              const int last_cl = cluster_of_pair(static_cast<long long>(og) * p.ST + p.ST - 1, p.GP, NC);
              const int nseg = last_cl - first_cl + 1;
              m2 = xd2; lsum = 0.f; best = -INFINITY; bestk = INT_MAX;
              for (int sgi = 0; sgi < nseg; ++sgi)
                fold(__ldcg(&p.seg_part[(static_cast<size_t>(ot) * p.maxseg + sgi) * kTile + trow]));
            }
          }
          if (last && ovalid) {
            const float Sd = fmaf(w, cd + eps, b);
            float per, stat, aux = 0.f;
            int ks = -1;
            if (VARIANT == GE2E_SOFTMAX) {
              close_softmax_row(m2 * kLn2, lsum, Sd, eps, stat, aux, per);   // s3:120-121
            } else {
              per = 1.f - 1.f / (1.f + expf(-Sd));
              stat = best;
              if (bestk != INT_MAX) { ks = bestk; per += 1.f / (1.f + expf(-best)); }
            }
            p.row_stat_out[orow] = stat;
            if (p.row_aux_out != nullptr) p.row_aux_out[orow] = aux;
            if (VARIANT == GE2E_CONTRAST && p.kstar_out != nullptr) p.kstar_out[orow] = ks;
            if (p.per_row_out != nullptr) p.per_row_out[orow] = per;
            loss_acc += per;
          }
        }
      } else {
        // drain the accumulator [128 x D] of this segment (columns split between the two halves)
        if (MODE == TC_BWD_DE && ovalid && s0 == 0 && half == 0) db_acc -= g * eps * ex2(-lse2);   // item 12
        mbar_wait(bar(BAR_ACC_FULL), sg & 1);
        tc_fence_after();
        float* out = p.acc_out + static_cast<size_t>(orow) * p.D;
        const int ch0 = half ? kslabs / 2 : 0, ch1 = half ? kslabs : kslabs / 2;
        for (int ch = ch0; ch < ch1; ++ch) {
          uint32_t v[32];
          tmem_ld32(tmem + lane_addr + 2 * kTile + ch * 32, v);
          tmem_ld_wait();
          if (ovalid) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 o = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                                     __uint_as_float(v[i + 3]));
              float4* dst = reinterpret_cast<float4*>(out + ch * 32 + i);
              if (full) *dst = o; else atomicAdd(dst, o);   // partial segments add into the zeroed output
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_ACC_EMPTY));
      }
      gp += s1 - s0;
      tr.mark();   // segment flushed
    }

    // ------------------------------------------------------------ scalar reductions (once per CTA)
    if (MODE == TC_FWD) {
      loss_acc = warp_sum(loss_acc);
      if (lane == 0) tail->red[ew] = loss_acc;
      named_bar_sync(1, kEpiThreads);
      if (ew == 0 && lane == 0) {
        float t = 0.f;
        for (int i = 0; i < kEpiWarps; ++i) t += tail->red[i];
        atomicAdd(p.loss_accum, t);
      }
    } else if (MODE == TC_BWD_DE) {
      dw_acc = warp_sum(dw_acc) * g;
      db_acc = warp_sum(db_acc);
      if (lane == 0) { tail->red[ew] = dw_acc; tail->red[kEpiWarps + ew] = db_acc; }
      named_bar_sync(1, kEpiThreads);
      if (ew == 0 && lane == 0) {
        float tw = 0.f, tb = 0.f;
        for (int i = 0; i < kEpiWarps; ++i) { tw += tail->red[i]; tb += tail->red[kEpiWarps + i]; }
        atomicAdd(p.dwdb + 0, tw);
        atomicAdd(p.dwdb + 1, tb);
      }
    }
  }

  // ------------------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();     // no CTA leaves while a peer may still multicast into it / arrive on it
  if (warp == 2) tmem_dealloc<kTmemCols>(tmem);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// 2-D map over X[rows, D] fp32: box = [box_rows][32 cols], 128-byte swizzle (K-major slabs).
int make_map_2d(CUtensorMap* m, const float* base, int rows, int D, int box_rows) {
  auto enc = get_encode();
  if (enc == nullptr) return GE2E_ERR_LAUNCH;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(D) * 4};
  cuuint32_t box[2] = {kSlabCols, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GE2E_OK : GE2E_ERR_LAUNCH;
}

// 3-D map over the same X viewed as [D/32][rows][32]: box = [box_slabs][16 rows][32 cols]
// (MN-major operand chunks for MMA2: 16 k-rows x D columns per ring stage).  32-bit MN-major
// operands must use the 32-byte-atom flavour of the 128B swizzle (UMMA SWIZZLE_128B_BASE32B).
int make_map_3d(CUtensorMap* m, const float* base, int rows, int D, int box_slabs) {
  auto enc = get_encode();
  if (enc == nullptr) return GE2E_ERR_LAUNCH;
  cuuint64_t dims[3] = {kSlabCols, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(D / kSlabCols)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(D) * 4, kSlabCols * 4};
  cuuint32_t box[3] = {kSlabCols, kMma2Rows, static_cast<cuuint32_t>(box_slabs)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GE2E_OK : GE2E_ERR_LAUNCH;
}

unsigned long long* g_trace = nullptr;   // set through tc_set_trace (debug only)
int g_trace_mode = -1;                   // -1: every kernel, else only TC_FWD / TC_BWD_DE / TC_BWD_DC

constexpr size_t kSmemBytes = 1024 + kMaxSlabs * kSlabBytes + kStages * kStageBytes + sizeof(SharedTail);

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Cluster size: GE2E_TC_CLUSTER overrides (1, 2 or 4); it must divide the number of 32-column
// chunks of D so that an MMA2 stage splits evenly between the CTAs.
int pick_cluster(int D) {
  static int env = -1;
  if (env < 0) {
    const char* s = getenv("GE2E_TC_CLUSTER");
    env = s ? atoi(s) : 0;
  }
  int C = (env == 1 || env == 2 || env == 4) ? env : 1;
  while (C > 1 && (D / kSlabCols) % C != 0) C >>= 1;
  return C;
}

// How many clusters of size C can be co-resident (1 CTA per SM: the kernel needs ~225 KB smem).
// Cluster size 4 strands a few SMs per GPC; measured by the occupancy API when available.
template <int MODE, int VARIANT>
int max_clusters(int C) {
  static int cache[kMaxCluster + 1] = {0, 0, 0, 0, 0};
  if (cache[C] == 0) {
    int n = 0;
    if (C > 1) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(sm_count() / C * C);
      cfg.blockDim = dim3(kThreadsTc);
      cfg.dynamicSmemBytes = kSmemBytes;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaFuncSetAttribute(tc_strip_kernel<MODE, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)kSmemBytes);
      if (cudaOccupancyMaxActiveClusters(&n, tc_strip_kernel<MODE, VARIANT>, &cfg) != cudaSuccess) n = 0;
      (void)cudaGetLastError();
    }
    if (n <= 0) n = sm_count() / C;
    cache[C] = n;
  }
  return cache[C];
}

struct Layout {
  int OT, ST, C, OG, NC, maxseg;
  long long GP;
  size_t done_bytes, part_bytes;
  bool whole;       // every cluster's range is a whole number of owner groups: no partial flushes
};

Layout make_layout(int n_own, int n_str, int D, int max_cl_of_C(int)) {
  Layout L{};
  L.OT = (n_own + kTile - 1) / kTile;
  L.ST = (n_str + kTile - 1) / kTile;
  L.C = pick_cluster(D);
  L.OG = (L.OT + L.C - 1) / L.C;
  L.GP = static_cast<long long>(L.OG) * L.ST;
  L.NC = static_cast<int>(std::min<long long>(max_cl_of_C(L.C), L.GP));
  const long long per = L.GP / L.NC;                 // >= 1
  L.maxseg = static_cast<int>(std::min<long long>(L.ST, L.ST / per + 2));
  L.done_bytes = (static_cast<size_t>(L.OT) * sizeof(int) + 255) & ~static_cast<size_t>(255);
  L.part_bytes = static_cast<size_t>(L.OT) * L.maxseg * kTile * sizeof(float2);
  L.whole = (L.GP % L.NC == 0) && (per % L.ST == 0);
  return L;
}

template <int MODE, int VARIANT>
int launch_tc(const CUtensorMap& own, const CUtensorMap& s2, const CUtensorMap& s3, const TcParams& p, int NC,
              cudaStream_t st) {
  auto kern = tc_strip_kernel<MODE, VARIANT>;
  GE2E_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  TcParams q = p;
  q.trace = (g_trace_mode < 0 || g_trace_mode == MODE) ? g_trace : nullptr;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("GE2E_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    q.dbg = dbg;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(NC * p.C);
  cfg.blockDim = dim3(kThreadsTc);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  GE2E_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, own, s2, s3, q));
  GE2E_LAUNCHED();
  return GE2E_OK;
}

void fill_common(TcParams& p, const RowsArgs& a, const Layout& L) {
  p.D = a.D; p.kslabs = a.D / kSlabCols; p.M = a.M; p.spk_offset = a.spk_offset;
  p.OT = L.OT; p.ST = L.ST; p.C = L.C; p.OG = L.OG; p.GP = L.GP;
  p.cos_diag = a.cos_diag; p.w = a.w; p.b = a.b; p.eps = a.eps;
}

}  // namespace

void tc_set_trace(unsigned long long* device_buf, int mode) { g_trace = device_buf; g_trace_mode = mode; }

bool tc_supported(int n_local, int n_total, int M, int D, int variant) {
  (void)variant;
  if (D % kSlabCols != 0 || D < kSlabCols || D > kMaxSlabs * kSlabCols) return false;
  if (n_total < 256 || static_cast<long long>(n_local) * M < 256) return false;
  return get_encode() != nullptr;
}

size_t tc_workspace_bytes(int n_local, int n_total, int M, int D, int variant) {
  const Layout L = (variant == GE2E_SOFTMAX) ? make_layout(n_local * M, n_total, D, max_clusters<TC_FWD, GE2E_SOFTMAX>)
                                             : make_layout(n_local * M, n_total, D, max_clusters<TC_FWD, GE2E_CONTRAST>);
  return L.done_bytes + L.part_bytes;
}

int tc_fwd_rows(const RowsArgs& a, float* row_stat, int32_t* row_kstar, float* row_aux, float* loss_accum,
                float* per_row_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int U = a.n_local * a.M;
  const Layout L = (a.variant == GE2E_SOFTMAX) ? make_layout(U, a.n_total, a.D, max_clusters<TC_FWD, GE2E_SOFTMAX>)
                                               : make_layout(U, a.n_total, a.D, max_clusters<TC_FWD, GE2E_CONTRAST>);
  if (ws_bytes < L.done_bytes + L.part_bytes) return GE2E_ERR_WORKSPACE;
  CUtensorMap tmE, tmC;
  int rc = make_map_2d(&tmE, a.e_hat, U, a.D, kTile);
  if (rc != GE2E_OK) return rc;
  if ((rc = make_map_2d(&tmC, a.c_hat_all, a.n_total, a.D, kTile / L.C)) != GE2E_OK) return rc;
  TcParams p{};
  fill_common(p, a, L);
  p.n_own = U; p.n_str = a.n_total;
  p.row_stat_out = row_stat; p.kstar_out = row_kstar; p.row_aux_out = row_aux; p.loss_accum = loss_accum;
  p.per_row_out = per_row_out;
  p.seg_done = static_cast<int*>(ws);
  p.seg_part = reinterpret_cast<float2*>(static_cast<uint8_t*>(ws) + L.done_bytes);
  p.maxseg = L.maxseg;
  if (!L.whole) GE2E_CUDA_TRY(cudaMemsetAsync(ws, 0, L.done_bytes, st));
  if (a.variant == GE2E_SOFTMAX) return launch_tc<TC_FWD, GE2E_SOFTMAX>(tmE, tmC, tmC, p, L.NC, st);
  return launch_tc<TC_FWD, GE2E_CONTRAST>(tmE, tmC, tmC, p, L.NC, st);
}

int tc_bwd_rows(const RowsArgs& a, const float* row_stat, const int32_t* row_kstar, const float* row_aux,
                const float* grad_out, float* dE_hat, float* dC_hat_partial, float* dwdb_accum, void* ws,
                size_t ws_bytes, cudaStream_t st) {
  (void)row_kstar; (void)ws; (void)ws_bytes;
  const int U = a.n_local * a.M;
  const size_t dc_elems = static_cast<size_t>(a.n_total) * a.D;
  if (dwdb_accum == dC_hat_partial + dc_elems) {
    GE2E_CUDA_TRY(cudaMemsetAsync(dC_hat_partial, 0, (dc_elems + 2) * sizeof(float), st));
  } else {
    GE2E_CUDA_TRY(cudaMemsetAsync(dC_hat_partial, 0, dc_elems * sizeof(float), st));
    GE2E_CUDA_TRY(cudaMemsetAsync(dwdb_accum, 0, 2 * sizeof(float), st));
  }
  const Layout Le = make_layout(U, a.n_total, a.D, max_clusters<TC_BWD_DE, GE2E_SOFTMAX>);
  const Layout Lc = make_layout(a.n_total, U, a.D, max_clusters<TC_BWD_DC, GE2E_SOFTMAX>);
  const int slabs = a.D / kSlabCols;
  CUtensorMap tmE_own, tmC_own, tmE_s2, tmC_s2, tmE_s3, tmC_s3;
  int rc;
  if ((rc = make_map_2d(&tmE_own, a.e_hat, U, a.D, kTile)) != GE2E_OK) return rc;
  if ((rc = make_map_2d(&tmC_own, a.c_hat_all, a.n_total, a.D, kTile)) != GE2E_OK) return rc;
  if ((rc = make_map_2d(&tmC_s2, a.c_hat_all, a.n_total, a.D, kTile / Le.C)) != GE2E_OK) return rc;
  if ((rc = make_map_3d(&tmC_s3, a.c_hat_all, a.n_total, a.D, slabs / Le.C)) != GE2E_OK) return rc;
  if ((rc = make_map_2d(&tmE_s2, a.e_hat, U, a.D, kTile / Lc.C)) != GE2E_OK) return rc;
  if ((rc = make_map_3d(&tmE_s3, a.e_hat, U, a.D, slabs / Lc.C)) != GE2E_OK) return rc;

  // dE_hat = (wG) C_hat: owner = utterance tiles; partial owner-group ranges add into a zeroed output
  TcParams p{};
  fill_common(p, a, Le);
  p.row_stat = row_stat; p.row_aux = row_aux; p.grad_out = grad_out;
  p.n_own = U; p.n_str = a.n_total;
  p.acc_out = dE_hat; p.dwdb = dwdb_accum;
  if (!Le.whole) GE2E_CUDA_TRY(cudaMemsetAsync(dE_hat, 0, static_cast<size_t>(U) * a.D * sizeof(float), st));
  rc = launch_tc<TC_BWD_DE, GE2E_SOFTMAX>(tmE_own, tmC_s2, tmC_s3, p, Le.NC, st);
  if (rc != GE2E_OK) return rc;

  // dC_hat = (wG)^T E_hat: owner = centroid tiles, the utterance range is cut stream-K style
  fill_common(p, a, Lc);
  p.n_own = a.n_total; p.n_str = U;
  p.acc_out = dC_hat_partial; p.dwdb = nullptr;
  return launch_tc<TC_BWD_DC, GE2E_SOFTMAX>(tmC_own, tmE_s2, tmE_s3, p, Lc.NC, st);
}

}  // namespace ge2e
