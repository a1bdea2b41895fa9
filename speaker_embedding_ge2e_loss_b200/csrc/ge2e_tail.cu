// Model tail feeding the loss: last-frame select -> Linear(H -> D) -> L2
// normalise, reference embedding_model_GE2E/s2_model_GE2E_loss_speach_embed.py:28-34
//     x = x[:, x.size(1) - 1];  x = self.projection(x);  x = x / torch.norm(x, dim=1).unsqueeze(1)
//
// Forward, one kernel: Y[128 rows x D] = X_tile[128 x H] * W[D x H]^T on tcgen05 (kind::tf32, fp32
// accumulation in tensor memory), K streamed in 32-column slabs through a 4-stage TMA ring (X and W are
// both K-major, 128B swizzle -- the same operand layout as the loss's E * C^T), and an epilogue that is
// the loss's prologue: + bias, row sum of squares, scale to unit length, rows staged through shared
// memory and written with TMA.  The last-frame select is the row stride of the X tensor map (no copy).
// Y never reaches HBM; what is kept for the backward is E and 1 / ||y||.
//
//   warp 0      TMA producer           warp 1      TMEM allocation + MMA issue (one elected lane)
//   warps 2-5   epilogue, one thread per row of the tile (TMEM lane quarter = warp % 4)
//
// Backward of the normalisation (+ bias gradient) is a SIMT row kernel; the two gradient GEMMs of the Linear
// layer (dX = dY W, dW = dY^T X -- what autograd runs under s2:31) are one more tcgen05 kernel, tail_gemm_kernel:
// C[128 x 256 tile] (+)= A B with B (and for dW also A) taken MN-major straight from the row-major tensors, K
// streamed through a TMA ring, K split over CTAs for dW (partial tiles added at the L2 with TMA reduce-add).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>

#include "ge2e_common.cuh"
#include "ge2e_tc_ptx.cuh"

namespace ge2e {

namespace {

using namespace ptx;

constexpr int kTailRows = 128;               // rows of X per CTA
constexpr int kTailKc = 32;                  // K columns per ring stage (one 128-byte swizzled slab)
constexpr int kTailStages = 4;               // (6 in the CTA-pair instantiation, same ring bytes)
constexpr int kTailMaxStages = 6;
constexpr int kTailABytes = kTailRows * 128;             // 16 KB
constexpr int kTailBBytes = 256 * 128;                   // 32 KB (D <= 256 rows of W)
constexpr int kTailStageBytes = kTailABytes + kTailBBytes;
constexpr int kTailThreads = 192;
constexpr int kTailEpiThreads = 128;
constexpr int kTailSmemBytes = kTailStages * kTailStageBytes + 1024 /*alignment*/ + 2048 /*tail struct*/;

struct TailShared {
  unsigned long long full[kTailMaxStages], empty[kTailMaxStages], acc_full;
  uint32_t tmem_base;
  float bias[256];
};

struct TailParams {
  int U, H, D;
  const float* bias;     // nullable
  float* inv_norm;       // [U]
};

// CG = 2: two CTAs of a cluster own two consecutive 128-row tiles and run every MMA as a pair (cta_group::2, UMMA
// M = 256); each fetches its own X slab and HALF of the W slab (rows [rank D/2, +D/2)).  Halving the W traffic per SM
// alone changed nothing (24.6 us at U = 10240 either way: the K loop is a latency chain on the strided X slabs, not
// an L2-feed limit); what the smaller stage buys is a ring of 6 stages instead of 4 in the same 192 KB (22.5 us).
// The leader CTA issues the MMAs; the barriers it waits on live in the leader and are signalled by both CTAs (as in
// ge2e_tc.cu).
template <int CG>
__global__ void __launch_bounds__(kTailThreads, 1)
embed_tail_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                      const __grid_constant__ CUtensorMap tm_e, const TailParams p) {
  extern __shared__ uint8_t tail_raw[];
  const uint32_t raw = smem_u32(tail_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;                       // 1024-byte aligned (swizzle atom)
  TailShared* sh = reinterpret_cast<TailShared*>(tail_raw + (ring - raw) + kTailStages * kTailStageBytes);   // (same offset in both instantiations)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kTailRows;
  const int nchunks = (p.H + kTailKc - 1) / kTailKc;
  const int D = p.D;
  // ring: 4 stages of [X 16 KB | W 32 KB] alone, 6 stages of [X 16 KB | W half 16 KB] as a pair (same 192 KB): the
  // K loop is a latency chain (a stage comes back ~2.8 us after it was requested from HBM), depth is what it needs
  constexpr int kStages = (CG == 2) ? 6 : kTailStages;
  constexpr int kStageBytes = (CG == 2) ? kTailABytes + kTailBBytes / 2 : kTailStageBytes;
  const int cr = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = cr == 0;
  const int Dw = D / CG;                                              // rows of W this CTA fetches per stage
  auto lbar = [&](const unsigned long long* b) { return (CG == 2) ? mapa(smem_u32(b), 0) : smem_u32(b); };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_x);
    prefetch_tmap(&tm_w);
    prefetch_tmap(&tm_e);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&sh->full[s]), 1);
      mbar_init(smem_u32(&sh->empty[s]), 1);
    }
    mbar_init(smem_u32(&sh->acc_full), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CG == 1) tmem_alloc<256>(smem_u32(&sh->tmem_base)); else tmem_alloc_2cta<256>(smem_u32(&sh->tmem_base));
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();      // the peer's barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  pdl_wait();          // X comes from the stream predecessor (the LSTM)
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      const uint32_t bytes = static_cast<uint32_t>(kTailABytes + Dw * 128);
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % kStages;
        mbar_wait(smem_u32(&sh->empty[s]), ((c / kStages) & 1) ^ 1);
        const uint32_t fb = lbar(&sh->full[s]);
        if (leader) mbar_expect_tx(smem_u32(&sh->full[s]), bytes * CG);
        if (CG == 1) {
          tma_load_2d(ring + s * kStageBytes, &tm_x, c * kTailKc, row0, fb);
          tma_load_2d(ring + s * kStageBytes + kTailABytes, &tm_w, c * kTailKc, 0, fb);
        } else {
          tma_load_2d_2cta(ring + s * kStageBytes, &tm_x, c * kTailKc, row0, fb);
          tma_load_2d_2cta(ring + s * kStageBytes + kTailABytes, &tm_w, c * kTailKc, cr * Dw, fb);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issue (leader CTA of a pair)
    if (leader) {
      const uint32_t idesc = idesc_tf32(kTailRows * CG, D, 0, 0);
      const uint64_t dk = smem_desc(0, 16, 1024, kLayoutSw128);          // K-major, 8-row groups 1024 B apart
      bool ready = false;
      uint32_t lead_lane = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % kStages;
        const uint32_t ph = (c / kStages) & 1;
        if (!ready) mbar_wait(smem_u32(&sh->full[s]), ph);
        tc_fence_after();
        const int sn = (s + 1 == kStages) ? 0 : s + 1;
        const uint32_t pn = (s + 1 == kStages) ? (ph ^ 1) : ph;
        uint32_t probe = 0;
        if (elect_leader(lead_lane)) {
          const uint64_t da = dk | ((ring + s * kStageBytes) >> 4);
          const uint64_t db = dk | ((ring + s * kStageBytes + kTailABytes) >> 4);
          probe = umma_stage_ss<CG, 1>(tmem, da, db, 0, 0, idesc, c != 0, smem_u32(&sh->empty[s]),
                                       smem_u32(&sh->full[sn]), pn);
        }
        __syncwarp();
        ready = (c + 1 < nchunks) && __shfl_sync(0xffffffffu, probe, lead_lane) != 0;
      }
      if (elect_one()) {
        if (CG == 1) umma_commit(smem_u32(&sh->acc_full)); else umma_commit_2cta(smem_u32(&sh->acc_full), 3);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue: one thread per row
    const int q = warp & 3;                           // TMEM lane quarter this warp may read
    const int trow = q * 32 + lane;
    const int et = threadIdx.x - 64;                  // 0..127
    for (int i = et; i < 256; i += kTailEpiThreads) sh->bias[i] = (p.bias != nullptr && i < D) ? p.bias[i] : 0.f;
    named_bar_sync(1, kTailEpiThreads);
    mbar_wait(smem_u32(&sh->acc_full), 0);
    tc_fence_after();
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int nslab = D / 32;
    float ss = 0.f;
    for (int ch = 0; ch < nslab; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem + lane_addr + ch * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float y = __uint_as_float(v[i]) + sh->bias[ch * 32 + i];
        ss = fmaf(y, y, ss);
      }
    }
    // x / torch.norm(x, dim=1): no epsilon in the reference (a zero row gives NaN there and here)
    const float inv = 1.0f / sqrtf(ss);
    if (row0 + trow < p.U && p.inv_norm != nullptr) p.inv_norm[row0 + trow] = inv;
    // every MMA has completed (acc_full), so the ring is free: stage E in the swizzled slab layout
    for (int ch = 0; ch < nslab; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem + lane_addr + ch * 32, v);
      tmem_ld_wait();
      const uint32_t row_smem = ring + ch * kTailABytes + trow * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = (__uint_as_float(v[4 * c + i]) + sh->bias[ch * 32 + 4 * c + i]) * inv;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_smem + ((c ^ (trow & 7)) << 4)),
                     "r"(__float_as_uint(o[0])), "r"(__float_as_uint(o[1])), "r"(__float_as_uint(o[2])),
                     "r"(__float_as_uint(o[3]))
                     : "memory");
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    named_bar_sync(1, kTailEpiThreads);
    if (et == 0) {
      for (int ch = 0; ch < nslab; ++ch) tma_store_2d(&tm_e, ch * 32, row0, ring + ch * kTailABytes);   // clips rows >= U
      tma_store_commit();
      tma_store_wait_read();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();      // no CTA leaves while its peer may still arrive on it / read its smem
  if (warp == 1) {
    if (CG == 1) tmem_dealloc<256>(tmem); else tmem_dealloc_2cta<256>(tmem);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 tail_encode() {
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

// [rows, cols] fp32 with a row stride (floats); box = [box_rows][32 cols], 128-byte swizzle
int tail_map(CUtensorMap* m, const float* base, long long rows, int cols, long long row_stride, int box_rows) {
  auto enc = tail_encode();
  if (enc == nullptr) return GE2E_ERR_LAUNCH;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(row_stride) * 4};
  cuuint32_t box[2] = {kTailKc, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GE2E_OK : GE2E_ERR_LAUNCH;
}

// ---- backward of  e = y / ||y||  (+ the bias gradient): one warp per row ------------------------
//   dY = (dE - e (e . dE)) / ||y||,   dbias += column sums of dY
constexpr int kTailBwdWarps = 8;
__global__ void __launch_bounds__(kTailBwdWarps * 32)
embed_tail_bwd_rows_kernel(const float* __restrict__ dE, const float* __restrict__ E,
                           const float* __restrict__ inv_norm, int U, int D, float* __restrict__ dY,
                           float* __restrict__ dbias) {
  __shared__ float colsum[256];
  pdl_wait();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) colsum[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_lane = D / 32;                      // D in {64, 128, 256}: 2, 4 or 8 columns per lane
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int row = blockIdx.x * kTailBwdWarps + warp; row < U; row += gridDim.x * kTailBwdWarps) {
    const float* g = dE + (size_t)row * D;
    const float* e = E + (size_t)row * D;
    float gv[8], ev[8];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per_lane) {
        gv[i] = __ldg(g + lane + 32 * i);
        ev[i] = __ldg(e + lane + 32 * i);
        dot = fmaf(gv[i], ev[i], dot);
      }
    dot = warp_sum(dot);
    const float inv = __ldg(inv_norm + row);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per_lane) {
        const float d = (gv[i] - ev[i] * dot) * inv;
        dY[(size_t)row * D + lane + 32 * i] = d;
        acc[i] += d;
      }
  }
  if (dbias != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per_lane) atomicAdd(&colsum[lane + 32 * i], acc[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) atomicAdd(dbias + i, colsum[i]);
  }
}

// ---- gradient GEMMs of the Linear layer ---------------------------------------------------------
//   dX[U x H] = dY[U x D] W[D x H]          A = dY rows, K-major (K = D);   B = W,  K x N row-major = MN-major
//   dW[D x H] = dY^T[D x U] X[U x H]        A = dY,      K x M row-major = MN-major (K = U);  B = X, MN-major
// One CTA per (128-row M tile, 256-column N tile, K range); TF32 operands, fp32 accumulation in 256 TMEM
// columns; ring stage = 32 k: A 16 KB + B 32 KB.  MN-major TF32 operands use the 32-byte-atom 128B swizzle
// (ge2e_tc_ptx.cuh): chunks of 32 columns 4096 B apart (LBO), groups of 4 k-rows 512 B apart (SBO), one MMA
// (k = 8) consumes 1024 B.  warp 0 = TMA, warp 1 = MMA + TMEM, warps 2-5 = epilogue (a row each).
constexpr int kGemmStages = 4;
constexpr int kGemmK = 32;                        // k per ring stage
constexpr int kGemmN = 256;
constexpr int kGemmABytes = 128 * kGemmK * 4;     // 16 KB
constexpr int kGemmBBytes = kGemmN * kGemmK * 4;  // 32 KB
constexpr int kGemmStageBytes = kGemmABytes + kGemmBBytes;
constexpr int kGemmSmemBytes = kGemmStages * kGemmStageBytes + 1024 + 256;

struct GemmShared {
  unsigned long long full[kGemmStages], empty[kGemmStages], acc_full;
  uint32_t tmem_base;
};

struct GemmParams {
  int k_total;       // K
  int k_per_cta;     // multiple of kGemmK; blockIdx.z covers [z * k_per_cta, min(K, (z + 1) * k_per_cta))
  int a_mn;          // A is MN-major (K x M row-major)
  int reduce;        // add into C (K split) instead of storing
  int n_slabs;       // 32-column slabs of C in this launch's N tiles that exist at all (for the last tile: clipped by TMA)
};

__global__ void __launch_bounds__(kTailThreads, 1)
tail_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_c, const GemmParams p) {
  extern __shared__ uint8_t gemm_raw[];
  const uint32_t raw = smem_u32(gemm_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  GemmShared* sh = reinterpret_cast<GemmShared*>(gemm_raw + (ring - raw) + kGemmStages * kGemmStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kGemmN, m0 = blockIdx.y * 128;
  const int k0 = blockIdx.z * p.k_per_cta;
  const int k1 = min(p.k_total, k0 + p.k_per_cta);
  const int nst = (k1 - k0 + kGemmK - 1) / kGemmK;      // >= 1 by construction of the grid

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_a); prefetch_tmap(&tm_b); prefetch_tmap(&tm_c);
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(smem_u32(&sh->full[s]), 1);
      mbar_init(smem_u32(&sh->empty[s]), 1);
    }
    mbar_init(smem_u32(&sh->acc_full), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(smem_u32(&sh->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;

  if (warp == 0) {
    if (elect_one()) {
      for (int c = 0; c < nst; ++c) {
        const int s = c % kGemmStages;
        mbar_wait(smem_u32(&sh->empty[s]), ((c / kGemmStages) & 1) ^ 1);
        const uint32_t fb = smem_u32(&sh->full[s]);
        mbar_expect_tx(fb, kGemmStageBytes);          // out-of-range parts of a box are zero-filled and counted
        const uint32_t sa = ring + s * kGemmStageBytes, sb = sa + kGemmABytes;
        const int k = k0 + c * kGemmK;
        if (p.a_mn) tma_load_3d(sa, &tm_a, 0, k, m0 / 32, fb);       // [4 chunks of 32 m][32 k][32]
        else tma_load_2d(sa, &tm_a, k, m0, fb);                      // [128 m][32 k]
        tma_load_3d(sb, &tm_b, 0, k, n0 / 32, fb);                   // [8 chunks of 32 n][32 k][32]
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t idesc = idesc_tf32(128, kGemmN, p.a_mn, 1);
    const uint64_t dk = smem_desc(0, 16, 1024, kLayoutSw128);
    const uint64_t dmn = smem_desc(0, kGemmK * 128, 512, kLayoutSw128Base32);
    const uint64_t da0 = p.a_mn ? dmn : dk;
    const uint32_t a_step = p.a_mn ? 64u : 2u;          // descriptor units (16 B) per k = 8
    for (int c = 0; c < nst; ++c) {
      const int s = c % kGemmStages;
      mbar_wait(smem_u32(&sh->full[s]), (c / kGemmStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = ring + s * kGemmStageBytes, sb = sa + kGemmABytes;
        uint64_t da = da0 | (sa >> 4), db = dmn | (sb >> 4);
#pragma unroll
        for (int j = 0; j < kGemmK / 8; ++j) {
          umma_tf32_ss(tmem, da, db, idesc, (c | j) != 0);
          da += a_step; db += 64;
        }
        umma_commit(smem_u32(&sh->empty[s]));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&sh->acc_full));
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int trow = q * 32 + lane;
    const int et = threadIdx.x - 64;
    mbar_wait(smem_u32(&sh->acc_full), 0);
    tc_fence_after();
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    // every MMA has completed, so every load has landed and been consumed: the ring is free to stage C
    for (int ch = 0; ch < p.n_slabs; ++ch) {
      uint32_t v[32];
      tmem_ld32(tmem + lane_addr + ch * 32, v);
      tmem_ld_wait();
      const uint32_t row_smem = ring + ch * kTailABytes + trow * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_smem + ((c ^ (trow & 7)) << 4)),
                     "r"(v[4 * c]), "r"(v[4 * c + 1]), "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                     : "memory");
    }
    tc_fence_before();
    fence_proxy_async_smem();
    named_bar_sync(1, kTailEpiThreads);
    if (et == 0) {
      for (int ch = 0; ch < p.n_slabs; ++ch) {          // rows / columns past the end of C are clipped
        if (p.reduce) tma_reduce_add_2d(&tm_c, n0 + ch * 32, m0, ring + ch * kTailABytes);
        else tma_store_2d(&tm_c, n0 + ch * 32, m0, ring + ch * kTailABytes);
      }
      tma_store_commit();
      tma_store_wait_read();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

// [rows, cols] fp32 with a row stride, viewed as [cols / 32][rows][32]: box = [box_chunks][32 rows][32 cols],
// 32-byte-atom 128B swizzle -- an MN-major TF32 operand block of 32 k-rows (cols % 32 == 0)
int tail_map_mn(CUtensorMap* m, const float* base, long long rows, int cols, long long row_stride, int box_chunks) {
  auto enc = tail_encode();
  if (enc == nullptr) return GE2E_ERR_LAUNCH;
  cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(cols / 32)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(row_stride) * 4, 128};
  cuuint32_t box[3] = {32, kGemmK, static_cast<cuuint32_t>(box_chunks)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GE2E_OK : GE2E_ERR_LAUNCH;
}

int launch_tail_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p, int m_tiles,
                     int n_tiles, int k_splits, cudaStream_t st) {
  static thread_local int attr_dev = -1;
  int dev = 0;
  GE2E_CUDA_TRY(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    GE2E_CUDA_TRY(cudaFuncSetAttribute(tail_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    attr_dev = dev;
  }
  tail_gemm_kernel<<<dim3(n_tiles, m_tiles, k_splits), kTailThreads, kGemmSmemBytes, st>>>(ta, tb, tc, p);
  GE2E_LAUNCHED();
  return GE2E_OK;
}

}  // namespace

bool tail_bwd_gemms_supported(int U, int H, int D, long long x_row_stride, long long dx_row_stride) {
  return U > 0 && H >= 32 && H % 32 == 0 && (D == 64 || D == 128 || D == 256) && x_row_stride >= H &&
         x_row_stride % 4 == 0 && dx_row_stride >= H && dx_row_stride % 4 == 0;
}

int tail_bwd_gemms(const float* dY, const float* W, const float* X, long long x_row_stride, int U, int H, int D,
                   float* dX, long long dx_row_stride, float* dW, cudaStream_t st) {
  if (!tail_bwd_gemms_supported(U, H, D, x_row_stride, dx_row_stride)) return GE2E_ERR_UNSUPPORTED;
  const int n_tiles = (H + kGemmN - 1) / kGemmN;
  const int n_slabs = std::min(kGemmN, H) / 32;     // H < 256: fewer slabs; a clipped last tile stores nothing past H
  int rc;
  if (dX != nullptr) {
    if ((reinterpret_cast<uintptr_t>(dY) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(dX)) & 15)
      return GE2E_ERR_UNSUPPORTED;
    CUtensorMap ta, tb, tc;
    if ((rc = tail_map(&ta, dY, U, D, D, 128)) != GE2E_OK) return rc;                   // K-major A: [128 m][32 k]
    if ((rc = tail_map_mn(&tb, W, D, H, H, kGemmN / 32)) != GE2E_OK) return rc;         // W[k = d][n = h]
    if ((rc = tail_map(&tc, dX, U, H, dx_row_stride, 128)) != GE2E_OK) return rc;
    GemmParams p{D, D, 0, 0, n_slabs};
    if ((rc = launch_tail_gemm(ta, tb, tc, p, (U + 127) / 128, n_tiles, 1, st)) != GE2E_OK) return rc;
  }
  if (dW != nullptr) {
    if (X == nullptr) return GE2E_ERR_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(dY) | reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(dW)) & 15)
      return GE2E_ERR_UNSUPPORTED;
    CUtensorMap ta, tb, tc;
    if ((rc = tail_map_mn(&ta, dY, U, D, D, 4)) != GE2E_OK) return rc;                   // dY[k = u][m = d]
    if ((rc = tail_map_mn(&tb, X, U, H, x_row_stride, kGemmN / 32)) != GE2E_OK) return rc;   // X[k = u][n = h]
    if ((rc = tail_map(&tc, dW, D, H, H, 128)) != GE2E_OK) return rc;
    const int m_tiles = (D + 127) / 128;
    // K = U split so that the grid covers the SMs about once; a split is a whole number of ring stages
    const int stages = (U + kGemmK - 1) / kGemmK;
    int want = std::max(1, 148 / (m_tiles * n_tiles));
    int per = (stages + want - 1) / want;                       // stages per CTA
    const int splits = (stages + per - 1) / per;
    GemmParams p{U, per * kGemmK, 1, splits > 1 ? 1 : 0, n_slabs};
    if (splits > 1) GE2E_CUDA_TRY(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)D * H, st));
    if ((rc = launch_tail_gemm(ta, tb, tc, p, m_tiles, n_tiles, splits, st)) != GE2E_OK) return rc;
  }
  return GE2E_OK;
}

bool tail_supported(int U, int H, int D, long long x_row_stride) {
  return U > 0 && H > 0 && H % 4 == 0 && (D == 64 || D == 128 || D == 256) && x_row_stride >= H &&
         x_row_stride % 4 == 0;
}

int tail_fwd(const float* X, long long x_row_stride, const float* W, const float* bias, int U, int H, int D,
             float* E, float* inv_norm, cudaStream_t st) {
  if (!tail_supported(U, H, D, x_row_stride)) return GE2E_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(E)) & 15)
    return GE2E_ERR_UNSUPPORTED;
  CUtensorMap tm_x, tm_w, tm_e;
  int rc;
  if ((rc = tail_map(&tm_x, X, U, H, x_row_stride, kTailRows)) != GE2E_OK) return rc;
  const int tiles = (U + kTailRows - 1) / kTailRows;
  // CTA pairs from two tiles on (one W half per CTA); a single tile runs alone
  const int cg = tiles >= 2 ? 2 : 1;
  if ((rc = tail_map(&tm_w, W, D, H, H, D / cg)) != GE2E_OK) return rc;
  if ((rc = tail_map(&tm_e, E, U, D, D, kTailRows)) != GE2E_OK) return rc;
  static thread_local int attr_dev = -1;
  int dev = 0;
  GE2E_CUDA_TRY(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    GE2E_CUDA_TRY(cudaFuncSetAttribute(embed_tail_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kTailSmemBytes));
    GE2E_CUDA_TRY(cudaFuncSetAttribute(embed_tail_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kTailSmemBytes));
    attr_dev = dev;
  }
  TailParams p{U, H, D, bias, inv_norm};
  if (cg == 1) {
    launch_pdl(embed_tail_fwd_kernel<1>, dim3(tiles), dim3(kTailThreads), (size_t)kTailSmemBytes, st, true, tm_x, tm_w,
               tm_e, p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((tiles + 1) / 2 * 2); cfg.blockDim = dim3(kTailThreads);
    cfg.dynamicSmemBytes = kTailSmemBytes; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    GE2E_CUDA_TRY(cudaLaunchKernelEx(&cfg, embed_tail_fwd_kernel<2>, tm_x, tm_w, tm_e, p));
  }
  GE2E_LAUNCHED();
  return GE2E_OK;
}

int tail_bwd_rows(const float* dE, const float* E, const float* inv_norm, int U, int D, float* dY, float* dbias,
                  cudaStream_t st) {
  if (U <= 0 || !(D == 64 || D == 128 || D == 256)) return GE2E_ERR_UNSUPPORTED;
  if (dbias != nullptr) GE2E_CUDA_TRY(cudaMemsetAsync(dbias, 0, sizeof(float) * D, st));
  int grid = (U + kTailBwdWarps - 1) / kTailBwdWarps;
  if (grid > 148 * 4) grid = 148 * 4;
  embed_tail_bwd_rows_kernel<<<grid, kTailBwdWarps * 32, 0, st>>>(dE, E, inv_norm, U, D, dY, dbias);
  GE2E_LAUNCHED();
  return GE2E_OK;
}

}  // namespace ge2e
