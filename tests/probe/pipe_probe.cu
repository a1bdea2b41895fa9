// Timing probe for the single-thread costs of the tcgen05 pipeline (debug aid, not product code):
// how long one thread spends in mbarrier.try_wait / tcgen05.fence / tcgen05.commit / tcgen05.mma
// issue, how fast the tensor pipe retires TF32 MMAs of a given shape (SS and TS, cta_group 1 and
// 2), and what a TMA-less producer/consumer ring costs per stage.  All numbers are SM clocks.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I ../../speaker_embedding_ge2e_loss_b200/csrc \
//        pipe_probe.cu -o pipe_probe
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ge2e_tc_ptx.cuh"

using namespace ge2e::ptx;

__device__ __forceinline__ long long clk() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}

template <int CG>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (CG == 1) {
    umma_tf32_ss(d, da, db, idesc, acc);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (CG == 1) {
    umma_tf32_ts(d, a, db, idesc, acc);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
        "r"(a), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t bar) {
  if (CG == 1) umma_commit(bar);
  else
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"((uint16_t)1)
        : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

constexpr int kStages = 6;
enum { R_TRYWAIT, R_FENCE, R_COMMIT_ISSUE, R_COMMIT_RT, R_SS128_ISSUE, R_SS128_DONE, R_SS256_ISSUE, R_SS256_DONE,
       R_TS256_ISSUE, R_TS256_DONE, R_SS64_ISSUE, R_SS64_DONE, R_LOOP0, R_COUNT = R_LOOP0 + 8 * 6 + 12 };

struct LoopCfg { int nslab, nmma, n; };

constexpr int R2_BASE = 12 + 8 * 6;
template <int CG, int NMMA, int N>
__device__ __forceinline__ void ring2(int c, long long* out, uint64_t* bars, uint32_t a_smem, uint32_t b_smem, uint32_t tmem,
                                      uint32_t rank, int& g_stage_p, int& g_phase_p, int& g_stage_c, int& g_phase_c) {
  const int tid = threadIdx.x, warp = tid >> 5;
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  const uint64_t dk = smem_desc(0, 16, 1024, kLayoutSw128);
  const int MM = CG == 2 ? 256 : 128;
  const int B_DONE = 2 * kStages;
  __syncthreads();
  if (rank == 0 && warp == 1) {
    int stage = g_stage_p, phase = g_phase_p;
    for (int s = 0; s < 64; ++s) {
      mbar_wait(bar(kStages + stage), phase ^ 1);
      if (elect_one()) mbar_arrive(bar(stage));
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    __syncwarp();
    if (tid == 32) { g_stage_p = stage; g_phase_p = phase; }
  } else if (rank == 0 && warp == 0) {
    int stage = g_stage_c, phase = g_phase_c;
    const uint32_t idesc = idesc_tf32(MM, N, 0, 0);
    const long long tb = clk();
    for (int s = 0; s < 64; ++s) {
      mbar_wait(bar(stage), phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = dk | (a_smem >> 4);
        const uint64_t db = dk | (b_smem >> 4);
#pragma unroll
        for (int k = 0; k < NMMA; ++k) mma_ss<CG>(tmem, da + 2 * (k & 3), db + 2 * (k & 3), idesc, (s | k) != 0);
        commit<CG>(bar(kStages + stage));
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    const long long te = clk();
    if (elect_one()) commit<CG>(bar(B_DONE));
    __syncwarp();
    mbar_wait(bar(B_DONE), c & 1);
    const long long td = clk();
    if (tid == 0) { g_stage_c = stage; g_phase_c = phase; out[R2_BASE + c * 2] = te - tb; out[R2_BASE + c * 2 + 1] = td - tb; }
  }
  __syncthreads();
}

template <int CG>
__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bars[2 * kStages + 4];
  __shared__ uint32_t tmem_holder;
  __shared__ int g_stage_p, g_phase_p, g_stage_c, g_phase_c;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 16384;   // A [128][32], B [256][32], both sw128 K-major
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  auto bar = [&](int i) { return smem_u32(&bars[i]); };
  const int B_DONE = 2 * kStages, B_PRE = 2 * kStages + 1;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128)
    reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0.001f * (i & 63);
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(bar(i), 1); mbar_init(bar(kStages + i), 1); }
    mbar_init(bar(B_DONE), 1);
    mbar_init(bar(B_PRE), 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    if (CG == 1) tmem_alloc<512>(smem_u32(&tmem_holder));
    else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const uint64_t dk = smem_desc(0, 16, 1024, kLayoutSw128);
  const int MM = CG == 2 ? 256 : 128;

  if (rank == 0 && tid == 0) {
    long long t0, t1, t2;
    int done_phase = 0;
    // ---- try_wait on an already-complete phase
    mbar_arrive(bar(B_PRE));
    mbar_wait(bar(B_PRE), 0);
    t0 = clk();
    for (int i = 0; i < 32; ++i) mbar_wait(bar(B_PRE), 0);
    t1 = clk();
    out[R_TRYWAIT] = (t1 - t0) / 32;
    // ---- fence::after_thread_sync
    t0 = clk();
    for (int i = 0; i < 32; ++i) tc_fence_after();
    t1 = clk();
    out[R_FENCE] = (t1 - t0) / 32;
    // ---- commit with nothing pending: issue cost and round trip to the barrier
    long long ci = 0, crt = 0;
    for (int i = 0; i < 8; ++i) {
      t0 = clk();
      commit<CG>(bar(B_DONE));
      t1 = clk();
      mbar_wait(bar(B_DONE), done_phase);
      done_phase ^= 1;
      t2 = clk();
      ci += t1 - t0; crt += t2 - t0;
    }
    out[R_COMMIT_ISSUE] = ci / 8;
    out[R_COMMIT_RT] = crt / 8;
    // ---- 32 back-to-back MMAs: issue time and completion time
    auto run = [&](int N, bool ts, int ri, int rd) {
      const uint32_t idesc = idesc_tf32(MM, N, 0, 0);
      t0 = clk();
#pragma unroll 1
      for (int k = 0; k < 32; ++k) {
        const uint64_t da = dk | ((a_smem + (k & 3) * 32) >> 4);
        const uint64_t db = dk | ((b_smem + (k & 3) * 32) >> 4);
        if (!ts) mma_ss<CG>(tmem, da, db, idesc, k != 0);
        else mma_ts<CG>(tmem, tmem + 256 + (k & 15) * 8, db, idesc, k != 0);
      }
      t1 = clk();
      commit<CG>(bar(B_DONE));
      mbar_wait(bar(B_DONE), done_phase);
      done_phase ^= 1;
      t2 = clk();
      out[ri] = t1 - t0;
      out[rd] = t2 - t0;
    };
    for (int rep = 0; rep < 2; ++rep) {
      run(128, false, R_SS128_ISSUE, R_SS128_DONE);
      run(256, false, R_SS256_ISSUE, R_SS256_DONE);
      run(256, true, R_TS256_ISSUE, R_TS256_DONE);
      run(64, false, R_SS64_ISSUE, R_SS64_DONE);
    }
  }
  __syncthreads();
  if (CG == 2) cluster_sync_all();

  // ---- TMA-less ring: thread 32 plays the producer (wait EMPTY -> arrive FULL), thread 0 the MMA
  // issuer (wait FULL, fence, nmma MMAs, commit EMPTY); 8 configurations
  const LoopCfg cfgs[8] = {{64, 0, 128}, {64, 4, 128}, {64, 8, 128}, {64, 4, 256}, {64, 8, 256}, {64, 16, 128},
                           {64, 16, 256}, {64, 2, 128}};
  for (int c = 0; c < 8; ++c) {
    const LoopCfg L = cfgs[c];
    // barrier phases continue across configurations: every config runs nslab = 64 slabs; 64 is not a
    // multiple of kStages, so track stage/phase globally
    if (tid == 0 && c == 0) { g_stage_p = g_phase_p = g_stage_c = g_phase_c = 0; }
    __syncthreads();
    if (rank == 0 && tid == 32) {
      int stage = g_stage_p, phase = g_phase_p;
      for (int s = 0; s < L.nslab; ++s) {
        mbar_wait(bar(kStages + stage), phase ^ 1);
        mbar_arrive(bar(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      g_stage_p = stage; g_phase_p = phase;
    } else if (rank == 0 && tid == 0) {
      int stage = g_stage_c, phase = g_phase_c;
      const uint32_t idesc = idesc_tf32(MM, L.n, 0, 0);
      long long tw = 0, tf = 0, tm = 0, tc = 0;
      const long long tb = clk();
      for (int s = 0; s < L.nslab; ++s) {
        long long a0 = clk();
        mbar_wait(bar(stage), phase);
        long long a1 = clk();
        tc_fence_after();
        long long a2 = clk();
#pragma unroll 1
        for (int k = 0; k < L.nmma; ++k) {
          const uint64_t da = dk | ((a_smem + (k & 3) * 32) >> 4);
          const uint64_t db = dk | ((b_smem + (k & 3) * 32) >> 4);
          mma_ss<CG>(tmem, da, db, idesc, (s | k) != 0);
        }
        long long a3 = clk();
        commit<CG>(bar(kStages + stage));
        long long a4 = clk();
        tw += a1 - a0; tf += a2 - a1; tm += a3 - a2; tc += a4 - a3;
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      const long long te = clk();
      // drain: the last kStages commits must land before the next configuration reuses the ring
      g_stage_c = stage; g_phase_c = phase;
      long long* o = out + R_LOOP0 + c * 6;
      o[0] = te - tb; o[1] = tw; o[2] = tf; o[3] = tm; o[4] = tc;
      commit<CG>(bar(B_DONE));
      // B_DONE phase bookkeeping: 8 + 8 completions so far (even) -> parity of this one is c & 1
      mbar_wait(bar(B_DONE), c & 1);
      o[5] = clk() - tb;
    }
    __syncthreads();
  }

  // ---- same ring, whole-warp roles + elect.sync issue (the CUTLASS idiom), unrolled MMAs
  ring2<CG, 4, 128>(0, out, bars, a_smem, b_smem, tmem, rank, g_stage_p, g_phase_p, g_stage_c, g_phase_c);
  ring2<CG, 8, 128>(1, out, bars, a_smem, b_smem, tmem, rank, g_stage_p, g_phase_p, g_stage_c, g_phase_c);
  ring2<CG, 4, 256>(2, out, bars, a_smem, b_smem, tmem, rank, g_stage_p, g_phase_p, g_stage_c, g_phase_c);
  ring2<CG, 8, 256>(3, out, bars, a_smem, b_smem, tmem, rank, g_stage_p, g_phase_p, g_stage_c, g_phase_c);
  ring2<CG, 2, 256>(4, out, bars, a_smem, b_smem, tmem, rank, g_stage_p, g_phase_p, g_stage_c, g_phase_c);
  ring2<CG, 16, 64>(5, out, bars, a_smem, b_smem, tmem, rank, g_stage_p, g_phase_p, g_stage_c, g_phase_c);

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    if (CG == 1) tmem_dealloc<512>(tmem);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

template <int CG>
static int run_one() {
  long long* d;
  cudaMalloc(&d, R_COUNT * sizeof(long long));
  cudaMemset(d, 0, R_COUNT * sizeof(long long));
  const size_t smem = 1024 + 16384 + 32768;
  cudaFuncSetAttribute(probe<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CG);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, probe<CG>, d);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  printf("=== cta_group::%d  (%s)\n", CG, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<long long> h(R_COUNT);
  cudaMemcpy(h.data(), d, R_COUNT * sizeof(long long), cudaMemcpyDeviceToHost);
  printf("try_wait(complete)      %lld cyc\n", h[R_TRYWAIT]);
  printf("fence::after            %lld cyc\n", h[R_FENCE]);
  printf("commit issue / roundtrip %lld / %lld cyc\n", h[R_COMMIT_ISSUE], h[R_COMMIT_RT]);
  printf("32x MMA SS N=64 : issue %lld  done %lld  (%.1f cyc/MMA)\n", h[R_SS64_ISSUE], h[R_SS64_DONE], h[R_SS64_DONE] / 32.0);
  printf("32x MMA SS N=128: issue %lld  done %lld  (%.1f cyc/MMA)\n", h[R_SS128_ISSUE], h[R_SS128_DONE], h[R_SS128_DONE] / 32.0);
  printf("32x MMA SS N=256: issue %lld  done %lld  (%.1f cyc/MMA)\n", h[R_SS256_ISSUE], h[R_SS256_DONE], h[R_SS256_DONE] / 32.0);
  printf("32x MMA TS N=256: issue %lld  done %lld  (%.1f cyc/MMA)\n", h[R_TS256_ISSUE], h[R_TS256_DONE], h[R_TS256_DONE] / 32.0);
  const LoopCfg cfgs[8] = {{64, 0, 128}, {64, 4, 128}, {64, 8, 128}, {64, 4, 256}, {64, 8, 256}, {64, 16, 128},
                           {64, 16, 256}, {64, 2, 128}};
  for (int c = 0; c < 8; ++c) {
    const long long* o = h.data() + R_LOOP0 + c * 6;
    printf("ring 64 stages x %2d MMA N=%3d: per stage total %.0f (wait %.0f fence %.0f mma %.0f commit %.0f)  drained %.0f  ideal-mma %.0f\n",
           cfgs[c].nmma, cfgs[c].n, o[0] / 64.0, o[1] / 64.0, o[2] / 64.0, o[3] / 64.0, o[4] / 64.0, o[5] / 64.0,
           cfgs[c].nmma * cfgs[c].n / 2.0);
  }
  const int r2[6][2] = {{4, 128}, {8, 128}, {4, 256}, {8, 256}, {2, 256}, {16, 64}};
  for (int c = 0; c < 6; ++c)
    printf("elect-ring 64 stages x %2d MMA N=%3d: per stage issue %.0f  drained %.0f  ideal-mma %.0f\n", r2[c][0], r2[c][1],
           h[R_LOOP0 + 48 + c * 2] / 64.0, h[R_LOOP0 + 48 + c * 2 + 1] / 64.0, r2[c][0] * r2[c][1] / 2.0);
  cudaFree(d);
  return 0;
}

int main() {
  if (run_one<1>()) return 1;
  if (run_one<2>()) return 1;
  return 0;
}
