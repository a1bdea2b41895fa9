// Hardware probe for the tcgen05 building blocks used by ge2e_tc.cu (debug aid, not product code):
// one CTA computes C[128 x N] = A[128 x K] * B with
//   mode 0: A smem K-major,  B smem K-major   (B given as [N][K])
//   mode 1: A smem K-major,  B smem MN-major  (B given as [K][N], 3-D TMA chunks of 16 k-rows)
//   mode 2: A in TMEM,       B smem K-major
//   mode 3: A in TMEM,       B smem MN-major
// and compares with a CPU reference (inputs are small integers: exact in TF32).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I ../../speaker_embedding_ge2e_loss_b200/csrc \
//        -I ../../include umma_probe.cu -o umma_probe
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ge2e_tc_ptx.cuh"

using namespace ge2e::ptx;

constexpr int K = 64;     // 2 slabs of 32
constexpr int NMAX = 256;

struct Cfg { int mode, N, lbo, sbo, use2d, ltype, tmasw; };

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB2,
             const __grid_constant__ CUtensorMap tmB3, const float* __restrict__ A, float* __restrict__ out, float* __restrict__ dbg, Cfg c) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_holder;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base;                 // 2 slabs x 16 KB
  const uint32_t b_smem = base + 2 * 16384;     // K-major: 2 slabs x (N*128 B); MN-major: 4 chunks x (N/32*2048 B)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
  if (tid == 0) { mbar_init(bar_tma, 1); mbar_init(bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_holder));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;
  const bool a_tmem = c.mode >= 2, b_mn = (c.mode & 1) != 0;

  if (tid == 0) {
    uint32_t bytes = 0;
    if (!a_tmem) bytes += 2 * 16384;
    bytes += b_mn ? (K / 16) * (c.N / 32) * 2048 : 2 * c.N * 128;
    mbar_expect_tx(bar_tma, bytes);
    if (!a_tmem) for (int ks = 0; ks < 2; ++ks) tma_load_2d(a_smem + ks * 16384, &tmA, ks * 32, 0, bar_tma);
    if (!b_mn) {
      for (int ks = 0; ks < 2; ++ks) tma_load_2d(b_smem + ks * c.N * 128, &tmB2, ks * 32, 0, bar_tma);
    } else {
      for (int kc = 0; kc < K / 16; ++kc) {
        if (!c.use2d) tma_load_3d(b_smem + kc * (c.N / 32) * 2048, &tmB3, 0, kc * 16, 0, bar_tma);
        else for (int ds = 0; ds < c.N / 32; ++ds)   // tmB2 here maps Bmn[K][NMAX] with box {32, 16}
          tma_load_2d(b_smem + kc * (c.N / 32) * 2048 + ds * 2048, &tmB2, ds * 32, kc * 16, bar_tma);
      }
    }
  }
  if (a_tmem) {   // A[128][K] -> TMEM columns [256, 256 + K): lane = row, column = k
    for (int ch = 0; ch < K / 32; ++ch) {
      uint32_t v[32];
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(A[tid * K + ch * 32 + i]);
      tmem_st32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 256 + ch * 32, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  {   // sentinel in the accumulator: detects an MMA that never wrote
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(7.0f);
    for (int ch = 0; ch < c.N / 32; ++ch) tmem_st32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, v);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  mbar_wait(bar_tma, 0);
  {   // dump the head of the B region and count non-zeros
    const float* bs = reinterpret_cast<const float*>(smem_raw + (b_smem - smem_u32(smem_raw)));
    if (tid < 64) dbg[tid] = bs[tid];
    int nz = 0;
    for (int i = tid; i < 4 * 8 * 512; i += 128) nz += (bs[i] != 0.f);
    atomicAdd(reinterpret_cast<int*>(dbg + 64), nz);
  }
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = idesc_tf32(128, c.N, 0, b_mn ? 1 : 0);
    for (int k = 0; k < K / 8; ++k) {    // one MMA per 8 k
      uint64_t db;
      if (!b_mn) db = smem_desc_sw128(b_smem + (k / 4) * c.N * 128 + (k % 4) * 32, 16, 1024);
      else db = smem_desc(b_smem + (k / 2) * (c.N / 32) * 2048 + (k % 2) * 1024, c.lbo, c.sbo, c.ltype);
      if (!a_tmem) {
        const uint64_t da = smem_desc_sw128(a_smem + (k / 4) * 16384 + (k % 4) * 32, 16, 1024);
        umma_tf32_ss(tmem, da, db, idesc, k != 0);
      } else {
        umma_tf32_ts(tmem, tmem + 256 + k * 8, db, idesc, k != 0);
      }
    }
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  for (int ch = 0; ch < c.N / 32; ++ch) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[tid * c.N + ch * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
  (void)lane;
}

static PFN_cuTensorMapEncodeTiled_v12000 enc() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
}
static void map2d(CUtensorMap* m, float* base, int rows, int cols, int box_rows, CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t str[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
  CUresult r = enc()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("map2d failed %d\n", (int)r);
}
static void map3d(CUtensorMap* m, float* base, int rows, int cols) {
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)cols / 32}; cuuint64_t str[2] = {(cuuint64_t)cols * 4, 128};
  cuuint32_t box[3] = {32, 16, (cuuint32_t)cols / 32}; cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("map3d failed %d\n", (int)r);
}

int main() {
  std::vector<float> hA(128 * K), hBk(NMAX * K), hBmn(K * NMAX);
  srand(1);
  for (auto& x : hA) x = (float)(rand() % 9 - 4);
  for (int n = 0; n < NMAX; ++n) for (int k = 0; k < K; ++k) { float v = (float)(rand() % 9 - 4); hBk[n * K + k] = v; hBmn[k * NMAX + n] = v; }
  float *dA, *dBk, *dBmn, *dOut;
  cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dBk, hBk.size() * 4); cudaMalloc(&dBmn, hBmn.size() * 4);
  cudaMalloc(&dOut, 128 * NMAX * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dBk, hBk.data(), hBk.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dBmn, hBmn.data(), hBmn.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 1024 + 2 * 16384 + 4 * 8 * 2048 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  Cfg cfgs[] = {{0, 128, 0, 0, 0, 2, 0}, {2, 256, 0, 0, 0, 2, 0},
                {1, 32, 2048, 512, 0, 1, 1}, {1, 32, 512, 2048, 0, 1, 1}, {1, 32, 2048, 512, 1, 1, 1},
                {1, 256, 2048, 512, 0, 1, 1}, {1, 256, 512, 2048, 0, 1, 1}, {1, 256, 2048, 512, 1, 1, 1},
                {1, 256, 2048, 1024, 0, 1, 1}, {1, 256, 2048, 512, 0, 1, 0}, {1, 256, 2048, 1024, 0, 1, 0},
                {1, 128, 2048, 512, 0, 1, 1}, {1, 64, 2048, 512, 0, 1, 1},
                {3, 256, 2048, 512, 0, 1, 1}, {3, 128, 2048, 512, 0, 1, 1}, {3, 32, 2048, 512, 0, 1, 1}};
  float* dDbg; cudaMalloc(&dDbg, 128 * 4);
  for (const Cfg& c : cfgs) {
    CUtensorMap tmA, tmB2, tmB3;
    map2d(&tmA, dA, 128, K, 128);
    const CUtensorMapSwizzle sw = c.tmasw ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    if (c.use2d) map2d(&tmB2, dBmn, K, NMAX, 16, sw); else map2d(&tmB2, dBk, NMAX, K, c.N);
    // MN-major source: [K][NMAX] row-major, only the first c.N columns are used (box covers c.N/32 chunks)
    {
      cuuint64_t dims[3] = {32, (cuuint64_t)K, (cuuint64_t)c.N / 32}; cuuint64_t str[2] = {(cuuint64_t)NMAX * 4, 128};
      cuuint32_t box[3] = {32, 16, (cuuint32_t)c.N / 32}; cuuint32_t es[3] = {1, 1, 1};
      CUresult r = enc()(&tmB3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dBmn, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) printf("map3d failed %d\n", (int)r);
    }
    cudaMemset(dOut, 0xFF, 128 * NMAX * 4);
    cudaMemset(dDbg, 0, 128 * 4);
    probe_kernel<<<1, 128, smem>>>(tmA, tmB2, tmB3, dA, dOut, dDbg, c);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> hOut(128 * c.N);
    cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, zeros = 0; int fm = -1, fn = -1;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < c.N; ++n) {
      float ref = 0; for (int k = 0; k < K; ++k) ref += hA[m * K + k] * hBk[n * K + k];
      float got = hOut[m * c.N + n];
      if (got == 0.f) ++zeros;
      if (!(got == ref)) { if (!bad) { fm = m; fn = n; } ++bad; }
    }
    float hDbg[128]; cudaMemcpy(hDbg, dDbg, sizeof(hDbg), cudaMemcpyDeviceToHost);
    int sevens = 0; for (float x : hOut) sevens += (x == 7.0f);
    printf("mode=%d N=%3d lbo=%4d sbo=%4d 2d=%d lt=%d tsw=%d : %s bad=%d/%d zeros=%d sevens=%d first_bad=(%d,%d) Bnz=%d B[0..7]=%g %g %g %g %g %g %g %g err=%s\n",
           c.mode, c.N, c.lbo, c.sbo, c.use2d, c.ltype, c.tmasw, bad ? "FAIL" : "ok  ", bad, 128 * c.N, zeros, sevens, fm, fn,
           *reinterpret_cast<int*>(&hDbg[64]), hDbg[0], hDbg[1], hDbg[2], hDbg[3], hDbg[4], hDbg[5], hDbg[6], hDbg[7],
           cudaGetErrorString(e));
    if (bad && c.N == 32) { printf("   got row0: "); for (int n = 0; n < 8; ++n) printf("%g ", hOut[n]);
      printf("\n   ref row0: "); for (int n = 0; n < 8; ++n) { float r = 0; for (int k = 0; k < K; ++k) r += hA[k] * hBk[n * K + k]; printf("%g ", r); } printf("\n"); }
    if (e != cudaSuccess) return 1;
  }
  (void)map3d;
  return 0;
}
