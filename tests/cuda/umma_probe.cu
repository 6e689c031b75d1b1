// Stand-alone probe of the tcgen05 building blocks used by kp_flash_umma.cu:
//   case 0: D[128,N] = A[128,K] * B[N,K]^T        A, B K-major, 128B swizzle (TMA-loaded)
//   case 1: D[128,N] = A[128,K] * Bt[K,N]         B MN-major (rows of Bt are K, N contiguous)
// bf16 inputs, fp32 accumulate in TMEM, read back with tcgen05.ld.  Prints max |err|.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cmath>
#include "../../kelpie_b200/csrc/kp_ptx.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2);} } while (0)

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static enc_fn get_enc() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  return (enc_fn)p;
}
static CUtensorMap make_map(void* base, int rows, int cols, int box_rows, int box_cols) {
  CUtensorMap m; cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t str[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
  CUresult r = get_enc()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(2); }
  return m;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

// K = 128 (two 64-wide k-blocks).  mode 0: B K-major [N rows x K]; mode 1: B MN-major [K rows x N]
template <int N>
__global__ void probe(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, int mode, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = sm;                  // 2 k-blocks x [128 x 128B] = 32 KB
  uint8_t* sB = sm + 32768;          // mode 0: 2 k-blocks x [N x 128B]; mode 1: (N/64) boxes x [128 k-rows x 128B]
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(&full, 1); ptx::mbar_init(&done, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base, 256); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    uint32_t bytes = 32768 + (mode == 0 ? 2 * N * 128 : (N / 64) * 128 * 128);
    ptx::mbar_arrive_expect_tx(&full, bytes);
    for (int kb = 0; kb < 2; ++kb) ptx::tma_load_2d(sA + kb * 16384, &amap, &full, kb * 64, 0);
    if (mode == 0) for (int kb = 0; kb < 2; ++kb) ptx::tma_load_2d(sB + kb * N * 128, &bmap, &full, kb * 64, 0);
    else for (int nb = 0; nb < N / 64; ++nb) ptx::tma_load_2d(sB + nb * 16384, &bmap, &full, nb * 64, 0);
    ptx::mbar_wait(&full, 0);
    ptx::tc_fence_after();
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (mode == 1) idesc |= (1u << 16);  // B is MN-major
    for (int ks = 0; ks < 8; ++ks) {     // 8 x K16
      const int kb = ks >> 2, kk = ks & 3;
      uint64_t ad = make_desc(ptx::smem_u32(sA + kb * 16384) + kk * 32, 16, 1024);
      uint64_t bd;
      if (mode == 0) bd = make_desc(ptx::smem_u32(sB + kb * N * 128) + kk * 32, 16, 1024);
      else bd = make_desc(ptx::smem_u32(sB) + ks * 2048, 16384, 1024);  // 16 k-rows per step; LBO = next 64-wide N box
      ptx::umma_bf16(tm, ad, bd, idesc, ks > 0);
    }
    ptx::umma_commit(&done);
  }
  __syncthreads();
  ptx::mbar_wait(&done, 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    ptx::tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 256);
}

// A [128 x 128] bf16 held in TMEM (two bf16 per 32-bit column, lane = row), B from smem.
// mode 2: B K-major; mode 3: B MN-major.
template <int N>
__global__ void probe_ts(const __nv_bfloat16* A, const __grid_constant__ CUtensorMap bmap, int mode, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = sm;
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(&full, 1); ptx::mbar_init(&done, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_base;
  const uint32_t TA = tm + 256;  // A region: 64 columns
  {  // every thread stores its row: 128 bf16 = 64 packed columns
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t r[32];
      for (int c = 0; c < 32; ++c) {
        const uint32_t lo = __bfloat16_as_ushort(A[row * 128 + 2 * (c0 + c)]);
        const uint32_t hi = __bfloat16_as_ushort(A[row * 128 + 2 * (c0 + c) + 1]);
        r[c] = lo | (hi << 16);
      }
      ptx::tmem_st_32x32(TA + ((uint32_t)(warp * 32) << 16) + c0, r);
    }
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (threadIdx.x == 0) {
    uint32_t bytes = (mode == 2 ? 2 * N * 128 : (N / 64) * 128 * 128);
    ptx::mbar_arrive_expect_tx(&full, bytes);
    if (mode == 2) for (int kb = 0; kb < 2; ++kb) ptx::tma_load_2d(sB + kb * N * 128, &bmap, &full, kb * 64, 0);
    else for (int nb = 0; nb < N / 64; ++nb) ptx::tma_load_2d(sB + nb * 16384, &bmap, &full, nb * 64, 0);
    ptx::mbar_wait(&full, 0);
    ptx::tc_fence_after();
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (mode == 3) idesc |= (1u << 16);
    for (int ks = 0; ks < 8; ++ks) {
      const int kb = ks >> 2, kk = ks & 3;
      uint64_t bd;
      if (mode == 2) bd = make_desc(ptx::smem_u32(sB + kb * N * 128) + kk * 32, 16, 1024);
      else bd = make_desc(ptx::smem_u32(sB) + ks * 2048, 16384, 1024);
      ptx::umma_bf16_ts(tm, TA + ks * 8, bd, idesc, ks > 0);
    }
    ptx::umma_commit(&done);
  }
  __syncthreads();
  ptx::mbar_wait(&done, 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    ptx::tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

// A [128 x 128] bf16 written to shared memory by the threads in the K-major NO-SWIZZLE layout
// (16-byte chunk c = 8 k-elements of row r at c*2048 + r*16), B K-major 128B-swizzled by TMA.  mode 4.
template <int N>
__global__ void probe_ns(const __nv_bfloat16* A, const __grid_constant__ CUtensorMap bmap, uint32_t lbo, uint32_t sbo, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = sm;          // 32 KB
  uint8_t* sB = sm + 32768;  // 2 k-blocks x [N x 128B]
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(&full, 1); ptx::mbar_init(&done, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base, 256); ptx::tmem_relinquish(); }
  {
    const int row = warp * 32 + lane;
    for (int c = 0; c < 16; ++c) {
      uint4 v = *reinterpret_cast<const uint4*>(A + row * 128 + c * 8);
      *reinterpret_cast<uint4*>(sA + c * 2048 + row * 16) = v;
    }
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&full, 2 * N * 128);
    for (int kb = 0; kb < 2; ++kb) ptx::tma_load_2d(sB + kb * N * 128, &bmap, &full, kb * 64, 0);
    ptx::mbar_wait(&full, 0);
    ptx::tc_fence_after();
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int ks = 0; ks < 8; ++ks) {
      const int kb = ks >> 2, kk = ks & 3;
      uint64_t ad = 0;
      ad |= (uint64_t)(((ptx::smem_u32(sA) + ks * 4096) >> 4) & 0x3fff);
      ad |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
      ad |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
      ad |= (uint64_t)1 << 46;  // layout type 0: no swizzle
      uint64_t bd = make_desc(ptx::smem_u32(sB + kb * N * 128) + kk * 32, 16, 1024);
      ptx::umma_bf16(tm, ad, bd, idesc, ks > 0);
    }
    ptx::umma_commit(&done);
  }
  __syncthreads();
  ptx::mbar_wait(&done, 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    ptx::tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 256);
}

template <int N>
static double run(int mode, uint32_t lbo = 0, uint32_t sbo = 0) {
  const int M = 128, K = 128;
  std::vector<__nv_bfloat16> A(M * K), B(N * K), Bt(K * N);
  std::vector<float> Af(M * K), Bf(N * K);
  srand(1234 + mode + N);
  for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; A[i] = __float2bfloat16(v); Af[i] = __bfloat162float(A[i]); }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { float v = (rand() % 2001 - 1000) / 1000.f; __nv_bfloat16 b = __float2bfloat16(v); B[n * K + k] = b; Bt[k * N + n] = b; Bf[n * K + k] = __bfloat162float(b); }
  __nv_bfloat16 *dA, *dB; float* dO;
  CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dO, M * N * 4));
  CK(cudaMemcpy(dA, A.data(), M * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, (mode == 0 || mode == 2 || mode == 4) ? B.data() : Bt.data(), N * K * 2, cudaMemcpyHostToDevice));
  CUtensorMap am = make_map(dA, M, K, 128, 64);
  const bool kmajor = (mode == 0 || mode == 2 || mode == 4);
  CUtensorMap bm = kmajor ? make_map(dB, N, K, N, 64) : make_map(dB, K, N, 128, 64);
  if (mode == 4) {
    CK(cudaFuncSetAttribute(probe_ns<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    probe_ns<N><<<1, 128, 100 * 1024>>>(dA, bm, lbo, sbo, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe_ns lbo=%u sbo=%u: %s\n", lbo, sbo, cudaGetErrorString(e)); exit(3); }
  } else if (mode < 2) {
    CK(cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    probe<N><<<1, 128, 100 * 1024>>>(am, bm, mode, dO);
  } else {
    CK(cudaFuncSetAttribute(probe_ts<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    probe_ts<N><<<1, 128, 100 * 1024>>>(dA, bm, mode, dO);
  }
  CK(cudaDeviceSynchronize());
  std::vector<float> O(M * N);
  CK(cudaMemcpy(O.data(), dO, M * N * 4, cudaMemcpyDeviceToHost));
  double err = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0; for (int k = 0; k < K; ++k) ref += (double)Af[m * K + k] * Bf[n * K + k];
    err = fmax(err, fabs(ref - O[m * N + n]));
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
  return err;
}

// cta_group::2 (cluster of two CTAs): D[256, N] = A[256, 128] * Bt[128, N].  Each CTA writes ITS 128 rows of A into its
// shared memory in the K-major no-swizzle layout and TMA-loads ITS half of Bt's N columns (MN-major, N/2 per CTA as
// 64-wide boxes `atom_stride` bytes apart).  Issued by the even CTA.  mode 5.
template <int N>
__global__ void __cluster_dims__(2, 1, 1) probe_pair(const __nv_bfloat16* A, const __grid_constant__ CUtensorMap bmap, float* out, int a_off) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = sm + 65536 + a_off;  // 32 KB
  uint8_t* sB = sm;                  // (N/2/64) boxes x [128 k-rows x 128 B]
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  if (threadIdx.x == 0) { ptx::mbar_init(&full, 1); ptx::mbar_init(&done, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc2(&tmem_base, 256); ptx::tmem_relinquish2(); }
  {
    const int row = warp * 32 + lane;
    for (int c = 0; c < 16; ++c) {
      uint4 v = *reinterpret_cast<const uint4*>(A + ((int)crank * 128 + row) * 128 + c * 8);
      *reinterpret_cast<uint4*>(sA + c * 2048 + row * 16) = v;
    }
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tm = tmem_base;
  constexpr int NB = N / 2 / 64;  // boxes per CTA
  if (threadIdx.x == 0) {
    if (crank == 0) ptx::mbar_arrive_expect_tx(&full, 2 * NB * 16384);
    const uint32_t bar = ptx::mapa_u32(ptx::smem_u32(&full), 0);
    for (int nb = 0; nb < NB; ++nb) ptx::tma_load_2d_pair(sB + nb * 16384, &bmap, bar, ((int)crank * NB + nb) * 64, 0);
    if (crank == 0) {
      ptx::mbar_wait(&full, 0);
      ptx::tc_fence_after();
      uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      for (int ks = 0; ks < 8; ++ks) {
        uint64_t ad = 0;
        ad |= (uint64_t)(((ptx::smem_u32(sA) + ks * 4096) >> 4) & 0x3fff);
        ad |= (uint64_t)((2048 >> 4) & 0x3fff) << 16;
        ad |= (uint64_t)((128 >> 4) & 0x3fff) << 32;
        ad |= (uint64_t)1 << 46;
        uint64_t bd = make_desc(ptx::smem_u32(sB) + ks * 2048, 16384, 1024);
        ptx::umma2_bf16(tm, ad, bd, idesc, ks > 0);
      }
      ptx::umma2_commit_mc(&done, 3);
    }
  }
  __syncthreads();
  ptx::mbar_wait(&done, 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    ptx::tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[((int)crank * 128 + warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 0) ptx::tmem_dealloc2(tm, 256);
}

template <int N>
static double run_pair(int a_off = 0) {
  const int M = 256, K = 128;
  std::vector<__nv_bfloat16> A(M * K), Bt(K * N);
  std::vector<float> Af(M * K), Bf(N * K);
  srand(4321 + N);
  for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; A[i] = __float2bfloat16(v); Af[i] = __bfloat162float(A[i]); }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { float v = (rand() % 2001 - 1000) / 1000.f; __nv_bfloat16 b = __float2bfloat16(v); Bt[k * N + n] = b; Bf[n * K + k] = __bfloat162float(b); }
  __nv_bfloat16 *dA, *dB; float* dO;
  CK(cudaMalloc(&dA, M * K * 2)); CK(cudaMalloc(&dB, N * K * 2)); CK(cudaMalloc(&dO, M * N * 4));
  CK(cudaMemcpy(dA, A.data(), M * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bt.data(), N * K * 2, cudaMemcpyHostToDevice));
  CUtensorMap bm = make_map(dB, K, N, 128, 64);
  CK(cudaFuncSetAttribute(probe_pair<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
  probe_pair<N><<<2, 128, 226 * 1024>>>(dA, bm, dO, a_off);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe_pair N=%d: %s\n", N, cudaGetErrorString(e)); exit(3); }
  std::vector<float> O(M * N);
  CK(cudaMemcpy(O.data(), dO, M * N * 4, cudaMemcpyDeviceToHost));
  double err = 0; int badrow_lo = -1, badrow_hi = -1, badcol_lo = -1, badcol_hi = -1;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0; for (int k = 0; k < K; ++k) ref += (double)Af[m * K + k] * Bf[n * K + k];
    const double d = fabs(ref - O[m * N + n]);
    if (d > 1e-3) { if (badrow_lo < 0) badrow_lo = m; badrow_hi = m; if (badcol_lo < 0 || n < badcol_lo) badcol_lo = n; if (n > badcol_hi) badcol_hi = n; }
    err = fmax(err, d);
  }
  if (badrow_lo >= 0) printf("  bad rows %d..%d, bad cols %d..%d\n", badrow_lo, badrow_hi, badcol_lo, badcol_hi);
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
  return err;
}

int main() {
  int bad = 0;
  double e;
  e = run<64>(0);  printf("K-major  N=64  max_err=%g\n", e); bad += e > 1e-3;
  e = run<128>(0); printf("K-major  N=128 max_err=%g\n", e); bad += e > 1e-3;
  e = run<64>(1);  printf("MN-major N=64  max_err=%g\n", e); bad += e > 1e-3;
  e = run<128>(1); printf("MN-major N=128 max_err=%g\n", e); bad += e > 1e-3;
  e = run<64>(2);  printf("A-in-TMEM, B K-major  N=64  max_err=%g\n", e); bad += e > 1e-3;
  e = run<64>(3);  printf("A-in-TMEM, B MN-major N=64  max_err=%g\n", e); bad += e > 1e-3;
  e = run<128>(3); printf("A-in-TMEM, B MN-major N=128 max_err=%g\n", e); bad += e > 1e-3;
  e = run<64>(4, 2048, 128);  printf("A no-swizzle (LBO 2048 = K chunks, SBO 128 = 8-row groups), N=64 max_err=%g\n", e); bad += e > 1e-3;
  e = run<128>(4, 2048, 128); printf("A no-swizzle (LBO 2048, SBO 128), N=128 max_err=%g\n", e); bad += e > 1e-3;
  e = run_pair<128>(); printf("cta_group::2 M=256 N=128: A no-swizzle from smem, B MN-major max_err=%g\n", e); bad += e > 1e-3;
  e = run_pair<256>(); printf("cta_group::2 M=256 N=256: A no-swizzle from smem, B MN-major (two atoms per CTA) max_err=%g\n", e); bad += e > 1e-3;
  for (int off : {32768, 65536, 98304, 131072}) {
    e = run_pair<256>(off); printf("cta_group::2 N=256, A at smem offset %d: max_err=%g\n", 65536 + off, e); bad += e > 1e-3;
  }
  if (getenv("PROBE_SWAPPED")) { e = run<64>(4, 128, 2048); printf("A no-swizzle SWAPPED (LBO 128, SBO 2048) max_err=%g\n", e); }
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad ? 1 : 0;
}
