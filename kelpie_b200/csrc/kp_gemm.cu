// Plain fp32 GEMM on CUDA cores for the frozen ConvE network's Linear layer
// (conve.py:50-52,150; forward x = feat W^T, backward dfeat = dh W).  128x64x16 tiles, 8x4
// register tile per thread.  (These two GEMMs are ~10 % of the ConvE entity-projection work;
// moving them to tcgen05 with the bf16x3 split is listed under "next" in DESIGN.md.)
#include "kp_internal.h"

namespace {

constexpr int GM = 128, GN = 64, GK = 16, GT = 256;

// C[M,N] = A[M,K] * op(B);  TRANSB: B is [N,K] row-major (C = A B^T), else B is [K,N] row-major
template <bool TRANSB>
__global__ void __launch_bounds__(GT) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                   const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc, int kb,
                                                   int k_per_split) {
  __shared__ float As[GK][GM + 4];
  __shared__ float Bs[GK][GN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
  const int tx = tid & 15, ty = tid >> 4;  // thread tile: rows ty*8.., cols tx*4..
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int a_row = tid >> 1, a_k = (tid & 1) * 8;
  const int k_begin = k_per_split > 0 ? (int)blockIdx.z * k_per_split : 0;
  if (k_per_split > 0) K = min(K, k_begin + k_per_split);  // split-K: this CTA's range (a multiple of GK long)
  for (int k0 = k_begin; k0 < K; k0 += GK) {
    {  // A tile: [128 rows x 16 k], stored transposed
      const int m = m0 + a_row;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = k0 + a_k + h * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < M && k < K) v = *reinterpret_cast<const float4*>(A + (size_t)m * lda + k);
        As[a_k + h * 4 + 0][a_row] = v.x;
        As[a_k + h * 4 + 1][a_row] = v.y;
        As[a_k + h * 4 + 2][a_row] = v.z;
        As[a_k + h * 4 + 3][a_row] = v.w;
      }
    }
    if (TRANSB) {  // B tile from [N,K]: 64 rows x 16 k
      const int n = n0 + (tid >> 2), k = k0 + (tid & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < N && k < K) v = *reinterpret_cast<const float4*>(B + (size_t)n * ldb + k);
      const int kk = (tid & 3) * 4, nn = tid >> 2;
      Bs[kk + 0][nn] = v.x;
      Bs[kk + 1][nn] = v.y;
      Bs[kk + 2][nn] = v.z;
      Bs[kk + 3][nn] = v.w;
    } else {  // B tile from [K,N]: 16 k-rows x 64 cols
      const int k = k0 + (tid >> 4), n = n0 + (tid & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < kb) {  // rows of B beyond kb do not exist (A is zero-padded there)
        if (n + 3 < N) {
          v = *reinterpret_cast<const float4*>(B + (size_t)k * ldb + n);
        } else {
          if (n < N) v.x = B[(size_t)k * ldb + n];
          if (n + 1 < N) v.y = B[(size_t)k * ldb + n + 1];
          if (n + 2 < N) v.z = B[(size_t)k * ldb + n + 2];
        }
      }
      *reinterpret_cast<float4*>(&Bs[tid >> 4][(tid & 15) * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) {
        if (k_per_split > 0) atomicAdd(C + (size_t)m * ldc + n, acc[i][j]);
        else C[(size_t)m * ldc + n] = acc[i][j];
      }
    }
  }
}

// Skinny C[M, N] = A[M, K] B[N, K]^T for a handful of rows and a long K (the Linear layer's forward for the few
// (s, p) pairs of an explain-path batch: M ~ 10, N = 200, K = 9728).  The tiled kernel above gives such a product
// ceil(N / 64) = 4 CTAs that walk K serially (860 us); here a CTA owns 8 rows x 4 columns, its threads stride K
// with 128-bit loads and the partial sums are folded in a fixed order (shuffles, then the warps in order), so the
// result is reproducible run to run -- unlike an atomic split-K.
constexpr int SK_THREADS = 512, SK_ROWS = 8, SK_COLS = 4;  // 512 threads: K = 9728 is 4.75 trips of the strided loop
__global__ void __launch_bounds__(SK_THREADS) skinny_nt_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                               const float* __restrict__ B, int ldb, float* __restrict__ C,
                                                               int ldc) {
  __shared__ float red[SK_THREADS / 32][SK_ROWS * SK_COLS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * SK_COLS, m0 = blockIdx.y * SK_ROWS;
  float acc[SK_ROWS][SK_COLS];
#pragma unroll
  for (int r = 0; r < SK_ROWS; ++r)
#pragma unroll
    for (int c = 0; c < SK_COLS; ++c) acc[r][c] = 0.f;
  for (int k = tid * 4; k < K; k += SK_THREADS * 4) {
    float4 b[SK_COLS];
#pragma unroll
    for (int c = 0; c < SK_COLS; ++c)
      b[c] = (n0 + c < N) ? *reinterpret_cast<const float4*>(B + (size_t)(n0 + c) * ldb + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r) {
      if (m0 + r >= M) break;
      const float4 a = *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * lda + k);
#pragma unroll
      for (int c = 0; c < SK_COLS; ++c) {
        acc[r][c] = __fmaf_rn(a.x, b[c].x, acc[r][c]);
        acc[r][c] = __fmaf_rn(a.y, b[c].y, acc[r][c]);
        acc[r][c] = __fmaf_rn(a.z, b[c].z, acc[r][c]);
        acc[r][c] = __fmaf_rn(a.w, b[c].w, acc[r][c]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < SK_ROWS; ++r)
#pragma unroll
    for (int c = 0; c < SK_COLS; ++c) {
      float v = acc[r][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][r * SK_COLS + c] = v;
    }
  __syncthreads();
  if (tid < SK_ROWS * SK_COLS) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < SK_THREADS / 32; ++w) sum += red[w][tid];
    const int m = m0 + tid / SK_COLS, n = n0 + tid % SK_COLS;
    if (m < M && n < N) C[(size_t)m * ldc + n] = sum;
  }
}

}  // namespace

int kp_sgemm_skinny_nt(kp_ctx* ctx, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                       cudaStream_t st) {
  if (M <= 0 || N <= 0) return KP_OK;
  if (K % 4 != 0 || lda % 4 != 0 || ldb % 4 != 0) KP_FAIL(ctx, KP_EINVAL, "skinny GEMM needs K and leading dimensions multiple of 4");
  dim3 grid((N + SK_COLS - 1) / SK_COLS, (M + SK_ROWS - 1) / SK_ROWS);
  KpTimer timer(ctx, kp_ctx::T_CONV, st);
  skinny_nt_kernel<<<grid, SK_THREADS, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc);
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}

int kp_sgemm(kp_ctx* ctx, bool transb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
             int ldc, cudaStream_t st, int k_rows_b, bool split_k) {
  if (M <= 0 || N <= 0) return KP_OK;
  if (K % 4 != 0 || lda % 4 != 0 || ldb % 4 != 0) KP_FAIL(ctx, KP_EINVAL, "sgemm needs K and leading dimensions multiple of 4");
  dim3 grid((M + GM - 1) / GM, (N + GN - 1) / GN);
  // opt-in (the sum order becomes unordered): few output tiles and a long K -> cut K over blockIdx.z, partial sums
  // reduced with fp32 atomics into a zeroed C
  int k_per_split = 0;
  const long long tiles = (long long)grid.x * grid.y;
  if (split_k && tiles * 2 <= ctx->sm_count && K >= 64 * GK) {
    int splits = (int)((2LL * ctx->sm_count + tiles - 1) / tiles);
    if (splits > K / (8 * GK)) splits = K / (8 * GK);
    if (splits > 1) {
      k_per_split = ((K + splits - 1) / splits + GK - 1) / GK * GK;
      grid.z = (K + k_per_split - 1) / k_per_split;
      KP_CUDA(ctx, cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st));
    }
  }
  KpTimer timer(ctx, kp_ctx::T_CONV, st);
  if (transb)
    sgemm_kernel<true><<<grid, GT, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, K, k_per_split);
  else
    sgemm_kernel<false><<<grid, GT, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, k_rows_b >= 0 ? k_rows_b : K, k_per_split);
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
