// fp32 GEMM  C[M, N] = A[M, K] * B[N, K]^T  on the tensor cores for the frozen ConvE network's Linear
// layer (conve.py:50-52,150: forward x = feat W^T with K = 9728; backward dfeat = dh W, run as
// dh (W^T)^T with K = 200), in the same bf16x3 split arithmetic as the fused pass: every fp32 operand
// is hi + lo (two bf16), products hi*hi + hi*lo + lo*hi accumulate in fp32 in TMEM.
//
// cta_group::2 pairs: 256 rows of A (two 128-row tiles, one per SM) x 128 rows of B per accumulator
// tile; each SM stages its own A rows and HALF of every B box; K-major operands by TMA through a
// 6-slot ring; the accumulator is double-buffered in TMEM so the epilogue (one thread per row,
// tcgen05.ld -> 128-bit global stores) overlaps the MMAs of the next tile.  B is split once
// (kp_umma_b_prepare, weights are frozen); A is split per call into workspace arena 1.
#include <cuda_bf16.h>

#include "kp_internal.h"
#include "kp_ptx.cuh"

namespace {

constexpr int GT = 192;
constexpr int SLOT = 32768;
constexpr int NSLOT = 6;

struct GCtl {
  uint64_t full[NSLOT], empty[NSLOT];
  uint64_t s_full[2], s_free[2];
  uint32_t tmem_base;
};
constexpr int STAGE = 4096;  // per epilogue warp: 32 rows x 32 columns of C on their way from TMEM to row-contiguous global stores
constexpr size_t G_SMEM = (size_t)NSLOT * SLOT + 4 * STAGE + sizeof(GCtl) + 1024;
static_assert(G_SMEM <= 232448, "shared memory budget of one CTA");

struct GK_ {
  int M, N, KB, ksteps, n_tiles, tiles_per_strip, kb_per_split;  // kb_per_split > 0: split-K over blockIdx.z, C accumulated with reductions
  long long part_stride;  // > 0 (with kb_per_split): split z stores its partial sums at C + z * part_stride instead (summed in a fixed order by the caller)
  int nt;  // columns of C per accumulator tile: 128, or up to 256 (a multiple of 16): A is then streamed once per 256 columns and an
           // MMA reads 4 KB of A + nt / 64 KB of B per nt / 2 cycles -- 64 B/clk at nt = 256 instead of 96, under the shared-memory port
  float* C;
  long long ldc;
};

__device__ __forceinline__ uint64_t udesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// TF32 = false: operands are bf16 hi / lo tables (64 k per 128-byte block, error ~2^-17 per product);
// TF32 = true: operands are tf32 hi / lo tables stored as fp32 (32 k per block, kind::tf32 at half the MMA rate, error ~2^-21).
template <bool TF32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GT, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap bh64_map, const __grid_constant__ CUtensorMap bl64_map,
                 const __grid_constant__ CUtensorMap ah_map, const __grid_constant__ CUtensorMap al_map, const GK_ p) {
  extern __shared__ uint8_t graw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(graw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = sm;
  uint8_t* stage = sm + (size_t)NSLOT * SLOT;
  GCtl* ctl = reinterpret_cast<GCtl*>(stage + 4 * STAGE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int strip = blockIdx.y, mtile = blockIdx.x;
  const int t0 = strip * p.tiles_per_strip;
  const int t1 = min(t0 + p.tiles_per_strip, p.n_tiles);
  const int ntile = t1 - t0;
  if (ntile <= 0) return;  // uniform over the pair
  const uint32_t crank = ptx::cluster_ctarank();
  const bool leader = crank == 0;
  const int kb0 = p.kb_per_split > 0 ? (int)blockIdx.z * p.kb_per_split : 0;
  const int kb1 = p.kb_per_split > 0 ? min(kb0 + p.kb_per_split, p.KB) : p.KB;
  if (kb0 >= kb1) return;  // uniform over the pair

  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&ctl->s_full[b], 1);
      ptx::mbar_init(&ctl->s_free[b], 256);
    }
    ptx::fence_barrier_init();
  }
  const bool wide = p.nt > 128;
  const uint32_t tm_stride = wide ? 256u : 128u, tm_cols = wide ? 512u : 256u;
  if (warp == 1) {
    ptx::tmem_alloc2(&ctl->tmem_base, tm_cols);
    ptx::tmem_relinquish2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tm = ctl->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&bh64_map);
      ptx::prefetch_tmap(&bl64_map);
      ptx::prefetch_tmap(&ah_map);
      ptx::prefetch_tmap(&al_map);
      uint32_t use = 0;
      auto load = [&](const CUtensorMap* hi, const CUtensorMap* lo, int col, int row, uint32_t lo_off, uint32_t bytes_pair) {
        const int s = use % NSLOT;
        ptx::mbar_wait(&ctl->empty[s], ((use / NSLOT) & 1) ^ 1);
        if (leader) ptx::mbar_arrive_expect_tx(&ctl->full[s], bytes_pair);
        const uint32_t bar = ptx::mapa_u32(ptx::smem_u32(&ctl->full[s]), 0);
        uint8_t* dst = ring + (size_t)s * SLOT;
        ptx::tma_load_2d_pair(dst, hi, bar, col, row);
        ptx::tma_load_2d_pair(dst + lo_off, lo, bar, col, row);
        ++use;
      };
      for (int i = 0; i < ntile; ++i)
        for (int kb = kb0; kb < kb1; ++kb) {
          load(&ah_map, &al_map, kb * (TF32 ? 32 : 64), mtile * 128, 16384, 2 * 32768);                          // my 128 rows of A
          if (!wide) {
            load(&bh64_map, &bl64_map, kb * (TF32 ? 32 : 64), (t0 + i) * 128 + (int)crank * 64, 8192, 2 * 16384);  // my half of the B tile
          } else {  // my nt / 2 <= 128 rows of the B tile as two 64-row boxes (rows past my half are loaded and not read)
            const int s = use % NSLOT, col = kb * (TF32 ? 32 : 64), row = (t0 + i) * p.nt + (int)crank * (p.nt / 2);
            ptx::mbar_wait(&ctl->empty[s], ((use / NSLOT) & 1) ^ 1);
            if (leader) ptx::mbar_arrive_expect_tx(&ctl->full[s], 2 * 32768);
            const uint32_t bar = ptx::mapa_u32(ptx::smem_u32(&ctl->full[s]), 0);
            uint8_t* dst = ring + (size_t)s * SLOT;
            ptx::tma_load_2d_pair(dst, &bh64_map, bar, col, row);
            ptx::tma_load_2d_pair(dst + 8192, &bh64_map, bar, col, row + 64);
            ptx::tma_load_2d_pair(dst + 16384, &bl64_map, bar, col, row);
            ptx::tma_load_2d_pair(dst + 24576, &bl64_map, bar, col, row + 64);
            ++use;
          }
        }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      const uint32_t fmt = TF32 ? 2u : 1u;  // cute::UMMA::F16F32Format: BF16 = 1, TF32 = 2
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (((uint32_t)p.nt >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t ring_a = ptx::smem_u32(ring);
      const uint64_t DK = udesc(0, 16, 1024);
      uint32_t use = 0;
      for (int i = 0; i < ntile; ++i) {
        const int sb = i & 1;
        const uint32_t d_s = tm + sb * tm_stride;
        if (i >= 2) {
          ptx::mbar_wait_cluster(&ctl->s_free[sb], ((i >> 1) - 1) & 1);
          ptx::tc_fence_after();
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&ctl->full[use % NSLOT], (use / NSLOT) & 1);
          ptx::mbar_wait(&ctl->full[(use + 1) % NSLOT], ((use + 1) / NSLOT) & 1);
          ptx::tc_fence_after();
          const uint32_t a_hi = ring_a + (use % NSLOT) * SLOT, a_lo = a_hi + 16384;
          const uint32_t b_hi = ring_a + ((use + 1) % NSLOT) * SLOT, b_lo = b_hi + (wide ? 16384 : 8192);
          const uint64_t ah = DK + (a_hi >> 4), al = DK + (a_lo >> 4), bh = DK + (b_hi >> 4), bl = DK + (b_lo >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kb * 4 + kk >= p.ksteps) break;  // only zero padding beyond K
            if (TF32) {
              ptx::umma2_tf32(d_s, ah + kk * 2, bh + kk * 2, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
              ptx::umma2_tf32(d_s, ah + kk * 2, bl + kk * 2, idesc, 1u);
              ptx::umma2_tf32(d_s, al + kk * 2, bh + kk * 2, idesc, 1u);
            } else {
              ptx::umma2_bf16(d_s, ah + kk * 2, bh + kk * 2, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
              ptx::umma2_bf16(d_s, ah + kk * 2, bl + kk * 2, idesc, 1u);
              ptx::umma2_bf16(d_s, al + kk * 2, bh + kk * 2, idesc, 1u);
            }
          }
          ptx::umma2_commit_mc(&ctl->empty[use % NSLOT], 3);
          ptx::umma2_commit_mc(&ctl->empty[(use + 1) % NSLOT], 3);
          use += 2;
        }
        ptx::umma2_commit_mc(&ctl->s_full[sb], 3);
      }
    }
  } else {
    // A thread reads one row of the accumulator from TMEM; the warp's 32 x 32 block goes through shared memory (16-byte chunks
    // XOR-swizzled by the row: conflict-free both ways) so that every store instruction writes 4 rows x 128 contiguous bytes of C
    // instead of 16 bytes in each of 32 rows.
    const int sub = warp & 3;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    const long long m0 = (long long)mtile * 128 + sub * 32;
    float4* st4 = reinterpret_cast<float4*>(stage + (size_t)sub * STAGE);
    const int rr = lane >> 3, cc = lane & 7;
    const uint32_t s_free_leader0 = ptx::mapa_u32(ptx::smem_u32(&ctl->s_free[0]), 0);
    for (int i = 0; i < ntile; ++i) {
      const int sb = i & 1;
      const int n0 = (t0 + i) * p.nt;
      ptx::mbar_wait(&ctl->s_full[sb], (i >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t s_addr = tm + sb * tm_stride + lane_off;
#pragma unroll 1
      for (int c0 = 0; c0 < p.nt; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(s_addr + c0, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          st4[lane * 8 + (c ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]),
                                                          __uint_as_float(r[4 * c + 2]), __uint_as_float(r[4 * c + 3]));
        __syncwarp();
        const int col = n0 + c0 + cc * 4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int R = it * 4 + rr;
          const float4 v = st4[R * 8 + (cc ^ (R & 7))];
          if (m0 + R < p.M && col < p.N) {

            float* dst = p.C + (m0 + R) * p.ldc + col;
            if (p.part_stride > 0)
              *reinterpret_cast<float4*>(dst + (long long)blockIdx.z * p.part_stride) = v;
            else if (p.kb_per_split > 0)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            else
              *reinterpret_cast<float4*>(dst) = v;
          }
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive_cluster(s_free_leader0 + 8u * sb);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc2(tm, tm_cols);
}

// fp32 [rows, cols] (ld) -> bf16 hi / lo [rows_pad, cols_pad] (zero padded)
__global__ void split2_kernel(const float* __restrict__ src, long long rows, int cols, long long ld, long long rows_pad, int cols_pad,
                              __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long total = rows_pad * (long long)(cols_pad / 2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (cols_pad / 2);
    const int c = (int)(i % (cols_pad / 2)) * 2;
    float a = 0.f, b = 0.f;
    if (r < rows) {
      if (c < cols) a = src[r * ld + c];
      if (c + 1 < cols) b = src[r * ld + c + 1];
    }
    const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
    reinterpret_cast<uint32_t*>(hi)[i] = pack2(ah, bh);
    reinterpret_cast<uint32_t*>(lo)[i] = pack2(__float2bfloat16_rn(a - __bfloat162float(ah)), __float2bfloat16_rn(b - __bfloat162float(bh)));
  }
}
// dst[c, r] = src[r, c]
__global__ void transpose_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
    if (r0 + j < rows && c0 + threadIdx.x < cols) t[j][threadIdx.x] = src[(size_t)(r0 + j) * cols + c0 + threadIdx.x];
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y)
    if (c0 + j < cols && r0 + threadIdx.x < rows) dst[(size_t)(c0 + j) * rows + r0 + threadIdx.x] = t[threadIdx.x][j];
}

// fp32 src [cols, rows] (ld): the TRANSPOSE -> bf16 hi / lo [rows_pad, cols_pad] (zero padded); 32x32 tiles through shared memory
__global__ void split2t_kernel(const float* __restrict__ src, long long rows, int cols, long long ld, long long rows_pad, int cols_pad,
                               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float t[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {  // src row = c0 + j (a column of the result), src col = r0 + x
    const long long sr = c0 + j, sc = r0 + threadIdx.x;
    t[j][threadIdx.x] = (sr < cols && sc < rows) ? src[sr * ld + sc] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {  // result row r0 + j, col c0 + x
    const long long r = r0 + j;
    const int c = c0 + threadIdx.x;
    if (r < rows_pad && c < cols_pad) {
      const float a = t[threadIdx.x][j];
      const __nv_bfloat16 ah = __float2bfloat16_rn(a);
      hi[r * cols_pad + c] = ah;
      lo[r * cols_pad + c] = __float2bfloat16_rn(a - __bfloat162float(ah));
    }
  }
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// fp32 [rows, cols] (ld) -> tf32 hi / lo [rows_pad, cols_pad] as fp32 words (zero padded): x = hi + lo + O(2^-22 |x|)
__global__ void split2_tf32_kernel(const float* __restrict__ src, long long rows, int cols, long long ld, long long rows_pad, int cols_pad,
                                   float* __restrict__ hi, float* __restrict__ lo) {
  const long long total = rows_pad * (long long)cols_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols_pad;
    const int c = (int)(i - r * cols_pad);
    const float a = (r < rows && c < cols) ? src[r * ld + c] : 0.f;
    const float h = to_tf32(a);
    hi[i] = h;
    lo[i] = to_tf32(a - h);
  }
}
// the same from the TRANSPOSE: src is [cols, rows] (ld)
__global__ void split2t_tf32_kernel(const float* __restrict__ src, long long rows, int cols, long long ld, long long rows_pad, int cols_pad,
                                    float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float t[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long sr = c0 + j, sc = r0 + threadIdx.x;
    t[j][threadIdx.x] = (sr < cols && sc < rows) ? src[sr * ld + sc] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long r = r0 + j;
    const int c = c0 + threadIdx.x;
    if (r < rows_pad && c < cols_pad) {
      const float a = t[threadIdx.x][j];
      const float h = to_tf32(a);
      hi[r * cols_pad + c] = h;
      lo[r * cols_pad + c] = to_tf32(a - h);
    }
  }
}

}  // namespace

// C[M, N] = opA(A) * opB(B)^T with both operands given per call (fp32, split into workspace arena 1 behind `ws_offset`):
// opA(A) = A [M, K] (lda) or, transA, the transpose of A [K, M];  opB(B) = B [N, K] (ldb) or, transB, the transpose of B [K, N].
// A long K with few output tiles is cut over blockIdx.z and accumulated with fp32 reductions into a zeroed C.
int kp_gemm_umma_dyn(kp_ctx* ctx, const float* A, long long lda, bool transA, int M, const float* B, long long ldb, bool transB,
                     int N, int K, float* C, long long ldc, size_t ws_offset, cudaStream_t st, bool tf32) {
  if (M <= 0 || N <= 0 || K <= 0) return KP_OK;
  const int Nc = (N + 3) & ~3;  // columns are stored four at a time: the (zero) padding up to Nc is written too
  if (ldc % 4 != 0 || ldc < Nc) KP_FAIL(ctx, KP_EINVAL, "tcgen05 GEMM needs ldc a multiple of 4 and >= N rounded up to 4");
  const int n_mt = ((M + 255) / 256) * 2;
  const long long Mpad = (long long)n_mt * 128, Npad = ((long long)N + 127) / 128 * 128;
  const int epb = tf32 ? 32 : 64, esz = tf32 ? 4 : 2;  // elements per 128-byte k-block, bytes per element
  const int Kpad = (K + epb - 1) / epb * epb;
  const size_t abytes = ((size_t)Mpad * Kpad * esz + 1023) & ~size_t(1023), bbytes = ((size_t)Npad * Kpad * esz + 1023) & ~size_t(1023);
  int rc;
  ws_offset = (ws_offset + 1023) & ~size_t(1023);
  if ((rc = kp_ws_reserve(ctx, ws_offset + 2 * abytes + 2 * bbytes + 2048, 1)) != KP_OK) return rc;
  char* base = ctx->ws_arena[1] + ws_offset;
  void *ah = base, *al = base + abytes, *bh = base + 2 * abytes, *bl = base + 2 * abytes + bbytes;
  auto split = [&](const float* src, long long ld, bool trans, long long rows, long long rows_pad, void* h, void* l) {
    const dim3 tgrid((unsigned)((rows_pad + 31) / 32), (unsigned)((Kpad + 31) / 32));
    if (tf32) {
      if (trans) split2t_tf32_kernel<<<tgrid, dim3(32, 8), 0, st>>>(src, rows, K, ld, rows_pad, Kpad, (float*)h, (float*)l);
      else split2_tf32_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(src, rows, K, ld, rows_pad, Kpad, (float*)h, (float*)l);
    } else {
      if (trans) split2t_kernel<<<tgrid, dim3(32, 8), 0, st>>>(src, rows, K, ld, rows_pad, Kpad, (__nv_bfloat16*)h, (__nv_bfloat16*)l);
      else split2_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(src, rows, K, ld, rows_pad, Kpad, (__nv_bfloat16*)h, (__nv_bfloat16*)l);
    }
  };
  split(A, lda, transA, M, Mpad, ah, al);
  split(B, ldb, transB, N, Npad, bh, bl);
  KP_LAUNCHED(ctx, 2);
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap ah_map, al_map, bh_map, bl_map;
  if ((rc = kp_encode_2d(ctx, &ah_map, ah, dt, esz, Mpad, Kpad, Kpad, 128, epb, true)) != KP_OK) return rc;
  if ((rc = kp_encode_2d(ctx, &al_map, al, dt, esz, Mpad, Kpad, Kpad, 128, epb, true)) != KP_OK) return rc;
  if ((rc = kp_encode_2d(ctx, &bh_map, bh, dt, esz, Npad, Kpad, Kpad, 64, epb, true)) != KP_OK) return rc;
  if ((rc = kp_encode_2d(ctx, &bl_map, bl, dt, esz, Npad, Kpad, Kpad, 64, epb, true)) != KP_OK) return rc;
  GK_ p;
  p.M = M;
  p.N = Nc;
  p.KB = Kpad / epb;
  p.ksteps = (K + epb / 4 - 1) / (epb / 4);  // one MMA covers a quarter of a k-block (32 bytes): 16 bf16 or 8 tf32
  p.nt = 128;
  p.n_tiles = (N + 127) / 128;
  const int s = kp_plan_strips(n_mt / 2, ctx->sm_count / 2, p.n_tiles);
  p.tiles_per_strip = (p.n_tiles + s - 1) / s;
  const int n_strips = (p.n_tiles + p.tiles_per_strip - 1) / p.tiles_per_strip;
  int ksplit = 1;
  const long long clusters = (long long)(n_mt / 2) * n_strips;
  if (clusters * 2 <= ctx->sm_count / 2 && p.KB >= 16) {  // under half a wave and a long K: split it
    ksplit = (int)((ctx->sm_count / 2) / clusters);
    if (ksplit > p.KB / 8) ksplit = p.KB / 8;
    if (ksplit < 1) ksplit = 1;
  }
  p.kb_per_split = ksplit > 1 ? (p.KB + ksplit - 1) / ksplit : 0;
  if (ksplit > 1) {
    ksplit = (p.KB + p.kb_per_split - 1) / p.kb_per_split;
    KP_CUDA(ctx, cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)Nc * 4, (size_t)M, st));
  }
  p.part_stride = 0;
  p.C = C;
  p.ldc = ldc;
  KP_SMEM_ONCE(ctx, (gemm_umma_kernel<false>), G_SMEM);
  KP_SMEM_ONCE(ctx, (gemm_umma_kernel<true>), G_SMEM);
  {
    KpTimer timer(ctx, kp_ctx::T_CONV, st);
    if (tf32) gemm_umma_kernel<true><<<dim3(n_mt, n_strips, ksplit), GT, G_SMEM, st>>>(bh_map, bl_map, ah_map, al_map, p);
    else gemm_umma_kernel<false><<<dim3(n_mt, n_strips, ksplit), GT, G_SMEM, st>>>(bh_map, bl_map, ah_map, al_map, p);
  }
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}

// B[N, K] fp32 (row-major, or its transpose when `transpose`: then the source is [K, N]) -> split tables + half-tile maps
int kp_umma_b_prepare(kp_ctx* ctx, const float* B, int N, int K, bool transpose, kp_umma_b* out, cudaStream_t st) {
  const long long Npad = ((long long)N + 127) / 128 * 128;
  const int Kpad = (K + 63) / 64 * 64;
  void *h = nullptr, *l = nullptr, *tmp = nullptr;
  if (cudaMalloc(&h, (size_t)Npad * Kpad * 2) != cudaSuccess || cudaMalloc(&l, (size_t)Npad * Kpad * 2) != cudaSuccess)
    KP_FAIL(ctx, KP_ENOMEM, "cannot allocate split GEMM operand (%lld x %d)", Npad, Kpad);
  ctx->owned.push_back(h);
  ctx->owned.push_back(l);
  const float* src = B;
  if (transpose) {
    if (cudaMalloc(&tmp, (size_t)N * K * 4) != cudaSuccess) KP_FAIL(ctx, KP_ENOMEM, "cannot allocate transpose scratch");
    transpose_kernel<<<dim3((N + 31) / 32, (K + 31) / 32), dim3(32, 8), 0, st>>>(B, K, N, (float*)tmp);  // src [K, N] -> [N, K]
    KP_LAUNCHED(ctx, 1);
    src = (const float*)tmp;
  }
  split2_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(src, N, K, K, Npad, Kpad, (__nv_bfloat16*)h, (__nv_bfloat16*)l);
  KP_LAUNCHED(ctx, 1);
  if (tmp) {
    KP_CUDA(ctx, cudaStreamSynchronize(st));
    cudaFree(tmp);
  }
  int rc;
  if ((rc = kp_encode_2d(ctx, &out->hi64, h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Npad, Kpad, Kpad, 64, 64, true)) != KP_OK) return rc;
  if ((rc = kp_encode_2d(ctx, &out->lo64, l, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Npad, Kpad, Kpad, 64, 64, true)) != KP_OK) return rc;
  out->N = N;
  out->K = K;
  out->Kpad = Kpad;
  out->ready = true;
  return KP_OK;
}

size_t kp_gemm_umma_a_bytes(int M, const kp_umma_b& B) {  // bytes of ONE split half (hi or lo) of an A operand of M rows
  const long long Mpad = (long long)((M + 255) / 256) * 256;
  return ((size_t)Mpad * B.Kpad * 2 + 1023) & ~size_t(1023);
}

int kp_gemm_umma(kp_ctx* ctx, const float* A, long long lda, int M, const kp_umma_b& B, float* C, long long ldc, size_t ws_offset,
                 cudaStream_t st) {
  if (M <= 0) return KP_OK;
  const long long Mpad = (long long)((M + 255) / 256) * 256;
  const size_t abytes = kp_gemm_umma_a_bytes(M, B);
  int rc;
  ws_offset = (ws_offset + 1023) & ~size_t(1023);  // the caller's own scratch at the start of arena 1 (may hold A itself)
  if ((rc = kp_ws_reserve(ctx, ws_offset + 2 * abytes + 2048, 1)) != KP_OK) return rc;
  __nv_bfloat16* ah = reinterpret_cast<__nv_bfloat16*>(ctx->ws_arena[1] + ws_offset);
  __nv_bfloat16* al = reinterpret_cast<__nv_bfloat16*>(ctx->ws_arena[1] + ws_offset + abytes);
  split2_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(A, M, B.K, lda, Mpad, B.Kpad, ah, al);
  KP_LAUNCHED(ctx, 1);
  return kp_gemm_umma_split(ctx, ah, al, M, B, C, ldc, st);
}

// The same with A already split by its producer: bf16 hi / lo [M rounded up to 256, B.Kpad], columns K..Kpad zero (rows beyond M
// are read but their results are not stored).
// Few rows and a long K (the forward Linear layer of an explain-sized batch: one pair of CTAs, 152 k-blocks): K is cut over up to half
// the SMs' worth of CTA pairs; every part stores its own partial result and the caller adds them in a fixed order (no atomics: the
// sum does not depend on the schedule).
int kp_gemm_umma_ksplit(const kp_ctx* ctx, int M, const kp_umma_b& B) {
  const int clusters = (M + 255) / 256, KB = B.Kpad / 64;
  if (!ctx->gemm_ksplit || B.N > 256 || KB < 16) return 1;
  int ks = (ctx->sm_count / 2) / clusters;
  if (ks > KB / 4) ks = KB / 4;
  if (ks < 2) return 1;
  const int per = (KB + ks - 1) / ks;
  return (KB + per - 1) / per;
}

int kp_gemm_umma_split(kp_ctx* ctx, const void* ah, const void* al, int M, const kp_umma_b& B, float* C, long long ldc, cudaStream_t st,
                       float* parts) {
  if (M <= 0) return KP_OK;
  if (B.N % 4 != 0 || ldc % 4 != 0) KP_FAIL(ctx, KP_EINVAL, "tcgen05 GEMM needs N and ldc multiple of 4");
  const int n_mt = ((M + 255) / 256) * 2;
  const long long Mpad = (long long)n_mt * 128;
  int rc;
  CUtensorMap ah_map, al_map;
  if ((rc = kp_encode_2d(ctx, &ah_map, const_cast<void*>(ah), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Mpad, B.Kpad, B.Kpad, 128, 64, true)) != KP_OK) return rc;
  if ((rc = kp_encode_2d(ctx, &al_map, const_cast<void*>(al), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Mpad, B.Kpad, B.Kpad, 128, 64, true)) != KP_OK) return rc;
  GK_ p;
  p.M = M;
  p.N = B.N;
  p.KB = B.Kpad / 64;
  p.ksteps = (B.K + 15) / 16;
  p.kb_per_split = 0;
  // wide accumulator tiles (option gemm_wide): one tile of N rounded up to 16 when N <= 256 (the forward Linear layer: N = 200 ->
  // 208 columns instead of two tiles of 128), 256 columns otherwise (the backward one: N = 9728)
  p.nt = (ctx->gemm_wide && B.N > 128) ? (B.N <= 256 ? ((B.N + 15) / 16) * 16 : 256) : 128;
  p.n_tiles = (B.N + p.nt - 1) / p.nt;
  const int s = kp_plan_strips(n_mt / 2, ctx->sm_count / 2, p.n_tiles);
  p.tiles_per_strip = (p.n_tiles + s - 1) / s;
  const int n_strips = (p.n_tiles + p.tiles_per_strip - 1) / p.tiles_per_strip;
  p.C = C;
  p.ldc = ldc;
  p.part_stride = 0;
  const int ks = parts ? kp_gemm_umma_ksplit(ctx, M, B) : 1;  // parts: room for kp_gemm_umma_ksplit() x [M, ldc]; then C is NOT written
  if (ks > 1) {
    if (n_strips != 1) KP_FAIL(ctx, KP_EINVAL, "split-K GEMM expects one accumulator tile per row block");
    p.kb_per_split = (p.KB + ks - 1) / ks;
    p.part_stride = (long long)M * ldc;
    p.C = parts;
  }
  KP_SMEM_ONCE(ctx, (gemm_umma_kernel<false>), G_SMEM);
  {
    KpTimer timer(ctx, kp_ctx::T_CONV, st);
    gemm_umma_kernel<false><<<dim3(n_mt, n_strips, ks), GT, G_SMEM, st>>>(B.hi64, B.lo64, ah_map, al_map, p);
  }
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
