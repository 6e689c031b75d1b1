// Fused "score all entities -> normalise -> contract with the entity table" pass (CUDA cores).
//
// For G query rows q_g it computes, WITHOUT materialising the [G, N] logits,
//   softmax mode (ComplEx, multiclass_nll_optimizer.py:123-135 / complex.py:59-86):
//       m_g = max_j z_gj,  l_g = sum_j exp(z_gj - m_g),  O_g = sum_j exp(z_gj - m_g) * E_j
//   sigmoid mode (ConvE, bce_optimizer.py:194-208 / conve.py:155-158):
//       O_g = sum_j sigmoid(z_gj) * E_j
// with z_gj = q_g . E_j over the N frozen entities.  O_g is exactly the term the backward
// pass of the reference needs (d loss / d query = sum_j G_gj E_j); the reference forms it with
// a dense [B, N+1] logit matrix and four SGEMMs per step.
//
// Each CTA owns 16 query rows and a strip of entity tiles (online softmax inside the strip);
// strips are merged by the caller (kp_flash_merge_* device helpers in kp_flash.cuh).
#include "kp_flash.cuh"
#include "kp_internal.h"

namespace {

constexpr int FQ = 16, FTHREADS = 256;

struct FlashK {
  int G, N, D, DP, FN, mode, n_tiles, tiles_per_strip, n_strips;
  const float* qmat;
  const float* ent;
  float* part_m;
  float* part_l;
  float* part_O;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// OV = float4 accumulators per thread = ceil(D / 64)
template <int OV>
__global__ void __launch_bounds__(FTHREADS) flash_simt_kernel(const FlashK p) {
  extern __shared__ __align__(16) float fsm[];
  const int D = p.D, DP = p.DP, FN = p.FN;
  float* Qs = fsm;                    // [FQ][D]
  float* Es = Qs + FQ * D;            // [2][FN][DP]
  float* Ps = Es + 2 * FN * DP;       // [FQ][FN]
  float* red = Ps + FQ * FN;          // [FQ][16] scratch
  float* row_m = red + FQ * 16;       // [FQ]
  float* row_l = row_m + FQ;          // [FQ]
  float* row_sc = row_l + FQ;         // [FQ]

  const int tid = threadIdx.x;
  const int g0 = blockIdx.y * FQ;
  const int strip = blockIdx.x;
  const int t0 = strip * p.tiles_per_strip, t1 = min(t0 + p.tiles_per_strip, p.n_tiles);
  const int q = tid >> 4, sub = tid & 15;

  for (int i = tid; i < FQ * D / 4; i += FTHREADS) {
    const int r = i / (D / 4), c = i % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g0 + r < p.G) v = *reinterpret_cast<const float4*>(p.qmat + (size_t)(g0 + r) * D + c * 4);
    *reinterpret_cast<float4*>(Qs + r * D + c * 4) = v;
  }
  if (tid < FQ) {
    row_m[tid] = -INFINITY;
    row_l[tid] = 0.f;
  }
  float4 o[OV];
#pragma unroll
  for (int v = 0; v < OV; ++v) o[v] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto load_tile = [&](int t, int buf) {
    const int j0 = t * FN;
    float* dst = Es + buf * FN * DP;
    for (int i = tid; i < FN * D / 4; i += FTHREADS) {
      const int r = i / (D / 4), c = i % (D / 4);
      if (j0 + r < p.N)
        cp_async16(dst + r * DP + c * 4, p.ent + (size_t)(j0 + r) * D + c * 4);
      else
        *reinterpret_cast<float4*>(dst + r * DP + c * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };

  if (t0 < t1) load_tile(t0, 0);
  for (int t = t0; t < t1; ++t) {
    const int buf = (t - t0) & 1;
    if (t + 1 < t1) {
      load_tile(t + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* E = Es + buf * FN * DP;
    const int j0 = t * FN;
    // ---- logits: thread (q, sub) handles entities sub, sub+16, ... of the tile ----
    for (int jj = sub; jj < FN; jj += 16) {
      const float4* qv = reinterpret_cast<const float4*>(Qs + q * D);
      const float4* ev = reinterpret_cast<const float4*>(E + jj * DP);
      float acc = 0.f;
#pragma unroll 4
      for (int c = 0; c < D / 4; ++c) {
        const float4 a = qv[c], b = ev[c];
        acc = __fmaf_rn(a.x, b.x, acc);
        acc = __fmaf_rn(a.y, b.y, acc);
        acc = __fmaf_rn(a.z, b.z, acc);
        acc = __fmaf_rn(a.w, b.w, acc);
      }
      Ps[q * FN + jj] = (j0 + jj < p.N) ? acc : -INFINITY;
    }
    __syncthreads();
    // ---- normalisation: 16 threads per row ----
    if (p.mode == KP_FLASH_SOFTMAX) {
      float mx = -INFINITY;
      for (int jj = sub; jj < FN; jj += 16) mx = fmaxf(mx, Ps[q * FN + jj]);
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_old = row_m[q];
      const float m_new = fmaxf(m_old, mx);
      float sum = 0.f;
      for (int jj = sub; jj < FN; jj += 16) {
        const float e = (m_new == -INFINITY) ? 0.f : expf(Ps[q * FN + jj] - m_new);
        Ps[q * FN + jj] = e;
        sum += e;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      __syncwarp();
      if (sub == 0) {
        const float sc = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
        row_sc[q] = sc;
        row_l[q] = row_l[q] * sc + sum;
        row_m[q] = m_new;
      }
    } else {
      for (int jj = sub; jj < FN; jj += 16) {
        const float z = Ps[q * FN + jj];
        Ps[q * FN + jj] = (z == -INFINITY) ? 0.f : 1.f / (1.f + expf(-z));
      }
      if (sub == 0) row_sc[q] = 1.f;
    }
    __syncthreads();
    // ---- O[q, :] = O[q, :] * scale + P[q, :] @ E ----
    {
      const float sc = row_sc[q];
#pragma unroll
      for (int v = 0; v < OV; ++v) {
        o[v].x *= sc; o[v].y *= sc; o[v].z *= sc; o[v].w *= sc;
      }
      for (int jj = 0; jj < FN; ++jj) {
        const float pj = Ps[q * FN + jj];
        const float* er = E + jj * DP;
#pragma unroll
        for (int v = 0; v < OV; ++v) {
          const int k = (v * 16 + sub) * 4;
          if (k < D) {
            const float4 e = *reinterpret_cast<const float4*>(er + k);
            o[v].x = __fmaf_rn(pj, e.x, o[v].x);
            o[v].y = __fmaf_rn(pj, e.y, o[v].y);
            o[v].z = __fmaf_rn(pj, e.z, o[v].z);
            o[v].w = __fmaf_rn(pj, e.w, o[v].w);
          }
        }
      }
    }
    __syncthreads();
  }
  const int g = g0 + q;
  if (g < p.G) {
    const size_t slot = (size_t)strip * p.G + g;
#pragma unroll
    for (int v = 0; v < OV; ++v) {
      const int k = (v * 16 + sub) * 4;
      if (k < D) *reinterpret_cast<float4*>(p.part_O + slot * D + k) = o[v];
    }
    if (sub == 0) {
      p.part_m[slot] = row_m[q];
      p.part_l[slot] = row_l[q];
    }
  }
}

}  // namespace

int kp_flash_plan(kp_ctx* ctx, int G, int* n_strips) {
  const int D = ctx->D, DP = D + 4;
  int FN = 32;
  while (FN > 4 && (size_t)(FQ * D + 2 * FN * DP + FQ * FN + FQ * 19) * 4 > 200 * 1024) FN >>= 1;
  const int n_tiles = (int)((ctx->N + FN - 1) / FN);
  const int n_qt = (G + FQ - 1) / FQ;
  int s = (2 * ctx->sm_count + n_qt - 1) / n_qt;
  if (s > 64) s = 64;
  if (s > n_tiles) s = n_tiles;
  if (s < 1) s = 1;
  const int tps = (n_tiles + s - 1) / s;
  *n_strips = (n_tiles + tps - 1) / tps;
  return FN;
}

int kp_flash_simt(kp_ctx* ctx, const float* qmat, int G, int mode, float* part_m, float* part_l, float* part_O,
                  cudaStream_t st) {
  if (G <= 0) return KP_OK;
  FlashK p;
  p.G = G;
  p.N = (int)ctx->N;
  p.D = ctx->D;
  p.DP = ctx->D + 4;
  p.mode = mode;
  p.FN = kp_flash_plan(ctx, G, &p.n_strips);
  p.n_tiles = (int)((ctx->N + p.FN - 1) / p.FN);
  p.tiles_per_strip = (p.n_tiles + p.n_strips - 1) / p.n_strips;
  p.qmat = qmat;
  p.ent = ctx->ent;
  p.part_m = part_m;
  p.part_l = part_l;
  p.part_O = part_O;
  const size_t smem = (size_t)(FQ * p.D + 2 * p.FN * p.DP + FQ * p.FN + FQ * 19) * 4;
  dim3 grid(p.n_strips, (G + FQ - 1) / FQ);
  const int ov = (p.D + 63) / 64;
#define KP_FLASH_CASE(V)                                                                                           \
  {                                                                                                                \
    KP_SMEM_ONCE(ctx, (flash_simt_kernel<V>), 208 * 1024);                                                         \
    KpTimer timer(ctx, kp_ctx::T_FLASH, st);                                                                        \
    flash_simt_kernel<V><<<grid, FTHREADS, smem, st>>>(p);                                                         \
  }
  if (ov <= 2) KP_FLASH_CASE(2)
  else if (ov <= 4) KP_FLASH_CASE(4)
  else if (ov <= 8) KP_FLASH_CASE(8)
  else if (ov <= 16) KP_FLASH_CASE(16)
  else if (ov <= 32) KP_FLASH_CASE(32)
  else KP_FAIL(ctx, KP_EUNSUPPORTED, "dim %d too large for the CUDA-core softmax pass", p.D);
#undef KP_FLASH_CASE
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}

int kp_flash_max_strips(kp_ctx* ctx) {
  int a = 1, b = 1;
  kp_flash_plan(ctx, 16, &a);
  kp_flash_umma_plan(ctx, 32, &b);
  return a > b ? a : b;
}

// Upper bound on n_strips(g) * g over every row count g <= G a caller may run the pass with:
// CUDA-core path (g < 32), tcgen05 path with less than a wave per strip (strips * clusters <= ~4 waves, plus
// rounding: <= 5 SMs x 128 rows) and with more (<= 8 strips, chosen against wave quantisation).
size_t kp_flash_part_rows(kp_ctx* ctx, int G) {
  int a = 1;
  kp_flash_plan(ctx, 16, &a);
  size_t rows = (size_t)a * (size_t)(G < 32 ? G : 32);
  const size_t few = (size_t)64 * G < (size_t)ctx->sm_count * 640 ? (size_t)64 * G : (size_t)ctx->sm_count * 640;
  if (few > rows) rows = few;
  if ((size_t)8 * G > rows) rows = (size_t)8 * G;
  {  // the bound on tiles per strip (kp_internal.h) sets a minimum number of strips (memory-capped: non-increasing in the rows)
    const size_t smin = (size_t)kp_umma_min_strips(ctx, G) + 1;  // + 1: even-tile rounding
    if (smin * (size_t)G > rows) rows = smin * (size_t)G;
  }
  return rows + (size_t)G;  // slack for the CUDA-core plan's rounding (strips = ceil(2 SMs / query tiles))
}

int kp_flash_run(kp_ctx* ctx, const float* qmat, int G, int mode, float* part_m, float* part_l, float* part_O,
                 cudaStream_t st, int* n_strips) {
  if (kp_flash_umma_usable(ctx, G)) {
    kp_flash_umma_plan(ctx, G, n_strips);
    return kp_flash_umma(ctx, qmat, G, mode, part_m, part_l, part_O, st);
  }
  kp_flash_plan(ctx, G, n_strips);
  return kp_flash_simt(ctx, qmat, G, mode, part_m, part_l, part_O, st);
}

// ---- diagnostic entry: the fused pass alone (tests compare it with an fp64 restatement) ----------
namespace {
__global__ void flash_merge_kernel(const float* part_m, const float* part_l, const float* part_O, int n_strips, int G, int D,
                                   int mode, float* out_m, float* out_l, float* out_O) {
  const int g = blockIdx.x;
  float M, L;
  if (mode == KP_FLASH_SIGMOID) {  // no normaliser: strips simply add
    L = 0.f;
    for (int s = 0; s < n_strips; ++s) L += part_l[(size_t)s * G + g];
    if (threadIdx.x == 0) {
      out_m[g] = 0.f;
      out_l[g] = L;
    }
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
      float acc = 0.f;
      for (int s = 0; s < n_strips; ++s) acc += part_O[((size_t)s * G + g) * D + k];
      out_O[(size_t)g * D + k] = acc;
    }
    return;
  }
  kp_flash_merge_stats(part_m, part_l, n_strips, G, g, M, L);
  if (threadIdx.x == 0) {
    out_m[g] = M;
    out_l[g] = L;
  }
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < n_strips; ++s) {
      const float m = part_m[(size_t)s * G + g];
      if (m != -INFINITY) acc += part_O[((size_t)s * G + g) * D + k] * expf(m - M);
    }
    out_O[(size_t)g * D + k] = acc;
  }
}
}  // namespace

extern "C" int kp_debug_contract(kp_ctx* ctx, int32_t n_rows, const float* queries, int32_t mode, float* out_m, float* out_l,
                                 float* out_O, void* stream) {
  if (!ctx || n_rows <= 0 || !queries || !out_m || !out_l || !out_O) return KP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int G = n_rows, D = ctx->D;
  const size_t SG = kp_flash_part_rows(ctx, G);
  const size_t Gpad = ((size_t)G + 127) / 128 * 128;
  size_t need = WsCursor::need(Gpad * D, 4) + 2 * WsCursor::need(SG, 4) + WsCursor::need(SG * D, 4);
  int rc = kp_ws_reserve(ctx, need);
  if (rc != KP_OK) return rc;
  WsCursor ws{ctx->ws, ctx->ws + ctx->ws_bytes};
  float* q = ws.take<float>(Gpad * D);
  float* pm = ws.take<float>(SG);
  float* pl = ws.take<float>(SG);
  float* pO = ws.take<float>(SG * D);
  KP_CUDA(ctx, cudaMemsetAsync(q, 0, Gpad * D * 4, st));
  KP_CUDA(ctx, cudaMemcpyAsync(q, queries, (size_t)G * D * 4, cudaMemcpyDeviceToDevice, st));
  int n_strips = 1;
  if ((rc = kp_flash_run(ctx, q, G, mode, pm, pl, pO, st, &n_strips)) != KP_OK) return rc;
  flash_merge_kernel<<<G, 128, 0, st>>>(pm, pl, pO, n_strips, G, D, mode, out_m, out_l, out_O);
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
