// Batched TransE mimic post-training: one CTA per candidate explanation runs ALL epochs and
// steps of KelpiePairwiseRankingOptimizer (pairwise_ranking_optimizer.py:139-203) for its own
// mimic row in one launch.  The mimic row and the Adam moments live in shared memory; each
// warp takes one (positive, negative) row pair at a time with coalesced 128-bit row loads,
// reduces the two distances with warp shuffles, accumulates d(loss)/d(mimic) in registers and
// the step ends with a fused Adam update (torch.optim.Adam defaults, :46).  Only the mimic
// row's gradient is ever formed (the reference builds a dense [N+1, D] gradient per step).
//
// Loss per step (transe.py:67-75, regularizers.py:15-22):
//   mean_i max(0, pos_i - neg_i + margin) + (L2(pos factors) + L2(neg factors)) / 2
#include "kp_internal.h"
#include "kp_ptx.cuh"

namespace {

constexpr int TT_THREADS = 256;
constexpr int TT_WARPS = TT_THREADS / 32;

struct TrainK {
  int N, D, norm, C, static_epochs;
  const float* ent;
  const float* rel;
  const int64_t* row_off;
  const int32_t* rows_per_epoch;
  const int32_t* pos;
  const int32_t* neg;
  const float* init;
  float* out;
  const float2* adam;  // [epochs * steps per epoch] bias-correction table
  const int64_t* fact_off;  // compact tables (kp_pt_batch)
  const int32_t* facts;
  const uint16_t* pos_idx;
  const int32_t* neg_code;
  kp_hp hp;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

// Adam's bias corrections depend on the step number only: step_size = lr / (1 - beta1^t), sqrt(1 - beta2^t),
// evaluated in double as torch does on the host (optim/adam.py _single_tensor_adam).  One table for the whole batch
// instead of two double-precision pow() per thread and step inside the training kernel.
__global__ void adam_table_kernel(int steps, float lr, float beta1, float beta2, float2* __restrict__ tab) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= steps) return;
  const double bc1 = 1.0 - pow((double)beta1, (double)(t + 1));
  const double bc2 = 1.0 - pow((double)beta2, (double)(t + 1));
  tab[t] = make_float2((float)((double)lr / bc1), (float)sqrt(bc2));
}

// VPL: float4 vectors per lane (D <= 128 * VPL).  U: rows a warp has in flight at once (independent gathers).
// CP: compact index tables (kp_pt_batch.pos_idx / neg_code / facts) instead of pos / neg.
template <int VPL, int U, bool CP, bool L2N, bool EX>  // L2N: Euclidean norm (every shipped config), else L1; EX: D == 128 * VPL
__global__ void __launch_bounds__(TT_THREADS) transe_train_kernel(const TrainK p) {
  extern __shared__ float sm[];
  const int D = p.D;
  float* eM = sm;            // [D] mimic row
  float* am = eM + D;        // [D] Adam exp_avg
  float* av = am + D;        // [D] Adam exp_avg_sq
  float* gw = av + D;        // [TT_WARPS][D] per-warp gradient partials
  __shared__ int s_cnt[TT_WARPS];

  const int c = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = p.rows_per_epoch[c];
  const int64_t base0 = p.row_off[c];
  const int32_t* facts = CP ? p.facts + p.fact_off[c] * 3 : nullptr;
  for (int k = tid; k < D; k += TT_THREADS) {
    eM[k] = p.init[(size_t)c * D + k];
    am[k] = 0.f;
    av[k] = 0.f;
  }
  __syncthreads();

  const int M = p.N;
  // rows are addressed in float4 units with 32-bit offsets: one IMAD per gathered vector instead of 64-bit pointer
  // arithmetic and generic (shared-or-global) loads
  const float4* __restrict__ ent4 = reinterpret_cast<const float4*>(p.ent);
  const unsigned d4 = (unsigned)D >> 2;
  const float4* __restrict__ rel4 = reinterpret_cast<const float4*>(p.rel);
  auto ldrow = [&](int e, int k) -> float4 {
    if (e == M) return *reinterpret_cast<const float4*>(eM + k);  // warp-uniform
    return ent4[(unsigned)e * d4 + (unsigned)(k >> 2)];
  };
  const int bs = p.hp.batch_size;
  const int spe = (n + bs - 1) / bs;  // steps per epoch
  const long long n_steps = (long long)p.hp.epochs * spe;

  // The index rows of a step do not depend on the mimic row: they are fetched one step ahead, LPS lanes per row
  // (full tables: pos h, r, t, neg h, r, t; compact: the positive fact's h, r, t and the corruption code).
  constexpr int LPS = CP ? 4 : 6;
  constexpr int PF = 32 / LPS;
  // (epoch, first row) of a step advance incrementally: no 64-bit division per step
  auto row_base = [&](int ep, int b0) -> int64_t { return base0 + (p.static_epochs ? 0 : (int64_t)ep * n) + b0; };
  auto prefetch = [&](int ep, int b0) -> int {
    if (ep >= p.hp.epochs || lane >= LPS * PF) return 0;
    const int64_t rb = row_base(ep, b0);
    const int B = min(bs, n - b0);
    const int i = warp + TT_WARPS * (lane / LPS), k = lane % LPS;
    if (i >= B) return 0;
    if (CP) return (k == 3) ? p.neg_code[rb + i] : facts[(int)p.pos_idx[rb + i] * 3 + k];
    return (k < 3 ? p.pos : p.neg)[(rb + i) * 3 + (k % 3)];
  };
  int ep = 0, b0 = 0;
  int pf = (n_steps > 0) ? prefetch(0, 0) : 0;

  for (long long step = 0; step < n_steps; ++step) {
    const int64_t rb = row_base(ep, b0);
    const int B = min(bs, n - b0);
    int ep1 = ep, b1 = b0 + bs;  // the next step
    if (b1 >= n) ep1 = ep + 1, b1 = 0;
    float4 g[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) g[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
    int slot = 0;
    for (int i = warp; i < B; i += U * TT_WARPS, slot += U) {
      int h[U], r[U], t[U], h2[U], t2[U];
      bool valid[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int iu = i + u * TT_WARPS;
        valid[u] = iu < B;  // warp-uniform
        if (slot + u < PF) {
          const int l0 = LPS * (slot + u);
          h[u] = __shfl_sync(0xffffffffu, pf, l0 + 0);
          r[u] = __shfl_sync(0xffffffffu, pf, l0 + 1);
          t[u] = __shfl_sync(0xffffffffu, pf, l0 + 2);
          if (CP) {
            const int code = __shfl_sync(0xffffffffu, pf, l0 + 3);
            h2[u] = (code < 0) ? (code & 0x7fffffff) : h[u];
            t2[u] = (code < 0) ? t[u] : code;
          } else {
            h2[u] = __shfl_sync(0xffffffffu, pf, l0 + 3);
            t2[u] = __shfl_sync(0xffffffffu, pf, l0 + 5);
          }
        } else if (valid[u]) {
          if (CP) {
            const int32_t* f = facts + (int)p.pos_idx[rb + iu] * 3;
            const int code = p.neg_code[rb + iu];
            h[u] = f[0], r[u] = f[1], t[u] = f[2];
            h2[u] = (code < 0) ? (code & 0x7fffffff) : h[u];
            t2[u] = (code < 0) ? t[u] : code;
          } else {
            const int32_t* pr = p.pos + (rb + iu) * 3;
            const int32_t* nr = p.neg + (rb + iu) * 3;
            h[u] = pr[0], r[u] = pr[1], t[u] = pr[2], h2[u] = nr[0], t2[u] = nr[2];
          }
        }
        if (!valid[u]) h[u] = t[u] = h2[u] = t2[u] = M, r[u] = 0;  // harmless addresses; contributions masked below
      }
      float4 dp[U][VPL], dn[U][VPL];
      float sp[U], sn[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        sp[u] = sn[u] = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int k = (v * 32 + lane) * 4;
          dp[u][v] = dn[u][v] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (EX || k < D) {
            const float4 a = ldrow(h[u], k), rr = rel4[(unsigned)r[u] * d4 + (unsigned)(k >> 2)], b = ldrow(t[u], k);
            const float4 a2 = ldrow(h2[u], k), b2 = ldrow(t2[u], k);
            float4& x = dp[u][v];
            float4& y = dn[u][v];
            x.x = __fsub_rn(__fadd_rn(a.x, rr.x), b.x);
            x.y = __fsub_rn(__fadd_rn(a.y, rr.y), b.y);
            x.z = __fsub_rn(__fadd_rn(a.z, rr.z), b.z);
            x.w = __fsub_rn(__fadd_rn(a.w, rr.w), b.w);
            y.x = __fsub_rn(__fadd_rn(a2.x, rr.x), b2.x);
            y.y = __fsub_rn(__fadd_rn(a2.y, rr.y), b2.y);
            y.z = __fsub_rn(__fadd_rn(a2.z, rr.z), b2.z);
            y.w = __fsub_rn(__fadd_rn(a2.w, rr.w), b2.w);
            if (L2N) {
              sp[u] += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
              sn[u] += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
            } else {
              sp[u] += fabsf(x.x) + fabsf(x.y) + fabsf(x.z) + fabsf(x.w);
              sn[u] += fabsf(y.x) + fabsf(y.y) + fabsf(y.z) + fabsf(y.w);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        sp[u] = warp_sum(sp[u]);
        sn[u] = warp_sum(sn[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float spu = sp[u], snu = sn[u];
        if (L2N) {
          spu = sqrtf(spu);
          snu = sqrtf(snu);
        }
        const bool active = valid[u] && (spu - snu + p.hp.margin) > 0.f;
        const float cp = (float)((h[u] == M) - (t[u] == M));
        const float cn = (float)((h2[u] == M) - (t2[u] == M));
        if (valid[u]) cnt += (h[u] == M) + (t[u] == M) + (h2[u] == M) + (t2[u] == M);
        if (active) {
          // d||d||_2/dd = d/||d|| (0 at the origin, as torch); d||d||_1/dd = sign(d).  The mimic row's coefficient c is
          // -1, 0 or +1, so c / ||d|| == c * (1 / ||d||) exactly (one correctly rounded reciprocal, no division).
          const float ip = L2N ? ((spu > 0.f) ? cp * __frcp_rn(spu) : 0.f) : cp;
          const float in = L2N ? ((snu > 0.f) ? cn * __frcp_rn(snu) : 0.f) : cn;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const float4 x = dp[u][v], y = dn[u][v];
            if (L2N) {
              g[v].x += ip * x.x - in * y.x;
              g[v].y += ip * x.y - in * y.y;
              g[v].z += ip * x.z - in * y.z;
              g[v].w += ip * x.w - in * y.w;
            } else {
              g[v].x += ip * sgn(x.x) - in * sgn(y.x);
              g[v].y += ip * sgn(x.y) - in * sgn(y.y);
              g[v].z += ip * sgn(x.z) - in * sgn(y.z);
              g[v].w += ip * sgn(x.w) - in * sgn(y.w);
            }
          }
        }
      }
    }
    pf = prefetch(ep1, b1);  // in flight across the barrier and the optimizer update
    ep = ep1, b0 = b1;
    // all warps have finished READING eM for this step before anyone updates it
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int k = (v * 32 + lane) * 4;
      if (k < D) *reinterpret_cast<float4*>(gw + warp * D + k) = g[v];
    }
    if (lane == 0) s_cnt[warp] = cnt;
    const float2 bc = p.adam[step];  // (lr / bias_correction1, sqrt(bias_correction2))
    __syncthreads();
    int total_cnt = 0;
#pragma unroll
    for (int w = 0; w < TT_WARPS; ++w) total_cnt += s_cnt[w];
    const float reg = p.hp.reg_weight * (float)total_cnt / (3.f * (float)B * (float)D);
    const float invB = 1.f / (float)B;
    for (int k = tid; k < D; k += TT_THREADS) {
      float gs = 0.f;
#pragma unroll
      for (int w = 0; w < TT_WARPS; ++w) gs += gw[w * D + k];
      const float e = eM[k];
      const float grad = gs * invB + reg * e;
      const float m = am[k] + (1.f - p.hp.beta1) * (grad - am[k]);
      const float v2 = av[k] * p.hp.beta2 + (1.f - p.hp.beta2) * grad * grad;
      am[k] = m;
      av[k] = v2;
      const float denom = sqrtf(v2) / bc.y + p.hp.eps;
      eM[k] = e - bc.x * (m / denom);
    }
    __syncthreads();
  }
  for (int k = tid; k < D; k += TT_THREADS) p.out[(size_t)c * D + k] = eM[k];
}

template <bool CP, bool L2N>
int launch_train(kp_ctx* ctx, const TrainK& p, int vpl, size_t smem, cudaStream_t st) {
  const bool ex = ctx->D == 128 * vpl;
  if (vpl <= 1) {
    if (ex) transe_train_kernel<1, 2, CP, L2N, true><<<p.C, TT_THREADS, smem, st>>>(p);
    else transe_train_kernel<1, 2, CP, L2N, false><<<p.C, TT_THREADS, smem, st>>>(p);
  } else if (vpl <= 2) {
    if (ex) transe_train_kernel<2, 1, CP, L2N, true><<<p.C, TT_THREADS, smem, st>>>(p);
    else transe_train_kernel<2, 1, CP, L2N, false><<<p.C, TT_THREADS, smem, st>>>(p);
  } else if (vpl <= 4) {
    if (ex) transe_train_kernel<4, 1, CP, L2N, true><<<p.C, TT_THREADS, smem, st>>>(p);
    else transe_train_kernel<4, 1, CP, L2N, false><<<p.C, TT_THREADS, smem, st>>>(p);
  } else if (vpl <= 8) {
    if (ex) transe_train_kernel<8, 1, CP, L2N, true><<<p.C, TT_THREADS, smem, st>>>(p);
    else transe_train_kernel<8, 1, CP, L2N, false><<<p.C, TT_THREADS, smem, st>>>(p);
  } else {
    KP_FAIL(ctx, KP_EUNSUPPORTED, "TransE post-training supports dim <= 1024 (got %d)", ctx->D);
  }
  return KP_OK;
}

}  // namespace

int kp_transe_post_train(kp_ctx* ctx, const kp_pt_batch* b, const kp_hp* hp, cudaStream_t st) {
  if (b->pos_idx) {
    if (!b->neg_code || !b->facts || !b->fact_off) KP_FAIL(ctx, KP_EINVAL, "compact TransE tables need pos_idx, neg_code, facts and fact_off");
  } else if (!b->pos || !b->neg) {
    KP_FAIL(ctx, KP_EINVAL, "TransE post-training needs pos and neg rows (or the compact tables)");
  }
  TrainK p;
  p.N = (int)ctx->N;
  p.D = ctx->D;
  p.norm = ctx->norm;
  p.C = b->n_candidates;
  p.static_epochs = b->static_epochs;
  p.ent = ctx->ent;
  p.rel = ctx->rel;
  p.row_off = b->row_off;
  p.rows_per_epoch = b->rows_per_epoch;
  p.pos = b->pos;
  p.neg = b->neg;
  p.fact_off = b->fact_off;
  p.facts = b->facts;
  p.pos_idx = b->pos_idx;
  p.neg_code = b->neg_code;
  p.init = b->init_rows;
  p.out = b->out_rows;
  p.hp = *hp;
  const size_t smem = (size_t)(3 + TT_WARPS) * ctx->D * sizeof(float);
  const int vpl = (ctx->D + 127) / 128;
  if (hp->batch_size <= 0 || hp->epochs < 0) KP_FAIL(ctx, KP_EINVAL, "TransE post-training needs batch_size > 0");
  const long long max_steps = (long long)hp->epochs * ((b->max_rows_per_epoch + hp->batch_size - 1) / hp->batch_size);
  if (p.C <= 0 || max_steps <= 0) {  // nothing to train: the rows stay as initialised
    if (p.C > 0) KP_CUDA(ctx, cudaMemcpyAsync(b->out_rows, b->init_rows, (size_t)p.C * ctx->D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return KP_OK;
  }
  int rc;
  if ((rc = kp_ws_reserve(ctx, WsCursor::need((size_t)max_steps, sizeof(float2)), 1)) != KP_OK) return rc;
  float2* tab = reinterpret_cast<float2*>(ctx->ws_arena[1]);
  p.adam = tab;
  KpTimer timer(ctx, kp_ctx::T_TRANSE_TRAIN, st);
  adam_table_kernel<<<(unsigned)((max_steps + 255) / 256), 256, 0, st>>>((int)max_steps, hp->lr, hp->beta1, hp->beta2, tab);
  KP_LAUNCHED(ctx, 1);
  const bool compact = b->pos_idx != nullptr;
  if ((unsigned long long)(ctx->N + 1) * (unsigned long long)(ctx->D / 4) >= (1ull << 32) ||
      (unsigned long long)ctx->R2 * (unsigned long long)(ctx->D / 4) >= (1ull << 32))
    KP_FAIL(ctx, KP_EUNSUPPORTED, "TransE post-training addresses rows with 32-bit float4 offsets (table too large)");
  const bool l2n = ctx->norm == 2;
  rc = compact ? (l2n ? launch_train<true, true>(ctx, p, vpl, smem, st) : launch_train<true, false>(ctx, p, vpl, smem, st))
               : (l2n ? launch_train<false, true>(ctx, p, vpl, smem, st) : launch_train<false, false>(ctx, p, vpl, smem, st));
  if (rc != KP_OK) return rc;
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
