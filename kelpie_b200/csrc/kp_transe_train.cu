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
  kp_hp hp;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

template <int VPL>  // float4 vectors per lane; D <= 128 * VPL
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
  for (int k = tid; k < D; k += TT_THREADS) {
    eM[k] = p.init[(size_t)c * D + k];
    am[k] = 0.f;
    av[k] = 0.f;
  }
  __syncthreads();

  const int M = p.N;
  const int bs = p.hp.batch_size;
  long long step = 0;
  for (int ep = 0; ep < p.hp.epochs; ++ep) {
    const int64_t ebase = base0 + (p.static_epochs ? 0 : (int64_t)ep * n);
    for (int b0 = 0; b0 < n; b0 += bs) {
      const int B = min(bs, n - b0);
      float4 g[VPL];
#pragma unroll
      for (int v = 0; v < VPL; ++v) g[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
      for (int i = warp; i < B; i += TT_WARPS) {
        const int32_t* pr = p.pos + (ebase + b0 + i) * 3;
        const int32_t* nr = p.neg + (ebase + b0 + i) * 3;
        const int h = pr[0], r = pr[1], t = pr[2], h2 = nr[0], t2 = nr[2];
        const float* ph = (h == M) ? eM : p.ent + (size_t)h * D;
        const float* pt = (t == M) ? eM : p.ent + (size_t)t * D;
        const float* ph2 = (h2 == M) ? eM : p.ent + (size_t)h2 * D;
        const float* pt2 = (t2 == M) ? eM : p.ent + (size_t)t2 * D;
        const float* prl = p.rel + (size_t)r * D;
        float4 dp[VPL], dn[VPL];
        float sp = 0.f, sn = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int k = (v * 32 + lane) * 4;
          dp[v] = dn[v] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < D) {
            const float4 a = ld4(ph + k), rr = ld4(prl + k), b = ld4(pt + k);
            const float4 a2 = ld4(ph2 + k), b2 = ld4(pt2 + k);
            dp[v].x = __fsub_rn(__fadd_rn(a.x, rr.x), b.x);
            dp[v].y = __fsub_rn(__fadd_rn(a.y, rr.y), b.y);
            dp[v].z = __fsub_rn(__fadd_rn(a.z, rr.z), b.z);
            dp[v].w = __fsub_rn(__fadd_rn(a.w, rr.w), b.w);
            dn[v].x = __fsub_rn(__fadd_rn(a2.x, rr.x), b2.x);
            dn[v].y = __fsub_rn(__fadd_rn(a2.y, rr.y), b2.y);
            dn[v].z = __fsub_rn(__fadd_rn(a2.z, rr.z), b2.z);
            dn[v].w = __fsub_rn(__fadd_rn(a2.w, rr.w), b2.w);
            if (p.norm == 2) {
              sp += dp[v].x * dp[v].x + dp[v].y * dp[v].y + dp[v].z * dp[v].z + dp[v].w * dp[v].w;
              sn += dn[v].x * dn[v].x + dn[v].y * dn[v].y + dn[v].z * dn[v].z + dn[v].w * dn[v].w;
            } else {
              sp += fabsf(dp[v].x) + fabsf(dp[v].y) + fabsf(dp[v].z) + fabsf(dp[v].w);
              sn += fabsf(dn[v].x) + fabsf(dn[v].y) + fabsf(dn[v].z) + fabsf(dn[v].w);
            }
          }
        }
        sp = warp_sum(sp);
        sn = warp_sum(sn);
        if (p.norm == 2) {
          sp = sqrtf(sp);
          sn = sqrtf(sn);
        }
        const bool active = (sp - sn + p.hp.margin) > 0.f;
        const float cp = (float)((h == M) - (t == M));
        const float cn = (float)((h2 == M) - (t2 == M));
        cnt += (h == M) + (t == M) + (h2 == M) + (t2 == M);
        if (active) {
          // d||d||_2/dd = d/||d|| (0 at the origin, as torch); d||d||_1/dd = sign(d)
          const float ip = (p.norm == 2) ? ((sp > 0.f) ? cp / sp : 0.f) : cp;
          const float in = (p.norm == 2) ? ((sn > 0.f) ? cn / sn : 0.f) : cn;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            if (p.norm == 2) {
              g[v].x += ip * dp[v].x - in * dn[v].x;
              g[v].y += ip * dp[v].y - in * dn[v].y;
              g[v].z += ip * dp[v].z - in * dn[v].z;
              g[v].w += ip * dp[v].w - in * dn[v].w;
            } else {
              g[v].x += ip * sgn(dp[v].x) - in * sgn(dn[v].x);
              g[v].y += ip * sgn(dp[v].y) - in * sgn(dn[v].y);
              g[v].z += ip * sgn(dp[v].z) - in * sgn(dn[v].z);
              g[v].w += ip * sgn(dp[v].w) - in * sgn(dn[v].w);
            }
          }
        }
      }
      // all warps have finished READING eM for this step before anyone updates it
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int k = (v * 32 + lane) * 4;
        if (k < D) *reinterpret_cast<float4*>(gw + warp * D + k) = g[v];
      }
      if (lane == 0) s_cnt[warp] = cnt;
      __syncthreads();
      ++step;
      const double bc1 = 1.0 - pow((double)p.hp.beta1, (double)step);
      const double bc2 = 1.0 - pow((double)p.hp.beta2, (double)step);
      const float step_size = (float)((double)p.hp.lr / bc1);
      const float bc2_sqrt = (float)sqrt(bc2);
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
        const float denom = sqrtf(v2) / bc2_sqrt + p.hp.eps;
        eM[k] = e - step_size * (m / denom);
      }
      __syncthreads();
    }
  }
  for (int k = tid; k < D; k += TT_THREADS) p.out[(size_t)c * D + k] = eM[k];
}

}  // namespace

int kp_transe_post_train(kp_ctx* ctx, const kp_pt_batch* b, const kp_hp* hp, cudaStream_t st) {
  if (!b->pos || !b->neg) KP_FAIL(ctx, KP_EINVAL, "TransE post-training needs pos and neg rows");
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
  p.init = b->init_rows;
  p.out = b->out_rows;
  p.hp = *hp;
  const size_t smem = (size_t)(3 + TT_WARPS) * ctx->D * sizeof(float);
  const int vpl = (ctx->D + 127) / 128;
  KpTimer timer(ctx, kp_ctx::T_TRANSE_TRAIN, st);
  if (vpl <= 1) {
    transe_train_kernel<1><<<p.C, TT_THREADS, smem, st>>>(p);
  } else if (vpl <= 2) {
    transe_train_kernel<2><<<p.C, TT_THREADS, smem, st>>>(p);
  } else if (vpl <= 4) {
    transe_train_kernel<4><<<p.C, TT_THREADS, smem, st>>>(p);
  } else if (vpl <= 8) {
    transe_train_kernel<8><<<p.C, TT_THREADS, smem, st>>>(p);
  } else {
    KP_FAIL(ctx, KP_EUNSUPPORTED, "TransE post-training supports dim <= 1024 (got %d)", ctx->D);
  }
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
