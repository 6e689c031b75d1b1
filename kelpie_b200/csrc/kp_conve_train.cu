// Batched ConvE mimic post-training (KelpieBCEOptimizer, bce_optimizer.py:161-208;
// extract_batch :98-112; ConvE.all_scores conve.py:133-158; BCELoss; Adam lr = 1e-3).
// SURVEY.md section 9.4 gives the arithmetic.
//
// A step of one candidate covers up to `batch_size` (lhs, rel) pairs:
//   * "A" pairs (lhs == mimic): x = phi(e_M, R[r]) is recomputed every step; one fused
//     score->sigmoid->contract pass gives O = sum_j sigmoid(z_j) E_j over the frozen entities,
//     hence d loss / d x = (O + s_M e_M - sum_j t_j E_j) / (B (N+1)) without the dense
//     [B, N+1] prediction / target matrices; it is pulled back through the frozen network
//     (BN3 -> Linear^T -> ReLU/BN2 -> Conv^T -> BN1) to the lhs half of the stacked input.
//   * "B" pairs (frozen lhs, the mimic is their only positive): x is computed once per
//     batch; they reach e_M only through score column M.
// Dropout (conve.py:34-36, active during post-training, model.py:114-125): masks come from a
// counter-based generator keyed by (seed, pair, step, element) (kp_dropout.cuh) and are
// recomputed in the backward kernels; with any rate > 0 the features of the "B" pairs are
// recomputed every step as well (the reference pushes every pair through the network with
// fresh masks each step).
#include "kp_flash.cuh"
#include <cuda_bf16.h>
#include "kp_internal.h"
#include "kp_plan.cuh"
#include "kp_dropout.cuh"

namespace {

constexpr int CT = 256;

__device__ __forceinline__ float blk_sum(float v, float* red) {
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < CT / 32; ++w) s += red[w];
  return s;
}

struct CvDx {
  int GA, N, D, n_strips;
  float ta, tb;  // target = ta * [j in positives] + tb
  const float* ent;
  const float* colsum;
  const float* bn3;
  const int32_t *a_cand, *a_pair;
  const int32_t *nA, *nB;
  const int64_t* pos_off;
  const int32_t* pos_ids;
  const float* xA;
  const float* pO;
  const float* mim;
  float* dh;       // [GA, D] gradient at the Linear output
  float* colcoef;  // [GA]
  unsigned long long seed;
  int step;
  float p_hid;
};

// One CTA per pair: for the few pairs of an explain-sized batch (every pair gets an SM's worth of latency hiding).
__global__ void __launch_bounds__(CT) cv_dx_block(const CvDx p) {
  extern __shared__ float dsm[];
  __shared__ float red[CT / 32];
  const int g = blockIdx.x, tid = threadIdx.x, D = p.D;
  float* acc = dsm;  // [D]
  const int c = p.a_cand[g];
  const float* eM = p.mim + (size_t)c * D;
  const float* x = p.xA + (size_t)g * D;
  float part = 0.f;
  for (int k = tid; k < D; k += CT) part = __fmaf_rn(x[k], eM[k], part);
  const float zM = blk_sum(part, red);
  const float sM = 1.f / (1.f + expf(-zM));
  const int pair = p.a_pair[g];
  const int64_t pb = p.pos_off[pair], pe = p.pos_off[pair + 1];
  bool m_pos = false;
  for (int64_t i = pb; i < pe; ++i) m_pos |= (p.pos_ids[i] == p.N);
  const float scale = 1.f / ((float)(p.nA[c] + p.nB[c]) * (float)(p.N + 1));
  for (int k = tid; k < D; k += CT) {
    float o = 0.f;
    for (int s = 0; s < p.n_strips; ++s) o += p.pO[((size_t)s * p.GA + g) * D + k];
    float tsum = 0.f;
    for (int64_t i = pb; i < pe; ++i) {
      const int e = p.pos_ids[i];
      tsum += (e == p.N) ? eM[k] : p.ent[(size_t)e * D + k];
    }
    const float dx = (o + sM * eM[k] - p.ta * tsum - p.tb * (p.colsum[k] + eM[k])) * scale;
    const float a3 = p.bn3[k] / sqrtf(p.bn3[3 * D + k] + 1e-5f);
    float d = (x[k] > 0.f) ? dx * a3 : 0.f;
    if (p.p_hid > 0.f) d *= kp_drop_scale(p.seed, pair, p.step, KP_DROP_HIDDEN + k, p.p_hid);
    p.dh[(size_t)g * D + k] = d;
  }
  if (tid == 0) p.colcoef[g] = (sM - (p.ta * (m_pos ? 1.f : 0.f) + p.tb)) * scale;
}

// One WARP per pair (D = 200: seven elements per lane, no block-wide barriers): from a few thousand pairs on.
__global__ void __launch_bounds__(CT) cv_dx(const CvDx p) {
  const int g = blockIdx.x * (CT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31, D = p.D;
  if (g >= p.GA) return;
  const int c = p.a_cand[g];
  const float* eM = p.mim + (size_t)c * D;
  const float* x = p.xA + (size_t)g * D;
  float zM = 0.f;
  for (int k = lane; k < D; k += 32) zM = __fmaf_rn(x[k], eM[k], zM);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) zM += __shfl_xor_sync(0xffffffffu, zM, o);
  const float sM = 1.f / (1.f + expf(-zM));
  const int pair = p.a_pair[g];
  const int64_t pb = p.pos_off[pair], pe = p.pos_off[pair + 1];
  bool m_pos = false;
  for (int64_t i = pb + lane; i < pe; i += 32) m_pos |= (p.pos_ids[i] == p.N);
  m_pos = __any_sync(0xffffffffu, m_pos);
  const float scale = 1.f / ((float)(p.nA[c] + p.nB[c]) * (float)(p.N + 1));
  for (int k = lane; k < D; k += 32) {
    float o = 0.f;
    for (int s = 0; s < p.n_strips; ++s) o += p.pO[((size_t)s * p.GA + g) * D + k];
    float tsum = 0.f;
    for (int64_t i = pb; i < pe; ++i) {
      const int e = p.pos_ids[i];
      tsum += (e == p.N) ? eM[k] : p.ent[(size_t)e * D + k];
    }
    const float dx = (o + sM * eM[k] - p.ta * tsum - p.tb * (p.colsum[k] + eM[k])) * scale;
    const float a3 = p.bn3[k] / sqrtf(p.bn3[3 * D + k] + 1e-5f);
    float d = (x[k] > 0.f) ? dx * a3 : 0.f;
    if (p.p_hid > 0.f) d *= kp_drop_scale(p.seed, pair, p.step, KP_DROP_HIDDEN + k, p.p_hid);
    p.dh[(size_t)g * D + k] = d;
  }
  if (lane == 0) p.colcoef[g] = (sM - (p.ta * (m_pos ? 1.f : 0.f) + p.tb)) * scale;
}

struct CvBack {
  int GA, D, H, F, hidden;
  const float *conv_w, *bn1, *bn2;
  const float* feat;   // [GA, hidden] post-ReLU (and post-dropout) feature maps saved by the forward kernel, or
  const __nv_bfloat16* feat_hi;  // their bf16 heads [GA, kpad] (the forward GEMM's operand): only the sign of the maps is used
  int kpad;
  const float* dfeat;  // [GA, hidden] = dh @ fc_w  (Linear^T, kp_gemm.cu)
  float* glhs;         // [GA, D] gradient w.r.t. the lhs embedding
  const int32_t* pair; // [GA] pair ids keying the dropout masks
  unsigned long long seed;
  int step;
  float p_in, p_fm;
};

// One CTA per pair: ReLU / BN2 (/Dropout2d) backward, Conv^T restricted to the lhs half of the
// 40 x H image, BN1 (/input dropout) backward.
// Phase 1 stages d conv (hidden floats) in shared memory with an ODD per-filter stride, so that phase 2 can put one
// filter on every lane without bank conflicts.  Phase 2: a warp owns one row y of the lhs image; lane c accumulates
// filter c's contribution to the H pixels of that row from the (up to) three conv-output rows y, y-1, y-2 held in
// registers, then the H sums are reduced over the lanes.
template <int H>
__global__ void __launch_bounds__(CT) cv_backward(const CvBack p) {
  extern __shared__ float bsm[];
  constexpr int W2 = H - 2, per_f = 38 * W2, stride = per_f | 1;
  const int D = p.D, hidden = p.hidden, F = p.F;
  float* dcv = bsm;              // [F][stride]
  float* wsm = dcv + F * stride;  // [9F]
  float* a2s = wsm + 9 * F;       // [F] BN2 scale (x Dropout2d scale)
  const int tid = threadIdx.x, g = blockIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int c = tid; c < F; c += CT) {
    float a2 = p.bn2[c] / sqrtf(p.bn2[3 * F + c] + 1e-5f);
    if (p.p_fm > 0.f) a2 *= kp_drop_scale(p.seed, p.pair[g], p.step, KP_DROP_FEATURE + c, p.p_fm);
    a2s[c] = a2;
  }
  for (int k = tid; k < 9 * F; k += CT) wsm[k] = p.conv_w[k];
  __syncthreads();
  const float4* feat4 = reinterpret_cast<const float4*>(p.feat + (size_t)g * hidden);
  const float4* dfeat4 = reinterpret_cast<const float4*>(p.dfeat + (size_t)g * hidden);
  const uint2* hi4 = reinterpret_cast<const uint2*>(p.feat_hi + (size_t)g * p.kpad);
  for (int i4 = tid; i4 < hidden / 4; i4 += CT) {  // per_f = 38 * W2 is even, hidden a multiple of 4: a float4 may straddle two filters
    const float4 d = dfeat4[i4];
    const float dv[4] = {d.x, d.y, d.z, d.w};
    bool on[4];
    if (p.feat_hi) {  // maps are >= 0: positive <=> a non-zero bf16 head
      const uint2 h = hi4[i4];
      on[0] = (h.x & 0x7fffu) != 0; on[1] = (h.x & 0x7fff0000u) != 0; on[2] = (h.y & 0x7fffu) != 0; on[3] = (h.y & 0x7fff0000u) != 0;
    } else {
      const float4 f = feat4[i4];
      on[0] = f.x > 0.f; on[1] = f.y > 0.f; on[2] = f.z > 0.f; on[3] = f.w > 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i4 * 4 + j, c = i / per_f;
      dcv[c * stride + (i - c * per_f)] = on[j] ? dv[j] * a2s[c] : 0.f;  // dropped channels have feat == 0
    }
  }
  __syncthreads();
  const float a1 = p.bn1[0] / sqrtf(p.bn1[3] + 1e-5f);
  for (int y = warp; y < 20; y += CT / 32) {
    float acc[H];
#pragma unroll
    for (int x = 0; x < H; ++x) acc[x] = 0.f;
    for (int c = lane; c < F; c += 32) {
      const float* w = wsm + 9 * c;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = y - dy;  // conv-output row; yy <= 19 < 38 always
        if (yy < 0) continue;
        const float* src = dcv + c * stride + yy * W2;
        float r[W2];
#pragma unroll
        for (int xx = 0; xx < W2; ++xx) r[xx] = src[xx];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float wv = w[dy * 3 + dx];
#pragma unroll
          for (int xx = 0; xx < W2; ++xx) acc[xx + dx] = __fmaf_rn(wv, r[xx], acc[xx + dx]);
        }
      }
    }
#pragma unroll
    for (int x = 0; x < H; ++x) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[x] += __shfl_xor_sync(0xffffffffu, acc[x], o);
    }
    if (lane < H) {
      float v = acc[0];
#pragma unroll
      for (int x = 1; x < H; ++x) v = lane == x ? acc[x] : v;
      const int k = y * H + lane;
      if (p.p_in > 0.f) v *= kp_drop_scale(p.seed, p.pair[g], p.step, KP_DROP_INPUT + k, p.p_in);
      p.glhs[(size_t)g * D + k] = v * a1;
    }
  }
}

// any other image width: one thread per lhs pixel
__global__ void __launch_bounds__(CT) cv_backward_generic(const CvBack p) {
  extern __shared__ float bsm[];
  const int D = p.D, H = p.H, W2 = H - 2, hidden = p.hidden, per_f = 38 * W2;
  float* dcv = bsm;  // [hidden]
  const int tid = threadIdx.x, g = blockIdx.x;
  for (int i = tid; i < hidden; i += CT) {
    const int c = i / per_f;
    const float a2 = p.bn2[c] / sqrtf(p.bn2[3 * p.F + c] + 1e-5f);
    const bool on = p.feat_hi ? (__bfloat16_as_ushort(p.feat_hi[(size_t)g * p.kpad + i]) & 0x7fffu) != 0
                              : p.feat[(size_t)g * hidden + i] > 0.f;  // dropped channels are 0
    float d = on ? p.dfeat[(size_t)g * hidden + i] * a2 : 0.f;
    if (on && p.p_fm > 0.f) d *= kp_drop_scale(p.seed, p.pair[g], p.step, KP_DROP_FEATURE + c, p.p_fm);
    dcv[i] = d;
  }
  __syncthreads();
  const float a1 = p.bn1[0] / sqrtf(p.bn1[3] + 1e-5f);
  for (int k = tid; k < D; k += CT) {
    const int y = k / H, x = k % H;
    float acc = 0.f;
    for (int c = 0; c < p.F; ++c) {
      const float* w = p.conv_w + c * 9;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = y - dy;
        if (yy < 0 || yy >= 38) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int xx = x - dx;
          if (xx < 0 || xx >= W2) continue;
          acc = __fmaf_rn(w[dy * 3 + dx], dcv[c * per_f + yy * W2 + xx], acc);
        }
      }
    }
    if (p.p_in > 0.f) acc *= kp_drop_scale(p.seed, p.pair[g], p.step, KP_DROP_INPUT + k, p.p_in);
    p.glhs[(size_t)g * D + k] = acc * a1;
  }
}

struct CvUpd {
  int C, N, D;
  long long step;
  float step_size, bc2_sqrt;  // Adam: lr / (1 - beta1^step), sqrt(1 - beta2^step), evaluated in double on the host like torch
  float lr, beta1, beta2, eps, ta, tb;
  const int32_t *nA, *nB;
  const int64_t *aoff, *boff;
  const float *xA, *glhs, *colcoef, *xB;
  float *mim, *st1, *st2;
};

__global__ void __launch_bounds__(CT) cv_update(const CvUpd p) {
  extern __shared__ float vsm[];
  __shared__ float red[CT / 32];
  const int D = p.D, c = blockIdx.x, tid = threadIdx.x;
  float* eM = vsm;
  float* grad = eM + D;
  const int nA = p.nA[c], nB = p.nB[c];
  const int B = nA + nB;
  if (B == 0) return;
  for (int k = tid; k < D; k += CT) {
    eM[k] = p.mim[(size_t)c * D + k];
    grad[k] = 0.f;
  }
  __syncthreads();
  for (int64_t g = p.aoff[c]; g < p.aoff[c] + nA; ++g) {
    const float cc = p.colcoef[g];
    for (int k = tid; k < D; k += CT) grad[k] += p.glhs[(size_t)g * D + k] + cc * p.xA[(size_t)g * D + k];
  }
  const float scale = 1.f / ((float)B * (float)(p.N + 1));
  for (int64_t b = p.boff[c]; b < p.boff[c] + nB; ++b) {
    const float* x = p.xB + (size_t)b * D;
    float part = 0.f;
    for (int k = tid; k < D; k += CT) part = __fmaf_rn(x[k], eM[k], part);
    const float z = blk_sum(part, red);
    const float coef = (1.f / (1.f + expf(-z)) - (p.ta + p.tb)) * scale;
    for (int k = tid; k < D; k += CT) grad[k] += coef * x[k];
  }
  __syncthreads();
  for (int k = tid; k < D; k += CT) {
    const size_t idx = (size_t)c * D + k;
    const float g = grad[k];
    const float m = p.st1[idx] + (1.f - p.beta1) * (g - p.st1[idx]);
    const float v = p.st2[idx] * p.beta2 + (1.f - p.beta2) * g * g;
    p.st1[idx] = m;
    p.st2[idx] = v;
    const float denom = sqrtf(v) / p.bc2_sqrt + p.eps;
    p.mim[idx] = eM[k] - p.step_size * (m / denom);
  }
}

}  // namespace

int kp_conve_post_train(kp_ctx* ctx, const kp_pt_batch* b, const kp_hp* hp, cudaStream_t st) {
  if (!b->pos || !b->pos_off || !b->pos_ids) KP_FAIL(ctx, KP_EINVAL, "ConvE post-training needs pairs and positives");
  const bool dropout = ctx->cv.drop_in != 0.f || ctx->cv.drop_fm != 0.f || ctx->cv.drop_hid != 0.f;
  const unsigned long long seed = b->dropout_seed;
  const int C = b->n_candidates, D = ctx->D, hidden = ctx->cv.hidden;
  const int bs = hp->batch_size, max_n = b->max_rows_per_epoch;
  const int spe_max = max_n > 0 ? (max_n + bs - 1) / bs : 0;
  const long long T = (long long)hp->epochs * spe_max;
  int64_t cap = (int64_t)C * (int64_t)(max_n < bs ? max_n : bs);
  if (b->total_rows < cap) cap = b->total_rows;
  if (cap < 1) cap = 1;
  if (cap > (int64_t)1 << 28) KP_FAIL(ctx, KP_EUNSUPPORTED, "batch too large (%lld pairs per step)", (long long)cap);
  const int G = (int)cap, Gpad = ((G + 63) / 64) * 64;
  const size_t SG = kp_flash_part_rows(ctx, G);  // rows of strip partials, worst case over steps of <= G rows

  // feature maps of a step's pairs, kept for the backward pass: fp32 [G, hidden], or the forward GEMM's split operand (kp_conve_fc_umma)
  size_t feat_bytes = (size_t)G * hidden * 4;
  if (kp_conve_fc_umma(ctx, G) && 2 * kp_conve_feat_half_bytes(ctx, G) > feat_bytes) feat_bytes = 2 * kp_conve_feat_half_bytes(ctx, G);
  size_t need = 3 * WsCursor::need((size_t)C * D, 4) + 3 * WsCursor::need(C, 4) + 2 * WsCursor::need(C + 1, 8) +
                7 * WsCursor::need(G, 4) + 2 * WsCursor::need((size_t)Gpad * D, 4) +
                WsCursor::need(feat_bytes, 1) + WsCursor::need((size_t)G * hidden, 4) + 2 * WsCursor::need((size_t)G * D, 4) + WsCursor::need(G, 4) +
                2 * WsCursor::need(SG, 4) + WsCursor::need(SG * D, 4);
  int rc = kp_ws_reserve(ctx, need);
  if (rc != KP_OK) return rc;
  WsCursor ws{ctx->ws, ctx->ws + ctx->ws_bytes};
  float* mim = ws.take<float>((size_t)C * D);
  float* st1 = ws.take<float>((size_t)C * D);
  float* st2 = ws.take<float>((size_t)C * D);
  CxPlan pl;
  pl.C = C; pl.N = (int)ctx->N; pl.D = D; pl.bs = bs; pl.epochs = hp->epochs; pl.static_epochs = 1; pl.truth_is_row = 1;
  pl.row_off = b->row_off; pl.rows_per_epoch = b->rows_per_epoch; pl.pos = b->pos;
  pl.nA = ws.take<int32_t>(C); pl.nB = ws.take<int32_t>(C); pl.nSelf = ws.take<int32_t>(C);
  pl.aoff = ws.take<int64_t>(C + 1); pl.boff = ws.take<int64_t>(C + 1);
  pl.a_cand = ws.take<int32_t>(G); pl.a_rel = ws.take<int32_t>(G); pl.a_truth = ws.take<int32_t>(G);
  pl.b_cand = ws.take<int32_t>(G); pl.b_lhs = ws.take<int32_t>(G); pl.b_rel = ws.take<int32_t>(G); pl.b_row = ws.take<int32_t>(G);
  float* xA = ws.take<float>((size_t)Gpad * D);
  float* xB = ws.take<float>((size_t)Gpad * D);
  float* feat = reinterpret_cast<float*>(ws.take<char>(feat_bytes));
  float* dfeat = ws.take<float>((size_t)G * hidden);
  float* dh = ws.take<float>((size_t)G * D);
  float* glhs = ws.take<float>((size_t)G * D);
  float* colcoef = ws.take<float>(G);
  float* pm = ws.take<float>(SG);
  float* plv = ws.take<float>(SG);
  float* pO = ws.take<float>(SG * D);

  KP_CUDA(ctx, cudaMemcpyAsync(mim, b->init_rows, (size_t)C * D * 4, cudaMemcpyDeviceToDevice, st));
  KP_CUDA(ctx, cudaMemsetAsync(st1, 0, (size_t)C * D * 4, st));
  KP_CUDA(ctx, cudaMemsetAsync(st2, 0, (size_t)C * D * 4, st));

  const float ls = hp->label_smoothing;
  const float ta = ls != 0.f ? 1.f - ls : 1.f;
  const float tb = ls != 0.f ? 1.f / (float)(ctx->N + 1) : 0.f;
  const int cb = (C + 127) / 128;
  int64_t GA = 0, GB = 0;
  const int Hc = ctx->cv.H, Fc = ctx->cv.n_filters;
  const size_t back_smem = ((size_t)Fc * ((38 * (Hc - 2)) | 1) + 10 * Fc) * sizeof(float);
  if (back_smem > 200 * 1024) KP_FAIL(ctx, KP_EUNSUPPORTED, "ConvE hidden size %d too large", hidden);
  KP_SMEM_ONCE(ctx, (cv_backward<10>), 200 * 1024);
  KP_SMEM_ONCE(ctx, (cv_backward<4>), 200 * 1024);
  KP_SMEM_ONCE(ctx, cv_backward_generic, 200 * 1024);
  const bool static_plan = spe_max <= 1;
  for (long long t = 0; t < T; ++t) {
    if (t == 0 || !static_plan) {
      cx_count<<<cb, 128, 0, st>>>(pl, (int)t);
      cx_scan<<<1, 1024, 0, st>>>(pl);
      cx_assign<<<cb, 128, 0, st>>>(pl, (int)t);
      KP_LAUNCHED(ctx, 3);
      int64_t tot[2];
      KP_CUDA(ctx, cudaMemcpyAsync(&tot[0], pl.aoff + C, 8, cudaMemcpyDeviceToHost, st));
      KP_CUDA(ctx, cudaMemcpyAsync(&tot[1], pl.boff + C, 8, cudaMemcpyDeviceToHost, st));
      KP_CUDA(ctx, cudaStreamSynchronize(st));
      GA = tot[0];
      GB = tot[1];
      if (GA > G || GB > G) KP_FAIL(ctx, KP_EINVAL, "step uses more pairs than the batch declares");
      if (GB > 0 && !dropout &&
          (rc = kp_conve_features_ex(ctx, (int)GB, pl.b_lhs, pl.b_rel, 1, nullptr, nullptr, xB, nullptr, st)) != KP_OK)
        return rc;
    }
    if (GB > 0 && dropout &&  // fresh masks every step for the frozen-lhs pairs too
        (rc = kp_conve_features_ex(ctx, (int)GB, pl.b_lhs, pl.b_rel, 1, nullptr, nullptr, xB, nullptr, st, pl.b_row, seed, (int)t)) != KP_OK)
      return rc;
    int ns = 1;
    if (GA > 0) {
      if ((rc = kp_conve_features_ex(ctx, (int)GA, nullptr, pl.a_rel, 1, mim, pl.a_cand, xA, feat, st,
                                     dropout ? pl.a_truth : nullptr, seed, (int)t)) != KP_OK)
        return rc;
      if ((rc = kp_flash_run(ctx, xA, (int)GA, KP_FLASH_SIGMOID, pm, plv, pO, st, &ns)) != KP_OK) return rc;
      CvDx d;
      d.GA = (int)GA; d.N = (int)ctx->N; d.D = D; d.n_strips = ns; d.ta = ta; d.tb = tb;
      d.ent = ctx->ent; d.colsum = ctx->cv.ent_colsum; d.bn3 = ctx->cv.bn3;
      d.a_cand = pl.a_cand; d.a_pair = pl.a_truth; d.nA = pl.nA; d.nB = pl.nB;
      d.pos_off = b->pos_off; d.pos_ids = b->pos_ids; d.xA = xA; d.pO = pO; d.mim = mim; d.dh = dh; d.colcoef = colcoef;
      d.seed = seed; d.step = (int)t; d.p_hid = ctx->cv.drop_hid;
      if (GA < 4096) cv_dx_block<<<(int)GA, CT, (size_t)D * 4, st>>>(d);
      else cv_dx<<<(int)((GA + CT / 32 - 1) / (CT / 32)), CT, 0, st>>>(d);
      CvBack k;
      k.GA = (int)GA; k.D = D; k.H = ctx->cv.H; k.F = ctx->cv.n_filters; k.hidden = hidden;
      k.conv_w = ctx->cv.conv_w; k.bn1 = ctx->cv.bn1; k.bn2 = ctx->cv.bn2;
      k.feat = feat; k.dfeat = dfeat; k.glhs = glhs;
      k.feat_hi = kp_conve_fc_umma(ctx, (int)GA) ? reinterpret_cast<const __nv_bfloat16*>(feat) : nullptr;
      k.kpad = kp_conve_feat_kpad(ctx);
      k.pair = pl.a_truth; k.seed = seed; k.step = (int)t; k.p_in = ctx->cv.drop_in; k.p_fm = ctx->cv.drop_fm;
      KP_LAUNCHED(ctx, 1);
      if ((rc = kp_conve_fc(ctx, false, (int)GA, dh, dfeat, 0, st)) != KP_OK) return rc;
      {
        KpTimer timer(ctx, kp_ctx::T_CONV, st);
        if (Hc == 10) cv_backward<10><<<(int)GA, CT, back_smem, st>>>(k);
        else if (Hc == 4) cv_backward<4><<<(int)GA, CT, back_smem, st>>>(k);
        else cv_backward_generic<<<(int)GA, CT, back_smem, st>>>(k);
      }
      KP_LAUNCHED(ctx, 1);
    }
    if (GA + GB > 0) {
      CvUpd u;
      u.C = C; u.N = (int)ctx->N; u.D = D; u.step = t + 1;
      u.step_size = (float)((double)hp->lr / (1.0 - pow((double)hp->beta1, (double)(t + 1))));
      u.bc2_sqrt = (float)sqrt(1.0 - pow((double)hp->beta2, (double)(t + 1)));
      u.lr = hp->lr; u.beta1 = hp->beta1; u.beta2 = hp->beta2; u.eps = hp->eps; u.ta = ta; u.tb = tb;
      u.nA = pl.nA; u.nB = pl.nB; u.aoff = pl.aoff; u.boff = pl.boff;
      u.xA = xA; u.glhs = glhs; u.colcoef = colcoef; u.xB = xB; u.mim = mim; u.st1 = st1; u.st2 = st2;
      cv_update<<<C, CT, (size_t)2 * D * 4, st>>>(u);
      KP_LAUNCHED(ctx, 1);
    }
  }
  KP_CUDA(ctx, cudaMemcpyAsync(b->out_rows, mim, (size_t)C * D * 4, cudaMemcpyDeviceToDevice, st));
  return KP_OK;
}
