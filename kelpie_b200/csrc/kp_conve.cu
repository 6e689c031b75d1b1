// ConvE frozen network on the device: x = phi(lhs, rel) (conve.py:133-156, eval-mode BN):
//   [lhs;rel] as a 40 x H image -> BN2d(1) -> Conv2d(1->F, 3x3) -> BN2d(F) -> ReLU -> flatten
//   -> Linear(hidden -> D) -> BN1d(D) -> ReLU.
// The entity projection x @ E^T + sigmoid is the generic all-entity pass (kp_pass.cu).
#include "kp_internal.h"
#include "kp_ptx.cuh"
#include "kp_dropout.cuh"
#include "kp_conve_head.cuh"

#include <cuda_bf16.h>

namespace {

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}


struct ConvK {
  int Q, N, R2, D, H, F, hidden, stride;
  const float* ent;
  const float* rel;
  const int32_t* lhs_ids;      // [Q*stride] entity id of the lhs (N = mimic); NULL = every lhs is a mimic
  const int32_t* rel_ids;      // [Q*stride]
  const float* mimic;          // mimic rows
  const int32_t* mimic_index;  // row of `mimic` used by query q (NULL: row q)
  const float *conv_w, *conv_b, *fc_w, *fc_b, *bn1, *bn2, *bn3;
  float* x_out;      // [Q, D]
  const float* parts;  // n_parts > 1: the Linear layer's output as n_parts partial sums [Q, D] each (K cut over CTAs), else in x_out
  int n_parts;
  float* feat_out;   // [Q, hidden]: post-ReLU feature maps (kept for the backward pass), unless hi_out is set:
  __nv_bfloat16 *hi_out, *lo_out;  // the same as the Linear GEMM's split operand, bf16 [.., kpad] each (x = hi + lo + O(2^-16 x))
  int kpad;
  // training-mode dropout (conve.py:34-36,140-152; active during post-training, model.py:114-125)
  const int32_t* drop_ids;  // NULL = eval mode; else the pair id keying the masks of query q
  unsigned long long seed;
  int step;
  float p_in, p_fm, p_hid;
};


// Stage 1: stacked image -> BN1 (+input dropout) -> Conv 3x3 -> BN2 -> ReLU (+Dropout2d) -> feat[q, hidden].
// Persistent CTAs walk the pairs (grid stride): the filter weights and BN2 affines are staged once per CTA and -- with 32 filters,
// one group of 4 per warp -- stay in registers over all of a CTA's pairs; the image is double-buffered in shared memory (one
// barrier per pair).  A warp owns a group of 4 filters and its lanes sweep the 38 x W2 output positions: every image tap read
// from shared memory feeds 4 (split output: 8) FMAs and every store instruction of the warp writes 128 consecutive bytes of one
// filter's plane.
constexpr int CV_CONV_THREADS = 256;
__global__ void __launch_bounds__(CV_CONV_THREADS, 3) conve_conv_kernel(const ConvK p) {
  extern __shared__ float sm[];
  const int H = p.H, W2 = H - 2, D = p.D, F = p.F;
  const int img_sz = 40 * H;
  float* imgs = sm;               // [2][40*H]
  float* wsm = imgs + 2 * img_sz;  // [9F] weights | [F] bias | [F] alpha2 | [F] beta2
  float* bsm = wsm + 9 * F;
  float* a2s = bsm + F;
  float* b2s = a2s + F;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float a1, b1;
  bn_affine(p.bn1, 1, 0, a1, b1);
  const bool drop = p.drop_ids != nullptr;
  for (int k = tid; k < 9 * F; k += CV_CONV_THREADS) wsm[k] = p.conv_w[k];
  for (int c = tid; c < F; c += CV_CONV_THREADS) {
    float a2, b2;
    bn_affine(p.bn2, F, c, a2, b2);
    bsm[c] = p.conv_b[c];
    a2s[c] = a2;
    b2s[c] = b2;
  }
  const int per_f = 38 * W2;
  const bool even = (H & 1) == 0;  // split output: both positions of a lane then sit in one image row (3 x 4 window, six 8-byte loads)
  float w[4][9], bias[4], a2b[4], b2b[4];
  int c0_regs = -1;
  int it = 0;
  for (int q = blockIdx.x; q < p.Q; q += gridDim.x, ++it) {
    float* img = imgs + (it & 1) * img_sz;
    const int s = p.lhs_ids ? p.lhs_ids[(size_t)q * p.stride] : p.N, r = p.rel_ids[(size_t)q * p.stride];
    const float* l = (s == p.N) ? p.mimic + (size_t)(p.mimic_index ? p.mimic_index[q] : q) * D : p.ent + (size_t)s * D;
    const int pid = drop ? p.drop_ids[q] : 0;
    for (int k = tid; k < img_sz; k += CV_CONV_THREADS) {
      float v = (k < D) ? l[k] : p.rel[(size_t)r * D + (k - D)];
      v = v * a1 + b1;
      if (drop && p.p_in > 0.f) v *= kp_drop_scale(p.seed, pid, p.step, KP_DROP_INPUT + k, p.p_in);
      img[k] = v;
    }
    __syncthreads();  // the image of this pair (and, first time, the weights); the other buffer is free once every thread got here
    for (int c0 = warp * 4; c0 < F; c0 += (CV_CONV_THREADS / 32) * 4) {
      if (c0 != c0_regs) {
#pragma unroll
        for (int f = 0; f < 4; ++f) {
#pragma unroll
          for (int k = 0; k < 9; ++k) w[f][k] = wsm[(c0 + f) * 9 + k];
          bias[f] = bsm[c0 + f];
          a2b[f] = a2s[c0 + f];
          b2b[f] = b2s[c0 + f];
        }
        c0_regs = c0;
      }
      float a2[4], b2[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        a2[f] = a2b[f];
        b2[f] = b2b[f];
        if (drop && p.p_fm > 0.f) {  // Dropout2d scales a whole channel; relu(a x + b) * s = relu(s a x + s b) for s >= 0
          const float sc = kp_drop_scale(p.seed, pid, p.step, KP_DROP_FEATURE + c0 + f, p.p_fm);
          a2[f] *= sc;
          b2[f] *= sc;
        }
      }
      if (p.hi_out) {  // a lane owns two neighbouring positions: one 4-byte store per filter and half (hi = bf16(x), lo = bf16(x - hi))
        uint32_t* hi = reinterpret_cast<uint32_t*>(p.hi_out + (size_t)q * p.kpad + (size_t)c0 * per_f);
        uint32_t* lo = reinterpret_cast<uint32_t*>(p.lo_out + (size_t)q * p.kpad + (size_t)c0 * per_f);
        for (int pp = lane; pp < per_f / 2; pp += 32) {  // per_f = 38 * W2 is even
          const int pos = 2 * pp, y0 = pos / W2, y1 = (pos + 1) / W2;
          const float* tap0 = img + pos + 2 * y0;
          const float* tap1 = img + pos + 1 + 2 * y1;
          float acc0[4] = {bias[0], bias[1], bias[2], bias[3]}, acc1[4] = {bias[0], bias[1], bias[2], bias[3]};
          if (even) {
            const float2* t2 = reinterpret_cast<const float2*>(tap0);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const float2 a = t2[dy * (H >> 1)], b = t2[dy * (H >> 1) + 1];
              const float v[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                  acc0[f] = __fmaf_rn(w[f][dy * 3 + dx], v[dx], acc0[f]);
                  acc1[f] = __fmaf_rn(w[f][dy * 3 + dx], v[dx + 1], acc1[f]);
                }
            }
          } else {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const float v0 = tap0[dy * H + dx], v1 = tap1[dy * H + dx];
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                  acc0[f] = __fmaf_rn(w[f][dy * 3 + dx], v0, acc0[f]);
                  acc1[f] = __fmaf_rn(w[f][dy * 3 + dx], v1, acc1[f]);
                }
              }
          }
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const float x0 = fmaxf(__fmaf_rn(acc0[f], a2[f], b2[f]), 0.f), x1 = fmaxf(__fmaf_rn(acc1[f], a2[f], b2[f]), 0.f);
            const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);  // .x (low half) = x0
            const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
            const __nv_bfloat162 lw = __floats2bfloat162_rn(x0 - __uint_as_float(hb << 16), x1 - __uint_as_float(hb & 0xffff0000u));
            hi[(size_t)f * (per_f / 2) + pp] = hb;
            lo[(size_t)f * (per_f / 2) + pp] = *reinterpret_cast<const uint32_t*>(&lw);
          }
        }
        continue;
      }
      float* out = p.feat_out + (size_t)q * p.hidden + (size_t)c0 * per_f;
      for (int pos = lane; pos < per_f; pos += 32) {
        const int y = pos / W2;
        const float* tap = img + pos + 2 * y;  // y * H + x with H = W2 + 2
        float acc[4] = {bias[0], bias[1], bias[2], bias[3]};
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float v = tap[dy * H + dx];
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[f] = __fmaf_rn(w[f][dy * 3 + dx], v, acc[f]);
          }
#pragma unroll
        for (int f = 0; f < 4; ++f) out[(size_t)f * per_f + pos] = fmaxf(__fmaf_rn(acc[f], a2[f], b2[f]), 0.f);
      }
    }
    if (p.hi_out)  // zero columns up to the GEMM's k-block boundary
      for (int k = p.hidden + tid; k < p.kpad; k += CV_CONV_THREADS) {
        p.hi_out[(size_t)q * p.kpad + k] = __float2bfloat16_rn(0.f);
        p.lo_out[(size_t)q * p.kpad + k] = __float2bfloat16_rn(0.f);
      }
  }
}

// Stage 3: x = ReLU(BN3(dropout(raw + fc_b)))   (stage 2 is the Linear GEMM, kp_gemm.cu)
__global__ void conve_head_kernel(const ConvK p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)p.Q * p.D) return;
  const int q = (int)(i / p.D), k = (int)(i % p.D);
  kp_conve_head hd{p.fc_b, p.bn3, p.drop_ids, p.seed, p.step, p.p_hid};
  float raw;
  if (p.n_parts > 1) {
    raw = p.parts[i];
    for (int z = 1; z < p.n_parts; ++z) raw += p.parts[(size_t)z * p.Q * p.D + i];
  } else {
    raw = p.x_out[i];
  }
  p.x_out[i] = kp_conve_head_apply(hd, p.D, q, k, raw);
}

__global__ void colsum_kernel(int N, int D, const float* __restrict__ ent, float* __restrict__ out) {
  // one block per 32 columns; double accumulation keeps the sum exact enough for N ~ 1e6
  const int k = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5, ny = blockDim.x >> 5;
  __shared__ double part[32][33];
  double acc = 0.0;
  if (k < D)
    for (int j = ry; j < N; j += ny) acc += (double)ent[(size_t)j * D + k];
  part[ry][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ry == 0 && k < D) {
    double s = 0.0;
    for (int y = 0; y < ny; ++y) s += part[y][threadIdx.x & 31];
    out[k] = (float)s;
  }
}

template <typename T>
int upload(kp_ctx* ctx, const T* src, size_t n, T** dst) {
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, n * sizeof(T) + 16);
  if (e != cudaSuccess) KP_FAIL(ctx, KP_ENOMEM, "cudaMalloc(%zu): %s", n * sizeof(T), cudaGetErrorString(e));
  ctx->owned.push_back(d);
  KP_CUDA(ctx, cudaMemcpy(d, src, n * sizeof(T), cudaMemcpyDefault));
  *dst = static_cast<T*>(d);
  return KP_OK;
}

}  // namespace

int kp_conve_setup(kp_ctx* ctx, const kp_conve_weights* w) {
  const int D = ctx->D, H = D / 20;
  if (w->n_filters % 4 != 0) KP_FAIL(ctx, KP_EUNSUPPORTED, "ConvE filter count %d must be a multiple of 4 (the reference fixes 32)", w->n_filters);
  if (w->n_filters <= 0 || w->hidden != w->n_filters * 38 * (H - 2))
    KP_FAIL(ctx, KP_EINVAL, "ConvE hidden size %d != %d * 38 * %d", w->hidden, w->n_filters, H - 2);
  if (w->hidden % 4 != 0) KP_FAIL(ctx, KP_EUNSUPPORTED, "ConvE hidden size must be a multiple of 4");
  if (!w->conv_w || !w->conv_b || !w->fc_w || !w->fc_b || !w->bn1 || !w->bn2 || !w->bn3)
    KP_FAIL(ctx, KP_EINVAL, "null ConvE weight pointer");
  ctx->cv.n_filters = w->n_filters;
  ctx->cv.hidden = w->hidden;
  ctx->cv.H = H;
  ctx->cv.drop_in = w->drop_input;
  ctx->cv.drop_fm = w->drop_feature;
  ctx->cv.drop_hid = w->drop_hidden;
  int rc;
  if ((rc = upload(ctx, w->conv_w, (size_t)w->n_filters * 9, &ctx->cv.conv_w))) return rc;
  if ((rc = upload(ctx, w->conv_b, (size_t)w->n_filters, &ctx->cv.conv_b))) return rc;
  if ((rc = upload(ctx, w->fc_w, (size_t)D * w->hidden, &ctx->cv.fc_w))) return rc;
  if ((rc = upload(ctx, w->fc_b, (size_t)D, &ctx->cv.fc_b))) return rc;
  if ((rc = upload(ctx, w->bn1, 4, &ctx->cv.bn1))) return rc;
  if ((rc = upload(ctx, w->bn2, (size_t)4 * w->n_filters, &ctx->cv.bn2))) return rc;
  if ((rc = upload(ctx, w->bn3, (size_t)4 * D, &ctx->cv.bn3))) return rc;
  void* cs = nullptr;
  KP_CUDA(ctx, cudaMalloc(&cs, (size_t)D * 4));
  ctx->owned.push_back(cs);
  ctx->cv.ent_colsum = static_cast<float*>(cs);
  colsum_kernel<<<(D + 31) / 32, 1024>>>((int)ctx->N, D, ctx->ent, ctx->cv.ent_colsum);
  KP_LAUNCHED(ctx, 1);
  if (D % 4 == 0 && w->hidden % 4 == 0) {  // split operands of the Linear layer for the tcgen05 GEMMs
    if ((rc = kp_umma_b_prepare(ctx, ctx->cv.fc_w, D, w->hidden, false, &ctx->cv.fc_fwd, 0))) return rc;
    if ((rc = kp_umma_b_prepare(ctx, ctx->cv.fc_w, w->hidden, D, true, &ctx->cv.fc_bwd, 0))) return rc;
  }
  return KP_OK;
}

bool kp_conve_fc_umma(const kp_ctx* ctx, int M) {
  return ctx->conv_split && ctx->umma_fc && !ctx->force_simt && ctx->cv.fc_fwd.ready && M >= ctx->umma_fc_min_rows;
}
size_t kp_conve_feat_half_bytes(const kp_ctx* ctx, int M) { return kp_gemm_umma_a_bytes(M, ctx->cv.fc_fwd); }
int kp_conve_feat_kpad(const kp_ctx* ctx) { return ctx->cv.fc_fwd.Kpad; }

int kp_conve_fc(kp_ctx* ctx, bool forward, int M, const float* A, float* C, size_t ws_offset, cudaStream_t st) {
  const int D = ctx->D, hidden = ctx->cv.hidden;
  const kp_umma_b& B = forward ? ctx->cv.fc_fwd : ctx->cv.fc_bwd;
  if (ctx->umma_fc && !ctx->force_simt && B.ready && M >= ctx->umma_fc_min_rows)
    return kp_gemm_umma(ctx, A, forward ? hidden : D, M, B, C, forward ? D : hidden, ws_offset, st);
  if (forward && M < 128 && ctx->skinny_fc) return kp_sgemm_skinny_nt(ctx, M, D, hidden, A, hidden, ctx->cv.fc_w, hidden, C, D, st);
  if (forward) return kp_sgemm(ctx, true, M, D, hidden, A, hidden, ctx->cv.fc_w, hidden, C, D, st);
  return kp_sgemm(ctx, false, M, hidden, D, A, D, ctx->cv.fc_w, hidden, C, hidden, st);
}

int kp_conve_features_ex(kp_ctx* ctx, int Q, const int32_t* lhs_ids, const int32_t* rel_ids, int stride,
                         const float* mimic, const int32_t* mimic_index, float* x_out, float* feat_out,
                         cudaStream_t st, const int32_t* drop_ids, unsigned long long seed, int step) {
  ConvK p;
  p.drop_ids = drop_ids;
  p.seed = seed;
  p.step = step;
  p.p_in = ctx->cv.drop_in;
  p.p_fm = ctx->cv.drop_fm;
  p.p_hid = ctx->cv.drop_hid;
  p.Q = Q;
  p.N = (int)ctx->N;
  p.R2 = (int)ctx->R2;
  p.D = ctx->D;
  p.H = ctx->cv.H;
  p.F = ctx->cv.n_filters;
  p.hidden = ctx->cv.hidden;
  p.stride = stride;
  p.ent = ctx->ent;
  p.rel = ctx->rel;
  p.lhs_ids = lhs_ids;
  p.rel_ids = rel_ids;
  p.mimic = mimic;
  p.mimic_index = mimic_index;
  p.conv_w = ctx->cv.conv_w;
  p.conv_b = ctx->cv.conv_b;
  p.fc_w = ctx->cv.fc_w;
  p.fc_b = ctx->cv.fc_b;
  p.bn1 = ctx->cv.bn1;
  p.bn2 = ctx->cv.bn2;
  p.bn3 = ctx->cv.bn3;
  p.x_out = x_out;
  p.feat_out = feat_out;
  // feature maps go to the caller's buffer (kept for the backward pass) or to scratch, in chunks
  const int chunk = feat_out ? Q : (Q < 16384 ? Q : 16384);
  const bool split_first = kp_conve_fc_umma(ctx, chunk);
  const size_t fp32_bytes = (size_t)chunk * p.hidden * sizeof(float) + 1024;
  const size_t half = split_first ? kp_conve_feat_half_bytes(ctx, chunk) : 0;
  char* scratch = reinterpret_cast<char*>(feat_out);
  // arena 1: [feature maps unless the caller keeps them | partial outputs of a Linear GEMM whose K is cut over CTAs]
  const size_t parts_off = feat_out ? 0 : (((fp32_bytes > 2 * half + 2048 ? fp32_bytes : 2 * half + 2048) + 1023) & ~size_t(1023));
  const int ks_first = split_first ? kp_gemm_umma_ksplit(ctx, chunk, ctx->cv.fc_fwd) : 1;
  const size_t parts_bytes = ks_first > 1 ? (size_t)ks_first * chunk * p.D * sizeof(float) : 0;
  if (parts_off + parts_bytes > 0) {
    int rc = kp_ws_reserve(ctx, parts_off + parts_bytes, 1);
    if (rc != KP_OK) return rc;
  }
  if (!feat_out) scratch = ctx->ws_arena[1];
  float* parts = ks_first > 1 ? reinterpret_cast<float*>(ctx->ws_arena[1] + parts_off) : nullptr;
  for (int q0 = 0; q0 < Q; q0 += chunk) {
    const int n = (Q - q0 < chunk) ? Q - q0 : chunk;
    const bool split = kp_conve_fc_umma(ctx, n);
    ConvK c = p;
    c.Q = n;
    c.lhs_ids = lhs_ids ? lhs_ids + (size_t)q0 * stride : nullptr;
    c.rel_ids = rel_ids + (size_t)q0 * stride;
    c.mimic = (mimic && !mimic_index) ? mimic + (size_t)q0 * p.D : mimic;
    c.mimic_index = mimic_index ? mimic_index + q0 : nullptr;
    c.drop_ids = drop_ids ? drop_ids + q0 : nullptr;
    c.x_out = x_out + (size_t)q0 * p.D;
    c.feat_out = reinterpret_cast<float*>(scratch);
    c.hi_out = c.lo_out = nullptr;
    // (a last, shorter chunk has fewer rows per part and at least as many parts: only the first chunk's count is provided for)
    const int ks = (split && parts && n == chunk) ? ks_first : 1;
    c.parts = parts;
    c.n_parts = ks;
    c.kpad = kp_conve_feat_kpad(ctx);
    if (split) {
      c.hi_out = reinterpret_cast<__nv_bfloat16*>(scratch);
      c.lo_out = reinterpret_cast<__nv_bfloat16*>(scratch + kp_conve_feat_half_bytes(ctx, n));
    }
    {
      KpTimer timer(ctx, kp_ctx::T_CONV, st);
      const int grid = n < ctx->sm_count * 6 ? n : ctx->sm_count * 6;  // persistent from a few pairs per CTA on
      conve_conv_kernel<<<grid, CV_CONV_THREADS, ((size_t)80 * p.H + 12 * p.F) * sizeof(float), st>>>(c);
    }
    KP_LAUNCHED(ctx, 1);
    // (the head -- bias, dropout, BN3, ReLU -- stays a kernel of its own: in the GEMM's epilogue, one tile per CTA here, its
    //  208 sqrt / divide / hash evaluations per thread are a serial tail: 175 us against 133 + 17 us)
    int rc = split ? kp_gemm_umma_split(ctx, c.hi_out, c.lo_out, n, ctx->cv.fc_fwd, c.x_out, p.D, st, ks > 1 ? parts : nullptr)
                   : kp_conve_fc(ctx, true, n, c.feat_out, c.x_out, feat_out ? 0 : fp32_bytes, st);
    if (rc != KP_OK) return rc;
    conve_head_kernel<<<(int)(((size_t)n * p.D + 255) / 256), 256, 0, st>>>(c);
    KP_LAUNCHED(ctx, 1);
  }
  return KP_OK;
}

int kp_conve_features(kp_ctx* ctx, int Q, const int32_t* triples, int stride, const float* mimic, float* x_out,
                      cudaStream_t st) {
  return kp_conve_features_ex(ctx, Q, triples, triples + 1, stride, mimic, nullptr, x_out, nullptr, st);
}

