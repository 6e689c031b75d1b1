// Batched ComplEx mimic post-training (KelpieMultiClassNLLOptimizer,
// multiclass_nll_optimizer.py:123-164; ComplEx.forward complex.py:59-86; N3 regularizers.py:37-46;
// torch.optim.Adagrad / Adam / SGD).  SURVEY.md section 9.3 gives the arithmetic.
//
// Every row (h, r, t) of a mimic's training set mentions the mimic M.  Per optimiser step:
//   * rows with h == M ("A rows"): the query q = e_M o R[r] changes every step.  One fused
//     score->softmax->contract pass over the whole entity table gives, per row, the softmax
//     statistics (m, l) and O = sum_j exp(z_j - m) E_j, i.e. d loss / d q without ever
//     forming the [B, N+1] logits or the dense [N+1, 2d] table gradient of the reference.
//   * rows with h != M ("B rows", t == M): their query and their log-sum-exp over the N
//     frozen columns never change; both are computed ONCE per batch, after which a step costs
//     one dot product with e_M (the gradient reaches e_M only through score column M).
// The update kernel (one CTA per candidate) assembles d loss / d e_M from both row kinds and
// column M, adds N3, and applies the optimiser step in place.
#include "kp_flash.cuh"
#include "kp_internal.h"
#include "kp_plan.cuh"

namespace {

// q = lhs o rel (complex product), lhs = mimic row of cand[g] when lhs_id == NULL (A rows)
__global__ void cx_build_queries(int G, int D, const float* __restrict__ ent, const float* __restrict__ rel,
                                 const float* __restrict__ mim, const int32_t* __restrict__ cand,
                                 const int32_t* __restrict__ lhs_id, const int32_t* __restrict__ rel_id,
                                 float* __restrict__ qmat) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (g >= G) return;
  const float* l = lhs_id ? ent + (size_t)lhs_id[g] * D : mim + (size_t)cand[g] * D;
  const float* r = rel + (size_t)rel_id[g] * D;
  float* out = qmat + (size_t)g * D;
  const int d = D >> 1;
  for (int k = lane; k < d; k += 32) {
    const float lr = l[k], li = l[d + k], rr = r[k], ri = r[d + k];
    out[k] = __fsub_rn(__fmul_rn(lr, rr), __fmul_rn(li, ri));
    out[d + k] = __fadd_rn(__fmul_rn(lr, ri), __fmul_rn(li, rr));
  }
}

// lse over the frozen columns of each B row, from the per-strip softmax statistics
__global__ void cx_lse(int G, int n_strips, const float* __restrict__ pm, const float* __restrict__ pl,
                       float* __restrict__ lse) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  float M, L;
  kp_flash_merge_stats(pm, pl, n_strips, G, g, M, L);
  lse[g] = M + logf(L);
}

// Strips of row g folded into strip 0 in place (same order and weights as a sequential merge): one CTA per row, so the
// per-candidate update below walks one strip per row instead of n_strips (it runs one CTA per candidate; with a
// handful of candidates per batch -- the explain path -- the serial merge was 57 % of the device time).
// One CTA per row, one thread per output dim (up to 1024 threads; the strips' maxima / sums are loaded by all threads in parallel and
// folded in strip order from shared memory, the partial rows 16 loads in flight per thread): with the 64 strips of an explain-sized
// batch the kernel is a few dependent round trips to the L2 long instead of ~100.
constexpr int MRG_MAX_THREADS = 1024;
__global__ void __launch_bounds__(MRG_MAX_THREADS) cx_merge_strips(int GA, int D, int n_strips, float* __restrict__ pm,
                                                                  float* __restrict__ pl, float* __restrict__ pO) {
  extern __shared__ float wgt[];  // [n_strips] e^{m_s - M}, < 0 = empty strip | [n_strips] l_s
  __shared__ float red[MRG_MAX_THREADS / 32];
  float* ls = wgt + n_strips;
  const int g = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  float M = -INFINITY;
  for (int s = tid; s < n_strips; s += nt) {
    const float ms = pm[(size_t)s * GA + g];
    wgt[s] = ms;
    ls[s] = pl[(size_t)s * GA + g];
    M = fmaxf(M, ms);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  if ((tid & 31) == 0) red[tid >> 5] = M;
  __syncthreads();
  for (int w = 0; w < (nt + 31) / 32; ++w) M = fmaxf(M, red[w]);
  for (int s = tid; s < n_strips; s += nt) {
    const float ms = wgt[s];
    wgt[s] = (ms != -INFINITY) ? expf(ms - M) : -1.f;
  }
  __syncthreads();
  for (int k = tid; k < D; k += nt) {
    float ok = 0.f;
    for (int s0 = 0; s0 < n_strips; s0 += 16) {  // 16 partials in flight per thread, folded in strip order
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (s0 + j < n_strips) ? pO[((size_t)(s0 + j) * GA + g) * D + k] : 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float w = (s0 + j < n_strips) ? wgt[s0 + j] : -1.f;
        if (w >= 0.f) ok += v[j] * w;
      }
    }
    pO[(size_t)g * D + k] = ok;
  }
  if (tid == 0) {
    float L = 0.f;
    for (int s = 0; s < n_strips; ++s) {
      const float w = wgt[s];
      if (w >= 0.f) L += ls[s] * w;
    }
    pm[g] = M;
    pl[g] = L;
  }
}

struct CxUpd {
  int C, N, D, GA, n_strips, optimizer;
  long long step;  // 1-based optimiser step (Adam bias correction)
  float step_size, bc2_sqrt;  // Adam: lr / (1 - beta1^step), sqrt(1 - beta2^step), evaluated in double on the host like torch
  float lr, beta1, beta2, eps, n3;
  int reg_kind;  // KP_REG_N3: sum |f|^3 per component; KP_REG_N2: |row|_2^3 (regularizers.py:25-46)
  const float* ent;
  const float* rel;
  const int32_t *nA, *nB, *nSelf;
  const int64_t *aoff, *boff;
  const int32_t *a_rel, *a_truth;
  const float* qA;      // [GA, D]
  const float* pm;      // [n_strips, GA]
  const float* pl;
  const float* pO;      // [n_strips, GA, D]
  const float* qB;      // [GB, D]
  const float* lseB;    // [GB]
  float* mim;           // [C, D]
  float* st1;           // [C, D] Adagrad sum / Adam exp_avg
  float* st2;           // [C, D] Adam exp_avg_sq
  // staged != 0: cx_row_grad ran first -- pO[g, :] holds row g's gradient contribution and coefB[b] the coefficient of B row b
  int staged;
  const float* coefB;   // [GB]
  const int32_t *a_cand, *b_cand;
};

constexpr int UPD_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < UPD_THREADS / 32; ++w) s += red[w];
  return s;
}

// Per-ROW stage of the update (one CTA per row of the step, A rows then B rows), so that the per-candidate kernel below is left with
// plain sums: for an A row the strips are merged (as cx_merge_strips does), then the row's softmax statistics, dq and its
// contribution to the mimic row's gradient are computed with the expressions -- and the 256-thread reduction order -- of the
// per-candidate loop it replaces, and the contribution overwrites pO[g, :]; for a B row only the coefficient is needed.
// With the 5-10 rows per candidate of an explain-sized batch this takes the row loop (a block reduction and two barriers per
// row) off the critical path of 6 CTAs and spreads it over ~100.
__global__ void __launch_bounds__(UPD_THREADS) cx_row_grad(const CxUpd p, int GB, float* __restrict__ pm, float* __restrict__ pl,
                                                          float* __restrict__ pO, float* __restrict__ coefB) {
  extern __shared__ float usm[];
  const int D = p.D, d = D >> 1, tid = threadIdx.x;
  float* eM = usm;        // [D]
  float* dq = eM + D;     // [D]
  float* qv = dq + D;     // [D]
  float* wgt = qv + D;    // [n_strips]
  float* ls = wgt + p.n_strips;
  __shared__ float red[UPD_THREADS / 32];
  if ((int)blockIdx.x >= p.GA) {
    const int b = blockIdx.x - p.GA;
    if (b >= GB) return;
    const int c = p.b_cand[b];
    const float invB = 1.f / (float)(p.nA[c] + p.nB[c]);
    float part = 0.f;
    for (int k = tid; k < D; k += UPD_THREADS) part = __fmaf_rn(p.qB[(size_t)b * D + k], p.mim[(size_t)c * D + k], part);
    const float z = block_sum(part, red);
    const float lse = p.lseB[b];
    const float mx = fmaxf(lse, z);
    const float eS = expf(z - mx);
    const float pM = eS / (expf(lse - mx) + eS);
    if (tid == 0) coefB[b] = (pM - 1.f) * invB;  // the truth of a B row is the mimic itself
    return;
  }
  const int g = blockIdx.x, c = p.a_cand[g], GA = p.GA, ns = p.n_strips;
  const float invB = 1.f / (float)(p.nA[c] + p.nB[c]);
  float M = -INFINITY;
  for (int s = tid; s < ns; s += UPD_THREADS) {
    const float ms = pm[(size_t)s * GA + g];
    wgt[s] = ms;
    ls[s] = pl[(size_t)s * GA + g];
    M = fmaxf(M, ms);
  }
  float part = 0.f;
  for (int k = tid; k < D; k += UPD_THREADS) {
    const float qk = p.qA[(size_t)g * D + k], ek = p.mim[(size_t)c * D + k];
    eM[k] = ek;
    qv[k] = qk;
    part = __fmaf_rn(qk, ek, part);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  if ((tid & 31) == 0) red[tid >> 5] = M;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < UPD_THREADS / 32; ++w) M = fmaxf(M, red[w]);
  for (int s = tid; s < ns; s += UPD_THREADS) {
    const float ms = wgt[s];
    wgt[s] = (ms != -INFINITY) ? expf(ms - M) : -1.f;
  }
  const float zM = block_sum(part, red);  // (its barriers also publish wgt / ls / eM)
  float L = 0.f;
  for (int s = 0; s < ns; ++s) {
    const float w = wgt[s];
    if (w >= 0.f) L += ls[s] * w;
  }
  const float mx = fmaxf(M, zM);
  const float eF = L * expf(M - mx), eS = expf(zM - mx);
  const float den = eF + eS;
  const float pM = eS / den;
  const int o = p.a_truth[g];
  const float* Eo = (o == p.N) ? eM : p.ent + (size_t)o * D;
  for (int k = tid; k < D; k += UPD_THREADS) {
    float ok = 0.f;
    for (int s0 = 0; s0 < ns; s0 += 16) {  // 16 partials in flight per thread, folded in strip order
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (s0 + j < ns) ? pO[((size_t)(s0 + j) * GA + g) * D + k] : 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float w = (s0 + j < ns) ? wgt[s0 + j] : -1.f;
        if (w >= 0.f) ok += v[j] * w;
      }
    }
    dq[k] = (ok * (expf(M - mx) / den) + pM * eM[k] - Eo[k]) * invB;
  }
  __syncthreads();
  const float* rho = p.rel + (size_t)p.a_rel[g] * D;
  const float colM = (pM - (o == p.N ? 1.f : 0.f)) * invB;
  for (int k = tid; k < d; k += UPD_THREADS) {
    const float rr = rho[k], ri = rho[d + k], da = dq[k], db = dq[d + k];
    pO[(size_t)g * D + k] = da * rr + db * ri + colM * qv[k];
    pO[(size_t)g * D + d + k] = -da * ri + db * rr + colM * qv[d + k];
  }
}

__global__ void __launch_bounds__(UPD_THREADS) cx_update(const CxUpd p) {
  extern __shared__ float usm[];
  const int D = p.D, d = D >> 1;
  float* eM = usm;         // [D]
  float* grad = eM + D;    // [D]
  float* dq = grad + D;    // [D]
  float* qv = dq + D;      // [D]
  __shared__ float red[UPD_THREADS / 32];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int nA = p.nA[c], nB = p.nB[c];
  const int B = nA + nB;
  if (B == 0) return;  // candidate has no step here (finished, or no facts)
  const float invB = 1.f / (float)B;
  for (int k = tid; k < D; k += UPD_THREADS) {
    eM[k] = p.mim[(size_t)c * D + k];
    grad[k] = 0.f;
  }
  __syncthreads();

  if (p.staged) {  // rows already reduced to their contributions: sum them in row order (a thread owns its dims: no barriers)
    for (int k = tid; k < D; k += UPD_THREADS) {
      float acc = 0.f;
      for (int64_t g = p.aoff[c]; g < p.aoff[c] + nA; ++g) acc += p.pO[(size_t)g * D + k];
      for (int64_t b = p.boff[c]; b < p.boff[c] + nB; ++b) acc += p.coefB[b] * p.qB[(size_t)b * D + k];
      grad[k] = acc;
    }
  }
  for (int64_t g = p.aoff[c]; !p.staged && g < p.aoff[c] + nA; ++g) {
    float M, L;
    kp_flash_merge_stats(p.pm, p.pl, p.n_strips, p.GA, (int)g, M, L);
    float part = 0.f;
    for (int k = tid; k < D; k += UPD_THREADS) {
      const float qk = p.qA[(size_t)g * D + k];
      qv[k] = qk;
      part = __fmaf_rn(qk, eM[k], part);
    }
    const float zM = block_sum(part, red);
    const float mx = fmaxf(M, zM);
    const float eF = L * expf(M - mx), eS = expf(zM - mx);
    const float den = eF + eS;
    const float pM = eS / den;
    const int o = p.a_truth[g];
    const float* Eo = (o == p.N) ? eM : p.ent + (size_t)o * D;
    for (int k = tid; k < D; k += UPD_THREADS) {
      float ok = 0.f;
      for (int s = 0; s < p.n_strips; ++s) {
        const float ms = p.pm[(size_t)s * p.GA + g];
        if (ms != -INFINITY) ok += p.pO[((size_t)s * p.GA + g) * D + k] * expf(ms - M);
      }
      dq[k] = (ok * (expf(M - mx) / den) + pM * eM[k] - Eo[k]) * invB;
    }
    __syncthreads();
    const float* rho = p.rel + (size_t)p.a_rel[g] * D;
    const float colM = (pM - (o == p.N ? 1.f : 0.f)) * invB;
    for (int k = tid; k < d; k += UPD_THREADS) {
      const float rr = rho[k], ri = rho[d + k], da = dq[k], db = dq[d + k];
      grad[k] += da * rr + db * ri + colM * qv[k];
      grad[d + k] += -da * ri + db * rr + colM * qv[d + k];
    }
    __syncthreads();
  }
  for (int64_t b = p.boff[c]; !p.staged && b < p.boff[c] + nB; ++b) {
    float part = 0.f;
    for (int k = tid; k < D; k += UPD_THREADS) {
      const float qk = p.qB[(size_t)b * D + k];
      qv[k] = qk;
      part = __fmaf_rn(qk, eM[k], part);
    }
    const float z = block_sum(part, red);
    const float lse = p.lseB[b];
    const float mx = fmaxf(lse, z);
    const float eS = expf(z - mx);
    const float pM = eS / (expf(lse - mx) + eS);
    const float coef = (pM - 1.f) * invB;  // the truth of a B row is the mimic itself
    for (int k = tid; k < D; k += UPD_THREADS) grad[k] += coef * qv[k];
    __syncthreads();
  }
  // N3 (regularizers.py:37-46): w/B * sum_rows |f|^3, f = sqrt(re^2 + im^2) of lhs and rhs rows;
  // N2 (regularizers.py:25-34): w/B * sum_rows ||f||_2^3 -- d/de = 3 ||e||_2 e per occurrence of the mimic
  const float cntM = (float)(nA + nB + p.nSelf[c]);
  float row_norm = 0.f;
  if (p.n3 != 0.f && p.reg_kind == KP_REG_N2) {
    float part = 0.f;
    for (int k = tid; k < D; k += UPD_THREADS) part = __fmaf_rn(eM[k], eM[k], part);
    row_norm = sqrtf(block_sum(part, red));
  }
  for (int k = tid; k < D; k += UPD_THREADS) {
    float g = grad[k];
    const float e = eM[k];
    if (p.n3 != 0.f) {
      const int kr = (k < d) ? k : k - d;
      const float f = (p.reg_kind == KP_REG_N2) ? row_norm : sqrtf(eM[kr] * eM[kr] + eM[kr + d] * eM[kr + d]);
      g += 3.f * p.n3 * invB * cntM * f * e;
    }
    const size_t idx = (size_t)c * D + k;
    float out;
    if (p.optimizer == KP_OPT_ADAGRAD) {
      const float s = p.st1[idx] + g * g;
      p.st1[idx] = s;
      out = e - p.lr * (g / (sqrtf(s) + p.eps));
    } else if (p.optimizer == KP_OPT_ADAM) {
      const float m = p.st1[idx] + (1.f - p.beta1) * (g - p.st1[idx]);
      const float v = p.st2[idx] * p.beta2 + (1.f - p.beta2) * g * g;
      p.st1[idx] = m;
      p.st2[idx] = v;
      const float denom = sqrtf(v) / p.bc2_sqrt + p.eps;
      out = e - p.step_size * (m / denom);
    } else {
      out = e - p.lr * g;
    }
    p.mim[idx] = out;
  }
}

}  // namespace

int kp_complex_post_train(kp_ctx* ctx, const kp_pt_batch* b, const kp_hp* hp, cudaStream_t st) {
  if (!b->pos) KP_FAIL(ctx, KP_EINVAL, "ComplEx post-training needs rows");
  const int C = b->n_candidates, D = ctx->D;
  const int bs = hp->batch_size;
  const int max_n = b->max_rows_per_epoch;
  if (max_n < 0 || b->total_rows < 0) KP_FAIL(ctx, KP_EINVAL, "bad batch totals");
  const int spe_max = max_n > 0 ? (max_n + bs - 1) / bs : 0;
  const long long T = (long long)hp->epochs * spe_max;
  const bool static_plan = b->static_epochs != 0 && spe_max <= 1;
  // upper bound on rows taking part in one step
  int64_t cap = (int64_t)C * (int64_t)(max_n < bs ? max_n : bs);
  if (b->static_epochs == 0 && hp->epochs > 0) {
    const int64_t per_epoch = b->total_rows / hp->epochs;
    if (per_epoch < cap) cap = per_epoch;
  } else if (b->total_rows < cap) {
    cap = b->total_rows;
  }
  if (cap < 1) cap = 1;
  if (cap > (int64_t)1 << 30) KP_FAIL(ctx, KP_EUNSUPPORTED, "batch too large (%lld rows per step)", (long long)cap);
  const int G = (int)cap;
  const int Gpad = ((G + 63) / 64) * 64;
  const size_t SG = kp_flash_part_rows(ctx, G);  // rows of strip partials, worst case over steps of <= G rows

  size_t need = 0;
  need += 3 * WsCursor::need((size_t)C * D, 4);         // mim, st1, st2
  need += 3 * WsCursor::need(C, 4) + 2 * WsCursor::need(C + 1, 8);
  need += 6 * WsCursor::need(G, 4);
  need += 2 * WsCursor::need((size_t)Gpad * D, 4);      // qA, qB
  need += 2 * WsCursor::need(G, 4);                      // lseB, coefB
  need += 2 * WsCursor::need(SG, 4) + WsCursor::need(SG * D, 4);
  int rc = kp_ws_reserve(ctx, need);
  if (rc != KP_OK) return rc;
  WsCursor ws{ctx->ws, ctx->ws + ctx->ws_bytes};
  float* mim = ws.take<float>((size_t)C * D);
  float* st1 = ws.take<float>((size_t)C * D);
  float* st2 = ws.take<float>((size_t)C * D);
  CxPlan pl;
  pl.C = C; pl.N = (int)ctx->N; pl.D = D; pl.bs = bs; pl.epochs = hp->epochs; pl.static_epochs = b->static_epochs; pl.truth_is_row = 0;
  pl.row_off = b->row_off; pl.rows_per_epoch = b->rows_per_epoch; pl.pos = b->pos;
  pl.nA = ws.take<int32_t>(C); pl.nB = ws.take<int32_t>(C); pl.nSelf = ws.take<int32_t>(C);
  pl.aoff = ws.take<int64_t>(C + 1); pl.boff = ws.take<int64_t>(C + 1);
  pl.a_cand = ws.take<int32_t>(G); pl.a_rel = ws.take<int32_t>(G); pl.a_truth = ws.take<int32_t>(G);
  pl.b_cand = ws.take<int32_t>(G); pl.b_lhs = ws.take<int32_t>(G); pl.b_rel = ws.take<int32_t>(G); pl.b_row = nullptr;
  float* qA = ws.take<float>((size_t)Gpad * D);
  float* qB = ws.take<float>((size_t)Gpad * D);
  float* lseB = ws.take<float>(G);
  float* coefB = ws.take<float>(G);
  float* pm = ws.take<float>(SG);
  float* plv = ws.take<float>(SG);
  float* pO = ws.take<float>(SG * D);

  KP_CUDA(ctx, cudaMemcpyAsync(mim, b->init_rows, (size_t)C * D * 4, cudaMemcpyDeviceToDevice, st));
  KP_CUDA(ctx, cudaMemsetAsync(st1, 0, (size_t)C * D * 4, st));
  KP_CUDA(ctx, cudaMemsetAsync(st2, 0, (size_t)C * D * 4, st));

  const int cb = (C + 127) / 128;
  int64_t GA = 0, GB = 0;
  auto make_plan = [&](int t) -> int {
    cx_count<<<cb, 128, 0, st>>>(pl, t);
    cx_scan<<<1, 1024, 0, st>>>(pl);
    cx_assign<<<cb, 128, 0, st>>>(pl, t);
    KP_LAUNCHED(ctx, 3);
    int64_t tot[2];
    KP_CUDA(ctx, cudaMemcpyAsync(&tot[0], pl.aoff + C, 8, cudaMemcpyDeviceToHost, st));
    KP_CUDA(ctx, cudaMemcpyAsync(&tot[1], pl.boff + C, 8, cudaMemcpyDeviceToHost, st));
    KP_CUDA(ctx, cudaStreamSynchronize(st));
    GA = tot[0];
    GB = tot[1];
    if (GA > G || GB > G) KP_FAIL(ctx, KP_EINVAL, "step uses more rows (%lld/%lld) than the batch declares (%d)", (long long)GA, (long long)GB, G);
    if (GB > 0) {
      cx_build_queries<<<(int)((GB + 7) / 8), 256, 0, st>>>((int)GB, D, ctx->ent, ctx->rel, nullptr, nullptr, pl.b_lhs, pl.b_rel, qB);
      KP_LAUNCHED(ctx, 1);
      int ns = 1;
      int r2 = kp_flash_run(ctx, qB, (int)GB, KP_FLASH_SOFTMAX, pm, plv, pO, st, &ns);
      if (r2 != KP_OK) return r2;
      cx_lse<<<(int)((GB + 255) / 256), 256, 0, st>>>((int)GB, ns, pm, plv, lseB);
      KP_LAUNCHED(ctx, 1);
    }
    return KP_OK;
  };

  for (long long t = 0; t < T; ++t) {
    if (t == 0 || !static_plan) {
      if ((rc = make_plan((int)t)) != KP_OK) return rc;
    }
    int ns = 1;
    if (GA > 0) {
      cx_build_queries<<<(int)((GA + 7) / 8), 256, 0, st>>>((int)GA, D, ctx->ent, ctx->rel, mim, pl.a_cand, nullptr, pl.a_rel, qA);
      KP_LAUNCHED(ctx, 1);
      if ((rc = kp_flash_run(ctx, qA, (int)GA, KP_FLASH_SOFTMAX, pm, plv, pO, st, &ns)) != KP_OK) return rc;
    }
    const bool staged = ctx->cx_rowgrad != 0 && ctx->cx_merge != 0;
    if (GA > 0 && ns > 1 && ctx->cx_merge && !staged) {
      cx_merge_strips<<<(int)GA, D >= MRG_MAX_THREADS ? MRG_MAX_THREADS : ((D + 31) / 32) * 32, (size_t)2 * ns * sizeof(float), st>>>((int)GA, D, ns, pm, plv, pO);
      KP_LAUNCHED(ctx, 1);
      ns = 1;
    }
    if (GA + GB > 0) {
      CxUpd u;
      u.C = C; u.N = (int)ctx->N; u.D = D; u.GA = (int)GA; u.n_strips = ns; u.optimizer = hp->optimizer;
      u.step = t + 1;
      u.step_size = (float)((double)hp->lr / (1.0 - pow((double)hp->beta1, (double)(t + 1))));
      u.bc2_sqrt = (float)sqrt(1.0 - pow((double)hp->beta2, (double)(t + 1)));
      u.lr = hp->lr; u.beta1 = hp->beta1; u.beta2 = hp->beta2; u.eps = hp->eps; u.n3 = hp->reg_weight; u.reg_kind = hp->regularizer;
      u.ent = ctx->ent; u.rel = ctx->rel;
      u.nA = pl.nA; u.nB = pl.nB; u.nSelf = pl.nSelf; u.aoff = pl.aoff; u.boff = pl.boff;
      u.a_rel = pl.a_rel; u.a_truth = pl.a_truth;
      u.qA = qA; u.pm = pm; u.pl = plv; u.pO = pO; u.qB = qB; u.lseB = lseB;
      u.mim = mim; u.st1 = st1; u.st2 = st2;
      u.staged = staged ? 1 : 0; u.coefB = coefB; u.a_cand = pl.a_cand; u.b_cand = pl.b_cand;
      if (staged) {
        cx_row_grad<<<(int)(GA + GB), UPD_THREADS, ((size_t)3 * D + 2 * ns) * sizeof(float), st>>>(u, (int)GB, pm, plv, pO, coefB);
        KP_LAUNCHED(ctx, 1);
      }
      cx_update<<<C, UPD_THREADS, (size_t)4 * D * sizeof(float), st>>>(u);
      KP_LAUNCHED(ctx, 1);
    }
  }
  KP_CUDA(ctx, cudaMemcpyAsync(b->out_rows, mim, (size_t)C * D * 4, cudaMemcpyDeviceToDevice, st));
  return KP_OK;
}
