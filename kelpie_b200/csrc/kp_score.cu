// Query preparation, target scores, filter ranges and rank finalisation around the
// all-entity pass.  Replaces Model.all_scores / get_triple_results / predict_tails
// (transe.py:48-65, complex.py:88-113, conve.py:133-184, post_training_engine.py:101-125,
// model.py:42-68).
#include "kp_internal.h"
#include "kp_ptx.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per query: qmat[q] = query vector of (s_q, p_q, .)
//   TransE  : l + r                                  (transe.py:49-57)
//   ComplEx : [l_re r_re - l_im r_im , l_re r_im + l_im r_re]   (complex.py:90-99)
__global__ void prep_queries(int kind, int Q, int N, int R2, int D, const float* __restrict__ ent,
                             const float* __restrict__ rel, const int32_t* __restrict__ triples,
                             const float* __restrict__ mimic, float* __restrict__ qmat) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  const int s = triples[3 * q], p = triples[3 * q + 1];
  const bool bad = s < 0 || s > N || (s == N && !mimic) || p < 0 || p >= R2;
  const float* l = (s == N) ? mimic + (size_t)q * D : ent + (size_t)s * D;
  const float* r = rel + (size_t)p * D;
  float* out = qmat + (size_t)q * D;
  if (bad) {
    for (int k = lane; k < D; k += 32) out[k] = __int_as_float(0x7fc00000);
    return;
  }
  if (kind == KP_TRANSE) {
    for (int k = lane; k < D; k += 32) out[k] = __fadd_rn(l[k], r[k]);
  } else {
    const int d = D >> 1;
    for (int k = lane; k < d; k += 32) {
      const float lr = l[k], li = l[d + k], rr = r[k], ri = r[d + k];
      out[k] = __fsub_rn(__fmul_rn(lr, rr), __fmul_rn(li, ri));
      out[d + k] = __fadd_rn(__fmul_rn(lr, ri), __fmul_rn(li, rr));
    }
  }
}

// Score of one (query, row) pair with EXACTLY the arithmetic of the pass kernel that will scan the
// table, so that the target's score equals the score that kernel computes for row o (entities
// with identical embeddings then tie exactly, as they do in the reference's single matmul):
//   tile kernel (kp_pass.cu)   : one sequential fp32 FMA chain over k = 0 .. D-1
//   stream kernel (kp_stream.cu): lane l accumulates the float4 at (v*32 + l)*4, v = 0,1,..; then a
//                                 butterfly over lane offsets 16, 8, 4, 2, 1
__device__ __forceinline__ float term(int op, float acc, float x, float y) {
  if (op == KP_OP_DOT) return __fmaf_rn(x, y, acc);
  if (op == KP_OP_L2) {
    const float d = __fsub_rn(x, y);
    return __fmaf_rn(d, d, acc);
  }
  return __fadd_rn(acc, fabsf(__fsub_rn(x, y)));
}

__device__ __forceinline__ float pair_score(int op, int act, bool stream, const float* a, const float* b, int D, int lane) {
  float acc = 0.f;
  if (!stream) {
    if (lane == 0)
      for (int k = 0; k < D; ++k) acc = term(op, acc, a[k], b[k]);
    acc = __shfl_sync(0xffffffffu, acc, 0);
  } else {
    for (int k = lane * 4; k < D; k += 128) {
      if (op == KP_OP_L1) {  // the stream kernel adds the four |d| of a float4 before accumulating
        acc += fabsf(a[k] - b[k]) + fabsf(a[k + 1] - b[k + 1]) + fabsf(a[k + 2] - b[k + 2]) + fabsf(a[k + 3] - b[k + 3]);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (op == KP_OP_DOT) acc = __fmaf_rn(a[k + c], b[k + c], acc);
          else { const float d = a[k + c] - b[k + c]; acc = __fmaf_rn(d, d, acc); }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  }
  if (op == KP_OP_L2) acc = sqrtf(acc);
  if (act == KP_ACT_SIGMOID) acc = 1.f / (1.f + expf(-acc));
  return acc;
}

// One warp per query: score of the target o_q and of the query's own mimic row (column N).
__global__ void target_scores(int op, int act, int stream, int Q, int N, int D, const float* __restrict__ ent,
                              const int32_t* __restrict__ triples, const float* __restrict__ mimic,
                              const float* __restrict__ qmat, float* __restrict__ target,
                              float* __restrict__ self, int32_t* __restrict__ tgt_ent) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= Q) return;
  const int o = triples[3 * q + 2];
  const float* a = qmat + (size_t)q * D;
  float t = __int_as_float(0x7fc00000), sf = 0.f;
  if (o >= 0 && (o < N || (o == N && mimic))) {
    const float* b = (o == N) ? mimic + (size_t)q * D : ent + (size_t)o * D;
    t = pair_score(op, act, stream != 0, a, b, D, lane);
  }
  if (mimic) sf = pair_score(op, act, stream != 0, a, mimic + (size_t)q * D, D, lane);
  if (lane == 0) {
    target[q] = t;
    self[q] = sf;
    tgt_ent[q] = o;
  }
}

// Per-query filter range: explicit CSR (engine overlay) or lookup of key (s,p) in the resident CSR.
__global__ void filter_ranges(int Q, int64_t R2, const int32_t* __restrict__ triples,
                              const int64_t* __restrict__ flt_off, int64_t n_keys,
                              const int64_t* __restrict__ keys, const int64_t* __restrict__ koff,
                              int64_t* __restrict__ beg, int64_t* __restrict__ end) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  if (flt_off) {
    beg[q] = flt_off[q];
    end[q] = flt_off[q + 1];
    return;
  }
  const int64_t key = (int64_t)triples[3 * q] * R2 + triples[3 * q + 1];
  int64_t lo = 0, hi = n_keys;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < key) lo = mid + 1; else hi = mid;
  }
  if (lo < n_keys && keys[lo] == key) {
    beg[q] = koff[lo];
    end[q] = koff[lo + 1];
  } else {
    beg[q] = end[q] = 0;
  }
}

__device__ __forceinline__ int64_t lower_bound_ids(const int32_t* ids, int64_t lo, int64_t hi, int v) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (ids[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// One thread per query: fold the mimic column in and turn counters into the requested rank.
__global__ void finalize_ranks(int Q, int N, int mode, int minimize, int has_mimic,
                               const int32_t* __restrict__ cnt, const uint32_t* __restrict__ best,
                               const float* __restrict__ target, const float* __restrict__ self,
                               const int32_t* __restrict__ tgt_ent, const int64_t* __restrict__ beg,
                               const int64_t* __restrict__ end, const int32_t* __restrict__ ids,
                               float* __restrict__ out_target, float* __restrict__ out_best,
                               int64_t* __restrict__ out_rank, int32_t* __restrict__ out_cnt) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  int64_t strict = cnt[4 * q], tie = cnt[4 * q + 1], tie_lo = cnt[4 * q + 2];
  float b = kp_unord(best[q]);
  const float t = target[q];
  const int o = tgt_ent[q];
  const int64_t fb = beg[q], fe = end[q];
  const int64_t po = lower_bound_ids(ids, fb, fe, o);
  const bool o_in_f = po < fe && ids[po] == o;
  if (has_mimic && o != N) {
    const int64_t pn = lower_bound_ids(ids, fb, fe, N);
    const bool n_in_f = pn < fe && ids[pn] == N;
    if (!n_in_f) {
      const float s = self[q];
      strict += minimize ? (s < t) : (s > t);
      tie += (s == t);
      b = minimize ? fminf(b, s) : fmaxf(b, s);
    }
  }
  const int64_t n_f = (fe - fb) - (o_in_f ? 1 : 0);  // filtered entities other than the target
  const int64_t n_f_lo = po - fb;                     // ... of which with id < o
  int64_t rank;
  float bs;
  if (mode == KP_RANK_CONVE_SORT) {
    const float fv = 0.f;
    rank = 1 + strict + tie_lo + (fv > t ? n_f : 0) + (fv == t ? n_f_lo : 0);
    bs = fmaxf(fmaxf(b, t), n_f > 0 ? fv : -INFINITY);
  } else if (minimize) {
    const float fv = 1e6f;
    rank = strict + tie + 1 + (fv <= t ? n_f : 0);
    bs = fminf(fminf(b, t), n_f > 0 ? fv : INFINITY);
  } else {
    const float fv = -1e6f;
    const bool engine = (mode == KP_RANK_ENGINE_MAX);
    const int64_t self_count = (engine && o_in_f) ? (fv >= t ? 1 : 0) : 1;
    rank = strict + tie + self_count + (fv >= t ? n_f : 0);
    bs = fmaxf(fmaxf(b, (engine && o_in_f) ? fv : t), n_f > 0 ? fv : -INFINITY);
  }
  if (out_target) out_target[q] = t;
  if (out_best) out_best[q] = bs;
  if (out_rank) out_rank[q] = rank;
  if (out_cnt) {
    out_cnt[4 * q] = (int32_t)strict;
    out_cnt[4 * q + 1] = (int32_t)tie;
    out_cnt[4 * q + 2] = (int32_t)tie_lo;
    out_cnt[4 * q + 3] = o_in_f ? 1 : 0;
  }
}

// out[q, N] = self score (the mimic column of KelpieModel.all_scores)
__global__ void store_self(int Q, int N, const float* __restrict__ self, float* __restrict__ out, long long ld) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < Q) out[(long long)q * ld + N] = self[q];
}

}  // namespace

int kp_score_impl(kp_ctx* ctx, int Q, const int32_t* triples, const float* mimic, float* out, int64_t out_ld,
                  const int64_t* flt_off, const int32_t* flt_ids, int mode, float* target_score,
                  float* best_score, int64_t* rank, int32_t* counters, bool want_rank, cudaStream_t st) {
  const int D = ctx->D, N = (int)ctx->N;
  const int Qpad = ((Q + 63) / 64) * 64;
  size_t need = WsCursor::need((size_t)Qpad * D, 4) + 3 * WsCursor::need(Q, 4) + 2 * WsCursor::need(Q, 8) +
                WsCursor::need((size_t)Q * 4, 4) + WsCursor::need(Q, 4);
  int rc = kp_ws_reserve(ctx, need);
  if (rc != KP_OK) return rc;
  WsCursor ws{ctx->ws, ctx->ws + ctx->ws_bytes};
  float* qmat = ws.take<float>((size_t)Qpad * D);
  float* target = ws.take<float>(Q);
  float* self = ws.take<float>(Q);
  int32_t* tgt_ent = ws.take<int32_t>(Q);
  int64_t* fbeg = ws.take<int64_t>(Q);
  int64_t* fend = ws.take<int64_t>(Q);
  int32_t* cnt = ws.take<int32_t>((size_t)Q * 4);
  uint32_t* best = ws.take<uint32_t>(Q);

  const int wpb = 8;  // warps per block
  const int nb = (Q + wpb - 1) / wpb;
  int op = KP_OP_DOT, act = KP_ACT_NONE;
  if (ctx->kind == KP_TRANSE) op = (ctx->norm == 1) ? KP_OP_L1 : KP_OP_L2;
  if (ctx->kind == KP_CONVE) {
    act = KP_ACT_SIGMOID;
    if ((rc = kp_conve_features(ctx, Q, triples, 3, mimic, qmat, st)) != KP_OK) return rc;
  } else {
    prep_queries<<<nb, wpb * 32, 0, st>>>(ctx->kind, Q, N, (int)ctx->R2, D, ctx->ent, ctx->rel, triples, mimic, qmat);
    KP_LAUNCHED(ctx, 1);
  }
  const bool minimize = (ctx->kind == KP_TRANSE);

  kp_pass_args a{};
  a.qmat = qmat;
  a.Qn = Q;
  a.op = op;
  a.act = act;
  a.minimize = minimize;
  if (want_rank || mimic) {
    target_scores<<<nb, wpb * 32, 0, st>>>(op, act, kp_stream_usable(ctx, Q) ? 1 : 0, Q, N, D, ctx->ent, triples, mimic, qmat, target, self, tgt_ent);
    KP_LAUNCHED(ctx, 1);
  }
  if (!want_rank) {
    a.out = out;
    a.out_ld = out_ld;
    if ((rc = (kp_stream_usable(ctx, Q) ? kp_stream_launch(ctx, a, st) : kp_pass_launch(ctx, a, st))) != KP_OK) return rc;
    if (mimic) {
      store_self<<<(Q + 255) / 256, 256, 0, st>>>(Q, N, self, out, out_ld);
      KP_LAUNCHED(ctx, 1);
    }
    return KP_OK;
  }
  filter_ranges<<<(Q + 255) / 256, 256, 0, st>>>(Q, ctx->R2, triples, flt_off, ctx->n_keys, ctx->f_keys, ctx->f_off, fbeg, fend);
  KP_LAUNCHED(ctx, 1);
  const int32_t* ids = flt_off ? flt_ids : ctx->f_ids;
  KP_CUDA(ctx, cudaMemsetAsync(cnt, 0, (size_t)Q * 16, st));
  KP_CUDA(ctx, cudaMemsetAsync(best, minimize ? 0xff : 0x00, (size_t)Q * 4, st));
  a.rank = true;
  a.target = target;
  a.tgt_ent = tgt_ent;
  a.flt_beg = fbeg;
  a.flt_end = fend;
  a.flt_ids = ids;
  a.cnt = cnt;
  a.best = best;
  if (kp_rank_umma_usable(ctx, a))
    rc = kp_rank_umma_launch(ctx, a, st);
  else
    rc = kp_stream_usable(ctx, Q) ? kp_stream_launch(ctx, a, st) : kp_pass_launch(ctx, a, st);
  if (rc != KP_OK) return rc;
  finalize_ranks<<<(Q + 255) / 256, 256, 0, st>>>(Q, N, mode, minimize ? 1 : 0, mimic ? 1 : 0, cnt, best, target, self,
                                                 tgt_ent, fbeg, fend, ids, target_score, best_score, rank, counters);
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}

int kp_score_rows_impl(kp_ctx* ctx, int Q, const int32_t* triples, const float* mimic, float* out, cudaStream_t st) {
  const int D = ctx->D, N = (int)ctx->N;
  const int Qpad = ((Q + 63) / 64) * 64;
  int rc = kp_ws_reserve(ctx, WsCursor::need((size_t)Qpad * D, 4) + 2 * WsCursor::need(Q, 4));
  if (rc != KP_OK) return rc;
  WsCursor ws{ctx->ws, ctx->ws + ctx->ws_bytes};
  float* qmat = ws.take<float>((size_t)Qpad * D);
  float* self = ws.take<float>(Q);
  int32_t* tgt_ent = ws.take<int32_t>(Q);
  const int wpb = 8, nb = (Q + wpb - 1) / wpb;
  int op = KP_OP_DOT, act = KP_ACT_NONE;
  if (ctx->kind == KP_TRANSE) op = (ctx->norm == 1) ? KP_OP_L1 : KP_OP_L2;
  if (ctx->kind == KP_CONVE) {
    act = KP_ACT_SIGMOID;
    if ((rc = kp_conve_features(ctx, Q, triples, 3, mimic, qmat, st)) != KP_OK) return rc;
  } else {
    prep_queries<<<nb, wpb * 32, 0, st>>>(ctx->kind, Q, N, (int)ctx->R2, D, ctx->ent, ctx->rel, triples, mimic, qmat);
    KP_LAUNCHED(ctx, 1);
  }
  // the lane-strided accumulation of the streaming pass: 32 lanes per row instead of one sequential chain
  target_scores<<<nb, wpb * 32, 0, st>>>(op, act, (D % 4 == 0) ? 1 : 0, Q, N, D, ctx->ent, triples, mimic, qmat, out, self, tgt_ent);
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
