// HBM-streaming all-entity pass for FEW queries (Q <= 8): the regime where scoring is bound by
// reading the entity table once (N*D*4 bytes), e.g. get_triple_results of a single candidate
// (post_training_engine.py:101-125) or TransE.all_scores with a handful of queries
// (transe.py:48-65).  No shared-memory staging is needed: each warp owns a contiguous run of
// entity rows, every lane keeps its slice of the query vectors in registers, rows are read with
// coalesced 128-bit loads (4 rows in flight per warp), and the QT partial sums of a row are
// reduced across the warp with a halving butterfly (QT + 1 shuffles instead of 5 QT).
// STORE and fused filtered-RANK epilogues as in kp_pass.cu (per-query filter cursors live in lanes).
#include "kp_internal.h"
#include "kp_ptx.cuh"

namespace {

constexpr int ST_THREADS = 256;
constexpr int ROWS_PER_WARP_ITER = 4;

struct StreamK {
  int N, D, Qn, act, rank, minimize;
  long long rows_per_warp;
  const float* ent;
  const float* qmat;
  float* out;
  long long out_ld;
  const float* target;
  const int32_t* tgt_ent;
  const int64_t* flt_beg;
  const int64_t* flt_end;
  const int32_t* flt_ids;
  int32_t* cnt;
  uint32_t* best;
};

template <int OP>
__device__ __forceinline__ float acc4(float a, const float4& t, const float4& e) {
  if (OP == KP_OP_DOT) {
    a = __fmaf_rn(t.x, e.x, a); a = __fmaf_rn(t.y, e.y, a); a = __fmaf_rn(t.z, e.z, a); a = __fmaf_rn(t.w, e.w, a);
  } else if (OP == KP_OP_L2) {
    float d;
    d = t.x - e.x; a = __fmaf_rn(d, d, a); d = t.y - e.y; a = __fmaf_rn(d, d, a);
    d = t.z - e.z; a = __fmaf_rn(d, d, a); d = t.w - e.w; a = __fmaf_rn(d, d, a);
  } else {
    a += fabsf(t.x - e.x) + fabsf(t.y - e.y) + fabsf(t.z - e.z) + fabsf(t.w - e.w);
  }
  return a;
}

// Reduce QT per-lane partials over the 32 lanes; on return lane l (l < QT) holds the total of
// query l in v[0].  Halving butterfly: each round sends half of the live values.
template <int QT>
__device__ __forceinline__ void warp_reduce_multi(float (&v)[QT], int lane) {
  int live = QT, off = 16;
#pragma unroll
  for (int round = 0; round < 5; ++round, off >>= 1) {
    if (live > 1) {
      const int half = live >> 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < QT / 2; ++i) {
        if (i < half) {
          const float send = upper ? v[i] : v[i + half];
          const float recv = __shfl_xor_sync(0xffffffffu, send, off);
          v[i] = (upper ? v[i + half] : v[i]) + recv;
        }
      }
      live = half;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
}

// VPL = float4 per lane (D <= 128 * VPL)
template <int OP, int QT, int VPL>
__global__ void __launch_bounds__(ST_THREADS, 2) stream_kernel(const StreamK p) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (ST_THREADS / 32) + (threadIdx.x >> 5);
  const long long j_begin = warp * p.rows_per_warp;
  const long long j_end = min(j_begin + p.rows_per_warp, (long long)p.N);
  if (j_begin >= j_end) return;

  float4 q[QT][VPL];
#pragma unroll
  for (int i = 0; i < QT; ++i)
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int k = (v * 32 + lane) * 4;
      q[i][v] = (i < p.Qn && k < p.D) ? *reinterpret_cast<const float4*>(p.qmat + (size_t)i * p.D + k)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  // after the butterfly, lane `owner(i)` holds query i: with QT a power of two the value of query i
  // ends in the lanes whose high log2(QT) bits, read as the halving order, select i; recover it by
  // reducing a one-hot probe once.
  int my_query;
  {
    float probe[QT];
#pragma unroll
    for (int i = 0; i < QT; ++i) probe[i] = (lane == 0) ? (float)(i + 1) : 0.f;
    warp_reduce_multi<QT>(probe, lane);
    my_query = (int)probe[0] - 1;  // every lane ends up owning exactly one query id (replicated 32/QT times)
  }
  const bool owner = my_query >= 0 && my_query < p.Qn && (lane % (32 / QT) == 0 || QT == 32);
  float thr = 0.f;
  int tgt = -1;
  long long cur = 0, fend = 0;
  int c_strict = 0, c_tie = 0, c_tlo = 0;
  float bst = p.minimize ? INFINITY : -INFINITY;
  if (p.rank && owner) {
    thr = p.target[my_query];
    tgt = p.tgt_ent[my_query];
    long long lo = p.flt_beg[my_query], hi = p.flt_end[my_query];
    fend = hi;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (p.flt_ids[mid] < j_begin) lo = mid + 1; else hi = mid;
    }
    cur = lo;
  }

  for (long long j0 = j_begin; j0 < j_end; j0 += ROWS_PER_WARP_ITER) {
    float4 e[ROWS_PER_WARP_ITER][VPL];
#pragma unroll
    for (int r = 0; r < ROWS_PER_WARP_ITER; ++r)
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int k = (v * 32 + lane) * 4;
        const long long j = j0 + r;
        e[r][v] = (j < j_end && k < p.D) ? __ldg(reinterpret_cast<const float4*>(p.ent + (size_t)j * p.D + k))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int r = 0; r < ROWS_PER_WARP_ITER; ++r) {
      const long long j = j0 + r;
      if (j >= j_end) break;
      float s[QT];
#pragma unroll
      for (int i = 0; i < QT; ++i) {
        float a = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) a = acc4<OP>(a, q[i][v], e[r][v]);
        s[i] = a;
      }
      warp_reduce_multi<QT>(s, lane);
      float sc = s[0];
      if (OP == KP_OP_L2) sc = sqrtf(sc);
      if (p.act == KP_ACT_SIGMOID) sc = 1.f / (1.f + expf(-sc));
      if (owner) {
        if (p.rank) {
          bool masked = false;
          while (cur < fend) {
            const int id = p.flt_ids[cur];
            if (id > j) break;
            if (id == j) masked = true;
            ++cur;
          }
          if (!masked && (int)j != tgt) {
            const bool better = p.minimize ? (sc < thr) : (sc > thr);
            c_strict += better;
            c_tie += (sc == thr);
            c_tlo += (sc == thr && (int)j < tgt);
            bst = p.minimize ? fminf(bst, sc) : fmaxf(bst, sc);
          }
        } else {
          p.out[(long long)my_query * p.out_ld + j] = sc;
        }
      }
    }
  }
  if (p.rank && owner) {
    if (c_strict) atomicAdd(&p.cnt[my_query * 4 + 0], c_strict);
    if (c_tie) atomicAdd(&p.cnt[my_query * 4 + 1], c_tie);
    if (c_tlo) atomicAdd(&p.cnt[my_query * 4 + 2], c_tlo);
    if (p.minimize) atomicMin(&p.best[my_query], kp_ord(bst)); else atomicMax(&p.best[my_query], kp_ord(bst));
  }
}

template <int OP, int QT>
int launch_vpl(kp_ctx* ctx, const StreamK& p, int grid, cudaStream_t st) {
  const int vpl = (ctx->D + 127) / 128;
  KpTimer timer(ctx, kp_ctx::T_PASS, st);
  if (vpl <= 1) stream_kernel<OP, QT, 1><<<grid, ST_THREADS, 0, st>>>(p);
  else if (vpl <= 2) stream_kernel<OP, QT, 2><<<grid, ST_THREADS, 0, st>>>(p);
  else if (vpl <= 4) stream_kernel<OP, QT, 4><<<grid, ST_THREADS, 0, st>>>(p);
  else return -100;
  return KP_OK;
}

template <int OP>
int launch_qt(kp_ctx* ctx, const StreamK& p, int grid, cudaStream_t st) {
  if (p.Qn <= 1) return launch_vpl<OP, 1>(ctx, p, grid, st);
  if (p.Qn <= 2) return launch_vpl<OP, 2>(ctx, p, grid, st);
  if (p.Qn <= 4) return launch_vpl<OP, 4>(ctx, p, grid, st);
  return launch_vpl<OP, 8>(ctx, p, grid, st);
}

}  // namespace

bool kp_stream_usable(kp_ctx* ctx, int Qn) { return Qn <= 8 && ctx->D <= 512 && !ctx->force_tile; }

int kp_stream_launch(kp_ctx* ctx, const kp_pass_args& a, cudaStream_t st) {
  StreamK p;
  p.N = (int)ctx->N;
  p.D = ctx->D;
  p.Qn = a.Qn;
  p.act = a.act;
  p.rank = a.rank ? 1 : 0;
  p.minimize = a.minimize ? 1 : 0;
  p.ent = ctx->ent;
  p.qmat = a.qmat;
  p.out = a.out;
  p.out_ld = a.out_ld;
  p.target = a.target;
  p.tgt_ent = a.tgt_ent;
  p.flt_beg = a.flt_beg;
  p.flt_end = a.flt_end;
  p.flt_ids = a.flt_ids;
  p.cnt = a.cnt;
  p.best = a.best;
  // 8 CTAs of 8 warps per SM, every warp a contiguous run of rows (multiple of the unroll)
  const long long warps = (long long)ctx->sm_count * 8 * (ST_THREADS / 32);
  long long rpw = (ctx->N + warps - 1) / warps;
  rpw = ((rpw + ROWS_PER_WARP_ITER - 1) / ROWS_PER_WARP_ITER) * ROWS_PER_WARP_ITER;
  p.rows_per_warp = rpw;
  const long long used = (ctx->N + rpw - 1) / rpw;
  const int grid = (int)((used + ST_THREADS / 32 - 1) / (ST_THREADS / 32));
  int rc;
  switch (a.op) {
    case KP_OP_DOT: rc = launch_qt<KP_OP_DOT>(ctx, p, grid, st); break;
    case KP_OP_L2: rc = launch_qt<KP_OP_L2>(ctx, p, grid, st); break;
    default: rc = launch_qt<KP_OP_L1>(ctx, p, grid, st); break;
  }
  if (rc == -100) KP_FAIL(ctx, KP_EUNSUPPORTED, "streaming pass supports dim <= 512");
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
