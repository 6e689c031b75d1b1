// All-entity scoring pass on CUDA cores: S[q, j] = op(qmat[q, :], ent[j, :]) for a tile of
// 64 queries x 128 entities per CTA, entity and query chunks staged in shared memory by TMA
// through a 6-deep mbarrier ring, with either a STORE epilogue (Model.all_scores) or the fused
// filtered-RANK epilogue (post_training_engine.py:101-125, model.py:42-68) that never
// materialises the scores.
//
//   op  = DOT (ComplEx complex.py:88-113, ConvE projection conve.py:155)
//       | L2 / L1 distance (TransE transe.py:48-65, the table is streamed once for 64 queries)
//
// Layout: ent is [N, D] fp32 row-major in HBM; a stage holds a [128 rows x 32 floats] entity
// chunk written by TMA with the 128-byte swizzle (so that lane <-> entity row reads are
// bank-conflict free) and a [64 x 32] query chunk (read as warp broadcasts).
#include "kp_internal.h"
#include "kp_ptx.cuh"

namespace {

constexpr int QT = 64, NT = 128, KC = 32, RQ = 8, RJ = 4;
constexpr int N_CONSUMERS = 256, N_THREADS = N_CONSUMERS + 32;
constexpr int STAGE_E = NT * KC * 4, STAGE_Q = QT * KC * 4, STAGE = STAGE_E + STAGE_Q;
constexpr int NSTAGE = 6;

struct Ctl {
  uint64_t full[NSTAGE];
  uint64_t empty[NSTAGE];
  float thr[QT];
  int tgt[QT];
  uint32_t mask[2][QT][NT / 32];
};
constexpr size_t SMEM_BYTES = size_t(NSTAGE) * STAGE + sizeof(Ctl) + 1024;

struct PassK {
  int N, D, Qn, nchunks, n_jtiles, tiles_per_strip;
  int act, rank, minimize;
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
__device__ __forceinline__ void accum(float& a, float t, float e) {
  if (OP == KP_OP_DOT) {
    a = __fmaf_rn(t, e, a);
  } else if (OP == KP_OP_L2) {
    float d = __fsub_rn(t, e);
    a = __fmaf_rn(d, d, a);
  } else {
    a = __fadd_rn(a, fabsf(__fsub_rn(t, e)));
  }
}

template <int OP>
__global__ void __launch_bounds__(N_THREADS, 1)
pass_kernel(const __grid_constant__ CUtensorMap emap, const __grid_constant__ CUtensorMap qmap, const PassK p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + size_t(NSTAGE) * STAGE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.y * QT;
  const int jt0 = blockIdx.x * p.tiles_per_strip;
  const int jt1 = min(jt0 + p.tiles_per_strip, p.n_jtiles);
  if (jt0 >= jt1) return;
  const int total_iters = (jt1 - jt0) * p.nchunks;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], N_CONSUMERS / 32);
    }
    ptx::fence_barrier_init();
  }
  if (tid < QT) {
    int q = q0 + tid;
    ctl->thr[tid] = (p.rank && q < p.Qn) ? p.target[q] : 0.f;
    ctl->tgt[tid] = (p.rank && q < p.Qn) ? p.tgt_ent[q] : -1;
  }
  __syncthreads();

  if (warp == N_CONSUMERS / 32) {
    // ---------------- producer warp: one lane drives TMA ----------------
    if (lane == 0) {
      ptx::prefetch_tmap(&emap);
      ptx::prefetch_tmap(&qmap);
      for (int it = 0; it < total_iters; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (it / NSTAGE) & 1;
        ptx::mbar_wait(&ctl->empty[s], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&ctl->full[s], STAGE);
        const int jt = jt0 + it / p.nchunks, c = it % p.nchunks;
        uint8_t* st = smem + size_t(s) * STAGE;
        ptx::tma_load_2d(st, &emap, &ctl->full[s], c * KC, jt * NT);
        ptx::tma_load_2d(st + STAGE_E, &qmap, &ctl->full[s], c * KC, q0);
      }
    }
    return;
  }

  // ---------------- consumer warps ----------------
  const int qg = warp;  // queries qg*RQ .. qg*RQ+RQ-1 of the tile
  int cs[RQ], ct[RQ], cl[RQ];
  float bst[RQ];
#pragma unroll
  for (int i = 0; i < RQ; ++i) {
    cs[i] = ct[i] = cl[i] = 0;
    bst[i] = p.minimize ? INFINITY : -INFINITY;
  }
  // filter cursor of query `tid` (threads 0..QT-1)
  long long cur = 0, fend = 0;
  if (p.rank && tid < QT && q0 + tid < p.Qn) {
    long long lo = p.flt_beg[q0 + tid], hi = p.flt_end[q0 + tid];
    fend = hi;
    const int first = jt0 * NT;
    while (lo < hi) {  // lower_bound(first)
      long long mid = (lo + hi) >> 1;
      if (p.flt_ids[mid] < first) lo = mid + 1; else hi = mid;
    }
    cur = lo;
  }

  int it = 0;
  for (int jt = jt0; jt < jt1; ++jt) {
    float acc[RQ][RJ];
#pragma unroll
    for (int i = 0; i < RQ; ++i)
#pragma unroll
      for (int r = 0; r < RJ; ++r) acc[i][r] = 0.f;

    for (int c = 0; c < p.nchunks; ++c, ++it) {
      const int s = it % NSTAGE;
      const uint32_t ph = (it / NSTAGE) & 1;
      ptx::mbar_wait(&ctl->full[s], ph);
      const float4* E4 = reinterpret_cast<const float4*>(smem + size_t(s) * STAGE);
      const float4* Q4 = reinterpret_cast<const float4*>(smem + size_t(s) * STAGE + STAGE_E);
#pragma unroll
      for (int kk = 0; kk < KC / 4; ++kk) {
        float4 e[RJ];
#pragma unroll
        for (int r = 0; r < RJ; ++r) {
          const int row = lane + 32 * r;
          e[r] = E4[row * (KC / 4) + (kk ^ (row & 7))];
        }
#pragma unroll
        for (int i = 0; i < RQ; ++i) {
          const float4 t = Q4[(qg * RQ + i) * (KC / 4) + kk];
#pragma unroll
          for (int r = 0; r < RJ; ++r) {
            accum<OP>(acc[i][r], t.x, e[r].x);
            accum<OP>(acc[i][r], t.y, e[r].y);
            accum<OP>(acc[i][r], t.z, e[r].z);
            accum<OP>(acc[i][r], t.w, e[r].w);
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ctl->empty[s]);
    }

    // ---- epilogue of tile jt ----
    const int j0 = jt * NT;
    if (p.rank) {
      if (tid < QT) {
        uint32_t w[NT / 32] = {0, 0, 0, 0};
        const int jend = j0 + NT;
        while (cur < fend) {
          const int id = p.flt_ids[cur];
          if (id >= jend) break;
          if (id >= j0) w[(id - j0) >> 5] |= 1u << ((id - j0) & 31);
          ++cur;
        }
#pragma unroll
        for (int k = 0; k < NT / 32; ++k) ctl->mask[jt & 1][tid][k] = w[k];
      }
      ptx::bar_sync(1, N_CONSUMERS);
    }
#pragma unroll
    for (int i = 0; i < RQ; ++i) {
      const int ql = qg * RQ + i;
      const int q = q0 + ql;
      const float thr = ctl->thr[ql];
      const int tgt = ctl->tgt[ql];
#pragma unroll
      for (int r = 0; r < RJ; ++r) {
        const int j = j0 + lane + 32 * r;
        float sc = acc[i][r];
        if (OP == KP_OP_L2) sc = sqrtf(sc);
        if (p.act == KP_ACT_SIGMOID) sc = 1.f / (1.f + expf(-sc));
        if (p.rank) {
          const bool masked = (ctl->mask[jt & 1][ql][r] >> lane) & 1u;
          const bool valid = (j < p.N) && !masked && (j != tgt);
          const bool better = p.minimize ? (sc < thr) : (sc > thr);
          const bool tie = (sc == thr);
          cs[i] += (valid && better);
          ct[i] += (valid && tie);
          cl[i] += (valid && tie && j < tgt);
          if (valid) bst[i] = p.minimize ? fminf(bst[i], sc) : fmaxf(bst[i], sc);
        } else if (j < p.N && q < p.Qn) {
          p.out[(long long)q * p.out_ld + j] = sc;
        }
      }
    }
  }

  if (p.rank) {
#pragma unroll
    for (int i = 0; i < RQ; ++i) {
      int a = cs[i], b = ct[i], c = cl[i];
      float m = bst[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
        const float mo = __shfl_xor_sync(0xffffffffu, m, o);
        m = p.minimize ? fminf(m, mo) : fmaxf(m, mo);
      }
      const int q = q0 + qg * RQ + i;
      if (lane == 0 && q < p.Qn) {
        if (a) atomicAdd(&p.cnt[q * 4 + 0], a);
        if (b) atomicAdd(&p.cnt[q * 4 + 1], b);
        if (c) atomicAdd(&p.cnt[q * 4 + 2], c);
        if (p.minimize)
          atomicMin(&p.best[q], kp_ord(m));
        else
          atomicMax(&p.best[q], kp_ord(m));
      }
    }
  }
}

template <int OP>
int launch(kp_ctx* ctx, const CUtensorMap& qmap, const PassK& p, dim3 grid, cudaStream_t st) {
  KP_SMEM_ONCE(ctx, (pass_kernel<OP>), SMEM_BYTES);
  KpTimer timer(ctx, kp_ctx::T_PASS, st);
  pass_kernel<OP><<<grid, N_THREADS, SMEM_BYTES, st>>>(ctx->ent_map, qmap, p);
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}

}  // namespace

int kp_pass_launch(kp_ctx* ctx, const kp_pass_args& a, cudaStream_t st) {
  PassK p;
  p.N = (int)ctx->N;
  p.D = ctx->D;
  p.Qn = a.Qn;
  p.nchunks = (ctx->D + KC - 1) / KC;
  p.n_jtiles = (int)((ctx->N + NT - 1) / NT);
  const int n_qtiles = (a.Qn + QT - 1) / QT;
  int n_strips = ctx->sm_count / n_qtiles;
  if (n_strips < 1) n_strips = 1;
  if (n_strips > p.n_jtiles) n_strips = p.n_jtiles;
  p.tiles_per_strip = (p.n_jtiles + n_strips - 1) / n_strips;
  n_strips = (p.n_jtiles + p.tiles_per_strip - 1) / p.tiles_per_strip;
  p.act = a.act;
  p.rank = a.rank ? 1 : 0;
  p.minimize = a.minimize ? 1 : 0;
  p.out = a.out;
  p.out_ld = a.out_ld;
  p.target = a.target;
  p.tgt_ent = a.tgt_ent;
  p.flt_beg = a.flt_beg;
  p.flt_end = a.flt_end;
  p.flt_ids = a.flt_ids;
  p.cnt = a.cnt;
  p.best = a.best;
  CUtensorMap qmap;
  // qmat is allocated with its row count padded to a multiple of QT (kp_pass_args contract)
  int rc = kp_encode_2d_f32(ctx, &qmap, a.qmat, (int64_t)n_qtiles * QT, ctx->D, ctx->D, QT, KC, false);
  if (rc != KP_OK) return rc;
  dim3 grid(n_strips, n_qtiles);
  switch (a.op) {
    case KP_OP_DOT:
      return launch<KP_OP_DOT>(ctx, qmap, p, grid, st);
    case KP_OP_L2:
      return launch<KP_OP_L2>(ctx, qmap, p, grid, st);
    default:
      return launch<KP_OP_L1>(ctx, qmap, p, grid, st);
  }
}
