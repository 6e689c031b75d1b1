// Filtered rank of many DOT queries (ComplEx complex.py:88-113, ConvE projection conve.py:155-158;
// rank semantics post_training_engine.py:101-125, model.py:42-68, conve.py:160-184) on the tensor
// cores, with ranks that are BIT-IDENTICAL to the exact fp32 pass (kp_pass.cu).
//
// S = Q E^T runs as bf16x3 split products on tcgen05 (cta_group::2: 256 queries x 128 entities per
// MMA, operands by TMA, S double-buffered in TMEM).  The tensor-core score S_tc of a pair differs from
// the fp32 FMA chain S_seq the exact pass evaluates by at most
//     margin(q, j) = kappa * |q|_2 * |E_j|_2 ,   kappa = 2^-16 + D * 2^-22
// (split residuals <= 2^-16 sum_k |q_k E_jk|, two fp32 accumulations of D terms, Cauchy-Schwarz).
// The epilogue thread of a query row therefore decides every entity whose score is further than the
// margin from the target's threshold on the spot and appends the few others to a re-check list; a
// second kernel evaluates those pairs with the exact pass's arithmetic (one sequential fp32 FMA
// chain) and adds them to the same counters.  Thresholds are pre-activation: for ConvE the largest
// logit whose sigmoid is still below the target's and the smallest whose sigmoid is above it are
// found by bisection per query (sigmoid saturates, so equal sigmoids of different logits are ties,
// exactly as in the reference); without activation they are the target's fp32 neighbours.
//
// Counters per query (as kp_pass.cu): strictly better / tied / tied with a smaller id, over entities
// that are neither filtered nor the target, and the best other score.
//
// TransE with the L2 norm (transe.py:48-65, a minimiser) runs through the same kernel: the epilogue turns
// the tensor-core dot product into the negated squared distance  x = 2 q.E_j - |q|^2 - |E_j|^2, which
// differs from MINUS the exact pass's fp32 chain  sum_k fl(q_k - E_jk)^2  by at most
//     margin(q, j) = kappa2 * (|q|^2 + |E_j|^2) ,   kappa2 = kappa + (4 D + 64) * 2^-24
// (2 kappa |q||E_j| <= kappa (|q|^2 + |E_j|^2) from the dot product; (D + 8) 2^-24 relative from each
// fp32 norm, 3 ulp from the epilogue arithmetic, 2 (D + 3) 2^-24 from the exact chain's own rounding).
// The "activation" is then f(x) = -sqrtf(-x) (monotone, sqrtf correctly rounded), the target -t, and
// "strictly better" means f(x) > -t exactly as for the maximisers; undecided pairs are re-evaluated
// with the exact pass's sequential  d = q_k - E_jk; acc = fma(d, d, acc); sqrtf(acc).
#include <cuda_bf16.h>

#include "kp_internal.h"
#include "kp_ptx.cuh"

namespace {

constexpr int RT = 192;      // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (one thread per query row)
constexpr int SLOT = 32768;  // {hi 16 KB | lo 16 KB}
constexpr int NSLOT = 6;

struct RCtl {
  uint64_t full[NSLOT], empty[NSLOT];
  uint64_t s_full[2];  // S(i) complete in buffer i&1         (MMA commit, both CTAs)
  uint64_t s_free[2];  // buffer read by the epilogue threads (256 arrivals, leader only)
  uint32_t tmem_base;
  float enorm[2][128];
};
constexpr size_t R_SMEM = (size_t)NSLOT * SLOT + sizeof(RCtl) + 1024;

constexpr int ACT_NEGSQRT = 2;  // internal: f(x) = -sqrtf(-x) on x <= 0 (L2 distance as a maximiser)

struct RK {
  int Qn, N, D, KB, n_tiles, tiles_per_strip;
  float kappa;             // DOT: kappa; L2: kappa2
  const float* enorm;      // [Npad]
  const float* qnorm;      // [Qpad]
  const float* xlo;        // [Qpad] largest pre-activation score that ranks strictly below the target
  const float* xhi;        // [Qpad] smallest that ranks strictly above
  const int32_t* tgt_ent;  // [Qn]
  const int64_t* flt_beg;
  const int64_t* flt_end;
  const int32_t* flt_ids;
  int32_t* cnt;            // [Qn, 4]
  uint32_t* best;          // [Qn] kp_ord of the best other PRE-activation score (approximate; finalised by the caller)
  int2* pairs;             // re-check list (query, entity)
  unsigned long long* n_pairs;
  unsigned long long cap_pairs;
  // exact re-evaluation in place when the list is full (degenerate scores): no host round trip decides anything
  const float* qmat;       // [Qn, D] fp32 queries
  const float* ent;        // [N, D] fp32 table
  const float* target;     // [Qn]
  uint32_t* best_act;      // [Qn] kp_ord of the best re-checked ACTIVATED score
  int act;
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

__device__ __forceinline__ float act_apply(int act, float x) {
  if (act == KP_ACT_SIGMOID) return 1.f / (1.f + expf(-x));
  if (act == ACT_NEGSQRT) return -sqrtf(-x);
  return x;
}

// Exact score of (query q, entity j) with the arithmetic of the exact pass (kp_pass.cu: one sequential fp32 FMA chain),
// folded into the counters.
__device__ __forceinline__ void recheck_pair(int q, int j, int D, int act, const float* __restrict__ qmat, const float* __restrict__ ent,
                                             const float* __restrict__ target, const int32_t* __restrict__ tgt_ent,
                                             int32_t* __restrict__ cnt, uint32_t* __restrict__ best_act) {
  const float4* a = reinterpret_cast<const float4*>(qmat + (size_t)q * D);
  const float4* b = reinterpret_cast<const float4*>(ent + (size_t)j * D);
  float acc = 0.f;
  if (act == ACT_NEGSQRT) {  // kp_pass.cu accum<KP_OP_L2>
    for (int k = 0; k < D / 4; ++k) {
      const float4 x = a[k], y = b[k];
      float d = __fsub_rn(x.x, y.x);
      acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(x.y, y.y);
      acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(x.z, y.z);
      acc = __fmaf_rn(d, d, acc);
      d = __fsub_rn(x.w, y.w);
      acc = __fmaf_rn(d, d, acc);
    }
    acc = -acc;
  } else {
    for (int k = 0; k < D / 4; ++k) {
      const float4 x = a[k], y = b[k];
      acc = __fmaf_rn(x.x, y.x, acc);
      acc = __fmaf_rn(x.y, y.y, acc);
      acc = __fmaf_rn(x.z, y.z, acc);
      acc = __fmaf_rn(x.w, y.w, acc);
    }
  }
  const float sc = act_apply(act, acc);  // L2: minus the distance, compared with minus the target's
  const float t = (act == ACT_NEGSQRT) ? -target[q] : target[q];
  if (sc > t) atomicAdd(&cnt[q * 4 + 0], 1);
  if (sc == t) {
    atomicAdd(&cnt[q * 4 + 1], 1);
    if (j < tgt_ent[q]) atomicAdd(&cnt[q * 4 + 2], 1);
  }
  atomicMax(&best_act[q], kp_ord(sc));
}

template <bool L2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(RT, 1)
rank_umma_kernel(const __grid_constant__ CUtensorMap eh64_map, const __grid_constant__ CUtensorMap el64_map,
                 const __grid_constant__ CUtensorMap qh_map, const __grid_constant__ CUtensorMap ql_map, const RK p) {
  extern __shared__ uint8_t rraw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(rraw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = sm;
  RCtl* ctl = reinterpret_cast<RCtl*>(sm + (size_t)NSLOT * SLOT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int strip = blockIdx.y, qtile = blockIdx.x;
  const int t0 = strip * p.tiles_per_strip;
  const int t1 = min(t0 + p.tiles_per_strip, p.n_tiles);
  const int ntile = t1 - t0;
  if (ntile <= 0) return;  // uniform over the pair
  const uint32_t crank = ptx::cluster_ctarank();
  const bool leader = crank == 0;

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
  if (warp == 1) {
    ptx::tmem_alloc2(&ctl->tmem_base, 256);
    ptx::tmem_relinquish2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tm = ctl->tmem_base;

  if (warp == 0) {
    // ------------------------------- TMA producer (both CTAs) -------------------------------
    if (lane == 0) {
      ptx::prefetch_tmap(&eh64_map);
      ptx::prefetch_tmap(&el64_map);
      ptx::prefetch_tmap(&qh_map);
      ptx::prefetch_tmap(&ql_map);
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
        for (int kb = 0; kb < p.KB; ++kb) {
          load(&qh_map, &ql_map, kb * 64, qtile * 128, 16384, 2 * 32768);                         // my query tile
          load(&eh64_map, &el64_map, kb * 64, (t0 + i) * 128 + (int)crank * 64, 8192, 2 * 16384);  // my half of the entities
        }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA only) -------------------------------
    if (lane == 0 && leader) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t ring_a = ptx::smem_u32(ring);
      const uint64_t DK = udesc(0, 16, 1024);
      uint32_t use = 0;
      for (int i = 0; i < ntile; ++i) {
        const int sb = i & 1;
        const uint32_t d_s = tm + sb * 128;
        if (i >= 2) {  // the epilogue threads of both CTAs have read S(i-2) out of this buffer
          ptx::mbar_wait_cluster(&ctl->s_free[sb], ((i >> 1) - 1) & 1);
          ptx::tc_fence_after();
        }
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait(&ctl->full[use % NSLOT], (use / NSLOT) & 1);
          ptx::mbar_wait(&ctl->full[(use + 1) % NSLOT], ((use + 1) / NSLOT) & 1);
          ptx::tc_fence_after();
          const uint32_t q_hi = ring_a + (use % NSLOT) * SLOT, q_lo = q_hi + 16384;
          const uint32_t e_hi = ring_a + ((use + 1) % NSLOT) * SLOT, e_lo = e_hi + 8192;
          const uint64_t ah = DK + (q_hi >> 4), al = DK + (q_lo >> 4), bh = DK + (e_hi >> 4), bl = DK + (e_lo >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kb * 4 + kk >= (p.D + 15) / 16) break;  // only zero padding beyond D
            ptx::umma2_bf16(d_s, ah + kk * 2, bh + kk * 2, idesc_s, (kb > 0 || kk > 0) ? 1u : 0u);
            ptx::umma2_bf16(d_s, ah + kk * 2, bl + kk * 2, idesc_s, 1u);
            ptx::umma2_bf16(d_s, al + kk * 2, bh + kk * 2, idesc_s, 1u);
          }
          ptx::umma2_commit_mc(&ctl->empty[use % NSLOT], 3);
          ptx::umma2_commit_mc(&ctl->empty[(use + 1) % NSLOT], 3);
          use += 2;
        }
        ptx::umma2_commit_mc(&ctl->s_full[sb], 3);
      }
    }
  } else {
    // ------------------------------- epilogue: one thread per query row -------------------------------
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    const int q = qtile * 128 + row;
    const bool live = q < p.Qn;
    const uint32_t s_free_leader0 = ptx::mapa_u32(ptx::smem_u32(&ctl->s_free[0]), 0);
    const float xlo = live ? p.xlo[q] : INFINITY, xhi = live ? p.xhi[q] : INFINITY;
    const float qn = live ? p.qnorm[q] : 0.f;
    const float qm = qn * p.kappa;  // DOT margin factor
    const float qn2 = qn * qn;      // L2
    const int tgt = live ? p.tgt_ent[q] : -1;
    long long cur = 0, fend = 0;
    if (live) {
      long long lo = p.flt_beg[q], hi = p.flt_end[q];
      fend = hi;
      const int first = t0 * 128;
      while (lo < hi) {  // lower_bound(first)
        const long long mid = (lo + hi) >> 1;
        if (p.flt_ids[mid] < first) lo = mid + 1; else hi = mid;
      }
      cur = lo;
    }
    int strict = 0;
    float best = -INFINITY;
    for (int i = 0; i < ntile; ++i) {
      const int sb = i & 1;
      const int j0 = (t0 + i) * 128;
      // norms of this tile's entities (row r of the warp group fetches entity j0 + r)
      ctl->enorm[sb][row] = p.enorm[j0 + row];
      // filter bits of my query for this tile
      uint32_t msk[4] = {0, 0, 0, 0};
      {
        const int jend = j0 + 128;
        while (cur < fend) {
          const int id = p.flt_ids[cur];
          if (id >= jend) break;
          if (id >= j0) msk[(id - j0) >> 5] |= 1u << ((id - j0) & 31);
          ++cur;
        }
        if (tgt >= j0 && tgt < jend) msk[(tgt - j0) >> 5] |= 1u << ((tgt - j0) & 31);  // the target itself is not counted
        if (jend > p.N) {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (j0 + c >= p.N) msk[c >> 5] |= 1u << (c & 31);
        }
      }
      ptx::bar_sync(1, 128);  // enorm[sb] written by all four warps
      ptx::mbar_wait(&ctl->s_full[sb], (i >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t s_addr = tm + sb * 128 + lane_off;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(s_addr + c0, r);
        ptx::tmem_ld_wait();
        const uint32_t mw = msk[c0 >> 5];
        uint32_t unsure = 0;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float s = __uint_as_float(r[c]);
          float m;
          if (L2) {
            const float en = ctl->enorm[sb][c0 + c];
            const float nn = __fmaf_rn(en, en, qn2);
            s = __fmaf_rn(2.f, s, -nn);
            m = p.kappa * nn;
          } else {
            m = qm * ctl->enorm[sb][c0 + c];
          }
          const bool valid = !((mw >> c) & 1u);
          const bool above = (s - m >= xhi), below = (s + m <= xlo);
          strict += (valid && above);
          if (valid) best = fmaxf(best, s);
          if (valid && !above && !below) unsure |= 1u << c;
        }
        while (unsure) {  // rare: exact re-check by the second kernel
          const int c = __ffs(unsure) - 1;
          unsure &= unsure - 1;
          const unsigned long long at = atomicAdd(p.n_pairs, 1ull);
          if (at < p.cap_pairs)
            p.pairs[at] = make_int2(q, j0 + c0 + c);
          else  // list full (degenerate scores): evaluate here, slowly but exactly
            recheck_pair(q, j0 + c0 + c, p.D, p.act, p.qmat, p.ent, p.target, p.tgt_ent, p.cnt, p.best_act);
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive_cluster(s_free_leader0 + 8u * sb);
    }
    if (live) {
      if (strict) atomicAdd(&p.cnt[q * 4 + 0], strict);
      atomicMax(&p.best[q], kp_ord(best));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc2(tm, 256);
}


// Per query: |q|_2 and the pre-activation thresholds.  target[q] is the (activated) target score the
// exact pass would compare with; NaN (invalid triple) -> nothing is ever better or tied.
__global__ void rank_prepare(int Q, int Qpad, int D, int act, const float* __restrict__ qmat, const float* __restrict__ target,
                             float* __restrict__ qnorm, float* __restrict__ xlo, float* __restrict__ xhi) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= Qpad) return;
  float a = 0.f;
  if (q < Q)
    for (int k = lane; k < D; k += 32) a = __fmaf_rn(qmat[(size_t)q * D + k], qmat[(size_t)q * D + k], a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane != 0) return;
  qnorm[q] = sqrtf(a);
  // lo = max{x : f(x) < t-}, hi = min{x : f(x) > t+}; NaN = no such x (every comparison with it is false).
  // Without activation t- = t+ = t.  expf is accurate to 2 ulp but not guaranteed monotone, so for the
  // sigmoid the thresholds are taken 4 ulp of the target away: pairs inside go to the exact re-check.
  const float none = __int_as_float(0x7fc00000);
  float lo = INFINITY, hi = none;  // padding rows / invalid target: everything certainly below
  if (q < Q) {
    const float t = (act == ACT_NEGSQRT) ? -target[q] : target[q];
    const float top = (act == ACT_NEGSQRT) ? 0.f : INFINITY;  // largest pre-activation value of the domain
    if (t == t) {
      const int slack = (act == KP_ACT_SIGMOID) ? 4 : 0;
      const float tm = kp_unord(kp_ord(t) - slack), tp = kp_unord(kp_ord(t) + slack);
      const uint32_t a0 = kp_ord(-INFINITY), b0 = kp_ord(top);
      if (!(act_apply(act, -INFINITY) < tm)) {
        lo = none;
      } else if (act_apply(act, top) < tm) {
        lo = INFINITY;
      } else {
        uint32_t l = a0, h = b0;  // f(l) < t-, f(h) >= t-
        while (h - l > 1) {
          const uint32_t mid = l + ((h - l) >> 1);
          if (act_apply(act, kp_unord(mid)) < tm) l = mid; else h = mid;
        }
        lo = kp_unord(l);
      }
      if (!(act_apply(act, top) > tp)) {
        hi = none;
      } else if (act_apply(act, -INFINITY) > tp) {
        hi = -INFINITY;
      } else {
        uint32_t l = a0, h = b0;  // f(l) <= t+, f(h) > t+
        while (h - l > 1) {
          const uint32_t mid = l + ((h - l) >> 1);
          if (act_apply(act, kp_unord(mid)) > tp) h = mid; else l = mid;
        }
        hi = kp_unord(h);
      }
    }
  }
  xlo[q] = lo;
  xhi[q] = hi;
}

// Exact arithmetic of kp_pass.cu for the pairs the tensor-core pass could not decide: one thread per pair.
// The undecided pairs of the tensor-core pass, re-evaluated exactly.  The list length is read on the DEVICE (fixed
// grid-stride launch): the host never waits for the pass, so a caller can queue step k+1 behind step k.
__global__ void rank_recheck(const unsigned long long* __restrict__ n_ptr, unsigned long long cap, unsigned long long* __restrict__ total,
                             const int2* __restrict__ pairs, int D, int act, const float* __restrict__ qmat,
                             const float* __restrict__ ent, const float* __restrict__ target, const int32_t* __restrict__ tgt_ent,
                             int32_t* __restrict__ cnt, uint32_t* __restrict__ best_act) {
  const unsigned long long all = *n_ptr, n = all < cap ? all : cap;
  if (blockIdx.x == 0 && threadIdx.x == 0 && total) atomicAdd(total, all);
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x)
    recheck_pair(pairs[i].x, pairs[i].y, D, act, qmat, ent, target, tgt_ent, cnt, best_act);
}

// best[q]: pre-activation maximum from the tensor-core pass -> activated, merged with the re-checked exact scores
__global__ void rank_best_finish(int Q, int act, const uint32_t* __restrict__ best_pre, const uint32_t* __restrict__ best_act,
                                 uint32_t* __restrict__ best) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  float pre = kp_unord(best_pre[q]);
  if (act == ACT_NEGSQRT) pre = fminf(pre, 0.f);  // a tensor-core estimate of a tiny squared distance may come out negative
  const float a = (pre == -INFINITY || pre != pre) ? -INFINITY : act_apply(act, pre);  // no valid entity: as the exact pass
  const float b = best_act[q] ? kp_unord(best_act[q]) : -INFINITY;
  const float m = fmaxf(a, b);
  best[q] = kp_ord(act == ACT_NEGSQRT ? -m : m);  // L2: back to a distance (the caller keeps the minimum)
}

}  // namespace

bool kp_rank_umma_usable(kp_ctx* ctx, const kp_pass_args& a) {
  const bool dot = a.op == KP_OP_DOT && !a.minimize;
  const bool l2 = a.op == KP_OP_L2 && a.minimize && a.act == KP_ACT_NONE;
  return ctx->umma_rank != 0 && !ctx->force_simt && a.rank && (dot || l2) && a.Qn >= 128 && ctx->D <= 512 && ctx->D % 4 == 0;
}

int kp_rank_umma_launch(kp_ctx* ctx, const kp_pass_args& a, cudaStream_t st) {
  int rc;
  if ((rc = kp_umma_tables(ctx, st)) != KP_OK) return rc;
  const int Q = a.Qn, D = ctx->D;
  const bool l2 = a.op == KP_OP_L2;
  const int act = l2 ? ACT_NEGSQRT : a.act;
  const int n_qt = ((Q + 255) / 256) * 2;  // pairs of query tiles
  const long long Qpad = (long long)n_qt * 128;
  CUtensorMap qh_map, ql_map;
  // scratch behind the split queries (arena 1 is grow-only and the split sits at its start)
  const size_t split_bytes = 2 * (((size_t)Qpad * ctx->um.Dpad * 2 + 1023) & ~size_t(1023)) + 2048;
  unsigned long long cap = (unsigned long long)Q * 4096ull;
  if (cap > (64ull << 20)) cap = 64ull << 20;
  if (cap < (1ull << 16)) cap = 1ull << 16;
  const size_t extra = 3 * WsCursor::need(Qpad, 4) + 2 * WsCursor::need(Q, 4) + WsCursor::need(2, 8) + WsCursor::need(cap, 8);
  if ((rc = kp_ws_reserve(ctx, split_bytes + extra, 1)) != KP_OK) return rc;
  if ((rc = kp_umma_split_rows(ctx, a.qmat, Q, Qpad, &qh_map, &ql_map, st)) != KP_OK) return rc;
  WsCursor ws{ctx->ws_arena[1] + split_bytes, ctx->ws_arena[1] + ctx->ws_arena_bytes[1]};
  float* qnorm = ws.take<float>(Qpad);
  float* xlo = ws.take<float>(Qpad);
  float* xhi = ws.take<float>(Qpad);
  uint32_t* best_pre = ws.take<uint32_t>(Q);
  uint32_t* best_act = ws.take<uint32_t>(Q);
  unsigned long long* n_pairs = ws.take<unsigned long long>(2);
  int2* pairs = ws.take<int2>(cap);

  rank_prepare<<<(unsigned)((Qpad + 7) / 8), 256, 0, st>>>(Q, (int)Qpad, D, act, a.qmat, a.target, qnorm, xlo, xhi);
  KP_LAUNCHED(ctx, 1);
  KP_CUDA(ctx, cudaMemsetAsync(best_pre, 0, (size_t)Q * 4, st));
  KP_CUDA(ctx, cudaMemsetAsync(best_act, 0, (size_t)Q * 4, st));
  KP_CUDA(ctx, cudaMemsetAsync(n_pairs, 0, 16, st));

  RK p;
  p.Qn = Q;
  p.N = (int)ctx->N;
  p.D = D;
  p.KB = (D + 63) / 64;
  p.n_tiles = (int)((ctx->N + 127) / 128);
  const int s = kp_plan_strips(n_qt / 2, ctx->sm_count / 2, p.n_tiles);
  p.tiles_per_strip = (p.n_tiles + s - 1) / s;
  const int n_strips = (p.n_tiles + p.tiles_per_strip - 1) / p.tiles_per_strip;
  p.kappa = ldexpf(1.f, -16) + (float)D * ldexpf(1.f, -22);
  if (l2) p.kappa += (float)(4 * D + 64) * ldexpf(1.f, -24);
  p.enorm = ctx->um.enorm;
  p.qnorm = qnorm;
  p.xlo = xlo;
  p.xhi = xhi;
  p.tgt_ent = a.tgt_ent;
  p.flt_beg = a.flt_beg;
  p.flt_end = a.flt_end;
  p.flt_ids = a.flt_ids;
  p.cnt = a.cnt;
  p.best = best_pre;
  p.pairs = pairs;
  p.n_pairs = n_pairs;
  p.cap_pairs = cap;
  p.qmat = a.qmat;
  p.ent = ctx->ent;
  p.target = a.target;
  p.best_act = best_act;
  p.act = act;
  KP_SMEM_ONCE(ctx, (rank_umma_kernel<false>), R_SMEM);
  KP_SMEM_ONCE(ctx, (rank_umma_kernel<true>), R_SMEM);
  {
    KpTimer timer(ctx, kp_ctx::T_PASS, st);
    if (l2)
      rank_umma_kernel<true><<<dim3(n_qt, n_strips, 1), RT, R_SMEM, st>>>(ctx->um.eh64_map, ctx->um.el64_map, qh_map, ql_map, p);
    else
      rank_umma_kernel<false><<<dim3(n_qt, n_strips, 1), RT, R_SMEM, st>>>(ctx->um.eh64_map, ctx->um.el64_map, qh_map, ql_map, p);
  }
  KP_LAUNCHED(ctx, 1);
  // the undecided pairs (0.1 - 0.2 %) are re-evaluated by a fixed grid that reads the list length on the device
  if (!ctx->rank_recheck_total) {
    void* d = nullptr;
    if (cudaMalloc(&d, sizeof(unsigned long long)) != cudaSuccess) KP_FAIL(ctx, KP_ENOMEM, "re-check counter");
    ctx->owned.push_back(d);
    KP_CUDA(ctx, cudaMemsetAsync(d, 0, sizeof(unsigned long long), st));
    ctx->rank_recheck_total = static_cast<unsigned long long*>(d);
  }
  rank_recheck<<<(unsigned)ctx->sm_count * 8, 256, 0, st>>>(n_pairs, cap, ctx->rank_recheck_total, pairs, D, act, a.qmat, ctx->ent,
                                                            a.target, a.tgt_ent, a.cnt, best_act);
  KP_LAUNCHED(ctx, 1);
  rank_best_finish<<<(Q + 255) / 256, 256, 0, st>>>(Q, act, best_pre, best_act, a.best);
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
