// Fused score -> softmax/sigmoid -> contract pass on the 5th-gen tensor cores (tcgen05 + TMEM),
// operands staged by TMA.  Same contract and partial-result layout as kp_flash.cu (CUDA cores):
//   per query row g and entity strip:  m, l = softmax statistics,  O = sum_j p_gj * E_j.
//
// fp32 parity on bf16 tensor cores: every fp32 operand x is split as x = hi + lo (two bf16,
// ~16 mantissa bits) and each product is evaluated as hi*hi + hi*lo + lo*hi with fp32
// accumulation in TMEM (3 MMAs per algorithmic MMA; the dropped lo*lo term is ~2^-18 relative).
// The entity table is split once per context ([N,D] fp32 -> two [Npad,Dpad] bf16 tables, the
// same 4 bytes per element in HBM); query rows and the probabilities P are split on the fly.
//
// One CTA = 128 query rows x one strip of 128-entity tiles x one chunk of <= 256 output dims:
//   warp 0      TMA producer (6-slot ring of 32 KB {hi,lo} boxes of [128 rows x 64 dims])
//   warp 1      MMA issuer   S = Q E^T  (M128 N128, K-major A/B from smem)   -> TMEM S[2]
//                            O += P E   (M128 N64,  A = P from TMEM, B = E MN-major from smem)
//   warps 2..5  one thread per query row: TMEM S -> online softmax (lazy rescale of O in
//               TMEM) or sigmoid -> P split into bf16 hi/lo and written back to TMEM over the
//               S columns it was computed from (no shared-memory round trip for P)
// PV of tile i-1 is issued behind S of tile i so the softmax of a tile overlaps tensor work.
//
// Thread-block clusters: the CTAs of `cq` neighbouring query tiles x all dim chunks form one
// cluster and walk the same entity tiles in lock step.  Every operand box is fetched from L2 by
// ONE CTA of its sharer set and TMA-multicast into the same ring slot of every sharer (entity
// boxes of the S phase: all CTAs; query boxes: the CTAs of that query tile; entity boxes of the
// PV phase: the CTAs of that dim chunk); a slot is re-used once every CTA of the cluster has
// released it (tcgen05.commit multicast onto all `empty` barriers).
#include <cuda_bf16.h>

#include "kp_flash.cuh"
#include "kp_internal.h"
#include "kp_ptx.cuh"
#include "kp_umma_softmax.cuh"

namespace {

constexpr int UT = 192;               // threads
constexpr int SLOT = 32768;           // {hi 16 KB | lo 16 KB}
constexpr int NSLOT = 6;
constexpr float RESCALE_TAU = 8.0f;   // rescale O only when the row max grew by more than this

struct UCtl {
  uint64_t full[NSLOT], empty[NSLOT];
  uint64_t s_full[2];
  uint64_t p_full, pv_done, o_done;
  uint32_t tmem_base;
};
constexpr size_t U_SMEM = (size_t)NSLOT * SLOT + sizeof(UCtl) + 1024;

struct UK {
  int G, N, D, KB, n_tiles, tiles_per_strip, boxes_per_chunk, mode;
  int cq;  // query tiles per cluster (cluster = cq query tiles x all dim chunks)
  float* part_m;
  float* part_l;
  float* part_O;
};

__device__ __forceinline__ uint64_t udesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // Blackwell descriptor version
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

__global__ void __launch_bounds__(UT, 1)
flash_umma_kernel(const __grid_constant__ CUtensorMap eh_map, const __grid_constant__ CUtensorMap el_map,
                  const __grid_constant__ CUtensorMap qh_map, const __grid_constant__ CUtensorMap ql_map, const UK p) {
  extern __shared__ uint8_t uraw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(uraw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = sm;
  UCtl* ctl = reinterpret_cast<UCtl*>(sm + (size_t)NSLOT * SLOT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int strip = blockIdx.x, qtile = blockIdx.y, chunk = blockIdx.z;
  const int t0 = strip * p.tiles_per_strip;
  const int t1 = min(t0 + p.tiles_per_strip, p.n_tiles);
  const int ntile = t1 - t0;
  const int box0 = chunk * p.boxes_per_chunk;
  const int nbox = p.boxes_per_chunk;  // identical for every chunk (tables are zero-padded)
  if (ntile <= 0) return;              // uniform over the cluster (same strip)
  const uint32_t csize = ptx::cluster_nctarank(), crank = ptx::cluster_ctarank();
  const uint32_t cq = (uint32_t)p.cq, cc = csize / cq;
  const uint32_t qsel = crank % cq, csel = crank / cq;
  const uint16_t mask_all = (uint16_t)((1u << csize) - 1);
  uint16_t mask_q = 0, mask_c = 0;  // CTAs sharing my query tile / my dim chunk
  for (uint32_t c = 0; c < cc; ++c) mask_q |= (uint16_t)(1u << (qsel + cq * c));
  for (uint32_t q = 0; q < cq; ++q) mask_c |= (uint16_t)(1u << (q + cq * csel));

  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], csize);
    }
    for (int b = 0; b < 2; ++b) ptx::mbar_init(&ctl->s_full[b], 1);
    ptx::mbar_init(&ctl->p_full, 128);
    ptx::mbar_init(&ctl->pv_done, 1);
    ptx::mbar_init(&ctl->o_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&ctl->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (csize > 1) ptx::cluster_sync_all();  // barriers of every CTA initialised before any remote arrive
  ptx::tc_fence_after();
  const uint32_t tm = ctl->tmem_base;
  // O: cols [0,256).  S[b]: cols [256 + 128 b, +128); after the softmax the same columns hold P:
  // for every 32-column chunk c, cols [32c, +16) = bf16x2-packed hi and [32c+16, +16) = lo of the
  // 32 entities of that chunk (MMA K-step ks reads 8 packed columns at 32 (ks/2) + 8 (ks%2) (+16)).
  const uint32_t TM_O = tm, TM_S = tm + 256;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      ptx::prefetch_tmap(&eh_map);
      ptx::prefetch_tmap(&el_map);
      ptx::prefetch_tmap(&qh_map);
      ptx::prefetch_tmap(&ql_map);
      uint32_t use = 0, n_q = 0, n_es = 0, n_pv = 0;
      // every CTA arms its own `full` barrier; only the issuer of the sharer set fetches the box
      auto load = [&](const CUtensorMap* hi, const CUtensorMap* lo, int col, int row, bool issue, uint16_t mask) {
        const int s = use % NSLOT;
        ptx::mbar_wait(&ctl->empty[s], ((use / NSLOT) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&ctl->full[s], SLOT);
        if (issue) {
          uint8_t* dst = ring + (size_t)s * SLOT;
          if (csize == 1) {
            ptx::tma_load_2d(dst, hi, &ctl->full[s], col, row);
            ptx::tma_load_2d(dst + 16384, lo, &ctl->full[s], col, row);
          } else {
            ptx::tma_load_2d_mc(dst, hi, &ctl->full[s], col, row, mask);
            ptx::tma_load_2d_mc(dst + 16384, lo, &ctl->full[s], col, row, mask);
          }
        }
        ++use;
      };
      for (int i = 0; i <= ntile; ++i) {
        if (i < ntile)
          for (int kb = 0; kb < p.KB; ++kb) {
            load(&qh_map, &ql_map, kb * 64, qtile * 128, csel == (n_q % cc), mask_q);
            ++n_q;
            load(&eh_map, &el_map, kb * 64, (t0 + i) * 128, crank == (n_es % csize), mask_all);
            ++n_es;
          }
        if (i > 0)
          for (int b = 0; b < nbox; ++b) {
            load(&eh_map, &el_map, (box0 + b) * 64, (t0 + i - 1) * 128, qsel == (n_pv % cq), mask_c);
            ++n_pv;
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t ring_a = ptx::smem_u32(ring);
      // descriptor templates: only the 14-bit start-address field changes between MMAs
      const uint64_t DK = udesc(0, 16, 1024), DMN = udesc(0, 16384, 1024);
      uint32_t use = 0;
      auto wait_slot = [&](uint32_t u) { ptx::mbar_wait(&ctl->full[u % NSLOT], (u / NSLOT) & 1); };
      auto release = [&](uint32_t u) {  // slot free in THIS CTA once the MMAs issued so far have read it
        if (csize == 1)
          ptx::umma_commit(&ctl->empty[u % NSLOT]);
        else
          ptx::umma_commit_mc(&ctl->empty[u % NSLOT], mask_all);
      };
      auto pv = [&](int t) {
        ptx::mbar_wait(&ctl->p_full, t & 1);
        ptx::tc_fence_after();
        for (int b = 0; b < nbox; ++b) {
          wait_slot(use);
          ptx::tc_fence_after();
          const uint32_t e_hi = ring_a + (use % NSLOT) * SLOT, e_lo = e_hi + 16384;
          const uint32_t d_o = TM_O + b * 64;
          const uint32_t p_t = TM_S + (t & 1) * 128;
          const uint64_t bh = DMN + (e_hi >> 4), bl = DMN + (e_lo >> 4);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t a_hi = p_t + 32 * (ks >> 1) + 8 * (ks & 1), a_lo = a_hi + 16;
            const uint64_t b_hi = bh + ks * (2048 >> 4), b_lo = bl + ks * (2048 >> 4);
            ptx::umma_bf16_ts(d_o, a_hi, b_hi, idesc_pv, (t > 0 || ks > 0) ? 1u : 0u);
            ptx::umma_bf16_ts(d_o, a_hi, b_lo, idesc_pv, 1u);
            ptx::umma_bf16_ts(d_o, a_lo, b_hi, idesc_pv, 1u);
          }
          release(use);
          ++use;
        }
        ptx::umma_commit(&ctl->pv_done);
      };
      for (int i = 0; i < ntile; ++i) {
        // S(i) overwrites the buffer that held P(i-2): tcgen05.mma executes in issue order, and
        // PV(i-2) (its last reader) was issued before, so no extra barrier is needed here.
        const int sb = i & 1;
        const uint32_t d_s = TM_S + sb * 128;
        for (int kb = 0; kb < p.KB; ++kb) {
          wait_slot(use);
          wait_slot(use + 1);
          ptx::tc_fence_after();
          const uint32_t q_hi = ring_a + (use % NSLOT) * SLOT, q_lo = q_hi + 16384;
          const uint32_t e_hi = ring_a + ((use + 1) % NSLOT) * SLOT, e_lo = e_hi + 16384;
          const uint64_t ah = DK + (q_hi >> 4), al = DK + (q_lo >> 4), bh = DK + (e_hi >> 4), bl = DK + (e_lo >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            ptx::umma_bf16(d_s, ah + kk * 2, bh + kk * 2, idesc_s, (kb > 0 || kk > 0) ? 1u : 0u);
            ptx::umma_bf16(d_s, ah + kk * 2, bl + kk * 2, idesc_s, 1u);
            ptx::umma_bf16(d_s, al + kk * 2, bh + kk * 2, idesc_s, 1u);
          }
          release(use);
          release(use + 1);
          use += 2;
        }
        ptx::umma_commit(&ctl->s_full[sb]);
        if (i > 0) pv(i - 1);
      }
      pv(ntile - 1);
      ptx::umma_commit(&ctl->o_done);
    }
  } else {
    // ------------------------------- softmax / epilogue: one thread per query row -------------------------------
    const int sub = warp & 3;               // TMEM sub-partition this warp may access
    const int row = sub * 32 + lane;        // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    const int g = qtile * 128 + row;
    float m_ref = -INFINITY, l_run = 0.f;
    const int ocols = nbox * 64;
    for (int i = 0; i < ntile; ++i) {
      const int sb = i & 1;
      const int j0 = (t0 + i) * 128;
      ptx::mbar_wait(&ctl->s_full[sb], (i >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t s_addr = TM_S + sb * 128 + lane_off;
      float factor;
      if (p.mode == KP_FLASH_SOFTMAX)
        umma_sm::p_tile<true>(s_addr, j0, p.N, RESCALE_TAU, m_ref, l_run, factor);
      else
        umma_sm::p_tile<false>(s_addr, j0, p.N, RESCALE_TAU, m_ref, l_run, factor);
      if (__any_sync(0xffffffffu, factor != 1.f)) {
        // O holds tiles < i only once PV(i-1) has completed
        ptx::mbar_wait(&ctl->pv_done, (i & 1) ^ 1);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < ocols; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(TM_O + lane_off + c0, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) * factor);
          ptx::tmem_st_32x32(TM_O + lane_off + c0, r);
        }
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ctl->p_full);
    }
    ptx::mbar_wait(&ctl->o_done, 0);
    ptx::tc_fence_after();
    const size_t slot = (size_t)strip * p.G + (g < p.G ? g : 0);
#pragma unroll 1
    for (int c0 = 0; c0 < ocols; c0 += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(TM_O + lane_off + c0, r);
      ptx::tmem_ld_wait();
      if (g < p.G) {
        const int k0 = box0 * 64 + c0;
#pragma unroll
        for (int c = 0; c < 32; c += 4)
          if (k0 + c < p.D)
            *reinterpret_cast<float4*>(p.part_O + slot * p.D + k0 + c) =
                make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]), __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
      }
    }
    if (g < p.G && chunk == 0) {
      p.part_m[slot] = m_ref;
      p.part_l[slot] = l_run;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (csize > 1) ptx::cluster_sync_all();  // no CTA leaves while peers may still write its smem / barriers
  if (warp == 1) ptx::tmem_dealloc(tm, 512);
}

// fp32 [rows, D] -> bf16 hi / lo [rows_pad, Dpad] (zero padded)
__global__ void split_bf16_kernel(const float* __restrict__ src, long long rows, int D, long long rows_pad, int Dpad,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long total = rows_pad * (long long)(Dpad / 2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (Dpad / 2);
    const int c = (int)(i % (Dpad / 2)) * 2;
    float a = 0.f, b = 0.f;
    if (r < rows) {
      if (c < D) a = src[r * D + c];
      if (c + 1 < D) b = src[r * D + c + 1];
    }
    const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
    reinterpret_cast<uint32_t*>(hi)[i] = pack_bf16(ah, bh);
    reinterpret_cast<uint32_t*>(lo)[i] =
        pack_bf16(__float2bfloat16_rn(a - __bfloat162float(ah)), __float2bfloat16_rn(b - __bfloat162float(bh)));
  }
}

int encode_bf16(kp_ctx* ctx, CUtensorMap* map, const void* base, long long rows, int cols) {
  return kp_encode_2d(ctx, map, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, rows, cols, cols, 128, 64, true);
}

}  // namespace

bool kp_flash_umma_usable(kp_ctx* ctx, int G) {
  return !ctx->force_simt && G >= ctx->umma_min_rows && ctx->D <= 512 && ctx->D % 4 == 0;
}

// Fewest strips the tcgen05 fused pass may cut the entity range into for g rows: ceil(tiles / umma_max_tps), relaxed so that the
// strip partials (strips x g x D floats) stay within a quarter of the device memory (tables of many millions of entities).
int kp_umma_min_strips(kp_ctx* ctx, long long g) {
  if (ctx->umma_max_tps <= 0) return 1;
  const long long n_tiles = (ctx->N + 127) / 128;
  long long smin = (n_tiles + ctx->umma_max_tps - 1) / ctx->umma_max_tps;
  if (ctx->total_mem == 0) {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) {
      cudaGetLastError();
      tot = (size_t)64 << 30;
    }
    ctx->total_mem = (long long)tot;
  }
  const long long per_strip = (g > 0 ? g : 1) * (long long)ctx->D * 4;
  long long by_mem = (ctx->total_mem / 4) / per_strip;
  if (by_mem < 8) by_mem = 8;
  if (smin > by_mem) smin = by_mem;
  return (int)(smin < 1 ? 1 : smin);
}

namespace {
struct UPlan {
  int KBs;     // 64-wide k-blocks of the S phase = ceil(D / 64)
  int bpc;     // 64-wide output boxes per dim chunk (<= 4 : 256 TMEM columns)
  int cc;      // dim chunks
  int cq;      // query tiles per cluster
  int Dpad;    // padded row width of the split tables
  int n_qt;    // query tiles, padded to a multiple of cq
  int n_strips, n_tiles, tps;
  bool pair;   // cta_group::2 kernel (kp_flash_umma2.cu): SM pairs over 2 query tiles
  bool quad;   // clusters of two pairs sharing S across the dim chunks (kp_flash_umma4.cu)
  bool sv;     // ... with one pair computing S / softmax and the other contracting (kp_flash_umma_sv.cu)
};
UPlan umma_plan(kp_ctx* ctx, int G) {
  UPlan u;
  u.KBs = (ctx->D + 63) / 64;
  u.cc = u.KBs <= 4 ? 1 : 2;  // dim chunks (clustered together only when cq > 1)
  u.bpc = (u.KBs + u.cc - 1) / u.cc;
  u.bpc = (u.bpc + 1) & ~1;  // whole 128-dim groups (the cta_group::2 pass contracts two boxes per MMA)
  u.Dpad = (u.KBs > u.bpc * u.cc ? u.KBs : u.bpc * u.cc) * 64;
  const int n_qt = (G + 127) / 128;
  u.pair = ctx->umma_2sm != 0 && ctx->umma_cq <= 1 && n_qt >= 2;
  // default: independent CTAs.  Measured on B200 (1M x 512, 18k rows): clusters of 2x2 with TMA
  // multicast run at 159 TFLOP/s vs 219 without -- the per-SM shared-memory fill rate, which
  // multicast does not lower, is the limiter, and the lock-step slot release costs more than it saves.
  int cq = ctx->umma_cq > 0 ? (int)ctx->umma_cq : 1;
  while (cq > 1 && n_qt < cq) cq >>= 1;
  if (u.pair) cq = 2;
  u.cq = cq;
  u.n_qt = ((n_qt + cq - 1) / cq) * cq;
  u.n_tiles = (int)((ctx->N + 127) / 128);
  u.quad = u.pair && u.cc == 2 && ctx->umma_x4 != 0;
  u.sv = u.quad && ctx->umma_x4 >= 2;
  const int sms = u.sv ? kp_flash_umma_sv_sms(ctx) : (u.quad ? kp_flash_umma4_sms(ctx) : ctx->sm_count);
  // Strip count: the entity range is cut into `s` strips so that (query-tile clusters x strips) fills whole waves
  // of the SMs the launch can occupy.  unit = CTAs that must be co-resident, units = clusters per strip.
  const int unit = u.quad ? 4 : (u.pair ? 2 : 1);
  const long long units = ((long long)u.n_qt * u.cc + unit - 1) / unit;
  int s = kp_plan_strips(units, sms / unit, u.n_tiles);
  const int smin = kp_umma_min_strips(ctx, G);  // bound the length of one fp32 accumulation chain in TMEM (kp_internal.h)
  if (smin > s) s = smin;
  u.tps = (u.n_tiles + s - 1) / s;
  if (u.sv) u.tps = (u.tps + 1) & ~1;  // the S pair scores two entity tiles per MMA
  u.n_strips = (u.n_tiles + u.tps - 1) / u.tps;
  return u;
}
}  // namespace

int kp_flash_umma_plan(kp_ctx* ctx, int G, int* n_strips) {
  const UPlan u = umma_plan(ctx, G);
  *n_strips = u.n_strips;
  return u.bpc;
}

namespace {
// L2 norm of every entity row (used by the tcgen05 rank pass for its per-pair error margin)
__global__ void row_norm_kernel(const float* __restrict__ src, long long rows, int D, long long rows_pad, float* __restrict__ out) {
  const long long r = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows_pad) return;
  float a = 0.f;
  if (r < rows)
    for (int k = lane; k < D; k += 32) a = __fmaf_rn(src[r * D + k], src[r * D + k], a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[r] = sqrtf(a);
}
}  // namespace

// Split bf16 entity tables, their TMA maps, the row norms and the walk cursors (once per context).
int kp_umma_tables(kp_ctx* ctx, cudaStream_t st) {
  if (ctx->um.ready) return KP_OK;
  const UPlan u = umma_plan(ctx, 128);
  const int D = ctx->D, Dpad = u.Dpad;
  int rc;
  const long long Npad = ((ctx->N + 127) / 128) * 128;
  void *h = nullptr, *l = nullptr, *nrm = nullptr;
  if (cudaMalloc(&h, (size_t)Npad * Dpad * 2) != cudaSuccess || cudaMalloc(&l, (size_t)Npad * Dpad * 2) != cudaSuccess ||
      cudaMalloc(&nrm, (size_t)Npad * 4) != cudaSuccess)
    KP_FAIL(ctx, KP_ENOMEM, "cannot allocate the split bf16 entity tables (%lld x %d)", Npad, Dpad);
  ctx->owned.push_back(h);
  ctx->owned.push_back(l);
  ctx->owned.push_back(nrm);
  split_bf16_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->ent, ctx->N, D, Npad, Dpad, (__nv_bfloat16*)h, (__nv_bfloat16*)l);
  KP_LAUNCHED(ctx, 1);
  row_norm_kernel<<<(unsigned)((Npad + 7) / 8), 256, 0, st>>>(ctx->ent, ctx->N, D, Npad, (float*)nrm);
  KP_LAUNCHED(ctx, 1);
  if ((rc = encode_bf16(ctx, &ctx->um.eh_map, h, Npad, Dpad)) != KP_OK) return rc;
  if ((rc = encode_bf16(ctx, &ctx->um.el_map, l, Npad, Dpad)) != KP_OK) return rc;
  if ((rc = kp_encode_2d(ctx, &ctx->um.eh64_map, h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Npad, Dpad, Dpad, 64, 64, true)) != KP_OK) return rc;
  if ((rc = kp_encode_2d(ctx, &ctx->um.el64_map, l, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Npad, Dpad, Dpad, 64, 64, true)) != KP_OK) return rc;
  ctx->um.ent_hi = h;
  ctx->um.ent_lo = l;
  ctx->um.enorm = (float*)nrm;
  ctx->um.Dpad = Dpad;
  void* cur = nullptr;
  if (cudaMalloc(&cur, 64 * sizeof(int)) != cudaSuccess) KP_FAIL(ctx, KP_ENOMEM, "cannot allocate the walk cursors");
  ctx->owned.push_back(cur);
  KP_CUDA(ctx, cudaMemsetAsync(cur, 0, 64 * sizeof(int), st));
  ctx->umma_cursor = static_cast<int*>(cur);
  ctx->um.ready = true;
  return KP_OK;
}

// fp32 [G, D] rows -> bf16 hi / lo [Gpad, Dpad] in workspace arena 1 + their TMA maps (box {64, 128}).
int kp_umma_split_rows(kp_ctx* ctx, const float* mat, int G, long long Gpad, CUtensorMap* hi_map, CUtensorMap* lo_map,
                       cudaStream_t st) {
  const int Dpad = ctx->um.Dpad;
  const size_t qbytes = (size_t)Gpad * Dpad * 2;
  int rc;
  if ((rc = kp_ws_reserve(ctx, 2 * qbytes + 2048, 1)) != KP_OK) return rc;
  __nv_bfloat16* qh = reinterpret_cast<__nv_bfloat16*>(ctx->ws_arena[1]);
  __nv_bfloat16* ql = reinterpret_cast<__nv_bfloat16*>(ctx->ws_arena[1] + ((qbytes + 1023) & ~size_t(1023)));
  split_bf16_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(mat, G, ctx->D, Gpad, Dpad, qh, ql);
  KP_LAUNCHED(ctx, 1);
  if ((rc = encode_bf16(ctx, hi_map, qh, Gpad, Dpad)) != KP_OK) return rc;
  if ((rc = encode_bf16(ctx, lo_map, ql, Gpad, Dpad)) != KP_OK) return rc;
  return KP_OK;
}

int kp_flash_umma(kp_ctx* ctx, const float* qmat, int G, int mode, float* part_m, float* part_l, float* part_O,
                  cudaStream_t st) {
  if (G <= 0) return KP_OK;
  const UPlan u = umma_plan(ctx, G);
  const int D = ctx->D;
  int rc;
  if ((rc = kp_umma_tables(ctx, st)) != KP_OK) return rc;
  CUtensorMap qh_map, ql_map;
  if ((rc = kp_umma_split_rows(ctx, qmat, G, (long long)u.n_qt * 128, &qh_map, &ql_map, st)) != KP_OK) return rc;

  if (u.sv)
    return kp_flash_umma_sv_launch(ctx, qh_map, ql_map, G, u.KBs, u.bpc / 2, u.n_qt, u.n_strips, u.tps, mode, part_m, part_l,
                                   part_O, st);
  if (u.quad)
    return kp_flash_umma4_launch(ctx, qh_map, ql_map, G, u.KBs, u.bpc / 2, u.n_qt, u.n_strips, u.tps, mode, part_m, part_l,
                                 part_O, st);
  if (u.pair)
    return kp_flash_umma2_launch(ctx, qh_map, ql_map, G, u.KBs, u.bpc / 2, u.cc, u.n_qt, u.n_strips, u.tps, mode, part_m,
                                 part_l, part_O, st);
  UK p;
  p.G = G;
  p.N = (int)ctx->N;
  p.D = D;
  p.KB = u.KBs;
  p.n_tiles = u.n_tiles;
  p.boxes_per_chunk = u.bpc;
  p.tiles_per_strip = u.tps;
  p.mode = mode;
  p.cq = u.cq;
  p.part_m = part_m;
  p.part_l = part_l;
  p.part_O = part_O;
  KP_SMEM_ONCE(ctx, flash_umma_kernel, U_SMEM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(u.n_strips, u.n_qt, u.cc);
  cfg.blockDim = dim3(UT);
  cfg.dynamicSmemBytes = U_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = u.cq;
  attr[0].val.clusterDim.z = u.cq > 1 ? u.cc : 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  {
    KpTimer timer(ctx, kp_ctx::T_FLASH, st);
    KP_CUDA(ctx, cudaLaunchKernelEx(&cfg, flash_umma_kernel, ctx->um.eh_map, ctx->um.el_map, qh_map, ql_map, p));
  }
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
