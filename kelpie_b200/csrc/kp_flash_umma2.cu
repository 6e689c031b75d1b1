// cta_group::2 version of the fused score -> softmax/sigmoid -> contract pass (see
// kp_flash_umma.cu for the algorithm, the bf16x3 split and the TMEM layout).
//
// Two CTAs on an SM pair (cluster 2x1x1) own two neighbouring query tiles and walk the same entity
// tiles and dim chunk.  Every tcgen05.mma is issued by the even CTA with M = 256 (128 query rows
// in each CTA's TMEM); the B operand of an MMA is split across the pair, so each SM stages only
// HALF of every entity box:
//   S  = Q E^T : N = 128 entities, each CTA loads 64 entity rows of the k-block (16 KB instead of 32)
//   O += P E   : N = 128 dims,     each CTA loads one of the two 64-dim boxes     (2 boxes per tile instead of 4)
// i.e. 448 KB of shared-memory fill per 128x128 tile and CTA instead of 640 KB -- the measured
// limiter of the 1-SM kernel.  TMA loads of both CTAs complete on the even CTA's `full` barriers
// (.cta_group::2 form); tcgen05.commit multicasts slot releases / S-ready / PV-done to both CTAs;
// the odd CTA's softmax threads arrive remotely on the even CTA's P-ready barrier.
#include <cuda_bf16.h>

#include "kp_flash.cuh"
#include "kp_internal.h"
#include "kp_ptx.cuh"
#include "kp_umma_softmax.cuh"

namespace {

constexpr int UT = 192;
constexpr int SLOT = 32768;
constexpr int NSLOT = 6;   // ring slots when the query k-blocks are streamed with every entity tile
constexpr int NBUF = 7;    // 32 KB blocks of shared memory in all: resident query k-blocks (<= 4) + ring when they are kept
constexpr float RESCALE_TAU = 8.0f;

struct UCtl2 {
  uint64_t full[NBUF], empty[NBUF];
  uint64_t s_full[2];
  uint64_t p_full, pv_done, o_done;
  uint64_t q_full;  // resident-query mode: my query tile (all k-blocks, hi / lo) has landed
  uint32_t tmem_base;
  int start, start_local;
};
constexpr size_t U2_SMEM = (size_t)NBUF * SLOT + sizeof(UCtl2) + 1024;
static_assert(U2_SMEM <= 232448, "shared memory budget of one CTA");

struct UK2 {
  int G, N, D, KB, n_tiles, tiles_per_strip, groups_per_chunk, mode;
  int qres;  // keep the query tile (KB <= 4 k-blocks of 32 KB) resident instead of re-streaming it with every entity tile: with a
             // short K the S phase otherwise fills shared memory at 77 B/clk next to 96 B/clk of operand reads (port: 128 B/clk)
  float* part_m;
  float* part_l;
  float* part_O;
  int* cursor;  // optional [n_strips]: rotating start, see kp_flash_umma4.cu
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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UT, 1)
flash_umma2_kernel(const __grid_constant__ CUtensorMap eh_map, const __grid_constant__ CUtensorMap el_map,
                   const __grid_constant__ CUtensorMap eh64_map, const __grid_constant__ CUtensorMap el64_map,
                   const __grid_constant__ CUtensorMap qh_map, const __grid_constant__ CUtensorMap ql_map, const UK2 p) {
  extern __shared__ uint8_t uraw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(uraw) + 1023) & ~uintptr_t(1023));
  UCtl2* ctl = reinterpret_cast<UCtl2*>(sm + (size_t)NBUF * SLOT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int strip = blockIdx.y, qtile = blockIdx.x, chunk = blockIdx.z;  // the SM pair spans x
  const int t0 = strip * p.tiles_per_strip;
  const int t1 = min(t0 + p.tiles_per_strip, p.n_tiles);
  const int ntile = t1 - t0;
  if (ntile <= 0) return;  // uniform over the pair
  const uint32_t crank = ptx::cluster_ctarank();  // 0 = leader (issues every MMA)
  const bool leader = crank == 0;
  const int ngroup = p.groups_per_chunk;           // 128-dim groups of this chunk (<= 2)
  const int box0 = chunk * ngroup * 2;
  // Work that only touches zero padding is skipped: S needs ceil(D / 16) k-steps, and when the pass is a single
  // chunk its last dim group is contracted with N = the dims that are left, rounded up to 32 (each SM then
  // stages n_last / 2 dims, starting at the group's base + rank * n_last / 2).
  const int ksteps = (p.D + 15) / 16;
  const bool qres = p.qres != 0;
  uint8_t* qbuf = sm;                                           // resident mode: k-block kb of my query tile at kb * SLOT (hi | lo)
  uint8_t* ring = qres ? sm + (size_t)p.KB * SLOT : sm;
  const int nring = qres ? NBUF - p.KB : NSLOT;
  const int kb2n = (p.KB + 1) >> 1;                             // resident mode: entity k-blocks travel two per ring slot
  const int n_last = (gridDim.z == 1) ? min(128, ((p.D - 128 * (ngroup - 1) + 31) / 32) * 32) : 128;

  if (tid == 0) {
    for (int s = 0; s < NBUF; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    ptx::mbar_init(&ctl->q_full, 1);
    for (int b = 0; b < 2; ++b) ptx::mbar_init(&ctl->s_full[b], 1);
    ptx::mbar_init(&ctl->p_full, 256);  // 128 softmax threads of each CTA (used in the leader only)
    ptx::mbar_init(&ctl->pv_done, 1);
    ptx::mbar_init(&ctl->o_done, 1);
    if (crank == 0) ctl->start = p.cursor ? (int)((unsigned)*(volatile int*)&p.cursor[blockIdx.y] % (unsigned)ntile) : 0;
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc2(&ctl->tmem_base, 512);
    ptx::tmem_relinquish2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  if (tid == 0) ctl->start_local = (int)ptx::ld_cluster_u32(ptx::mapa_u32(ptx::smem_u32(&ctl->start), 0));
  __syncthreads();
  const int p0 = ctl->start_local;
  auto tile_of = [&](int t) {  // t-th tile of my walk -> tile index in the table (rotating start)
    const int x = p0 + t;
    return t0 + (x >= ntile ? x - ntile : x);
  };
  const uint32_t tm = ctl->tmem_base;
  const uint32_t TM_O = tm, TM_S = tm + 256;

  if (warp == 0) {
    // ------------------------------- TMA producer (both CTAs) -------------------------------
    if (lane == 0) {
      ptx::prefetch_tmap(&eh_map);
      ptx::prefetch_tmap(&el_map);
      ptx::prefetch_tmap(&eh64_map);
      ptx::prefetch_tmap(&el64_map);
      ptx::prefetch_tmap(&qh_map);
      ptx::prefetch_tmap(&ql_map);
      uint32_t use = 0;
      // bytes_pair = bytes the two CTAs together deliver for this slot use (armed on the leader's barrier)
      auto load = [&](const CUtensorMap* hi, const CUtensorMap* lo, int col, int row, uint32_t lo_off, uint32_t bytes_pair) {
        const int s = use % nring;
        ptx::mbar_wait(&ctl->empty[s], ((use / nring) & 1) ^ 1);
        if (leader) ptx::mbar_arrive_expect_tx(&ctl->full[s], bytes_pair);
        const uint32_t bar = ptx::mapa_u32(ptx::smem_u32(&ctl->full[s]), 0);
        uint8_t* dst = ring + (size_t)s * SLOT;
        ptx::tma_load_2d_pair(dst, hi, bar, col, row);
        ptx::tma_load_2d_pair(dst + lo_off, lo, bar, col, row);
        ++use;
      };
      if (qres) {  // my query tile, once
        if (leader) ptx::mbar_arrive_expect_tx(&ctl->q_full, (uint32_t)(2 * p.KB) * SLOT);
        const uint32_t qbar = ptx::mapa_u32(ptx::smem_u32(&ctl->q_full), 0);
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::tma_load_2d_pair(qbuf + (size_t)kb * SLOT, &qh_map, qbar, kb * 64, qtile * 128);
          ptx::tma_load_2d_pair(qbuf + (size_t)kb * SLOT + 16384, &ql_map, qbar, kb * 64, qtile * 128);
        }
      }
      for (int i = 0; i <= ntile; ++i) {
        if (i < ntile) {
          if (!qres) {
            for (int kb = 0; kb < p.KB; ++kb) {
              load(&qh_map, &ql_map, kb * 64, qtile * 128, 16384, 2 * 32768);                         // my query tile
              load(&eh64_map, &el64_map, kb * 64, tile_of(i) * 128 + (int)crank * 64, 8192, 2 * 16384);  // my half of the entities
            }
          } else {
            for (int k2 = 0; k2 < kb2n; ++k2) {  // my half of the entities, two k-blocks per slot: [hi | lo | hi | lo] x 8 KB
              const int nk = min(2, p.KB - 2 * k2);
              const int s = use % nring;
              ptx::mbar_wait(&ctl->empty[s], ((use / nring) & 1) ^ 1);
              if (leader) ptx::mbar_arrive_expect_tx(&ctl->full[s], (uint32_t)(2 * nk) * 16384);
              const uint32_t bar = ptx::mapa_u32(ptx::smem_u32(&ctl->full[s]), 0);
              uint8_t* dst = ring + (size_t)s * SLOT;
              for (int h = 0; h < nk; ++h) {
                ptx::tma_load_2d_pair(dst + h * 16384, &eh64_map, bar, (2 * k2 + h) * 64, tile_of(i) * 128 + (int)crank * 64);
                ptx::tma_load_2d_pair(dst + h * 16384 + 8192, &el64_map, bar, (2 * k2 + h) * 64, tile_of(i) * 128 + (int)crank * 64);
              }
              ++use;
            }
          }
        }
        if (i > 0)
          for (int g = 0; g < ngroup; ++g)  // my half of the group's dims (64, or n_last / 2 of a trimmed last group)
            load(&eh_map, &el_map, (box0 + 2 * g) * 64 + (int)crank * (g == ngroup - 1 ? n_last / 2 : 64), tile_of(i - 1) * 128, 16384,
                 2 * 32768);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA only) -------------------------------
    if (lane == 0 && leader) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t idesc_pv_last = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | (((uint32_t)n_last >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t ring_a = ptx::smem_u32(ring);
      // descriptor templates: only the 14-bit start-address field changes between MMAs
      const uint64_t DK = udesc(0, 16, 1024), DMN = udesc(0, 16384, 1024);
      uint32_t use = 0;
      auto wait_slot = [&](uint32_t u) { ptx::mbar_wait(&ctl->full[u % nring], (u / nring) & 1); };
      auto release = [&](uint32_t u) { ptx::umma2_commit_mc(&ctl->empty[u % nring], 3); };
      const uint32_t q_a = ptx::smem_u32(qbuf);
      if (qres) {
        ptx::mbar_wait(&ctl->q_full, 0);
        ptx::tc_fence_after();
      }
      auto pv = [&](int t) {
        ptx::mbar_wait_cluster(&ctl->p_full, t & 1);
        ptx::tc_fence_after();
        for (int g = 0; g < ngroup; ++g) {
          wait_slot(use);
          ptx::tc_fence_after();
          const uint32_t e_hi = ring_a + (use % nring) * SLOT, e_lo = e_hi + 16384;
          const uint32_t d_o = TM_O + g * 128;
          const uint32_t p_t = TM_S + (t & 1) * 128;
          const uint32_t idesc_g = (g == ngroup - 1) ? idesc_pv_last : idesc_pv;
          const uint64_t bh = DMN + (e_hi >> 4), bl = DMN + (e_lo >> 4);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t a_hi = p_t + 32 * (ks >> 1) + 8 * (ks & 1), a_lo = a_hi + 16;
            const uint64_t b_hi = bh + ks * (2048 >> 4), b_lo = bl + ks * (2048 >> 4);
            ptx::umma2_bf16_ts(d_o, a_hi, b_hi, idesc_g, (t > 0 || ks > 0) ? 1u : 0u);
            ptx::umma2_bf16_ts(d_o, a_hi, b_lo, idesc_g, 1u);
            ptx::umma2_bf16_ts(d_o, a_lo, b_hi, idesc_g, 1u);
          }
          release(use);
          ++use;
        }
        ptx::umma2_commit_mc(&ctl->pv_done, 3);
      };
      for (int i = 0; i < ntile; ++i) {
        const int sb = i & 1;
        const uint32_t d_s = TM_S + sb * 128;
        if (p.cursor && chunk == 0) *(volatile int*)&p.cursor[blockIdx.y] = tile_of(i) - t0;
        if (!qres) {
          for (int kb = 0; kb < p.KB; ++kb) {
            wait_slot(use);
            wait_slot(use + 1);
            ptx::tc_fence_after();
            const uint32_t q_hi = ring_a + (use % nring) * SLOT, q_lo = q_hi + 16384;
            const uint32_t e_hi = ring_a + ((use + 1) % nring) * SLOT, e_lo = e_hi + 8192;
            const uint64_t ah = DK + (q_hi >> 4), al = DK + (q_lo >> 4), bh = DK + (e_hi >> 4), bl = DK + (e_lo >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (kb * 4 + kk >= ksteps) break;  // only zero padding beyond D
              ptx::umma2_bf16(d_s, ah + kk * 2, bh + kk * 2, idesc_s, (kb > 0 || kk > 0) ? 1u : 0u);
              ptx::umma2_bf16(d_s, ah + kk * 2, bl + kk * 2, idesc_s, 1u);
              ptx::umma2_bf16(d_s, al + kk * 2, bh + kk * 2, idesc_s, 1u);
            }
            release(use);
            release(use + 1);
            use += 2;
          }
        } else {
          for (int k2 = 0; k2 < kb2n; ++k2) {
            wait_slot(use);
            ptx::tc_fence_after();
            const int nk = min(2, p.KB - 2 * k2);
            for (int h = 0; h < nk; ++h) {
              const int kb = 2 * k2 + h;
              const uint32_t q_hi = q_a + (uint32_t)kb * SLOT, q_lo = q_hi + 16384;
              const uint32_t e_hi = ring_a + (use % nring) * SLOT + (uint32_t)h * 16384u, e_lo = e_hi + 8192;
              const uint64_t ah = DK + (q_hi >> 4), al = DK + (q_lo >> 4), bh = DK + (e_hi >> 4), bl = DK + (e_lo >> 4);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (kb * 4 + kk >= ksteps) break;  // only zero padding beyond D
                ptx::umma2_bf16(d_s, ah + kk * 2, bh + kk * 2, idesc_s, (kb > 0 || kk > 0) ? 1u : 0u);
                ptx::umma2_bf16(d_s, ah + kk * 2, bl + kk * 2, idesc_s, 1u);
                ptx::umma2_bf16(d_s, al + kk * 2, bh + kk * 2, idesc_s, 1u);
              }
            }
            release(use);
            ++use;
          }
        }
        ptx::umma2_commit_mc(&ctl->s_full[sb], 3);
        if (i > 0) pv(i - 1);
      }
      pv(ntile - 1);
      ptx::umma2_commit_mc(&ctl->o_done, 3);
    }
  } else {
    // ------------------------------- softmax / epilogue (both CTAs, own rows) -------------------------------
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    const int g = qtile * 128 + row;
    const uint32_t p_full_leader = ptx::mapa_u32(ptx::smem_u32(&ctl->p_full), 0);
    float m_ref = -INFINITY, l_run = 0.f;
    const int ocols = ngroup * 128;
    for (int i = 0; i < ntile; ++i) {
      const int sb = i & 1;
      const int j0 = tile_of(i) * 128;
      ptx::mbar_wait(&ctl->s_full[sb], (i >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t s_addr = TM_S + sb * 128 + lane_off;
      float factor;
      if (p.mode == KP_FLASH_SOFTMAX)
        umma_sm::p_tile<true>(s_addr, j0, p.N, RESCALE_TAU, m_ref, l_run, factor);
      else
        umma_sm::p_tile<false>(s_addr, j0, p.N, RESCALE_TAU, m_ref, l_run, factor);
      if (__any_sync(0xffffffffu, factor != 1.f)) {
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
      ptx::mbar_arrive_cluster(p_full_leader);
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
  ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc2(tm, 512);
}

}  // namespace

// Launch for a plan prepared by kp_flash_umma (split tables / queries, maps).
int kp_flash_umma2_launch(kp_ctx* ctx, const CUtensorMap& qh_map, const CUtensorMap& ql_map, int G, int KBs,
                          int groups_per_chunk, int n_chunks, int n_qt, int n_strips, int tps, int mode, float* part_m,
                          float* part_l, float* part_O, cudaStream_t st) {
  UK2 p;
  p.G = G;
  p.N = (int)ctx->N;
  p.D = ctx->D;
  p.KB = KBs;
  p.n_tiles = (int)((ctx->N + 127) / 128);
  p.tiles_per_strip = tps;
  p.groups_per_chunk = groups_per_chunk;
  p.mode = mode;
  p.qres = (ctx->umma_qres != 0 && KBs <= 4) ? 1 : 0;
  p.part_m = part_m;
  p.part_l = part_l;
  p.part_O = part_O;
  // two dim chunks = two clusters per (query tiles, strip) whose lazy reference maxima must agree: they have to
  // visit the tiles in the same order, so the rotating start is only used by the single-chunk pass
  p.cursor = (ctx->umma_rotate && n_chunks == 1) ? ctx->umma_cursor : nullptr;
  KP_SMEM_ONCE(ctx, flash_umma2_kernel, U2_SMEM);
  if (n_qt % 2 != 0) KP_FAIL(ctx, KP_EINVAL, "pair kernel needs an even number of query tiles (%d)", n_qt);
  {
    int max_clusters = -1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_qt, n_strips, n_chunks);
    cfg.blockDim = dim3(UT);
    cfg.dynamicSmemBytes = U2_SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static bool reported = false;
    if (!reported) {
      cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, flash_umma2_kernel, &cfg);
      if (e != cudaSuccess || max_clusters <= 0)
        fprintf(stderr, "kelpie_b200: cudaOccupancyMaxActiveClusters -> %d (%s)\n", max_clusters, cudaGetErrorString(e));
      cudaGetLastError();
      reported = true;
    }
    KpTimer timer(ctx, kp_ctx::T_FLASH, st);
    flash_umma2_kernel<<<dim3(n_qt, n_strips, n_chunks), UT, U2_SMEM, st>>>(ctx->um.eh_map, ctx->um.el_map, ctx->um.eh64_map,
                                                                          ctx->um.el64_map, qh_map, ql_map, p);
  }
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
