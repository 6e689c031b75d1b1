// Fused score -> softmax/sigmoid -> contract pass for rows wider than 256 floats, SPECIALISED SMs.
//
// Cluster of 4 CTAs = two cta_group::2 pairs that own the SAME 256 query rows:
//   pair S (cluster ranks 0,1): S = Q E^T for TWO entity tiles per MMA (M256 N256) + the softmax
//   pair V (cluster ranks 2,3): O += P E over ALL 512 output dims (the whole TMEM of both SMs)
// Every probability tile P (bf16 hi / lo) travels once, from the softmax thread of a row on an S SM
// into the shared memory of the V SM that owns the same rows (st.shared::cluster, 64 KB per tile), in
// the K-major no-swizzle layout tcgen05.mma reads as its A operand (SS form).
//
// Why (round-1 profile of kp_flash_umma4.cu, where both pairs alternate S and PV phases): the S phase
// with N = 128 reads 96 B/clk of MMA operands from shared memory while TMA writes 62 B/clk into it
// (the query k-blocks are re-streamed for every 128-entity tile) -- more than the 128 B/clk port, and
// the probabilities cross TMEM three times (S, P written back, foreign P copied in).  Here
//   * an S SM streams the query k-blocks once per 256 entities: 64 B/clk of operand reads +
//     43 B/clk of TMA fill, S is double-buffered in all 512 TMEM columns and never written back;
//   * a V SM reads P from shared memory and E as an N = 256 MN-major operand: 64 B/clk + 21 B/clk of
//     TMA fill + 10 B/clk of incoming P; its TMEM holds only O (rescaled lazily, rarely);
//   * no pair waits for the other's softmax chain: the only cross-pair dependency is the P ring.
// MMA work per SM is unchanged (an S SM: 1 S per tile, a V SM: 1 PV over 512 dims per tile).
//
// Barriers (all in the CTA they are waited in):
//   S SM: full/empty[6] TMA ring | s_full[2] (MMA commit) | s_free[2] (256 softmax threads, leader) |
//         pin_empty[2] (128 remote arrivals of the V SM's row threads: P buffer b may be overwritten)
//   V SM: full/empty[3] | pin_full[2] (128 remote arrivals: P tile delivered) | p_ready[2] (256 row
//         threads, leader: O rescaled, operands visible to the async proxy) | pv_done[2] | o_done
#include <cuda_bf16.h>

#include "kp_flash.cuh"
#include "kp_internal.h"
#include "kp_ptx.cuh"
#include "kp_umma_softmax.cuh"

namespace {

constexpr int UT = 192;
constexpr int SLOT = 32768;
constexpr int NSLOT_S = 6;
constexpr int NSLOT_V = 3;
constexpr int PBUF = 65536;  // one P tile of 128 rows x 128 entities: hi 32 KB | lo 32 KB, chunk c (8 entities) of row r at c*2048 + r*16
constexpr int PHDR = 512;    // rescale factor of the tile per row
constexpr int V_DATA = NSLOT_V * SLOT + 2 * (PBUF + PHDR);
constexpr int S_DATA = NSLOT_S * SLOT;
constexpr int DATA_BYTES = V_DATA > S_DATA ? V_DATA : S_DATA;
constexpr float RESCALE_TAU = 8.0f;

struct SVCtl {
  uint64_t full[NSLOT_S], empty[NSLOT_S];
  uint64_t s_full[2], s_free[2], pin_empty[2];
  uint64_t pin_full[2], p_ready[2], pv_done[2];
  uint64_t o_done;
  uint32_t tmem_base;
  int start, start_local;
};
constexpr size_t SV_SMEM = (size_t)DATA_BYTES + sizeof(SVCtl) + 1024;
static_assert(SV_SMEM <= 232448, "shared memory budget of one CTA");

struct SVK {
  int G, N, D, KB, n_tiles, tiles_per_strip, ngroup, mode, dbg;
  float* part_m;
  float* part_l;
  float* part_O;
  int* cursor;               // optional [n_strips]: tile pair the running clusters of a strip are at (rotating start)
  unsigned long long* prof;  // optional [16]: MMA-thread wait cycles (S: slot, s_free, total; V: slot, p_ready, total)
};

// shared-memory matrix descriptor: layout 2 = SWIZZLE_128B, 0 = no swizzle (interleaved 8x16B core matrices)
__device__ __forceinline__ uint64_t sdesc(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// One 128-entity tile of this thread's row: logits r[128] (already in registers) -> probabilities, split into
// bf16 hi / lo and stored into the V SM's P buffer (pdst = cluster address of this row's first hi chunk).
template <bool SOFTMAX>
__device__ __forceinline__ void p_tile_ship(uint32_t (&r)[128], float& m_ref, float& l_run, float& factor, uint32_t pdst,
                                            uint64_t* pin_empty, bool wait_empty, uint32_t parity, bool ship) {
  using namespace umma_sm;
  factor = 1.f;
  float mneg = 0.f;
  if (SOFTMAX) {
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int c = 0; c < 128; c += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) mx[u] = fmaxf(mx[u], __uint_as_float(r[c + u]));
    }
    const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
    if (m_ref == -INFINITY) {
      m_ref = m;
    } else if (m > m_ref + RESCALE_TAU) {
      factor = ex2((m_ref - m) * LOG2E);
      m_ref = m;
    }
    mneg = (m_ref == -INFINITY) ? 0.f : m_ref * LOG2E;
  }
  if (wait_empty) ptx::mbar_wait_cluster(pin_empty, parity);  // the V SM has contracted the tile that was in this buffer
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t w[32];
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      float p0, p1;
      if (SOFTMAX) {
        p0 = ex2(__fmaf_rn(__uint_as_float(r[32 * q + c]), LOG2E, -mneg));
        p1 = ex2(__fmaf_rn(__uint_as_float(r[32 * q + c + 1]), LOG2E, -mneg));
      } else {
        p0 = rcp(1.f + ex2(-LOG2E * __uint_as_float(r[32 * q + c])));
        p1 = rcp(1.f + ex2(-LOG2E * __uint_as_float(r[32 * q + c + 1])));
      }
      sum0 += p0;
      sum1 += p1;
      const uint32_t hi = bf16x2(p0, p1);
      w[c >> 1] = hi;
      w[16 + (c >> 1)] = bf16x2(p0 - __uint_as_float(hi << 16), p1 - __uint_as_float(hi & 0xffff0000u));
    }
    if (ship) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const uint32_t a = pdst + (uint32_t)(4 * q + v) * 2048u;
        st_cluster_v4(a, w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
        st_cluster_v4(a + 32768u, w[16 + 4 * v], w[16 + 4 * v + 1], w[16 + 4 * v + 2], w[16 + 4 * v + 3]);
      }
    } else {  // timing experiment (option sv_dbg = 64): no DSMEM traffic, results invalid
      uint32_t x = 0;
#pragma unroll
      for (int v = 0; v < 32; ++v) x ^= w[v];
      if (x == 0x12345u) sum0 += 1.f;
    }
  }
  l_run = l_run * factor + (sum0 + sum1);
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(UT, 1)
flash_umma_sv_kernel(const __grid_constant__ CUtensorMap eh_map, const __grid_constant__ CUtensorMap el_map,
                     const __grid_constant__ CUtensorMap eh64_map, const __grid_constant__ CUtensorMap el64_map,
                     const __grid_constant__ CUtensorMap qh_map, const __grid_constant__ CUtensorMap ql_map, const SVK p) {
  extern __shared__ uint8_t uraw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(uraw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = sm;
  uint8_t* pbuf = sm + (size_t)NSLOT_V * SLOT;       // V role: two P tiles
  uint8_t* phdr = pbuf + 2 * (size_t)PBUF;           // V role: two x 128 factors
  SVCtl* ctl = reinterpret_cast<SVCtl*>(sm + DATA_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  const uint32_t half = crank & 1;    // which 128 of the cluster's 256 query rows / which half of every B operand
  const bool role_v = (crank >> 1) != 0;
  const uint32_t lead = crank & ~1u;  // MMA-issuing CTA of my pair
  const uint32_t peer = crank ^ 2u;   // CTA of the other pair that owns the same query rows
  const bool leader = half == 0;
  const int strip = blockIdx.y, qtile = (int)(blockIdx.x >> 2) * 2 + (int)half;
  const int t0 = strip * p.tiles_per_strip;
  const int t1 = min(t0 + p.tiles_per_strip, p.n_tiles);
  const int ntile = t1 - t0;
  if (ntile <= 0) return;  // uniform over the cluster
  const int npair = (ntile + 1) >> 1;
  const int nlim = min(p.N, t1 * 128);  // entities below this id belong to this strip's walk
  const int ngroup = p.ngroup;
  const uint16_t pair_mask = (uint16_t)(3u << lead);
  const int nslot = role_v ? NSLOT_V : NSLOT_S;

  if (tid == 0) {
    for (int s = 0; s < NSLOT_S; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&ctl->s_full[b], 1);
      ptx::mbar_init(&ctl->s_free[b], 256);
      ptx::mbar_init(&ctl->pin_empty[b], 128);
      ptx::mbar_init(&ctl->pin_full[b], 128);
      ptx::mbar_init(&ctl->p_ready[b], 256);
      ptx::mbar_init(&ctl->pv_done[b], 1);
    }
    ptx::mbar_init(&ctl->o_done, 1);
    // rotating start (see kp_flash_umma4.cu): clusters that begin mid-wave join the others where they are
    if (crank == 0) ctl->start = p.cursor ? (int)((unsigned)*(volatile int*)&p.cursor[blockIdx.y] % (unsigned)npair) : 0;
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
  auto pair_of = [&](int k) {  // k-th tile pair of my walk -> pair index inside the strip
    const int x = p0 + k;
    return x >= npair ? x - npair : x;
  };
  const uint32_t tm = ctl->tmem_base;

  if (warp == 0) {
    // ------------------------------- TMA producer (every CTA) -------------------------------
    if (lane == 0) {
      ptx::prefetch_tmap(&eh_map);
      ptx::prefetch_tmap(&el_map);
      ptx::prefetch_tmap(&eh64_map);
      ptx::prefetch_tmap(&el64_map);
      ptx::prefetch_tmap(&qh_map);
      ptx::prefetch_tmap(&ql_map);
      uint32_t use = 0;
      auto acquire = [&]() -> uint8_t* {
        const int s = use % nslot;
        ptx::mbar_wait(&ctl->empty[s], ((use / nslot) & 1) ^ 1);
        if (leader) ptx::mbar_arrive_expect_tx(&ctl->full[s], 2 * SLOT);
        return ring + (size_t)s * SLOT;
      };
      if (!role_v) {
        for (int k = 0; k < npair; ++k) {
          const int erow = (t0 + 2 * pair_of(k)) * 128 + (int)half * 128;
          for (int kb = 0; kb < p.KB; ++kb) {
            {
              uint8_t* dst = acquire();
              const uint32_t bar = ptx::mapa_u32(ptx::smem_u32(&ctl->full[use % nslot]), lead);
              ptx::tma_load_2d_pair(dst, &qh_map, bar, kb * 64, qtile * 128);
              ptx::tma_load_2d_pair(dst + 16384, &ql_map, bar, kb * 64, qtile * 128);
              ++use;
            }
            {
              uint8_t* dst = acquire();
              const uint32_t bar = ptx::mapa_u32(ptx::smem_u32(&ctl->full[use % nslot]), lead);
              ptx::tma_load_2d_pair(dst, &eh_map, bar, kb * 64, erow);
              ptx::tma_load_2d_pair(dst + 16384, &el_map, bar, kb * 64, erow);
              ++use;
            }
          }
        }
      } else {
        for (int t = 0; t < 2 * npair; ++t) {
          const int tile = t0 + 2 * pair_of(t >> 1) + (t & 1);
          for (int g = 0; g < ngroup; ++g)
            for (int eh = 0; eh < 2; ++eh) {
              uint8_t* dst = acquire();
              const uint32_t bar = ptx::mapa_u32(ptx::smem_u32(&ctl->full[use % nslot]), lead);
              const int d0 = g * 256 + (int)half * 128, r0 = tile * 128 + eh * 64;
              ptx::tma_load_2d_pair(dst, &eh64_map, bar, d0, r0);
              ptx::tma_load_2d_pair(dst + 8192, &eh64_map, bar, d0 + 64, r0);
              ptx::tma_load_2d_pair(dst + 16384, &el64_map, bar, d0, r0);
              ptx::tma_load_2d_pair(dst + 24576, &el64_map, bar, d0 + 64, r0);
              ++use;
            }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA of each pair) -------------------------------
    if (lane == 0 && leader) {
      // shared-window addresses carry the CTA's cluster rank above bit 24 (rank 2 -> 0x2000000); unmasked, that bit lands
      // in the descriptor's leading-offset field (harmless only where the field is unused: swizzled K-major, single MN atom)
      const uint32_t ring_a = ptx::smem_u32(ring) & 0x3ffffu;
      uint32_t use = 0;
      long long w_slot = 0, w_dep = 0;
      const long long t_begin = clock64();
      auto wait_slot = [&](uint32_t u) {
        if (ptx::mbar_try_wait(&ctl->full[u % nslot], (u / nslot) & 1)) return;
        const long long a = clock64();
        ptx::mbar_wait(&ctl->full[u % nslot], (u / nslot) & 1);
        w_slot += clock64() - a;
      };
      auto release = [&](uint32_t u) { ptx::umma2_commit_mc(&ctl->empty[u % nslot], pair_mask); };
      if (!role_v) {
        // S(k) = Q E^T for the two entity tiles of pair k: M256 N256, K-major A / B, SWIZZLE_128B
        const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
        const uint64_t DK = sdesc(16, 1024, 2);
        const int ksteps = (p.D + 15) / 16;
        for (int k = 0; k < npair; ++k) {
          const int j = k & 1;
          if (p.cursor && crank == 0) *(volatile int*)&p.cursor[blockIdx.y] = pair_of(k);
          if (k >= 2) {
            const long long a = clock64();
            ptx::mbar_wait_cluster(&ctl->s_free[j], ((k >> 1) - 1) & 1);
            w_dep += clock64() - a;
            ptx::tc_fence_after();
          }
          const uint32_t d_s = tm + j * 256;
          for (int kb = 0; kb < p.KB; ++kb) {
            wait_slot(use);
            wait_slot(use + 1);
            ptx::tc_fence_after();
            const uint32_t q_hi = ring_a + (use % nslot) * SLOT, q_lo = q_hi + 16384;
            const uint32_t e_hi = ring_a + ((use + 1) % nslot) * SLOT, e_lo = e_hi + 16384;
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
          ptx::umma2_commit_mc(&ctl->s_full[j], pair_mask);
        }
      } else {
        // O[:, g*256 .. +256) += P(t) E(t): M256 N256, A = P K-major without swizzle (written by the S SM's softmax
        // threads), B = E MN-major SWIZZLE_128B (two 64-dim boxes per SM, LBO = 8 KB apart)
        const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
        const uint64_t DA = sdesc(2048, 128, 0), DB = sdesc(8192, 1024, 2);
        const uint32_t pb = ptx::smem_u32(pbuf) & 0x3ffffu;
        for (int t = 0; t < 2 * npair; ++t) {
          const int b = t & 1;
          {
            const long long a = clock64();
            ptx::mbar_wait_cluster(&ctl->p_ready[b], (t >> 1) & 1);
            w_dep += clock64() - a;
          }
          ptx::tc_fence_after();
          for (int g = 0; g < ngroup; ++g)
            for (int eh = 0; eh < 2; ++eh) {
              wait_slot(use);
              ptx::tc_fence_after();
              const uint32_t e_hi = ring_a + (use % nslot) * SLOT, e_lo = e_hi + 16384;
              const uint32_t a_hi = pb + (uint32_t)b * PBUF + (uint32_t)eh * 8u * 2048u, a_lo = a_hi + 32768u;
              const uint32_t d_o = tm + g * 256;
              const uint64_t bh = DB + (e_hi >> 4), bl = DB + (e_lo >> 4), ah = DA + (a_hi >> 4), al = DA + (a_lo >> 4);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t a_h = ah + ks * (4096 >> 4), a_l = al + ks * (4096 >> 4);
                const uint64_t b_h = bh + ks * (2048 >> 4), b_l = bl + ks * (2048 >> 4);
                ptx::umma2_bf16(d_o, a_h, b_h, idesc_pv, (t > 0 || eh > 0 || ks > 0) ? 1u : 0u);
                ptx::umma2_bf16(d_o, a_h, b_l, idesc_pv, 1u);
                ptx::umma2_bf16(d_o, a_l, b_h, idesc_pv, 1u);
              }
              release(use);
              ++use;
            }
          ptx::umma2_commit_mc(&ctl->pv_done[b], pair_mask);
        }
        ptx::umma2_commit_mc(&ctl->o_done, pair_mask);
      }
      if (p.prof) {
        const int o = role_v ? 3 : 0;
        atomicAdd(p.prof + o + 0, (unsigned long long)w_slot);
        atomicAdd(p.prof + o + 1, (unsigned long long)w_dep);
        atomicAdd(p.prof + o + 2, (unsigned long long)(clock64() - t_begin));
      }
    }
  } else {
    // ------------------------------- row threads (every CTA, one thread per query row) -------------------------------
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    const int g = qtile * 128 + row;
    if (!role_v) {
      // softmax of both tiles of a pair, P shipped to the V SM that owns the same rows
      const uint32_t s_free_lead = ptx::mapa_u32(ptx::smem_u32(&ctl->s_free[0]), lead);
      const uint32_t pdst0 = ptx::mapa_u32(ptx::smem_u32(pbuf) + (uint32_t)row * 16u, peer);
      const uint32_t hdr0 = ptx::mapa_u32(ptx::smem_u32(phdr) + (uint32_t)row * 4u, peer);
      const uint32_t pin_full_peer = ptx::mapa_u32(ptx::smem_u32(&ctl->pin_full[0]), peer);
      float m_ref = -INFINITY, l_run = 0.f;
      for (int k = 0; k < npair; ++k) {
        const int j = k & 1;
        ptx::mbar_wait(&ctl->s_full[j], (k >> 1) & 1);
        ptx::tc_fence_after();
        const int jbase = (t0 + 2 * pair_of(k)) * 128;
#pragma unroll 1
        for (int u = 0; u < 2; ++u) {
          uint32_t r[128];
          const uint32_t s_addr = tm + lane_off + j * 256 + u * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) ptx::tmem_ld_32x32(s_addr + 32 * q, reinterpret_cast<uint32_t(&)[32]>(r[32 * q]));
          ptx::tmem_ld_wait();
          if (u == 1) {  // both tiles of S(k) are in registers: the buffer may take S(k+2)
            ptx::tc_fence_before();
            ptx::mbar_arrive_cluster(s_free_lead + (uint32_t)j * 8u);
          }
          const int j0 = jbase + u * 128;
          if (j0 + 128 > nlim) {
#pragma unroll
            for (int c = 0; c < 128; ++c)
              if (j0 + c >= nlim) r[c] = 0xff800000u;
          }
          float factor;
          const uint32_t pdst = pdst0 + (uint32_t)u * PBUF;
          if (p.mode == KP_FLASH_SOFTMAX)
            p_tile_ship<true>(r, m_ref, l_run, factor, pdst, &ctl->pin_empty[u], k > 0, (k - 1) & 1, !(p.dbg & 64));
          else
            p_tile_ship<false>(r, m_ref, l_run, factor, pdst, &ctl->pin_empty[u], k > 0, (k - 1) & 1, !(p.dbg & 64));
          st_cluster_f32(hdr0 + (uint32_t)u * PHDR, factor);
          fence_proxy_async_all();                                     // my stores precede the V SM's tcgen05.mma reads
          ptx::mbar_arrive_cluster(pin_full_peer + (uint32_t)u * 8u);  // release: my row of tile 2k+u is delivered
        }
      }
      if (g < p.G) {
        const size_t slot = (size_t)strip * p.G + g;
        p.part_m[slot] = m_ref;
        p.part_l[slot] = l_run;
      }
    } else {
      const uint32_t p_ready_lead = ptx::mapa_u32(ptx::smem_u32(&ctl->p_ready[0]), lead);
      const uint32_t pin_empty_peer = ptx::mapa_u32(ptx::smem_u32(&ctl->pin_empty[0]), peer);
      const uint32_t hdr_local = ptx::smem_u32(phdr) + (uint32_t)row * 4u;
      const int ocols = ngroup * 256;
      for (int t = 0; t < 2 * npair; ++t) {
        const int b = t & 1;
        ptx::mbar_wait_cluster(&ctl->pin_full[b], (t >> 1) & 1);
        float factor;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(factor) : "r"(hdr_local + (uint32_t)b * PHDR) : "memory");
        if (__any_sync(0xffffffffu, factor != 1.f)) {
          // O holds tiles < t only once PV(t-1) has completed (t >= 1 here: the first tile never rescales)
          ptx::mbar_wait(&ctl->pv_done[b ^ 1], ((t - 1) >> 1) & 1);
          ptx::tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < ocols; c0 += 32) {
            uint32_t r[32];
            ptx::tmem_ld_32x32(tm + lane_off + c0, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) * factor);
            ptx::tmem_st_32x32(tm + lane_off + c0, r);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
        }
        ptx::fence_proxy_async();
        ptx::mbar_arrive_cluster(p_ready_lead + (uint32_t)b * 8u);
        if (t >= 1) {  // PV(t-1) has read its P buffer: the S SM may overwrite it with tile t+1
          ptx::mbar_wait(&ctl->pv_done[b ^ 1], ((t - 1) >> 1) & 1);
          ptx::mbar_arrive_cluster(pin_empty_peer + (uint32_t)(b ^ 1) * 8u);
        }
      }
      ptx::mbar_wait(&ctl->o_done, 0);
      ptx::tc_fence_after();
      const size_t slot = (size_t)strip * p.G + (g < p.G ? g : 0);
#pragma unroll 1
      for (int c0 = 0; c0 < ocols; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tm + lane_off + c0, r);
        ptx::tmem_ld_wait();
        if (g < p.G) {
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            if (c0 + c < p.D)
              *reinterpret_cast<float4*>(p.part_O + slot * p.D + c0 + c) =
                  make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]), __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc2(tm, 512);
}

}  // namespace

// SMs a launch of this kernel can occupy (clusters of 4 do not tile every GPC)
int kp_flash_umma_sv_sms(kp_ctx* ctx) {
  static int cached = 0;
  if (cached) return cached;
  if (cudaFuncSetAttribute(flash_umma_sv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SV_SMEM) != cudaSuccess) {
    cudaGetLastError();
    return cached = ctx->sm_count;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(4 * 148, 1, 1);
  cfg.blockDim = dim3(UT);
  cfg.dynamicSmemBytes = SV_SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, flash_umma_sv_kernel, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return cached = ctx->sm_count;
  }
  return cached = 4 * n;
}

int kp_flash_umma_sv_launch(kp_ctx* ctx, const CUtensorMap& qh_map, const CUtensorMap& ql_map, int G, int KBs, int ngroup,
                            int n_qt, int n_strips, int tps, int mode, float* part_m, float* part_l, float* part_O,
                            cudaStream_t st) {
  SVK p;
  p.G = G;
  p.N = (int)ctx->N;
  p.D = ctx->D;
  p.KB = KBs;
  p.n_tiles = (int)((ctx->N + 127) / 128);
  p.tiles_per_strip = tps;
  p.ngroup = ngroup;
  p.mode = mode;
  p.dbg = (int)ctx->sv_dbg;
  p.part_m = part_m;
  p.part_l = part_l;
  p.part_O = part_O;
  p.prof = ctx->umma_prof;
  p.cursor = ctx->umma_rotate ? ctx->umma_cursor : nullptr;
  if (n_qt % 2 != 0 || tps % 2 != 0)
    KP_FAIL(ctx, KP_EINVAL, "S/V kernel needs an even number of query tiles (%d) and of tiles per strip (%d)", n_qt, tps);
  if (ngroup < 1 || ngroup > 2) KP_FAIL(ctx, KP_EINVAL, "S/V kernel holds at most 512 output dims (%d groups)", ngroup);
  KP_SMEM_ONCE(ctx, flash_umma_sv_kernel, SV_SMEM);
  {
    KpTimer timer(ctx, kp_ctx::T_FLASH, st);
    flash_umma_sv_kernel<<<dim3(2 * n_qt, n_strips, 1), UT, SV_SMEM, st>>>(ctx->um.eh_map, ctx->um.el_map, ctx->um.eh64_map,
                                                                          ctx->um.el64_map, qh_map, ql_map, p);
  }
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
