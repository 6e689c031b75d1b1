// Fused score -> softmax/sigmoid -> contract pass for rows wider than 256 floats WITHOUT the
// duplicated S = Q E^T of the two dim chunks (kp_flash_umma2.cu computes S once per chunk: x1.5 MMA
// work at D = 512).
//
// Cluster of 4 CTAs = two cta_group::2 pairs that own the SAME 256 query rows:
//   pair X (cluster ranks 0,1): output dims [0, 256)     pair Y (ranks 2,3): dims [256, 512)
// O of one chunk fills half of each SM's TMEM, so neither pair can hold both chunks -- but S does not
// depend on the chunk.  The pairs split the entity tiles by parity: X computes S and the
// probabilities P for even tiles, Y for odd tiles, and every P tile is shipped to the other pair
// through distributed shared memory (one 512-byte row per softmax thread, st.shared::cluster into
// the peer CTA that owns the same rows, release/acquire mbarriers).  Both pairs then contract EVERY
// P tile with their own 256 dims:  per entity tile and pair  1/2 S + 1 PV  instead of  1 S + 1 PV.
//
// Online-softmax state across the two producers: the reference max m_ref, the lazy rescale factor
// and the row sum travel in the message header, and every thread folds the tiles in index order
// (own tile: compute; foreign tile: adopt the header), so both pairs scale O identically and end
// with the same (m, l).  A thread needs the header of tile t-1 before the exponentials of tile t;
// the chain costs one softmax latency per tile and is hidden behind S (6144 MMA cycles per own tile).
//
// TMEM (per SM, 512 columns): O [0,256) | A [256,384): own S, overwritten by own P | B [384,512):
// foreign P (copied in from shared memory by the row's thread; both P tiles feed TS-form MMAs).
// MMA order of a pair (p = 0 for X, 1 for Y), mirrored by the TMA producer:
//   if p == 0: S(0);   for t = 0..: if t is foreign { if t+1 < ntile: S(t+1);  PV(t) } else PV(t)
// so the softmax of an own tile overlaps the PV of the foreign tile before it.
#include <cuda_bf16.h>

#include "kp_flash.cuh"
#include "kp_internal.h"
#include "kp_ptx.cuh"
#include "kp_umma_softmax.cuh"

namespace {

constexpr int UT = 192;
constexpr int SLOT = 32768;
constexpr int NSLOT = 4;
constexpr int PIN_BYTES = 33 * 2048;  // 32 x 16-byte chunks per row (chunk-major: chunk q of row r at q*2048 + r*16) + header chunk
constexpr float RESCALE_TAU = 8.0f;

struct UCtl4 {
  uint64_t full[NSLOT], empty[NSLOT];
  uint64_t s_full;                  // own S(k) complete                     (MMA commit, both CTAs of the pair)
  uint64_t p_own, p_for;            // own / foreign P(k) in TMEM            (256 thread arrivals, leader only)
  uint64_t opv_done, fpv_done;      // PV of the k-th own / foreign tile done (MMA commit, both CTAs)
  uint64_t pin_full, pin_empty;     // message k delivered here / my message k consumed by the peer (128 remote arrivals)
  uint64_t o_done;
  uint32_t tmem_base;
  int start, start_local;
};
constexpr size_t U4_SMEM = (size_t)NSLOT * SLOT + PIN_BYTES + sizeof(UCtl4) + 1024;

struct UK4 {
  int G, N, D, KB, n_tiles, tiles_per_strip, groups_per_chunk, mode;
  float* part_m;
  float* part_l;
  float* part_O;
  int* cursor;               // optional [n_strips]: tile the running clusters of a strip are at (rotating start)
  unsigned long long* prof;  // optional [16]: MMA-thread cycles waiting on TMA slots / own P / foreign P, total
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

__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// Own tile: S (TMEM A) -> P, written back over S and sent to the peer CTA (pin_row = cluster address
// of this row's first chunk in the peer's message buffer).  Same arithmetic as umma_sm::p_tile.
template <bool SOFTMAX>
__device__ __forceinline__ void p_tile_send(uint32_t s_addr, int j0, int N, float tau, float& m_ref, float& l_run,
                                            float& factor, uint32_t pin_row, uint64_t* pin_empty, int k) {
  using namespace umma_sm;
  uint32_t r[128];
#pragma unroll
  for (int q = 0; q < 4; ++q) ptx::tmem_ld_32x32(s_addr + 32 * q, reinterpret_cast<uint32_t(&)[32]>(r[32 * q]));
  ptx::tmem_ld_wait();
  if (j0 + 128 > N) {
#pragma unroll
    for (int c = 0; c < 128; ++c)
      if (j0 + c >= N) r[c] = 0xff800000u;
  }
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
    } else if (m > m_ref + tau) {
      factor = ex2((m_ref - m) * LOG2E);
      m_ref = m;
    }
    mneg = (m_ref == -INFINITY) ? 0.f : m_ref * LOG2E;
  }
  // the peer has consumed my previous message (normally long ago: it gates the peer's own next tile)
  if (k > 0) ptx::mbar_wait_cluster(pin_empty, (k - 1) & 1);
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
    ptx::tmem_st_32x32(s_addr + 32 * q, w);
#pragma unroll
    for (int v = 0; v < 8; ++v) st_cluster_v4(pin_row + (uint32_t)(8 * q + v) * 2048u, w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
  }
  const float sum = sum0 + sum1;
  st_cluster_v4(pin_row + 32u * 2048u, __float_as_uint(m_ref), __float_as_uint(factor), __float_as_uint(sum), 0u);
  l_run = l_run * factor + sum;
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(UT, 1)
flash_umma4_kernel(const __grid_constant__ CUtensorMap eh_map, const __grid_constant__ CUtensorMap el_map,
                   const __grid_constant__ CUtensorMap eh64_map, const __grid_constant__ CUtensorMap el64_map,
                   const __grid_constant__ CUtensorMap qh_map, const __grid_constant__ CUtensorMap ql_map, const UK4 p) {
  extern __shared__ uint8_t uraw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(uraw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = sm;
  uint8_t* pin = sm + (size_t)NSLOT * SLOT;
  UCtl4* ctl = reinterpret_cast<UCtl4*>(pin + PIN_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  const uint32_t half = crank & 1;   // which of the two query tiles / which half of every B operand
  const int par = (int)(crank >> 1); // pair: 0 = X (even tiles, first dim chunk), 1 = Y
  const uint32_t lead = crank & ~1u; // MMA-issuing CTA of my pair
  const uint32_t peer = crank ^ 2u;  // CTA of the other pair that owns the same query rows
  const bool leader = half == 0;
  const int strip = blockIdx.y, qtile = (int)(blockIdx.x >> 2) * 2 + (int)half, chunk = par;
  const int t0 = strip * p.tiles_per_strip;
  const int t1 = min(t0 + p.tiles_per_strip, p.n_tiles);
  const int ntile = t1 - t0;
  if (ntile <= 0) return;  // uniform over the cluster
  const int ngroup = p.groups_per_chunk;
  const int box0 = chunk * ngroup * 2;
  const uint16_t pair_mask = (uint16_t)(3u << lead);

  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    ptx::mbar_init(&ctl->s_full, 1);
    ptx::mbar_init(&ctl->p_own, 256);
    ptx::mbar_init(&ctl->p_for, 256);
    ptx::mbar_init(&ctl->opv_done, 1);
    ptx::mbar_init(&ctl->fpv_done, 1);
    ptx::mbar_init(&ctl->pin_full, 128);
    ptx::mbar_init(&ctl->pin_empty, 128);
    ptx::mbar_init(&ctl->o_done, 1);
    // Rotating start: a cluster that begins while others are mid-table joins them at the tile they are
    // at and wraps around (the online softmax is order-independent), so the resident clusters walk
    // the table together and every entity tile is fetched from HBM once per wave instead of once per
    // cluster (ncu at 4096 candidates: 508 GB of DRAM reads per launch without it).
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
  auto tile_of = [&](int t) {  // t-th tile of my walk -> tile index in the table
    const int x = p0 + t;
    return t0 + (x >= ntile ? x - ntile : x);
  };
  const uint32_t tm = ctl->tmem_base;
  const uint32_t TM_O = tm, TM_A = tm + 256, TM_B = tm + 384;

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
      auto load = [&](const CUtensorMap* hi, const CUtensorMap* lo, int col, int row, uint32_t lo_off, uint32_t bytes_pair) {
        const int s = use % NSLOT;
        ptx::mbar_wait(&ctl->empty[s], ((use / NSLOT) & 1) ^ 1);
        if (leader) ptx::mbar_arrive_expect_tx(&ctl->full[s], bytes_pair);
        const uint32_t bar = ptx::mapa_u32(ptx::smem_u32(&ctl->full[s]), lead);
        uint8_t* dst = ring + (size_t)s * SLOT;
        ptx::tma_load_2d_pair(dst, hi, bar, col, row);
        ptx::tma_load_2d_pair(dst + lo_off, lo, bar, col, row);
        ++use;
      };
      auto load_s = [&](int t) {
        for (int kb = 0; kb < p.KB; ++kb) {
          load(&qh_map, &ql_map, kb * 64, qtile * 128, 16384, 2 * 32768);
          load(&eh64_map, &el64_map, kb * 64, tile_of(t) * 128 + (int)half * 64, 8192, 2 * 16384);
        }
      };
      auto load_pv = [&](int t) {
        for (int g = 0; g < ngroup; ++g)
          load(&eh_map, &el_map, (box0 + 2 * g + (int)half) * 64, tile_of(t) * 128, 16384, 2 * 32768);
      };
      if (par == 0) load_s(0);
      for (int t = 0; t < ntile; ++t) {
        if ((t & 1) != par && t + 1 < ntile) load_s(t + 1);
        load_pv(t);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA of each pair) -------------------------------
    if (lane == 0 && leader) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t ring_a = ptx::smem_u32(ring);
      const uint64_t DK = udesc(0, 16, 1024), DMN = udesc(0, 16384, 1024);
      const int ksteps = (p.D + 15) / 16;
      uint32_t use = 0;
      long long w_slot = 0, w_own = 0, w_for = 0;
      const long long t_begin = clock64();
      auto wait_slot = [&](uint32_t u) {
        if (ptx::mbar_try_wait(&ctl->full[u % NSLOT], (u / NSLOT) & 1)) return;
        const long long a = clock64();
        ptx::mbar_wait(&ctl->full[u % NSLOT], (u / NSLOT) & 1);
        w_slot += clock64() - a;
      };
      auto release = [&](uint32_t u) { ptx::umma2_commit_mc(&ctl->empty[u % NSLOT], pair_mask); };
      auto mma_s = [&](int t) {
        if (p.cursor && crank == 0) *(volatile int*)&p.cursor[blockIdx.y] = tile_of(t) - t0;
        for (int kb = 0; kb < p.KB; ++kb) {
          wait_slot(use);
          wait_slot(use + 1);
          ptx::tc_fence_after();
          const uint32_t q_hi = ring_a + (use % NSLOT) * SLOT, q_lo = q_hi + 16384;
          const uint32_t e_hi = ring_a + ((use + 1) % NSLOT) * SLOT, e_lo = e_hi + 8192;
          const uint64_t ah = DK + (q_hi >> 4), al = DK + (q_lo >> 4), bh = DK + (e_hi >> 4), bl = DK + (e_lo >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kb * 4 + kk >= ksteps) break;  // only zero padding beyond D
            ptx::umma2_bf16(TM_A, ah + kk * 2, bh + kk * 2, idesc_s, (kb > 0 || kk > 0) ? 1u : 0u);
            ptx::umma2_bf16(TM_A, ah + kk * 2, bl + kk * 2, idesc_s, 1u);
            ptx::umma2_bf16(TM_A, al + kk * 2, bh + kk * 2, idesc_s, 1u);
          }
          release(use);
          release(use + 1);
          use += 2;
        }
        ptx::umma2_commit_mc(&ctl->s_full, pair_mask);
      };
      auto mma_pv = [&](int t) {
        const bool own = (t & 1) == par;
        const int k = t >> 1;
        {
          const long long a = clock64();
          ptx::mbar_wait_cluster(own ? &ctl->p_own : &ctl->p_for, k & 1);
          (own ? w_own : w_for) += clock64() - a;
        }
        ptx::tc_fence_after();
        const uint32_t p_t = own ? TM_A : TM_B;
        for (int g = 0; g < ngroup; ++g) {
          wait_slot(use);
          ptx::tc_fence_after();
          const uint32_t e_hi = ring_a + (use % NSLOT) * SLOT, e_lo = e_hi + 16384;
          const uint32_t d_o = TM_O + g * 128;
          const uint64_t bh = DMN + (e_hi >> 4), bl = DMN + (e_lo >> 4);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t a_hi = p_t + 32 * (ks >> 1) + 8 * (ks & 1), a_lo = a_hi + 16;
            const uint64_t b_hi = bh + ks * (2048 >> 4), b_lo = bl + ks * (2048 >> 4);
            ptx::umma2_bf16_ts(d_o, a_hi, b_hi, idesc_pv, (t > 0 || ks > 0) ? 1u : 0u);
            ptx::umma2_bf16_ts(d_o, a_hi, b_lo, idesc_pv, 1u);
            ptx::umma2_bf16_ts(d_o, a_lo, b_hi, idesc_pv, 1u);
          }
          release(use);
          ++use;
        }
        ptx::umma2_commit_mc(own ? &ctl->opv_done : &ctl->fpv_done, pair_mask);
      };
      if (par == 0) mma_s(0);
      for (int t = 0; t < ntile; ++t) {
        if ((t & 1) != par && t + 1 < ntile) mma_s(t + 1);
        mma_pv(t);
      }
      ptx::umma2_commit_mc(&ctl->o_done, pair_mask);
      if (p.prof) {
        atomicAdd(p.prof + 0, (unsigned long long)w_slot);
        atomicAdd(p.prof + 1, (unsigned long long)w_own);
        atomicAdd(p.prof + 2, (unsigned long long)w_for);
        atomicAdd(p.prof + 3, (unsigned long long)(clock64() - t_begin));
      }
    }
  } else {
    // ------------------------------- softmax / exchange / epilogue (every CTA, own rows) -------------------------------
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const uint32_t lane_off = (uint32_t)(sub * 32) << 16;
    const int g = qtile * 128 + row;
    const uint32_t p_own_lead = ptx::mapa_u32(ptx::smem_u32(&ctl->p_own), lead);
    const uint32_t p_for_lead = ptx::mapa_u32(ptx::smem_u32(&ctl->p_for), lead);
    const uint32_t pin_local = ptx::smem_u32(pin) + (uint32_t)row * 16u;
    const uint32_t pin_peer = ptx::mapa_u32(pin_local, peer);
    const uint32_t pin_full_peer = ptx::mapa_u32(ptx::smem_u32(&ctl->pin_full), peer);
    const uint32_t pin_empty_peer = ptx::mapa_u32(ptx::smem_u32(&ctl->pin_empty), peer);
    float m_ref = -INFINITY, l_run = 0.f;
    const int ocols = ngroup * 128;
    for (int t = 0; t < ntile; ++t) {
      const bool own = (t & 1) == par;
      const int k = t >> 1;
      float factor;
      if (own) {
        ptx::mbar_wait(&ctl->s_full, k & 1);
        ptx::tc_fence_after();
        const int j0 = tile_of(t) * 128;
        if (p.mode == KP_FLASH_SOFTMAX)
          p_tile_send<true>(TM_A + lane_off, j0, p.N, RESCALE_TAU, m_ref, l_run, factor, pin_peer, &ctl->pin_empty, k);
        else
          p_tile_send<false>(TM_A + lane_off, j0, p.N, RESCALE_TAU, m_ref, l_run, factor, pin_peer, &ctl->pin_empty, k);
        ptx::mbar_arrive_cluster(pin_full_peer);  // release: my row of the message is complete
      } else {
        ptx::mbar_wait_cluster(&ctl->pin_full, k & 1);
        uint32_t w[128];
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const uint4 v = ld_shared_v4(pin_local + (uint32_t)q * 2048u);
          w[4 * q] = v.x;
          w[4 * q + 1] = v.y;
          w[4 * q + 2] = v.z;
          w[4 * q + 3] = v.w;
        }
        const uint4 h = ld_shared_v4(pin_local + 32u * 2048u);
        ptx::mbar_arrive_cluster(pin_empty_peer);  // release: the buffer may be overwritten
        m_ref = __uint_as_float(h.x);
        factor = __uint_as_float(h.y);
        l_run = l_run * factor + __uint_as_float(h.z);
        if (k > 0) {  // B still feeds PV of the previous foreign tile
          ptx::mbar_wait(&ctl->fpv_done, (k - 1) & 1);
          ptx::tc_fence_after();
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) ptx::tmem_st_32x32(TM_B + lane_off + 32 * q, reinterpret_cast<uint32_t(&)[32]>(w[32 * q]));
      }
      if (__any_sync(0xffffffffu, factor != 1.f)) {
        // O holds tiles < t only once PV(t-1) has completed (t >= 1 here: the first tile never rescales)
        const bool prev_own = ((t - 1) & 1) == par;
        ptx::mbar_wait(prev_own ? &ctl->opv_done : &ctl->fpv_done, ((t - 1) >> 1) & 1);
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
      ptx::mbar_arrive_cluster(own ? p_own_lead : p_for_lead);
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

// SMs a launch of the 4-CTA-cluster kernel can occupy (clusters of 4 do not tile every GPC).
int kp_flash_umma4_sms(kp_ctx* ctx) {
  static int cached = 0;
  if (cached) return cached;
  if (cudaFuncSetAttribute(flash_umma4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)U4_SMEM) != cudaSuccess) {
    cudaGetLastError();
    return cached = ctx->sm_count;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(4 * 148, 1, 1);
  cfg.blockDim = dim3(UT);
  cfg.dynamicSmemBytes = U4_SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, flash_umma4_kernel, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return cached = ctx->sm_count;
  }
  return cached = 4 * n;
}

int kp_flash_umma4_launch(kp_ctx* ctx, const CUtensorMap& qh_map, const CUtensorMap& ql_map, int G, int KBs,
                          int groups_per_chunk, int n_qt, int n_strips, int tps, int mode, float* part_m, float* part_l,
                          float* part_O, cudaStream_t st) {
  UK4 p;
  p.G = G;
  p.N = (int)ctx->N;
  p.D = ctx->D;
  p.KB = KBs;
  p.n_tiles = (int)((ctx->N + 127) / 128);
  p.tiles_per_strip = tps;
  p.groups_per_chunk = groups_per_chunk;
  p.mode = mode;
  p.part_m = part_m;
  p.part_l = part_l;
  p.part_O = part_O;
  p.prof = ctx->umma_prof;
  p.cursor = ctx->umma_rotate ? ctx->umma_cursor : nullptr;
  KP_SMEM_ONCE(ctx, flash_umma4_kernel, U4_SMEM);
  if (n_qt % 2 != 0) KP_FAIL(ctx, KP_EINVAL, "cluster-4 kernel needs an even number of query tiles (%d)", n_qt);
  {
    KpTimer timer(ctx, kp_ctx::T_FLASH, st);
    flash_umma4_kernel<<<dim3(2 * n_qt, n_strips, 1), UT, U4_SMEM, st>>>(ctx->um.eh_map, ctx->um.el_map, ctx->um.eh64_map,
                                                                        ctx->um.el64_map, qh_map, ql_map, p);
  }
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
