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
//   warp 0      TMA producer (4-slot ring of 32 KB {hi,lo} boxes of [128 rows x 64 dims])
//   warp 1      MMA issuer   S = Q E^T  (M128 N128, K-major A/B)      -> TMEM S[2]
//                            O += P E   (M128 N64,  A = P K-major from smem, B = E MN-major)
//   warps 2..5  one thread per query row: TMEM S -> online softmax (lazy rescale of O in
//               TMEM) or sigmoid -> P split into bf16 hi/lo, written to smem in the
//               128B-swizzled K-major layout the MMA reads
// PV of tile i-1 is issued behind S of tile i so the softmax of a tile overlaps tensor work.
#include <cuda_bf16.h>

#include "kp_flash.cuh"
#include "kp_internal.h"
#include "kp_ptx.cuh"

namespace {

constexpr int UT = 192;               // threads
constexpr int SLOT = 32768;           // {hi 16 KB | lo 16 KB}
constexpr int NSLOT = 4;
constexpr int P_BYTES = 65536;        // Ph[2][16 KB] | Pl[2][16 KB]
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_TAU = 8.0f;   // rescale O only when the row max grew by more than this

struct UCtl {
  uint64_t full[NSLOT], empty[NSLOT];
  uint64_t s_full[2], s_empty[2];
  uint64_t p_full, p_empty, o_done;
  uint32_t tmem_base;
};
constexpr size_t U_SMEM = (size_t)NSLOT * SLOT + P_BYTES + sizeof(UCtl) + 1024;

struct UK {
  int G, N, D, KB, n_tiles, tiles_per_strip, boxes_per_chunk, mode;
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
  uint8_t* Pbuf = sm + (size_t)NSLOT * SLOT;
  UCtl* ctl = reinterpret_cast<UCtl*>(Pbuf + P_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int strip = blockIdx.x, qtile = blockIdx.y, chunk = blockIdx.z;
  const int t0 = strip * p.tiles_per_strip;
  const int t1 = min(t0 + p.tiles_per_strip, p.n_tiles);
  const int ntile = t1 - t0;
  const int box0 = chunk * p.boxes_per_chunk;
  const int nbox = min(p.boxes_per_chunk, p.KB - box0);
  if (ntile <= 0 || nbox <= 0) return;

  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&ctl->s_full[b], 1);
      ptx::mbar_init(&ctl->s_empty[b], 128);
    }
    ptx::mbar_init(&ctl->p_full, 128);
    ptx::mbar_init(&ctl->p_empty, 1);
    ptx::mbar_init(&ctl->o_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&ctl->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = ctl->tmem_base;
  const uint32_t TM_O = tm, TM_S = tm + 256;  // O: cols [0,256)   S[b]: cols [256 + 128 b, +128)

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      ptx::prefetch_tmap(&eh_map);
      ptx::prefetch_tmap(&el_map);
      ptx::prefetch_tmap(&qh_map);
      ptx::prefetch_tmap(&ql_map);
      uint32_t use = 0;
      auto load = [&](const CUtensorMap* hi, const CUtensorMap* lo, int col, int row) {
        const int s = use % NSLOT;
        ptx::mbar_wait(&ctl->empty[s], ((use / NSLOT) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&ctl->full[s], SLOT);
        ptx::tma_load_2d(ring + (size_t)s * SLOT, hi, &ctl->full[s], col, row);
        ptx::tma_load_2d(ring + (size_t)s * SLOT + 16384, lo, &ctl->full[s], col, row);
        ++use;
      };
      for (int i = 0; i <= ntile; ++i) {
        if (i < ntile)
          for (int kb = 0; kb < p.KB; ++kb) {
            load(&qh_map, &ql_map, kb * 64, qtile * 128);
            load(&eh_map, &el_map, kb * 64, (t0 + i) * 128);
          }
        if (i > 0)
          for (int b = 0; b < nbox; ++b) load(&eh_map, &el_map, (box0 + b) * 64, (t0 + i - 1) * 128);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t ring_a = ptx::smem_u32(ring), p_a = ptx::smem_u32(Pbuf);
      uint32_t use = 0;
      auto wait_slot = [&](uint32_t u) { ptx::mbar_wait(&ctl->full[u % NSLOT], (u / NSLOT) & 1); };
      auto pv = [&](int t) {
        ptx::mbar_wait(&ctl->p_full, t & 1);
        ptx::tc_fence_after();
        for (int b = 0; b < nbox; ++b) {
          wait_slot(use);
          ptx::tc_fence_after();
          const uint32_t e_hi = ring_a + (use % NSLOT) * SLOT, e_lo = e_hi + 16384;
          const uint32_t d_o = TM_O + b * 64;
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t pa = (ks >> 2) * 16384 + (ks & 3) * 32;
            const uint64_t a_hi = udesc(p_a + pa, 16, 1024), a_lo = udesc(p_a + 32768 + pa, 16, 1024);
            const uint64_t b_hi = udesc(e_hi + ks * 2048, 16384, 1024), b_lo = udesc(e_lo + ks * 2048, 16384, 1024);
            ptx::umma_bf16(d_o, a_hi, b_hi, idesc_pv, (t > 0 || ks > 0) ? 1u : 0u);
            ptx::umma_bf16(d_o, a_hi, b_lo, idesc_pv, 1u);
            ptx::umma_bf16(d_o, a_lo, b_hi, idesc_pv, 1u);
          }
          ptx::umma_commit(&ctl->empty[use % NSLOT]);
          ++use;
        }
        ptx::umma_commit(&ctl->p_empty);
      };
      for (int i = 0; i < ntile; ++i) {
        const int sb = i & 1;
        ptx::mbar_wait(&ctl->s_empty[sb], ((i >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_s = TM_S + sb * 128;
        for (int kb = 0; kb < p.KB; ++kb) {
          wait_slot(use);
          wait_slot(use + 1);
          ptx::tc_fence_after();
          const uint32_t q_hi = ring_a + (use % NSLOT) * SLOT, q_lo = q_hi + 16384;
          const uint32_t e_hi = ring_a + ((use + 1) % NSLOT) * SLOT, e_lo = e_hi + 16384;
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t a_hi = udesc(q_hi + kk * 32, 16, 1024), a_lo = udesc(q_lo + kk * 32, 16, 1024);
            const uint64_t b_hi = udesc(e_hi + kk * 32, 16, 1024), b_lo = udesc(e_lo + kk * 32, 16, 1024);
            ptx::umma_bf16(d_s, a_hi, b_hi, idesc_s, (kb > 0 || kk > 0) ? 1u : 0u);
            ptx::umma_bf16(d_s, a_hi, b_lo, idesc_s, 1u);
            ptx::umma_bf16(d_s, a_lo, b_hi, idesc_s, 1u);
          }
          ptx::umma_commit(&ctl->empty[use % NSLOT]);
          ptx::umma_commit(&ctl->empty[(use + 1) % NSLOT]);
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
      float factor = 1.f;
      if (p.mode == KP_FLASH_SOFTMAX) {
        float mx = -INFINITY;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(s_addr + c0, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (j0 + c0 + c < p.N) mx = fmaxf(mx, __uint_as_float(r[c]));
        }
        if (m_ref == -INFINITY) {
          m_ref = mx;
        } else if (mx > m_ref + RESCALE_TAU) {
          factor = exp2f((m_ref - mx) * LOG2E);
          m_ref = mx;
        }
      }
      // the previous tile's P has been consumed and O is up to date
      ptx::mbar_wait(&ctl->p_empty, (i & 1) ^ 1);
      ptx::tc_fence_after();
      if (__any_sync(0xffffffffu, factor != 1.f)) {
#pragma unroll 1
        for (int c0 = 0; c0 < ocols; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(TM_O + lane_off + c0, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) * factor);
          ptx::tmem_st_32x32(TM_O + lane_off + c0, r);
        }
        ptx::tmem_st_wait();
      }
      float sum = 0.f;
      const float mneg = (m_ref == -INFINITY) ? 0.f : m_ref * LOG2E;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(s_addr + c0, r);
        ptx::tmem_ld_wait();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float pv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float s = __uint_as_float(r[c + u]);
            float e;
            if (p.mode == KP_FLASH_SOFTMAX)
              e = exp2f(__fmaf_rn(s, LOG2E, -mneg));
            else
              e = 1.f / (1.f + expf(-s));
            pv[u] = (j0 + c0 + c + u < p.N) ? e : 0.f;
          }
          sum += pv[0] + pv[1];
          const __nv_bfloat16 h0 = __float2bfloat16_rn(pv[0]), h1 = __float2bfloat16_rn(pv[1]);
          hi[c >> 1] = pack_bf16(h0, h1);
          lo[c >> 1] = pack_bf16(__float2bfloat16_rn(pv[0] - __bfloat162float(h0)),
                                 __float2bfloat16_rn(pv[1] - __bfloat162float(h1)));
        }
        // entity columns c0..c0+31 -> k-block (c0 / 64), 16-byte chunks ((c0 % 64) / 8) .. +3, swizzled by row
        uint8_t* base = Pbuf + (c0 >> 6) * 16384 + row * 128;
        const int ch0 = (c0 & 63) >> 3;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int phys = ((ch0 + q4) ^ (row & 7)) << 4;
          *reinterpret_cast<uint4*>(base + phys) = make_uint4(hi[q4 * 4], hi[q4 * 4 + 1], hi[q4 * 4 + 2], hi[q4 * 4 + 3]);
          *reinterpret_cast<uint4*>(base + 32768 + phys) = make_uint4(lo[q4 * 4], lo[q4 * 4 + 1], lo[q4 * 4 + 2], lo[q4 * 4 + 3]);
        }
      }
      l_run = l_run * factor + sum;
      ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ctl->p_full);
      ptx::mbar_arrive(&ctl->s_empty[sb]);
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
  return !ctx->force_simt && G >= 32 && ctx->D <= 512 && ctx->D % 4 == 0;
}

int kp_flash_umma_plan(kp_ctx* ctx, int G, int* n_strips) {
  const int KB = (ctx->D + 63) / 64;
  const int boxes_per_chunk = KB <= 4 ? KB : (KB + 1) / 2 > 4 ? 4 : (KB + 1) / 2;
  const int n_chunks = (KB + boxes_per_chunk - 1) / boxes_per_chunk;
  const int n_tiles = (int)((ctx->N + 127) / 128);
  const int n_qt = (G + 127) / 128;
  int s = ctx->sm_count / (n_qt * n_chunks);
  if (s > 64) s = 64;
  if (s > n_tiles) s = n_tiles;
  if (s < 1) s = 1;
  const int tps = (n_tiles + s - 1) / s;
  *n_strips = (n_tiles + tps - 1) / tps;
  return boxes_per_chunk;
}

int kp_flash_umma(kp_ctx* ctx, const float* qmat, int G, int mode, float* part_m, float* part_l, float* part_O,
                  cudaStream_t st) {
  if (G <= 0) return KP_OK;
  const int D = ctx->D, Dpad = ((D + 63) / 64) * 64;
  int rc;
  if (!ctx->um.ready) {
    const long long Npad = ((ctx->N + 127) / 128) * 128;
    void *h = nullptr, *l = nullptr;
    if (cudaMalloc(&h, (size_t)Npad * Dpad * 2) != cudaSuccess || cudaMalloc(&l, (size_t)Npad * Dpad * 2) != cudaSuccess)
      KP_FAIL(ctx, KP_ENOMEM, "cannot allocate the split bf16 entity tables (%lld x %d)", Npad, Dpad);
    ctx->owned.push_back(h);
    ctx->owned.push_back(l);
    split_bf16_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->ent, ctx->N, D, Npad, Dpad, (__nv_bfloat16*)h, (__nv_bfloat16*)l);
    KP_LAUNCHED(ctx, 1);
    if ((rc = encode_bf16(ctx, &ctx->um.eh_map, h, Npad, Dpad)) != KP_OK) return rc;
    if ((rc = encode_bf16(ctx, &ctx->um.el_map, l, Npad, Dpad)) != KP_OK) return rc;
    ctx->um.ent_hi = h;
    ctx->um.ent_lo = l;
    ctx->um.ready = true;
  }
  const long long Gpad = ((G + 127) / 128) * 128;
  const size_t qbytes = (size_t)Gpad * Dpad * 2;
  if ((rc = kp_ws_reserve(ctx, 2 * qbytes + 2048, 1)) != KP_OK) return rc;
  __nv_bfloat16* qh = reinterpret_cast<__nv_bfloat16*>(ctx->ws_arena[1]);
  __nv_bfloat16* ql = reinterpret_cast<__nv_bfloat16*>(ctx->ws_arena[1] + ((qbytes + 1023) & ~size_t(1023)));
  split_bf16_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(qmat, G, D, Gpad, Dpad, qh, ql);
  KP_LAUNCHED(ctx, 1);
  CUtensorMap qh_map, ql_map;
  if ((rc = encode_bf16(ctx, &qh_map, qh, Gpad, Dpad)) != KP_OK) return rc;
  if ((rc = encode_bf16(ctx, &ql_map, ql, Gpad, Dpad)) != KP_OK) return rc;

  UK p;
  p.G = G;
  p.N = (int)ctx->N;
  p.D = D;
  p.KB = Dpad / 64;
  p.n_tiles = (int)((ctx->N + 127) / 128);
  int n_strips = 1;
  p.boxes_per_chunk = kp_flash_umma_plan(ctx, G, &n_strips);
  p.tiles_per_strip = (p.n_tiles + n_strips - 1) / n_strips;
  p.mode = mode;
  p.part_m = part_m;
  p.part_l = part_l;
  p.part_O = part_O;
  const int n_chunks = (p.KB + p.boxes_per_chunk - 1) / p.boxes_per_chunk;
  static bool configured = false;
  if (!configured) {
    KP_CUDA(ctx, cudaFuncSetAttribute(flash_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)U_SMEM));
    configured = true;
  }
  dim3 grid(n_strips, (G + 127) / 128, n_chunks);
  {
    KpTimer timer(ctx, kp_ctx::T_FLASH, st);
    flash_umma_kernel<<<grid, UT, U_SMEM, st>>>(ctx->um.eh_map, ctx->um.el_map, qh_map, ql_map, p);
  }
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
