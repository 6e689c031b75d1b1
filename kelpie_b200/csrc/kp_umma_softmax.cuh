// Per-row work of the tcgen05 fused pass between the two MMA phases (shared by kp_flash_umma.cu and
// kp_flash_umma2.cu): one thread owns one query row = one TMEM lane and turns the 128 fp32 logits of
// an entity tile into probabilities, split into bf16 hi / lo and written back over the same columns.
//
// Instruction diet (ncu of the first version showed the XU pipe saturated: every scalar
// float -> bf16 conversion is an F2F on the 16-lane XU pipe, 4 per element, next to the MUFU.EX2):
//   * cvt.rn.bf16x2.f32 packs two conversions into one F2FP on the ALU pipe; hi is widened back with
//     a shift / mask, so a pair of elements costs 2 F2FP + 2 ALU + 2 FADD instead of 8 F2F;
//   * ex2.approx.ftz directly (MUFU.EX2 without the denormal pre/post scaling of exp2f);
//   * the 128 logits stay in registers between the max scan and the exponentials (one TMEM read,
//     four loads in flight behind a single wait) and the entity bound is applied once, on the last
//     tile only, by overwriting the padding columns with -inf.
#pragma once
#include <stdint.h>

#include "kp_ptx.cuh"

namespace umma_sm {

constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// {e0 in the low half, e1 in the high half}, round to nearest even
__device__ __forceinline__ uint32_t bf16x2(float e0, float e1) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
  return r;
}

// S tile of this thread's row at TMEM address s_addr (128 columns; entity ids j0 .. j0+127, valid
// below N).  SOFTMAX: p = 2^((s - m_ref) log2 e) with the lazy reference max (m_ref only moves when
// the tile max exceeds it by more than tau; `factor` then rescales what was accumulated so far);
// otherwise p = sigmoid(s).  Adds sum(p) to l_run (after applying factor).  The caller still has
// to tcgen05.wait::st before signalling the MMA warp.
template <bool SOFTMAX>
__device__ __forceinline__ void p_tile(uint32_t s_addr, int j0, int N, float tau, float& m_ref, float& l_run,
                                       float& factor) {
  uint32_t r[128];
#pragma unroll
  for (int k = 0; k < 4; ++k) ptx::tmem_ld_32x32(s_addr + 32 * k, reinterpret_cast<uint32_t(&)[32]>(r[32 * k]));
  ptx::tmem_ld_wait();
  if (j0 + 128 > N) {  // last tile of the table (uniform over the CTA)
#pragma unroll
    for (int c = 0; c < 128; ++c)
      if (j0 + c >= N) r[c] = 0xff800000u;  // -inf -> p = 0 in both modes
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
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t w[32];
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      float p0, p1;
      if (SOFTMAX) {
        p0 = ex2(__fmaf_rn(__uint_as_float(r[32 * k + c]), LOG2E, -mneg));
        p1 = ex2(__fmaf_rn(__uint_as_float(r[32 * k + c + 1]), LOG2E, -mneg));
      } else {
        p0 = rcp(1.f + ex2(-LOG2E * __uint_as_float(r[32 * k + c])));
        p1 = rcp(1.f + ex2(-LOG2E * __uint_as_float(r[32 * k + c + 1])));
      }
      sum0 += p0;
      sum1 += p1;
      const uint32_t hi = bf16x2(p0, p1);
      w[c >> 1] = hi;
      w[16 + (c >> 1)] = bf16x2(p0 - __uint_as_float(hi << 16), p1 - __uint_as_float(hi & 0xffff0000u));
    }
    ptx::tmem_st_32x32(s_addr + 32 * k, w);
  }
  l_run = l_run * factor + (sum0 + sum1);
}

}  // namespace umma_sm
