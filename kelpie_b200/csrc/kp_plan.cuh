// Per-step row plan shared by the ComplEx and ConvE post-training drivers: which rows (pairs)
// of every candidate take part in global optimiser step t, split into "A" rows whose lhs is the
// mimic (query changes every step) and "B" rows with a frozen lhs, compacted into slot arrays.
#pragma once
#include <stdint.h>

namespace {

struct CxPlan {
  int C, N, D, bs, epochs, static_epochs;
  int truth_is_row;  // ConvE: a_truth holds the pair's global index (its positives CSR row)
  const int64_t* row_off;
  const int32_t* rows_per_epoch;
  const int32_t* pos;
  // per candidate
  int32_t* nA;    // [C]
  int32_t* nB;    // [C]
  int32_t* nSelf; // [C] A rows whose truth is M as well
  int64_t* aoff;  // [C+1]
  int64_t* boff;  // [C+1]
  // per slot
  int32_t* a_cand; int32_t* a_rel; int32_t* a_truth;
  int32_t* b_cand; int32_t* b_lhs; int32_t* b_rel;
  int32_t* b_row;  // nullable: global row / pair index of each B slot
};

// rows of candidate c used by global step t: returns B and the first row index
__device__ __forceinline__ int step_rows(const CxPlan& p, int c, int t, int64_t& first) {
  const int n = p.rows_per_epoch[c];
  if (n <= 0) return 0;
  const int spe = (n + p.bs - 1) / p.bs;
  if (t >= p.epochs * spe) return 0;
  const int ep = t / spe, b0 = (t % spe) * p.bs;
  first = p.row_off[c] + (p.static_epochs ? 0 : (int64_t)ep * n) + b0;
  return min(p.bs, n - b0);
}

__global__ void cx_count(const CxPlan p, int t) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  int64_t first = 0;
  const int B = step_rows(p, c, t, first);
  int a = 0, s = 0;
  for (int i = 0; i < B; ++i) {
    const int32_t* r = p.pos + (first + i) * 3;
    if (r[0] == p.N) {
      ++a;
      s += (r[2] == p.N);
    }
  }
  p.nA[c] = a;
  p.nB[c] = B - a;
  p.nSelf[c] = s;
}

// single-block exclusive scans of nA / nB (C up to a few hundred thousand)
__global__ void cx_scan(const CxPlan p) {
  __shared__ int64_t carry[2];
  __shared__ int64_t wsum[2][32];
  if (threadIdx.x == 0) carry[0] = carry[1] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < p.C; base += blockDim.x) {
    const int c = base + threadIdx.x;
    int64_t v[2] = {c < p.C ? p.nA[c] : 0, c < p.C ? p.nB[c] : 0};
    int64_t inc[2];
    for (int k = 0; k < 2; ++k) {
      int64_t x = v[k];
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      inc[k] = x;
      if (lane == 31) wsum[k][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
      for (int k = 0; k < 2; ++k) {
        int64_t x = lane < (blockDim.x >> 5) ? wsum[k][lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
          const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += y;
        }
        wsum[k][lane] = x;  // inclusive over warps
      }
    }
    __syncthreads();
    for (int k = 0; k < 2; ++k) {
      const int64_t before = carry[k] + (warp ? wsum[k][warp - 1] : 0) + inc[k] - v[k];
      if (c < p.C) (k ? p.boff : p.aoff)[c] = before;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      carry[0] += wsum[0][(blockDim.x >> 5) - 1];
      carry[1] += wsum[1][(blockDim.x >> 5) - 1];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    p.aoff[p.C] = carry[0];
    p.boff[p.C] = carry[1];
  }
}

__global__ void cx_assign(const CxPlan p, int t) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  int64_t first = 0;
  const int B = step_rows(p, c, t, first);
  int64_t a = p.aoff[c], b = p.boff[c];
  for (int i = 0; i < B; ++i) {
    const int32_t* r = p.pos + (first + i) * 3;
    if (r[0] == p.N) {
      p.a_cand[a] = c;
      p.a_rel[a] = r[1];
      p.a_truth[a] = p.truth_is_row ? (int32_t)(first + i) : r[2];
      ++a;
    } else {
      p.b_cand[b] = c;
      p.b_lhs[b] = r[0];
      p.b_rel[b] = r[1];
      if (p.b_row) p.b_row[b] = (int32_t)(first + i);
      ++b;
    }
  }
}

}  // namespace
