// Counter-based dropout masks for ConvE post-training.  The mask of an element depends only on
// (seed, pair id, optimiser step, element id), so it is independent of how candidates are batched
// and can be recomputed in the backward kernels instead of being stored.
// (The reference draws its masks from torch's Philox stream, conve.py:140-152; bit-level parity
// with dropout > 0 would need those very masks, so parity is defined at rate 0 -- SURVEY.md 8d
// config 3 -- and tests/test_gpu_conve_dropout.py checks this generator against a torch
// restatement that uses the same masks.)
#pragma once
#include <stdint.h>

enum { KP_DROP_INPUT = 0, KP_DROP_FEATURE = 1 << 20, KP_DROP_HIDDEN = 2 << 20 };

__host__ __device__ __forceinline__ uint32_t kp_drop_hash(unsigned long long seed, uint32_t pair, uint32_t step, uint32_t elem) {
  unsigned long long x = seed ^ (((unsigned long long)pair << 32) | step);
  x += 0x9E3779B97F4A7C15ull * ((unsigned long long)elem + 1ull);
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 32);
}

// 0 with probability p, 1 / (1 - p) otherwise (torch.nn.functional.dropout scaling)
__host__ __device__ __forceinline__ float kp_drop_scale(unsigned long long seed, int pair, int step, int elem, float p) {
  const uint32_t h = kp_drop_hash(seed, (uint32_t)pair, (uint32_t)step, (uint32_t)elem);
  const uint32_t thr = (uint32_t)((double)p * 4294967296.0);
  return (h < thr) ? 0.f : 1.f / (1.f - p);
}
