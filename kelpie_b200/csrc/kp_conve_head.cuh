// Head of the ConvE network behind the Linear layer (conve.py:150-155, eval-mode BN): x = ReLU(BN3(dropout(raw + fc_b))).
// One definition for the stand-alone kernel (kp_conve.cu) and the Linear GEMM's epilogue (kp_gemm_umma.cu): same roundings.
#pragma once
#include "kp_dropout.cuh"

struct kp_conve_head {
  const float* fc_b;        // [D]
  const float* bn3;         // {weight[D], bias[D], mean[D], var[D]}
  const int32_t* drop_ids;  // NULL = eval mode; else the pair id keying the masks of row q
  unsigned long long seed;
  int step;
  float p_hid;
};

__device__ __forceinline__ void bn_affine(const float* bn, int n, int i, float& alpha, float& beta) {
  // bn = {weight[n], bias[n], mean[n], var[n]}
  const float inv = 1.f / sqrtf(bn[3 * n + i] + 1e-5f);
  alpha = bn[i] * inv;
  beta = bn[n + i] - bn[2 * n + i] * alpha;
}

__device__ __forceinline__ float kp_conve_head_apply(const kp_conve_head& hd, int D, int q, int k, float raw) {
  float a3, b3;
  bn_affine(hd.bn3, D, k, a3, b3);
  float h = raw + hd.fc_b[k];
  if (hd.drop_ids && hd.p_hid > 0.f) h *= kp_drop_scale(hd.seed, hd.drop_ids[q], hd.step, KP_DROP_HIDDEN + k, hd.p_hid);
  return fmaxf(__fmaf_rn(h, a3, b3), 0.f);
}
