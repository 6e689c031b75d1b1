// Full-model ConvE training for verify_explanations' retrain-from-scratch:
// BCEOptimizer.step_on_batch (bce_optimizer.py:137-158) with ConvE.forward = all_scores (conve.py:133-158),
// BCELoss(mean) against label-smoothed multi-hot targets (:98-112) and optim.Adam over every parameter (:36).
//
// One step of B (lhs, rel) pairs; image = [lhs ; rel] as 40 x W floats (W = D / 20), F = 32 filters, conv output
// 38 x (W - 2) per filter, hidden = F * 38 * (W - 2).  Batch-norm layers run in TRAIN mode (batch statistics) unless
// B == 1 (:140-156).
//   vfit_gather      X0[i] = [E[lhs_i] ; R[rel_i]], batch sums for batch-norm 1                   CTA per pair
//   vfit_conv_fwd    in = dropout(bn1(X0));  C = conv3x3(in) + b;  per-filter sums for bn2          CTA per pair
//   vfit_bn2_relu    feat = dropout2d(relu(bn2(C)));  running statistics of bn1 / bn2               elementwise
//   GEMM             H[B, D] = feat W^T                                tcgen05, tf32x3 split   (kp_gemm_umma.cu)
//   vfit_bn3_fwd     H = dropout(H + b);  X = relu(bn3(H));  running statistics of bn3              CTA per 32 columns
//   GEMM             Z[B, N] = X E^T
//   vfit_bce         P = (sigmoid(Z) - target) / (B N) in place, loss                               CTA per pair
//   GEMM             dX[B, D] = P E            GEMM   gE[N, D] = P^T X   (the dense table gradient)
//   vfit_bn3_bwd     through relu / bn3 / dropout: dH, d gamma3, d beta3, d b                        CTA per 32 columns
//   GEMM             gW[D, hidden] = dH^T feat GEMM   dfeat[B, hidden] = dH W
//   vfit_bn2_bwd     dy = dfeat through dropout2d / relu (in place), per-filter sums of dy, dy * c^ CTA per pair
//   vfit_conv_bwd    dC through bn2; d conv_w, d conv_b; d in = conv-transpose; sums for bn1        CTA per pair
//   vfit_scatter     through bn1; gE[lhs] += ..., gR[rel] += ...; d gamma / d beta of bn1, bn2       CTA per pair
//   vfit_adam        torch.optim.Adam over all twelve parameter tensors                             segment table
// Batch statistics are accumulated in fp64 (sum, sum of squares) so the one-pass variance is exact to fp32.
#include "kp_dropout.cuh"
#include "kp_internal.h"

namespace {
constexpr int NSEG = 12;
enum { SEG_ENT = 0, SEG_REL, SEG_CONV_W, SEG_CONV_B, SEG_FC_W, SEG_FC_B, SEG_BN1_W, SEG_BN1_B, SEG_BN2_W, SEG_BN2_B, SEG_BN3_W, SEG_BN3_B };
struct VSeg {
  float* p;
  long long off, n;
};
struct VSegs {
  VSeg s[NSEG];
};
constexpr float BN_EPS = 1e-5f, BN_MOM = 0.1f;
}  // namespace

struct kp_vfit {
  kp_ctx* ctx = nullptr;  // plumbing for the GEMMs (workspace arenas, error text); tables borrowed
  int N = 0, R2 = 0, D = 0, W = 0, OW = 0, FS = 0, F = 0, hidden = 0, max_batch = 0;
  long long ldz = 0, n_params = 0, n_pairs = 0;
  float ls = 0, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f;
  unsigned long long seed = 0;
  kp_conve_params p;
  const int32_t* pairs = nullptr;
  const int64_t* pos_off = nullptr;
  const int32_t* pos_ids = nullptr;
  VSegs segs;
  float *g = nullptr, *m = nullptr, *v = nullptr;
  float *X0 = nullptr, *C = nullptr, *feat = nullptr, *H = nullptr, *X = nullptr, *Z = nullptr, *dX = nullptr, *dH = nullptr,
        *dfeat = nullptr, *dX0 = nullptr, *st3 = nullptr;
  double* stats = nullptr;  // S1[2] | S2[2F] | B2[2F] | B1[2]
  float* dHt = nullptr;     // [D, ceil4(max_batch)] transpose of dH for the fp32 weight-gradient GEMM
  // Precision of the six GEMMs (bit g: 0 H, 1 Z, 2 dX, 3 gE, 4 gW, 5 dfeat).  Measured (tools/debug_conve_fit.py):
  // relu gating + Adam's scale invariance turn the 1e-5 perturbations of bf16x3 products into 1e-3 L-inf deviations of
  // the trained tensors within 72 steps, while tf32x3 (error ~2^-21 per product, kind::tf32 at half the bf16 MMA rate)
  // stays at the reference's own fp32-vs-fp64 noise floor (1e-5), like exact fp32.  Default: the five GEMMs on the
  // activation path as tf32x3 on the tensor cores; gE = P^T X (no cancellation, the largest GEMM) as bf16x3.
  // KP_VFIT_TF32=<mask> / KP_VFIT_FP32=<mask> override (diagnostics): tf32x3 wins over fp32 CUDA cores wins over bf16x3.
  int tf32_mask = 55;
  int fp32_mask = 0;
  long long t = 0;
  int64_t launches = 0;
  std::string err;
};

namespace {

std::string g_vfit_error;

struct VDims {
  int B, D, W, OW, FS, F, hidden, train_bn, step;
  float p_in, p_fm, p_hid;
  unsigned long long seed;
};

__device__ __forceinline__ double warp_sum(double x) {
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ float warp_sumf(float x) {
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
// sum of (a, b) over the CTA; the result is valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b, double* red) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    red[2 * w] = a;
    red[2 * w + 1] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0.0;
    b = 0.0;
    for (int i = 0; i < nw; ++i) {
      a += red[2 * i];
      b += red[2 * i + 1];
    }
  }
}
__device__ __forceinline__ void bn_moments(double s, double ss, double M, float& mean, float& var) {
  const double mu = s / M;
  double va = ss / M - mu * mu;
  if (va < 0.0) va = 0.0;
  mean = (float)mu;
  var = (float)va;
}
__device__ __forceinline__ float inv_std(float var) { return 1.f / sqrtf(var + BN_EPS); }
__device__ __forceinline__ float drop(const VDims& d, float rate, int pid, int elem) {
  return rate > 0.f ? kp_drop_scale(d.seed, pid, d.step, elem, rate) : 1.f;
}
// batch-norm 1 statistics of this step: batch moments (train) or the running ones (eval)
__device__ __forceinline__ void bn1_stats(const VDims& d, const kp_conve_params& p, const double* S1, float& mean, float& istd) {
  float var;
  if (d.train_bn) bn_moments(S1[0], S1[1], (double)d.B * 2.0 * d.D, mean, var);
  else mean = p.bn1_mean[0], var = p.bn1_var[0];
  istd = inv_std(var);
}
__device__ __forceinline__ void bn2_stats(const VDims& d, const kp_conve_params& p, const double* S2, int f, float& mean, float& istd) {
  float var;
  if (d.train_bn) bn_moments(S2[2 * f], S2[2 * f + 1], (double)d.B * d.FS, mean, var);
  else mean = p.bn2_mean[f], var = p.bn2_var[f];
  istd = inv_std(var);
}

__global__ void __launch_bounds__(128) vfit_gather(VDims d, kp_conve_params p, const int32_t* __restrict__ pairs,
                                                   const int32_t* __restrict__ order, float* __restrict__ X0, double* __restrict__ S1) {
  __shared__ double red[8];
  const int i = blockIdx.x, pid = order[i];
  const float* l = p.ent + (size_t)pairs[2 * pid] * d.D;
  const float* r = p.rel + (size_t)pairs[2 * pid + 1] * d.D;
  double s = 0.0, ss = 0.0;
  for (int k = threadIdx.x; k < 2 * d.D; k += blockDim.x) {
    const float x = k < d.D ? l[k] : r[k - d.D];
    X0[(size_t)i * 2 * d.D + k] = x;
    s += x;
    ss += (double)x * x;
  }
  block_sum2(s, ss, red);
  if (threadIdx.x == 0 && d.train_bn) {
    atomicAdd(S1, s);
    atomicAdd(S1 + 1, ss);
  }
}

// dynamic smem: in[2D] | w[9F] | b[F]
__global__ void __launch_bounds__(256) vfit_conv_fwd(VDims d, kp_conve_params p, const int32_t* __restrict__ order,
                                                     const float* __restrict__ X0, float* __restrict__ C, const double* __restrict__ S1,
                                                     double* __restrict__ S2) {
  extern __shared__ float sm[];
  float* in = sm;
  float* w = in + 2 * d.D;
  float* b = w + 9 * d.F;
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, pid = order[i];
  float mean, istd;
  bn1_stats(d, p, S1, mean, istd);
  const float sc = istd * p.bn1_w[0], sh = p.bn1_b[0] - mean * sc;
  for (int k = tid; k < 2 * d.D; k += 256) in[k] = (X0[(size_t)i * 2 * d.D + k] * sc + sh) * drop(d, d.p_in, pid, KP_DROP_INPUT + k);
  for (int k = tid; k < 9 * d.F; k += 256) w[k] = p.conv_w[k];
  for (int k = tid; k < d.F; k += 256) b[k] = p.conv_b[k];
  __syncthreads();
  for (int f = warp; f < d.F; f += 8) {
    float wf[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wf[k] = w[9 * f + k];
    double s = 0.0, ss = 0.0;
    for (int pos = lane; pos < d.FS; pos += 32) {
      const int y = pos / d.OW, x = pos - y * d.OW;
      float acc = b[f];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) acc = fmaf(in[(y + ky) * d.W + x + kx], wf[3 * ky + kx], acc);
      C[(size_t)i * d.hidden + f * d.FS + pos] = acc;
      s += acc;
      ss += (double)acc * acc;
    }
    if (d.train_bn) {
      s = warp_sum(s);
      ss = warp_sum(ss);
      if (lane == 0) {
        atomicAdd(S2 + 2 * f, s);
        atomicAdd(S2 + 2 * f + 1, ss);
      }
    }
  }
}

// feat = dropout2d(relu(bn2(C))); CTA 0 also folds the batch moments of bn1 / bn2 into the running statistics
__global__ void __launch_bounds__(256) vfit_bn2_relu(VDims d, kp_conve_params p, const int32_t* __restrict__ order, const float* __restrict__ C,
                                                     float* __restrict__ feat, const double* __restrict__ S1, const double* __restrict__ S2) {
  __shared__ float sc[64], sh[64];
  const int tid = threadIdx.x;
  if (tid < d.F) {
    float mean, istd;
    bn2_stats(d, p, S2, tid, mean, istd);
    sc[tid] = istd * p.bn2_w[tid];
    sh[tid] = p.bn2_b[tid] - mean * sc[tid];
    if (blockIdx.x == 0 && d.train_bn) {
      float var;
      const double M = (double)d.B * d.FS;
      bn_moments(S2[2 * tid], S2[2 * tid + 1], M, mean, var);
      p.bn2_mean[tid] = (1.f - BN_MOM) * p.bn2_mean[tid] + BN_MOM * mean;
      p.bn2_var[tid] = (1.f - BN_MOM) * p.bn2_var[tid] + BN_MOM * (float)((double)var * M / (M - 1.0));
    }
  }
  if (blockIdx.x == 0 && tid == 64 && d.train_bn) {
    float mean, var;
    const double M = (double)d.B * 2.0 * d.D;
    bn_moments(S1[0], S1[1], M, mean, var);
    p.bn1_mean[0] = (1.f - BN_MOM) * p.bn1_mean[0] + BN_MOM * mean;
    p.bn1_var[0] = (1.f - BN_MOM) * p.bn1_var[0] + BN_MOM * (float)((double)var * M / (M - 1.0));
  }
  __syncthreads();
  const long long total = (long long)d.B * d.hidden;
  for (long long idx = blockIdx.x * 256ll + tid; idx < total; idx += (long long)gridDim.x * 256) {
    const int i = (int)(idx / d.hidden), c = (int)(idx - (long long)i * d.hidden), f = c / d.FS;
    float v = fmaxf(fmaf(C[idx], sc[f], sh[f]), 0.f);
    if (d.p_fm > 0.f) v *= drop(d, d.p_fm, order[i], KP_DROP_FEATURE + f);
    feat[idx] = v;
  }
}

// CTA = 32 columns x 32 row groups.  H <- dropout(H + b) (the batch-norm input, kept for the backward); X = relu(bn3(H))
__global__ void __launch_bounds__(1024) vfit_bn3_fwd(VDims d, kp_conve_params p, const int32_t* __restrict__ order, float* __restrict__ H,
                                                    float* __restrict__ X, float* __restrict__ st3) {
  __shared__ double red[2][32][32];
  __shared__ float bc[2][32];
  const int cx = threadIdx.x, ry = threadIdx.y, c = blockIdx.x * 32 + cx;
  const bool live = c < d.D;
  double s = 0.0, ss = 0.0;
  if (live) {
    const float bias = p.fc_b[c];
    for (int i = ry; i < d.B; i += 32) {
      float h = H[(size_t)i * d.D + c] + bias;
      if (d.p_hid > 0.f) h *= drop(d, d.p_hid, order[i], KP_DROP_HIDDEN + c);
      H[(size_t)i * d.D + c] = h;
      s += h;
      ss += (double)h * h;
    }
  }
  red[0][ry][cx] = s;
  red[1][ry][cx] = ss;
  __syncthreads();
  if (ry == 0 && live) {
    for (int k = 1; k < 32; ++k) s += red[0][k][cx], ss += red[1][k][cx];
    float mean, var;
    if (d.train_bn) {
      bn_moments(s, ss, (double)d.B, mean, var);
      p.bn3_mean[c] = (1.f - BN_MOM) * p.bn3_mean[c] + BN_MOM * mean;
      p.bn3_var[c] = (1.f - BN_MOM) * p.bn3_var[c] + BN_MOM * (float)((double)var * d.B / (d.B - 1.0));
    } else {
      mean = p.bn3_mean[c];
      var = p.bn3_var[c];
    }
    const float istd = inv_std(var);
    st3[c] = mean;
    st3[d.D + c] = istd;
    bc[0][cx] = istd * p.bn3_w[c];
    bc[1][cx] = p.bn3_b[c] - mean * bc[0][cx];
  }
  __syncthreads();
  if (live) {
    const float sc = bc[0][cx], sh = bc[1][cx];
    for (int i = ry; i < d.B; i += 32) X[(size_t)i * d.D + c] = fmaxf(fmaf(H[(size_t)i * d.D + c], sc, sh), 0.f);
  }
}

// row i of Z: P = (sigmoid(z) - t) / (B N) in place, t = w * [j in positives(i)] + base (bce_optimizer.py:104-110);
// loss += sum_j (softplus(z) - t z) / (B N).  Padding columns [N, ldz) are cleared (they are GEMM operands next).
__global__ void __launch_bounds__(256) vfit_bce(int B, int N, long long ldz, float* __restrict__ Z, const int32_t* __restrict__ order,
                                                const int64_t* __restrict__ pos_off, const int32_t* __restrict__ pos_ids, float w, float base,
                                                float* __restrict__ loss) {
  __shared__ double red[16];
  const int i = blockIdx.x, tid = threadIdx.x, pid = order[i];
  float* z = Z + (size_t)i * ldz;
  const int64_t e0 = pos_off[pid], e1 = pos_off[pid + 1];
  const float inv = 1.f / ((float)B * (float)N);
  float lsf = 0.f;  // per-thread partial in fp32 (<= ldz / 256 terms), reduced in fp64
  if (loss)
    for (int64_t e = e0 + tid; e < e1; e += 256) lsf -= w * z[pos_ids[e]];
  __syncthreads();
  for (int j = tid; j < (int)ldz; j += 256) {
    if (j < N) {
      const float x = z[j];
      const float en = expf(-fabsf(x));             // exp(-|x|) in (0, 1]
      const float sig_abs = 1.f / (1.f + en);       // sigmoid(|x|)
      if (loss) lsf += fmaxf(x, 0.f) + log1pf(en) - base * x;
      z[j] = ((x >= 0.f ? sig_abs : en * sig_abs) - base) * inv;
    } else {
      z[j] = 0.f;
    }
  }
  __syncthreads();
  for (int64_t e = e0 + tid; e < e1; e += 256) z[pos_ids[e]] -= w * inv;
  if (loss) {
    double ls = (double)lsf, unused = 0.0;
    block_sum2(ls, unused, red);
    if (tid == 0) atomicAdd(loss, (float)(ls * (double)inv));
  }
}

// backward through relu / bn3 / hidden dropout; CTA = 32 columns x 32 row groups
__global__ void __launch_bounds__(1024) vfit_bn3_bwd(VDims d, kp_conve_params p, const int32_t* __restrict__ order, const float* __restrict__ H,
                                                    const float* __restrict__ X, const float* __restrict__ dX, const float* __restrict__ st3,
                                                    float* __restrict__ dH, float* __restrict__ g_w, float* __restrict__ g_b,
                                                    float* __restrict__ g_fcb) {
  __shared__ double red[2][32][32];
  __shared__ float bc[2][32];
  const int cx = threadIdx.x, ry = threadIdx.y, c = blockIdx.x * 32 + cx;
  const bool live = c < d.D;
  const float mean = live ? st3[c] : 0.f, istd = live ? st3[d.D + c] : 0.f;
  double a = 0.0, b = 0.0;
  if (live)
    for (int i = ry; i < d.B; i += 32) {
      const size_t o = (size_t)i * d.D + c;
      const float dy = X[o] > 0.f ? dX[o] : 0.f;
      a += dy;
      b += (double)dy * ((H[o] - mean) * istd);
    }
  red[0][ry][cx] = a;
  red[1][ry][cx] = b;
  __syncthreads();
  if (ry == 0 && live) {
    for (int k = 1; k < 32; ++k) a += red[0][k][cx], b += red[1][k][cx];
    g_b[c] = (float)a;
    g_w[c] = (float)b;
    bc[0][cx] = (float)(a / d.B);
    bc[1][cx] = (float)(b / d.B);
  }
  __syncthreads();
  double sb = 0.0;
  if (live) {
    const float am = d.train_bn ? bc[0][cx] : 0.f, bm = d.train_bn ? bc[1][cx] : 0.f, gs = p.bn3_w[c] * istd;
    for (int i = ry; i < d.B; i += 32) {
      const size_t o = (size_t)i * d.D + c;
      const float dy = X[o] > 0.f ? dX[o] : 0.f;
      float dh = gs * (dy - am - (H[o] - mean) * istd * bm);
      if (d.p_hid > 0.f) dh *= drop(d, d.p_hid, order[i], KP_DROP_HIDDEN + c);
      dH[o] = dh;
      sb += dh;
    }
  }
  __syncthreads();
  red[0][ry][cx] = sb;
  __syncthreads();
  if (ry == 0 && live) {
    for (int k = 1; k < 32; ++k) sb += red[0][k][cx];
    g_fcb[c] = (float)sb;
  }
}

// dy = dfeat through dropout2d and relu (in place); per-filter sums of dy and dy * c^ (c^ = normalised conv output)
__global__ void __launch_bounds__(256) vfit_bn2_bwd(VDims d, kp_conve_params p, const int32_t* __restrict__ order, const float* __restrict__ C,
                                                    const float* __restrict__ feat, float* __restrict__ dfeat, const double* __restrict__ S2,
                                                    double* __restrict__ B2) {
  const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pid = order[i];
  for (int f = warp; f < d.F; f += 8) {
    float mean, istd;
    bn2_stats(d, p, S2, f, mean, istd);
    const float ds = drop(d, d.p_fm, pid, KP_DROP_FEATURE + f);
    double a = 0.0, b = 0.0;
    for (int pos = lane; pos < d.FS; pos += 32) {
      const size_t o = (size_t)i * d.hidden + f * d.FS + pos;
      const float dy = feat[o] > 0.f ? dfeat[o] * ds : 0.f;
      dfeat[o] = dy;
      a += dy;
      b += (double)dy * ((C[o] - mean) * istd);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) {
      atomicAdd(B2 + 2 * f, a);
      atomicAdd(B2 + 2 * f + 1, b);
    }
  }
}

// dynamic smem: dC[hidden] | in[2D] | w[9F] | k[5F] (gamma*istd, mean, istd, mean dy, mean dy c^)
__global__ void __launch_bounds__(256) vfit_conv_bwd(VDims d, kp_conve_params p, const int32_t* __restrict__ order, const float* __restrict__ X0,
                                                     const float* __restrict__ C, const float* __restrict__ dy, const double* __restrict__ S1,
                                                     const double* __restrict__ S2, const double* __restrict__ B2, double* __restrict__ B1,
                                                     float* __restrict__ g_cw, float* __restrict__ g_cb, float* __restrict__ dX0) {
  extern __shared__ float sm[];
  __shared__ double red[16];
  float* dC = sm;
  float* in = dC + d.hidden;
  float* w = in + 2 * d.D;
  float* kf = w + 9 * d.F;
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, pid = order[i];
  float mean1, istd1;
  bn1_stats(d, p, S1, mean1, istd1);
  const float sc1 = istd1 * p.bn1_w[0], sh1 = p.bn1_b[0] - mean1 * sc1;
  for (int k = tid; k < 2 * d.D; k += 256) in[k] = (X0[(size_t)i * 2 * d.D + k] * sc1 + sh1) * drop(d, d.p_in, pid, KP_DROP_INPUT + k);
  for (int k = tid; k < 9 * d.F; k += 256) w[k] = p.conv_w[k];
  if (tid < d.F) {
    float mean, istd;
    bn2_stats(d, p, S2, tid, mean, istd);
    const double M = (double)d.B * d.FS;
    kf[tid] = p.bn2_w[tid] * istd;
    kf[d.F + tid] = mean;
    kf[2 * d.F + tid] = istd;
    kf[3 * d.F + tid] = d.train_bn ? (float)(B2[2 * tid] / M) : 0.f;
    kf[4 * d.F + tid] = d.train_bn ? (float)(B2[2 * tid + 1] / M) : 0.f;
  }
  __syncthreads();
  for (int idx = tid; idx < d.hidden; idx += 256) {
    const int f = idx / d.FS;
    const size_t o = (size_t)i * d.hidden + idx;
    const float ch = (C[o] - kf[d.F + f]) * kf[2 * d.F + f];
    dC[idx] = kf[f] * (dy[o] - kf[3 * d.F + f] - ch * kf[4 * d.F + f]);
  }
  __syncthreads();
  for (int f = warp; f < d.F; f += 8) {  // d conv_w[f, :], d conv_b[f]
    float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, accb = 0.f;
    for (int pos = lane; pos < d.FS; pos += 32) {
      const int y = pos / d.OW, x = pos - y * d.OW;
      const float g = dC[f * d.FS + pos];
      accb += g;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) acc[3 * ky + kx] = fmaf(g, in[(y + ky) * d.W + x + kx], acc[3 * ky + kx]);
    }
    accb = warp_sumf(accb);
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = warp_sumf(acc[k]);
    if (lane == 0) {
      atomicAdd(g_cb + f, accb);
#pragma unroll
      for (int k = 0; k < 9; ++k) atomicAdd(g_cw + 9 * f + k, acc[k]);
    }
  }
  double a = 0.0, b = 0.0;
  const int OH = 2 * 20 - 2;
  for (int q = tid; q < 2 * d.D; q += 256) {  // d in = conv-transpose of dC
    const int yy = q / d.W, xx = q - yy * d.W;
    float s = 0.f;
    for (int f = 0; f < d.F; ++f) {
      const float* dc = dC + f * d.FS;
      const float* wf = w + 9 * f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int y = yy - ky;
        if (y < 0 || y >= OH) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int x = xx - kx;
          if (x >= 0 && x < d.OW) s = fmaf(dc[y * d.OW + x], wf[3 * ky + kx], s);
        }
      }
    }
    s *= drop(d, d.p_in, pid, KP_DROP_INPUT + q);
    dX0[(size_t)i * 2 * d.D + q] = s;
    a += s;
    b += (double)s * ((X0[(size_t)i * 2 * d.D + q] - mean1) * istd1);
  }
  block_sum2(a, b, red);
  if (tid == 0) {
    atomicAdd(B1, a);
    atomicAdd(B1 + 1, b);
  }
}

// backward through bn1 and the gather; CTA 0 also writes d gamma / d beta of bn1 and bn2
__global__ void __launch_bounds__(128) vfit_scatter(VDims d, kp_conve_params p, const int32_t* __restrict__ pairs, const int32_t* __restrict__ order,
                                                    const float* __restrict__ X0, const float* __restrict__ dX0, const double* __restrict__ S1,
                                                    const double* __restrict__ B1, const double* __restrict__ B2, float* __restrict__ gE,
                                                    float* __restrict__ gR, float* __restrict__ g_bn1w, float* __restrict__ g_bn1b,
                                                    float* __restrict__ g_bn2w, float* __restrict__ g_bn2b) {
  const int i = blockIdx.x, pid = order[i];
  float mean, istd;
  bn1_stats(d, p, S1, mean, istd);
  const double M = (double)d.B * 2.0 * d.D;
  const float am = d.train_bn ? (float)(B1[0] / M) : 0.f, bm = d.train_bn ? (float)(B1[1] / M) : 0.f, gs = p.bn1_w[0] * istd;
  float* ge = gE + (size_t)pairs[2 * pid] * d.D;
  float* gr = gR + (size_t)pairs[2 * pid + 1] * d.D;
  for (int k = threadIdx.x; k < 2 * d.D; k += blockDim.x) {
    const size_t o = (size_t)i * 2 * d.D + k;
    const float g = gs * (dX0[o] - am - (X0[o] - mean) * istd * bm);
    atomicAdd(k < d.D ? ge + k : gr + (k - d.D), g);
  }
  if (i == 0) {
    if (threadIdx.x == 0) {
      g_bn1b[0] = (float)B1[0];
      g_bn1w[0] = (float)B1[1];
    }
    if (threadIdx.x < d.F) {
      g_bn2b[threadIdx.x] = (float)B2[2 * threadIdx.x];
      g_bn2w[threadIdx.x] = (float)B2[2 * threadIdx.x + 1];
    }
  }
}

// dHt[c, i] = dH[i, c], zero for i in [B, Bpad)
__global__ void vfit_transpose(int B, int Bpad, int D, const float* __restrict__ dH, float* __restrict__ dHt) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = i0 + r, c = c0 + threadIdx.x;
    t[r][threadIdx.x] = (i < B && c < D) ? dH[(size_t)i * D + c] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int c = c0 + r, i = i0 + threadIdx.x;
    if (c < D && i < Bpad) dHt[(size_t)c * Bpad + i] = t[threadIdx.x][r];
  }
}

// torch.optim.Adam (no weight decay / amsgrad): blockIdx.y = parameter tensor
__global__ void __launch_bounds__(256) vfit_adam(VSegs segs, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                 float step_size, float inv_sqrt_bias2, float beta1, float beta2, float eps) {
  const VSeg s = segs.s[blockIdx.y];
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < s.n; i += (long long)gridDim.x * 256) {
    const float gi = g[s.off + i];
    const float mi = beta1 * m[s.off + i] + (1.f - beta1) * gi, vi = beta2 * v[s.off + i] + (1.f - beta2) * gi * gi;
    m[s.off + i] = mi;
    v[s.off + i] = vi;
    s.p[i] -= step_size * (mi / (sqrtf(vi) * inv_sqrt_bias2 + eps));
  }
}

int vfit_fail(kp_vfit* f, int code, const char* msg) {
  if (f) f->err = msg; else g_vfit_error = msg;
  return code;
}

}  // namespace

extern "C" int kp_conve_fit_destroy(kp_vfit* f) {
  if (!f) return KP_OK;
  if (f->ctx) {
    cudaSetDevice(f->ctx->device);
    cudaDeviceSynchronize();
  }
  float* bufs[] = {f->g, f->m, f->v, f->X0, f->C, f->feat, f->H, f->X, f->Z, f->dX, f->dH, f->dfeat, f->dX0, f->st3, f->dHt};
  for (float* b : bufs) cudaFree(b);
  cudaFree(f->stats);
  if (f->ctx) kp_ctx_destroy(f->ctx);
  delete f;
  return KP_OK;
}

extern "C" int kp_conve_fit_create(int device, int64_t n_entities, int64_t n_relations2, int32_t dim, const kp_conve_params* params,
                                   float label_smoothing, int32_t max_batch, int64_t n_pairs, const int32_t* pairs, const int64_t* pos_off,
                                   const int32_t* pos_ids, uint64_t dropout_seed, kp_vfit** out) {
  if (!out) return KP_EINVAL;
  *out = nullptr;
  if (!params || n_entities <= 0 || n_relations2 <= 0 || max_batch <= 0 || n_pairs <= 0 || !pairs || !pos_off || !pos_ids)
    return vfit_fail(nullptr, KP_EINVAL, "kp_conve_fit_create: null / empty argument");
  const int W = dim / 20;
  if (dim <= 0 || dim % 20 != 0 || W < 3 || params->n_filters <= 0 || params->n_filters > 64 || params->n_filters % 4 != 0 ||
      params->hidden != params->n_filters * 38 * (W - 2))
    return vfit_fail(nullptr, KP_EINVAL, "kp_conve_fit_create: dim = 20 * w (w >= 3), n_filters <= 64 (multiple of 4), hidden = n_filters * 38 * (w - 2)");
  const float* ptrs[] = {params->ent, params->rel, params->conv_w, params->conv_b, params->fc_w, params->fc_b, params->bn1_w, params->bn1_b,
                         params->bn1_mean, params->bn1_var, params->bn2_w, params->bn2_b, params->bn2_mean, params->bn2_var, params->bn3_w,
                         params->bn3_b, params->bn3_mean, params->bn3_var};
  for (const float* q : ptrs)
    if (!q) return vfit_fail(nullptr, KP_EINVAL, "kp_conve_fit_create: every parameter tensor must be given (device pointers)");
  if ((reinterpret_cast<uintptr_t>(params->fc_w) & 15) != 0)
    return vfit_fail(nullptr, KP_EINVAL, "kp_conve_fit_create: fc_w must be 16-byte aligned");
  kp_vfit* f = new kp_vfit();
  int rc = kp_ctx_create(device, KP_TRANSE, n_entities, n_relations2, dim, 2, params->ent, params->rel, nullptr, &f->ctx);
  if (rc != KP_OK) {
    g_vfit_error = kp_last_error(nullptr);
    delete f;
    return rc;
  }
  f->p = *params;
  f->N = (int)n_entities;
  f->R2 = (int)n_relations2;
  f->D = dim;
  f->W = W;
  f->OW = W - 2;
  f->FS = 38 * (W - 2);
  f->F = params->n_filters;
  f->hidden = params->hidden;
  f->max_batch = max_batch;
  f->ldz = ((long long)n_entities + 3) / 4 * 4;
  f->ls = label_smoothing;
  f->seed = dropout_seed;
  f->n_pairs = n_pairs;
  f->pairs = pairs;
  f->pos_off = pos_off;
  f->pos_ids = pos_ids;
  if (const char* e = getenv("KP_VFIT_FP32")) f->fp32_mask = atoi(e);
  if (const char* e = getenv("KP_VFIT_TF32")) f->tf32_mask = atoi(e);
  const long long cnt[NSEG] = {(long long)n_entities * dim, (long long)n_relations2 * dim, 9ll * f->F, f->F, (long long)dim * f->hidden, dim, 1, 1,
                               f->F, f->F, dim, dim};
  float* par[NSEG] = {params->ent, params->rel, params->conv_w, params->conv_b, params->fc_w, params->fc_b, params->bn1_w, params->bn1_b,
                      params->bn2_w, params->bn2_b, params->bn3_w, params->bn3_b};
  long long off = 0;
  for (int s = 0; s < NSEG; ++s) {
    f->segs.s[s].p = par[s];
    f->segs.s[s].off = off;
    f->segs.s[s].n = cnt[s];
    off += (cnt[s] + 3) / 4 * 4;
  }
  f->n_params = off;
  const size_t pb = (size_t)off * 4, B = (size_t)max_batch;
  bool ok = cudaMalloc(&f->g, pb) == cudaSuccess && cudaMalloc(&f->m, pb) == cudaSuccess && cudaMalloc(&f->v, pb) == cudaSuccess &&
            cudaMalloc(&f->X0, B * 2 * dim * 4) == cudaSuccess && cudaMalloc(&f->dX0, B * 2 * dim * 4) == cudaSuccess &&
            cudaMalloc(&f->C, B * f->hidden * 4) == cudaSuccess && cudaMalloc(&f->feat, B * f->hidden * 4) == cudaSuccess &&
            cudaMalloc(&f->dfeat, B * f->hidden * 4) == cudaSuccess && cudaMalloc(&f->H, B * dim * 4) == cudaSuccess &&
            cudaMalloc(&f->X, B * dim * 4) == cudaSuccess && cudaMalloc(&f->dX, B * dim * 4) == cudaSuccess &&
            cudaMalloc(&f->dH, B * dim * 4) == cudaSuccess && cudaMalloc(&f->Z, B * f->ldz * 4) == cudaSuccess &&
            cudaMalloc(&f->st3, (size_t)2 * dim * 4) == cudaSuccess && cudaMalloc(&f->dHt, (size_t)dim * ((B + 3) / 4 * 4) * 4) == cudaSuccess && cudaMalloc(&f->stats, (size_t)(4 + 4 * f->F) * 8) == cudaSuccess;
  const size_t smem_bwd = ((size_t)f->hidden + 2 * dim + 14 * f->F) * 4;
  if (ok && smem_bwd > 200 * 1024) {
    kp_conve_fit_destroy(f);
    return vfit_fail(nullptr, KP_EUNSUPPORTED, "kp_conve_fit_create: hidden layer too wide for the shared-memory backward (dim <= 760)");
  }
  if (!ok || cudaFuncSetAttribute(vfit_conv_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bwd) != cudaSuccess) {
    cudaGetLastError();
    kp_conve_fit_destroy(f);
    return vfit_fail(nullptr, KP_ENOMEM, "kp_conve_fit_create: cannot allocate the gradient / activation buffers");
  }
  cudaMemset(f->g, 0, pb);
  cudaMemset(f->m, 0, pb);
  cudaMemset(f->v, 0, pb);
  *out = f;
  return KP_OK;
}

extern "C" const char* kp_conve_fit_error(const kp_vfit* f) { return f ? f->err.c_str() : g_vfit_error.c_str(); }
extern "C" int64_t kp_conve_fit_launches(const kp_vfit* f) { return f ? f->launches + (f->ctx ? f->ctx->launches : 0) : 0; }

extern "C" int kp_conve_fit_steps(kp_vfit* f, int64_t n_steps, const int64_t* step_off, const int32_t* order, float lr, float* loss_out,
                                  void* stream) {
  if (!f || n_steps < 0 || !step_off || !order) return KP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  kp_ctx* ctx = f->ctx;
  cudaSetDevice(ctx->device);
  if (n_steps > 0 && (step_off[0] < 0 || step_off[n_steps] > f->n_pairs))
    return vfit_fail(f, KP_EINVAL, "kp_conve_fit_steps: step offsets outside the pair table");
  if (loss_out && cudaMemsetAsync(loss_out, 0, (size_t)n_steps * sizeof(float), st) != cudaSuccess)
    return vfit_fail(f, KP_ECUDA, "kp_conve_fit_steps: cannot clear the loss buffer");
  const int N = f->N, D = f->D, F = f->F, hidden = f->hidden;
  const VSeg* sg = f->segs.s;
  float* g = f->g;
  float *gE = g + sg[SEG_ENT].off, *gR = g + sg[SEG_REL].off, *gCW = g + sg[SEG_CONV_W].off, *gCB = g + sg[SEG_CONV_B].off,
        *gW = g + sg[SEG_FC_W].off, *gFB = g + sg[SEG_FC_B].off;
  double *S1 = f->stats, *S2 = S1 + 2, *B2 = S2 + 2 * F, *B1 = B2 + 2 * F;
  const size_t smem_fwd = ((size_t)2 * D + 10 * F) * 4, smem_bwd = ((size_t)hidden + 2 * D + 14 * F) * 4;
  const float w = f->ls != 0.f ? 1.f - f->ls : 1.f, base = f->ls != 0.f ? 1.f / (float)N : 0.f;
  long long max_n = 0;
  for (int s = 0; s < NSEG; ++s) max_n = sg[s].n > max_n ? sg[s].n : max_n;
  int adam_blocks = (int)((max_n + 255) / 256);
  if (adam_blocks > ctx->sm_count * 8) adam_blocks = ctx->sm_count * 8;
  const dim3 colgrid((D + 31) / 32), colblock(32, 32);
  for (int64_t k = 0; k < n_steps; ++k) {
    const int B = (int)(step_off[k + 1] - step_off[k]);
    if (B <= 0) continue;
    if (B > f->max_batch) return vfit_fail(f, KP_EINVAL, "kp_conve_fit_steps: a step has more pairs than max_batch");
    const int32_t* ord = order + step_off[k];
    ++f->t;
    VDims d{B, D, f->W, f->OW, f->FS, F, hidden, B > 1 ? 1 : 0, (int)f->t, f->p.drop_input, f->p.drop_feature, f->p.drop_hidden, f->seed};
    int rc = KP_OK;
    bool ok = cudaMemsetAsync(f->stats, 0, (size_t)(4 + 4 * F) * 8, st) == cudaSuccess &&
              cudaMemsetAsync(gR, 0, (size_t)(sg[SEG_FC_W].off - sg[SEG_REL].off) * 4, st) == cudaSuccess;  // gR, d conv_w, d conv_b
    if (!ok) return vfit_fail(f, KP_ECUDA, "kp_conve_fit_steps: memset failed");
    vfit_gather<<<B, 128, 0, st>>>(d, f->p, f->pairs, ord, f->X0, S1);
    vfit_conv_fwd<<<B, 256, smem_fwd, st>>>(d, f->p, ord, f->X0, f->C, S1, S2);
    {
      long long blocks = ((long long)B * hidden + 255) / 256;
      if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
      vfit_bn2_relu<<<(int)blocks, 256, 0, st>>>(d, f->p, ord, f->C, f->feat, S1, S2);
    }
    rc = (f->fp32_mask & ~f->tf32_mask & 1) ? kp_sgemm(ctx, true, B, D, hidden, f->feat, hidden, f->p.fc_w, hidden, f->H, D, st, -1, true)
                                            : kp_gemm_umma_dyn(ctx, f->feat, hidden, false, B, f->p.fc_w, hidden, false, D, hidden, f->H, D, 0, st,
                                                               f->tf32_mask & 1);
    if (rc != KP_OK) goto gemm_failed;
    vfit_bn3_fwd<<<colgrid, colblock, 0, st>>>(d, f->p, ord, f->H, f->X, f->st3);
    rc = (f->fp32_mask & ~f->tf32_mask & 2) ? kp_sgemm(ctx, true, B, N, D, f->X, D, f->p.ent, D, f->Z, (int)f->ldz, st, -1, true)
                                            : kp_gemm_umma_dyn(ctx, f->X, D, false, B, f->p.ent, D, false, N, D, f->Z, f->ldz, 0, st, f->tf32_mask & 2);
    if (rc != KP_OK) goto gemm_failed;
    vfit_bce<<<B, 256, 0, st>>>(B, N, f->ldz, f->Z, ord, f->pos_off, f->pos_ids, w, base, loss_out ? loss_out + k : nullptr);
    rc = (f->fp32_mask & ~f->tf32_mask & 4) ? kp_sgemm(ctx, false, B, D, (int)f->ldz, f->Z, (int)f->ldz, f->p.ent, D, f->dX, D, st, N, true)
                                            : kp_gemm_umma_dyn(ctx, f->Z, f->ldz, false, B, f->p.ent, D, true, D, N, f->dX, D, 0, st, f->tf32_mask & 4);
    if (rc != KP_OK || (rc = kp_gemm_umma_dyn(ctx, f->Z, f->ldz, true, N, f->X, D, true, D, B, gE, D, 0, st, f->tf32_mask & 8)) != KP_OK) goto gemm_failed;
    vfit_bn3_bwd<<<colgrid, colblock, 0, st>>>(d, f->p, ord, f->H, f->X, f->dX, f->st3, f->dH, g + sg[SEG_BN3_W].off, g + sg[SEG_BN3_B].off, gFB);
    if (f->fp32_mask & ~f->tf32_mask & 16) {
      const int Bpad = (B + 3) / 4 * 4;
      vfit_transpose<<<dim3((D + 31) / 32, (Bpad + 31) / 32), dim3(32, 8), 0, st>>>(B, Bpad, D, f->dH, f->dHt);
      rc = kp_sgemm(ctx, false, D, hidden, Bpad, f->dHt, Bpad, f->feat, hidden, gW, hidden, st, B, true);
    } else {
      rc = kp_gemm_umma_dyn(ctx, f->dH, D, true, D, f->feat, hidden, true, hidden, B, gW, hidden, 0, st, f->tf32_mask & 16);
    }
    if (rc != KP_OK) goto gemm_failed;
    rc = (f->fp32_mask & ~f->tf32_mask & 32) ? kp_sgemm(ctx, false, B, hidden, D, f->dH, D, f->p.fc_w, hidden, f->dfeat, hidden, st, -1, true)
                                             : kp_gemm_umma_dyn(ctx, f->dH, D, false, B, f->p.fc_w, hidden, true, hidden, D, f->dfeat, hidden, 0, st,
                                                                f->tf32_mask & 32);
    if (rc != KP_OK) goto gemm_failed;
    vfit_bn2_bwd<<<B, 256, 0, st>>>(d, f->p, ord, f->C, f->feat, f->dfeat, S2, B2);
    vfit_conv_bwd<<<B, 256, smem_bwd, st>>>(d, f->p, ord, f->X0, f->C, f->dfeat, S1, S2, B2, B1, gCW, gCB, f->dX0);
    vfit_scatter<<<B, 128, 0, st>>>(d, f->p, f->pairs, ord, f->X0, f->dX0, S1, B1, B2, gE, gR, g + sg[SEG_BN1_W].off, g + sg[SEG_BN1_B].off,
                                    g + sg[SEG_BN2_W].off, g + sg[SEG_BN2_B].off);
    {
      const double bias1 = 1.0 - pow((double)f->beta1, (double)f->t), bias2 = 1.0 - pow((double)f->beta2, (double)f->t);
      vfit_adam<<<dim3(adam_blocks, NSEG), 256, 0, st>>>(f->segs, f->g, f->m, f->v, (float)((double)lr / bias1), (float)(1.0 / sqrt(bias2)),
                                                         f->beta1, f->beta2, f->eps);
    }
    f->launches += 10;
    continue;
  gemm_failed:
    f->err = kp_last_error(ctx);
    return rc;
  }
  if (cudaGetLastError() != cudaSuccess) return vfit_fail(f, KP_ECUDA, "kp_conve_fit_steps: kernel launch failed");
  return KP_OK;
}
