// Full-model ComplEx training for verify_explanations' retrain-from-scratch:
// MultiClassNLLOptimizer.step_on_batch (multiclass_nll_optimizer.py:123-135) with ComplEx.forward
// (complex.py:58-86: 1-vs-all logits against the whole entity table), CrossEntropyLoss(mean) and
// Adagrad / Adam / SGD over BOTH embedding tables.  The regulariser weight is 0 in every shipped config
// (N3 / N2 contribute nothing) and a non-zero weight is rejected.
//
// One step of B rows (s, r, o), N entities, D = 2d floats per row (tables are updated in place):
//   cfit_queries   q_i = E[s_i] (complex *) R[r_i]                                    warp per row
//   GEMM           Z[B, N] = Q E^T                            tcgen05, bf16x3 split   (kp_gemm_umma.cu)
//   cfit_softmax   P = (softmax(Z) - onehot(o)) / B  in place, loss                    CTA per row
//   GEMM           dQ[B, D] = P E                             K = N, split-K
//   GEMM           gE[N, D] = P^T Q                           the dense table gradient through column j
//   cfit_scatter   gE[s_i] += dQ_i (*) conj R[r_i],  gR[r_i] += dQ_i (*) conj E[s_i]    warp per row, vector reductions
//   cfit_update    dense optimiser step over E and R (torch semantics), clears nothing (gE is overwritten, gR zeroed)
// At DBpedia50 shape the logits are 512 x 24 620 fp32 = 50 MB, so they are simply materialised in HBM.
#include "kp_internal.h"

struct kp_cfit {
  kp_ctx* ctx = nullptr;  // tables borrowed; used for the GEMM plumbing (workspace arenas, error text)
  int N = 0, R2 = 0, D = 0, optimizer = 0, max_batch = 0;
  long long ldz = 0;
  float lr = 0, beta1 = 0.9f, beta2 = 0.999f, eps = 0;
  float *ent = nullptr, *rel = nullptr;
  float *g = nullptr, *s1 = nullptr, *s2 = nullptr;  // [(N + R2) * D]: gradient, optimiser state (Adagrad: s1 = sum of squares)
  float *Q = nullptr, *dQ = nullptr, *Z = nullptr;
  long long t = 0;
  int64_t launches = 0;
  std::string err;
};

namespace {

std::string g_cfit_error;
enum { OPT_ADAGRAD = 0, OPT_ADAM = 1, OPT_SGD = 2 };

__device__ __forceinline__ void red_add_v4(float* addr, float4 x) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
}

// q_i = l (*) rho:  re = l_re rho_re - l_im rho_im,  im = l_re rho_im + l_im rho_re   (complex.py:74-75)
__global__ void cfit_queries(int B, int D, const float* __restrict__ ent, const float* __restrict__ rel, const int32_t* __restrict__ rows,
                             float* __restrict__ Q) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= B) return;
  const int d = D >> 1;
  const float* l = ent + (size_t)rows[3 * i] * D;
  const float* r = rel + (size_t)rows[3 * i + 1] * D;
  for (int k = lane; k < d; k += 32) {
    const float lr = l[k], li = l[d + k], rr = r[k], ri = r[d + k];
    Q[(size_t)i * D + k] = lr * rr - li * ri;
    Q[(size_t)i * D + d + k] = lr * ri + li * rr;
  }
}

// row i of Z: P = (softmax - onehot(o_i)) / B in place; loss += (logsumexp - z_o) / B
__global__ void __launch_bounds__(256) cfit_softmax(int B, int N, long long ldz, float* __restrict__ Z, const int32_t* __restrict__ rows,
                                                    float* __restrict__ loss) {
  __shared__ float red[8];
  __shared__ float bc;
  const int i = blockIdx.x, tid = threadIdx.x;
  float* z = Z + (size_t)i * ldz;
  float m = -INFINITY;
  for (int j = tid; j < N; j += 256) m = fmaxf(m, z[j]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  if (tid == 0) {
    float x = red[0];
    for (int w = 1; w < 8; ++w) x = fmaxf(x, red[w]);
    bc = x;
  }
  __syncthreads();
  m = bc;
  float s = 0.f;
  for (int j = tid; j < N; j += 256) s += expf(z[j] - m);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float x = 0.f;
    for (int w = 0; w < 8; ++w) x += red[w];
    bc = x;
  }
  __syncthreads();
  s = bc;
  const int o = rows[3 * i + 2];
  const float inv = 1.f / (s * (float)B), ib = 1.f / (float)B;
  if (tid == 0 && loss) atomicAdd(loss, (m + logf(s) - z[o]) * ib);
  __syncthreads();
  for (int j = tid; j < N; j += 256) z[j] = expf(z[j] - m) * inv - (j == o ? ib : 0.f);
}

// gradient through the query: dl = dq (*) conj(rho), drho = dq (*) conj(l)   
__global__ void cfit_scatter(int B, int D, const float* __restrict__ ent, const float* __restrict__ rel, const int32_t* __restrict__ rows,
                             const float* __restrict__ dQ, float* __restrict__ gE, float* __restrict__ gR) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= B) return;
  const int d = D >> 1;
  const int s = rows[3 * i], r = rows[3 * i + 1];
  const float* l = ent + (size_t)s * D;
  const float* rho = rel + (size_t)r * D;
  const float* dq = dQ + (size_t)i * D;
  for (int k = lane * 4; k < d; k += 128) {  // d is a multiple of 4 when D is a multiple of 8; the tail is handled below
    if (k + 3 < d) {
      const float4 a = *reinterpret_cast<const float4*>(dq + k), b = *reinterpret_cast<const float4*>(dq + d + k);
      const float4 lr = *reinterpret_cast<const float4*>(l + k), li = *reinterpret_cast<const float4*>(l + d + k);
      const float4 rr = *reinterpret_cast<const float4*>(rho + k), ri = *reinterpret_cast<const float4*>(rho + d + k);
      red_add_v4(gE + (size_t)s * D + k, make_float4(a.x * rr.x + b.x * ri.x, a.y * rr.y + b.y * ri.y, a.z * rr.z + b.z * ri.z, a.w * rr.w + b.w * ri.w));
      red_add_v4(gE + (size_t)s * D + d + k, make_float4(-a.x * ri.x + b.x * rr.x, -a.y * ri.y + b.y * rr.y, -a.z * ri.z + b.z * rr.z, -a.w * ri.w + b.w * rr.w));
      red_add_v4(gR + (size_t)r * D + k, make_float4(a.x * lr.x + b.x * li.x, a.y * lr.y + b.y * li.y, a.z * lr.z + b.z * li.z, a.w * lr.w + b.w * li.w));
      red_add_v4(gR + (size_t)r * D + d + k, make_float4(-a.x * li.x + b.x * lr.x, -a.y * li.y + b.y * lr.y, -a.z * li.z + b.z * lr.z, -a.w * li.w + b.w * lr.w));
    }
  }
}

__global__ void cfit_update(long long n_ent, long long n_all, int optimizer, float* __restrict__ ent, float* __restrict__ rel,
                            const float* __restrict__ g, float* __restrict__ s1, float* __restrict__ s2, float lr, float step_size,
                            float inv_sqrt_bias2, float beta1, float beta2, float eps) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_all; i += (long long)gridDim.x * blockDim.x) {
    float* p = i < n_ent ? ent + i : rel + (i - n_ent);
    const float gi = g[i];
    if (optimizer == OPT_ADAGRAD) {  // torch.optim.Adagrad: state_sum += g^2; p -= lr g / (sqrt(state_sum) + 1e-10)
      const float a = s1[i] + gi * gi;
      s1[i] = a;
      *p -= lr * gi / (sqrtf(a) + eps);
    } else if (optimizer == OPT_ADAM) {
      const float m = beta1 * s1[i] + (1.f - beta1) * gi, v = beta2 * s2[i] + (1.f - beta2) * gi * gi;
      s1[i] = m;
      s2[i] = v;
      *p -= step_size * (m / (sqrtf(v) * inv_sqrt_bias2 + eps));
    } else {
      *p -= lr * gi;
    }
  }
}

int cfit_fail(kp_cfit* f, int code, const char* msg) {
  if (f) f->err = msg; else g_cfit_error = msg;
  return code;
}

}  // namespace

extern "C" int kp_complex_fit_destroy(kp_cfit* f) {
  if (!f) return KP_OK;
  if (f->ctx) {
    cudaSetDevice(f->ctx->device);
    cudaDeviceSynchronize();
  }
  cudaFree(f->g);
  cudaFree(f->s1);
  cudaFree(f->s2);
  cudaFree(f->Q);
  cudaFree(f->dQ);
  cudaFree(f->Z);
  if (f->ctx) kp_ctx_destroy(f->ctx);
  delete f;
  return KP_OK;
}

// optimizer: 0 Adagrad (eps 1e-10), 1 Adam (betas decay1 / decay2, eps 1e-8), 2 SGD.  ent / rel: device tables, updated in place.
extern "C" int kp_complex_fit_create(int device, int64_t n_entities, int64_t n_relations2, int32_t dim, int32_t optimizer, float lr,
                                     float beta1, float beta2, float reg_weight, int32_t max_batch, float* ent, float* rel,
                                     kp_cfit** out) {
  if (!out) return KP_EINVAL;
  *out = nullptr;
  if (n_entities <= 0 || n_relations2 <= 0 || dim <= 0 || dim % 8 != 0 || max_batch <= 0 || optimizer < 0 || optimizer > 2 || !ent || !rel)
    return cfit_fail(nullptr, KP_EINVAL, "kp_complex_fit_create: dim must be a multiple of 8, optimizer 0..2, device tables");
  if (reg_weight != 0.f)
    return cfit_fail(nullptr, KP_EUNSUPPORTED, "kp_complex_fit_create: a non-zero regulariser weight is not supported (every shipped config uses 0)");
  kp_cfit* f = new kp_cfit();
  int rc = kp_ctx_create(device, KP_COMPLEX, n_entities, n_relations2, dim, 2, ent, rel, nullptr, &f->ctx);
  if (rc != KP_OK) {
    g_cfit_error = kp_last_error(nullptr);
    delete f;
    return rc;
  }
  f->N = (int)n_entities;
  f->R2 = (int)n_relations2;
  f->D = dim;
  f->optimizer = optimizer;
  f->max_batch = max_batch;
  f->ldz = ((long long)n_entities + 3) / 4 * 4;
  f->lr = lr;
  f->beta1 = beta1;
  f->beta2 = beta2;
  f->eps = optimizer == OPT_ADAGRAD ? 1e-10f : 1e-8f;
  f->ent = ent;
  f->rel = rel;
  const size_t pbytes = (size_t)(n_entities + n_relations2) * dim * sizeof(float);
  bool ok = cudaMalloc(&f->g, pbytes) == cudaSuccess && cudaMalloc(&f->s1, pbytes) == cudaSuccess &&
            (optimizer != OPT_ADAM || cudaMalloc(&f->s2, pbytes) == cudaSuccess) &&
            cudaMalloc(&f->Q, (size_t)max_batch * dim * 4) == cudaSuccess && cudaMalloc(&f->dQ, (size_t)max_batch * dim * 4) == cudaSuccess &&
            cudaMalloc(&f->Z, (size_t)max_batch * f->ldz * 4) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    kp_complex_fit_destroy(f);
    return cfit_fail(nullptr, KP_ENOMEM, "kp_complex_fit_create: cannot allocate the gradient / logit buffers");
  }
  cudaMemset(f->s1, 0, pbytes);
  if (f->s2) cudaMemset(f->s2, 0, pbytes);
  *out = f;
  return KP_OK;
}

extern "C" const char* kp_complex_fit_error(const kp_cfit* f) { return f ? f->err.c_str() : g_cfit_error.c_str(); }
extern "C" int64_t kp_complex_fit_launches(const kp_cfit* f) { return f ? f->launches + (f->ctx ? f->ctx->launches : 0) : 0; }

// n_steps consecutive steps; step k trains on rows [step_off[k], step_off[k+1]) of `rows` ([total, 3] int32, device).
// loss_out (device, nullable): [n_steps] mean cross-entropy of every step.
extern "C" int kp_complex_fit_steps(kp_cfit* f, int64_t n_steps, const int64_t* step_off, const int32_t* rows, float* loss_out,
                                    void* stream) {
  if (!f || n_steps < 0 || !step_off || !rows) return KP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  kp_ctx* ctx = f->ctx;
  cudaSetDevice(ctx->device);
  const int N = f->N, D = f->D;
  float* gE = f->g;
  float* gR = f->g + (size_t)N * D;
  if (loss_out && cudaMemsetAsync(loss_out, 0, (size_t)n_steps * sizeof(float), st) != cudaSuccess)
    return cfit_fail(f, KP_ECUDA, "kp_complex_fit_steps: cannot clear the loss buffer");
  const long long n_ent = (long long)N * D, n_all = (long long)(N + f->R2) * D;
  for (int64_t k = 0; k < n_steps; ++k) {
    const int B = (int)(step_off[k + 1] - step_off[k]);
    if (B <= 0) continue;
    if (B > f->max_batch) return cfit_fail(f, KP_EINVAL, "kp_complex_fit_steps: a step has more rows than max_batch");
    const int32_t* r = rows + 3 * step_off[k];
    int rc;
    cfit_queries<<<(B + 7) / 8, 256, 0, st>>>(B, D, f->ent, f->rel, r, f->Q);
    if ((rc = kp_gemm_umma_dyn(ctx, f->Q, D, false, B, f->ent, D, false, N, D, f->Z, f->ldz, 0, st)) != KP_OK) {
      f->err = kp_last_error(ctx);
      return rc;
    }
    cfit_softmax<<<B, 256, 0, st>>>(B, N, f->ldz, f->Z, r, loss_out ? loss_out + k : nullptr);
    if ((rc = kp_gemm_umma_dyn(ctx, f->Z, f->ldz, false, B, f->ent, D, true, D, N, f->dQ, D, 0, st)) != KP_OK ||
        (rc = kp_gemm_umma_dyn(ctx, f->Z, f->ldz, true, N, f->Q, D, true, D, B, gE, D, 0, st)) != KP_OK) {
      f->err = kp_last_error(ctx);
      return rc;
    }
    if (cudaMemsetAsync(gR, 0, (size_t)f->R2 * D * 4, st) != cudaSuccess) return cfit_fail(f, KP_ECUDA, "kp_complex_fit_steps: memset failed");
    cfit_scatter<<<(B + 7) / 8, 256, 0, st>>>(B, D, f->ent, f->rel, r, f->dQ, gE, gR);
    ++f->t;
    const double bias1 = 1.0 - pow((double)f->beta1, (double)f->t), bias2 = 1.0 - pow((double)f->beta2, (double)f->t);
    int blocks = (int)((n_all + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    cfit_update<<<blocks, 256, 0, st>>>(n_ent, n_all, f->optimizer, f->ent, f->rel, f->g, f->s1, f->s2, f->lr, (float)(f->lr / bias1),
                                        (float)(1.0 / sqrt(bias2)), f->beta1, f->beta2, f->eps);
    f->launches += 4;
  }
  if (cudaGetLastError() != cudaSuccess) return cfit_fail(f, KP_ECUDA, "kp_complex_fit_steps: kernel launch failed");
  return KP_OK;
}
