// Full-model TransE training for verify_explanations' retrain-from-scratch:
// PairwiseRankingOptimizer.step_on_batch (pairwise_ranking_optimizer.py:139-157) with
// TransE.forward (transe.py:67-75), L2 (regularizers.py:19-22) and torch.optim.Adam over BOTH
// embedding tables.  One step = two launches:
//
//   fit_grad  one warp per (positive, negative) row pair: gathers the six rows with 128-bit loads,
//             warp-shuffle norms, margin test, and accumulates the gradient of the six rows into dense
//             gradient tables with vector reductions (red.global.add.v4.f32)
//   fit_adam  fused dense Adam over the entity and relation tables (torch semantics: every row moves
//             every step once it has momentum, so the update is dense), clears the gradient it read
//
// loss = mean_i max(0, pos_i - neg_i + margin) + 1/2 (L2(pos factors) + L2(neg factors)),
// L2(f1..f3) = w/3 sum_k mean(f_k^2)  =>  every occurrence of a row as a factor adds w/(3 B D) row to its
// gradient.  Index tables (shuffles, corruptions) are drawn on the host in the reference's
// order; the kernels hold no random number generator.
#include "kp_internal.h"

struct kp_fit {
  int device = 0;
  int N = 0, R2 = 0, D = 0, norm = 2;
  float lr = 0, margin = 0, reg = 0, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f;
  float *ent = nullptr, *rel = nullptr;  // borrowed device tables, updated in place
  float *g = nullptr, *m = nullptr, *v = nullptr;  // [(N + R2) * D] each, entity rows first
  long long t = 0;                                   // Adam step counter
  int64_t launches = 0;
  std::string err;
};

namespace {

std::string g_fit_error;

__device__ __forceinline__ void red_add_v4(float* addr, float4 x) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

struct FitG {
  int B, N, D, norm;
  float margin, reg_coef;  // reg_coef = w / (3 B D)
  const float* ent;
  const float* rel;
  const int32_t* pos;
  const int32_t* neg;
  float* g_ent;
  float* g_rel;
  float* loss;  // nullable, one float
};

// V = float4 vectors per lane (D <= 128 * V, D a multiple of 4)
template <int V>
__global__ void __launch_bounds__(256) fit_grad_kernel(const FitG p) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= p.B) return;
  const int ps = p.pos[3 * i], pr = p.pos[3 * i + 1], po = p.pos[3 * i + 2];
  const int ns = p.neg[3 * i], nr = p.neg[3 * i + 1], no = p.neg[3 * i + 2];
  const float4* Es = reinterpret_cast<const float4*>(p.ent + (size_t)ps * p.D);
  const float4* Eo = reinterpret_cast<const float4*>(p.ent + (size_t)po * p.D);
  const float4* Rp = reinterpret_cast<const float4*>(p.rel + (size_t)pr * p.D);
  const float4* Ens = reinterpret_cast<const float4*>(p.ent + (size_t)ns * p.D);
  const float4* Eno = reinterpret_cast<const float4*>(p.ent + (size_t)no * p.D);
  const float4* Rn = reinterpret_cast<const float4*>(p.rel + (size_t)nr * p.D);
  float4 s[V], r[V], o[V], s2[V], r2[V], o2[V], d[V], d2[V];
  float ap = 0.f, an = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int c = lane + 32 * k;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool in = c * 4 < p.D;
    s[k] = in ? Es[c] : z; r[k] = in ? Rp[c] : z; o[k] = in ? Eo[c] : z;
    s2[k] = in ? Ens[c] : z; r2[k] = in ? Rn[c] : z; o2[k] = in ? Eno[c] : z;
    d[k] = make_float4(s[k].x + r[k].x - o[k].x, s[k].y + r[k].y - o[k].y, s[k].z + r[k].z - o[k].z, s[k].w + r[k].w - o[k].w);
    d2[k] = make_float4(s2[k].x + r2[k].x - o2[k].x, s2[k].y + r2[k].y - o2[k].y, s2[k].z + r2[k].z - o2[k].z,
                        s2[k].w + r2[k].w - o2[k].w);
    if (p.norm == 2) {
      ap += d[k].x * d[k].x + d[k].y * d[k].y + d[k].z * d[k].z + d[k].w * d[k].w;
      an += d2[k].x * d2[k].x + d2[k].y * d2[k].y + d2[k].z * d2[k].z + d2[k].w * d2[k].w;
    } else {
      ap += fabsf(d[k].x) + fabsf(d[k].y) + fabsf(d[k].z) + fabsf(d[k].w);
      an += fabsf(d2[k].x) + fabsf(d2[k].y) + fabsf(d2[k].z) + fabsf(d2[k].w);
    }
  }
  ap = warp_sum(ap);
  an = warp_sum(an);
  if (p.norm == 2) {
    ap = sqrtf(ap);
    an = sqrtf(an);
  }
  const float viol = ap - an + p.margin;
  const bool active = viol > 0.f;
  if (p.loss && lane == 0 && active) atomicAdd(p.loss, viol / (float)p.B);
  // d loss / d pos_i = 1/B, d loss / d neg_i = -1/B when active; d ||x||_2 / dx = x / ||x|| (0 at x = 0), d ||x||_1 / dx = sign(x)
  const float cp = active ? (p.norm == 2 ? (ap > 0.f ? 1.f / ((float)p.B * ap) : 0.f) : 1.f / (float)p.B) : 0.f;
  const float cn = active ? (p.norm == 2 ? (an > 0.f ? 1.f / ((float)p.B * an) : 0.f) : 1.f / (float)p.B) : 0.f;
  const float rc = p.reg_coef;
  auto dir = [&](float x) { return p.norm == 2 ? x : (x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f)); };
  float* gs = p.g_ent + (size_t)ps * p.D;
  float* go = p.g_ent + (size_t)po * p.D;
  float* gr = p.g_rel + (size_t)pr * p.D;
  float* gs2 = p.g_ent + (size_t)ns * p.D;
  float* go2 = p.g_ent + (size_t)no * p.D;
  float* gr2 = p.g_rel + (size_t)nr * p.D;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int c = (lane + 32 * k) * 4;
    if (c >= p.D) continue;
    const float4 u = make_float4(cp * dir(d[k].x), cp * dir(d[k].y), cp * dir(d[k].z), cp * dir(d[k].w));
    const float4 w = make_float4(cn * dir(d2[k].x), cn * dir(d2[k].y), cn * dir(d2[k].z), cn * dir(d2[k].w));
    red_add_v4(gs + c, make_float4(u.x + rc * s[k].x, u.y + rc * s[k].y, u.z + rc * s[k].z, u.w + rc * s[k].w));
    red_add_v4(gr + c, make_float4(u.x + rc * r[k].x, u.y + rc * r[k].y, u.z + rc * r[k].z, u.w + rc * r[k].w));
    red_add_v4(go + c, make_float4(-u.x + rc * o[k].x, -u.y + rc * o[k].y, -u.z + rc * o[k].z, -u.w + rc * o[k].w));
    red_add_v4(gs2 + c, make_float4(-w.x + rc * s2[k].x, -w.y + rc * s2[k].y, -w.z + rc * s2[k].z, -w.w + rc * s2[k].w));
    red_add_v4(gr2 + c, make_float4(-w.x + rc * r2[k].x, -w.y + rc * r2[k].y, -w.z + rc * r2[k].z, -w.w + rc * r2[k].w));
    red_add_v4(go2 + c, make_float4(w.x + rc * o2[k].x, w.y + rc * o2[k].y, w.z + rc * o2[k].z, w.w + rc * o2[k].w));
  }
}

// torch.optim.Adam (no weight decay, no amsgrad) over n4 float4 of parameters split across two tables
__global__ void fit_adam_kernel(long long n4_ent, long long n4_all, float4* __restrict__ ent, float4* __restrict__ rel,
                                float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v, float step_size,
                                float inv_sqrt_bias2, float beta1, float beta2, float eps) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4_all; i += (long long)gridDim.x * blockDim.x) {
    float4* pp = i < n4_ent ? ent + i : rel + (i - n4_ent);
    const float4 gi = g[i];
    g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 mi = m[i], vi = v[i], p = *pp;
#define KP_ADAM(c)                                               \
  mi.c = beta1 * mi.c + (1.f - beta1) * gi.c;                    \
  vi.c = beta2 * vi.c + (1.f - beta2) * gi.c * gi.c;             \
  p.c -= step_size * (mi.c / (sqrtf(vi.c) * inv_sqrt_bias2 + eps));
    KP_ADAM(x) KP_ADAM(y) KP_ADAM(z) KP_ADAM(w)
#undef KP_ADAM
    m[i] = mi;
    v[i] = vi;
    *pp = p;
  }
}

int fit_fail(kp_fit* f, int code, const char* msg) {
  if (f) f->err = msg; else g_fit_error = msg;
  return code;
}

}  // namespace

extern "C" int kp_transe_fit_create(int device, int64_t n_entities, int64_t n_relations2, int32_t dim, int32_t norm, float lr,
                                    float margin, float reg_weight, float* ent, float* rel, kp_fit** out) {
  if (!out) return KP_EINVAL;
  *out = nullptr;
  if (n_entities <= 0 || n_relations2 <= 0 || dim <= 0 || dim % 4 != 0 || dim > 512 || (norm != 1 && norm != 2) || !ent || !rel)
    return fit_fail(nullptr, KP_EUNSUPPORTED, "kp_transe_fit_create: dim must be a multiple of 4 (<= 512), norm 1 or 2, device tables");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cudaGetLastError();
    return fit_fail(nullptr, KP_ECUDA, "kp_transe_fit_create: no such CUDA device (kelpie_b200 has no CPU fallback)");
  }
  cudaSetDevice(device);
  kp_fit* f = new kp_fit();
  f->device = device;
  f->N = (int)n_entities;
  f->R2 = (int)n_relations2;
  f->D = dim;
  f->norm = norm;
  f->lr = lr;
  f->margin = margin;
  f->reg = reg_weight;
  f->ent = ent;
  f->rel = rel;
  const size_t bytes = (size_t)(n_entities + n_relations2) * dim * sizeof(float);
  if (cudaMalloc(&f->g, bytes) != cudaSuccess || cudaMalloc(&f->m, bytes) != cudaSuccess || cudaMalloc(&f->v, bytes) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(f->g);
    cudaFree(f->m);
    delete f;
    return fit_fail(nullptr, KP_ENOMEM, "kp_transe_fit_create: cannot allocate the optimiser state");
  }
  cudaMemset(f->g, 0, bytes);
  cudaMemset(f->m, 0, bytes);
  cudaMemset(f->v, 0, bytes);
  *out = f;
  return KP_OK;
}

extern "C" int kp_transe_fit_destroy(kp_fit* f) {
  if (!f) return KP_OK;
  cudaSetDevice(f->device);
  cudaFree(f->g);
  cudaFree(f->m);
  cudaFree(f->v);
  delete f;
  return KP_OK;
}

extern "C" const char* kp_transe_fit_error(const kp_fit* f) { return f ? f->err.c_str() : g_fit_error.c_str(); }
extern "C" int64_t kp_transe_fit_launches(const kp_fit* f) { return f ? f->launches : 0; }

// n_steps consecutive steps; step k uses rows [step_off[k], step_off[k+1]) of pos / neg (device, [rows, 3]).
// loss_out (device, nullable): [n_steps] fitting loss of every step (without the regulariser).
extern "C" int kp_transe_fit_steps(kp_fit* f, int64_t n_steps, const int64_t* step_off, const int32_t* pos, const int32_t* neg,
                                   float* loss_out, void* stream) {
  if (!f || n_steps < 0 || !step_off || !pos || !neg) return KP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  cudaSetDevice(f->device);
  if (loss_out && cudaMemsetAsync(loss_out, 0, (size_t)n_steps * sizeof(float), st) != cudaSuccess)
    return fit_fail(f, KP_ECUDA, "kp_transe_fit_steps: cannot clear the loss buffer");
  const long long n4_ent = (long long)f->N * f->D / 4, n4_all = (long long)(f->N + f->R2) * f->D / 4;
  for (int64_t k = 0; k < n_steps; ++k) {
    const int64_t b0 = step_off[k], b1 = step_off[k + 1];
    const int B = (int)(b1 - b0);
    if (B <= 0) continue;
    FitG p;
    p.B = B;
    p.N = f->N;
    p.D = f->D;
    p.norm = f->norm;
    p.margin = f->margin;
    p.reg_coef = f->reg / (3.f * (float)B * (float)f->D);
    p.ent = f->ent;
    p.rel = f->rel;
    p.pos = pos + 3 * b0;
    p.neg = neg + 3 * b0;
    p.g_ent = f->g;
    p.g_rel = f->g + (size_t)f->N * f->D;
    p.loss = loss_out ? loss_out + k : nullptr;
    const int blocks = (B + 7) / 8;
    switch ((f->D + 127) / 128) {
      case 1: fit_grad_kernel<1><<<blocks, 256, 0, st>>>(p); break;
      case 2: fit_grad_kernel<2><<<blocks, 256, 0, st>>>(p); break;
      case 3: fit_grad_kernel<3><<<blocks, 256, 0, st>>>(p); break;
      default: fit_grad_kernel<4><<<blocks, 256, 0, st>>>(p); break;
    }
    ++f->t;
    const double bias1 = 1.0 - pow((double)f->beta1, (double)f->t), bias2 = 1.0 - pow((double)f->beta2, (double)f->t);
    int ablocks = (int)((n4_all + 255) / 256);
    if (ablocks > 148 * 16) ablocks = 148 * 16;
    fit_adam_kernel<<<ablocks, 256, 0, st>>>(n4_ent, n4_all, reinterpret_cast<float4*>(f->ent), reinterpret_cast<float4*>(f->rel),
                                            reinterpret_cast<float4*>(f->g), reinterpret_cast<float4*>(f->m),
                                            reinterpret_cast<float4*>(f->v), (float)(f->lr / bias1), (float)(1.0 / sqrt(bias2)),
                                            f->beta1, f->beta2, f->eps);
    f->launches += 2;
  }
  if (cudaGetLastError() != cudaSuccess) return fit_fail(f, KP_ECUDA, "kp_transe_fit_steps: kernel launch failed");
  return KP_OK;
}
