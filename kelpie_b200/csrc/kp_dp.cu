// Data-poisoning relevance (src/relevance_engines/data_poisoning_engine.py:21-141) for ComplEx -- the one
// in-scope model that defines the score_embeddings() the reference engine calls (complex.py:47-56).
// One warp per job (prediction <ps, pp, po>, training fact <s, p, o>, perspective entity e in {ps, po}):
//   g   = d score(pred) / d E[e]                       get_gradient (:21-50), closed form of the autograd call
//   e'  = E[e] -/+ eps g                               necessary: worsen the score (:66-70), sufficient: improve it (:104-108)
//   out = +-(score(fact) - lambda * score(fact with E[e] := e'))        (:84-94, :122-131)
// ComplEx is a maximiser, so necessary = orig - lambda * perturbed, sufficient = -orig + lambda * perturbed.
#include "kp_internal.h"

namespace {

__global__ void __launch_bounds__(256) dp_kernel(int n, int D, const float* __restrict__ ent, const float* __restrict__ rel,
                                                 const int32_t* __restrict__ preds, const int32_t* __restrict__ facts,
                                                 const int32_t* __restrict__ entity, float eps, float lambd, int sufficient,
                                                 float* __restrict__ out) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= n) return;
  const int d = D >> 1;
  const int ps = preds[3 * j], pp = preds[3 * j + 1], po = preds[3 * j + 2];
  const int fs = facts[3 * j], fp = facts[3 * j + 1], fo = facts[3 * j + 2];
  const int e = entity[j];
  const bool grad_lhs = e == ps;   // entity_embedding = lhs if entity == s else rhs (:40)
  const bool swap_lhs = fs == e;   // if s == entity: lhs[1] = perturbed else rhs[1] = perturbed (:78-81)
  const float* L = ent + (size_t)ps * D;
  const float* R = rel + (size_t)pp * D;
  const float* O = ent + (size_t)po * D;
  const float* Ee = ent + (size_t)e * D;
  const float* FL = ent + (size_t)fs * D;
  const float* FR = rel + (size_t)fp * D;
  const float* FO = ent + (size_t)fo * D;
  const float sgn = sufficient ? eps : -eps;
  float orig = 0.f, pert = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float lr = L[k], li = L[d + k], rr = R[k], ri = R[d + k], orr = O[k], oi = O[d + k];
    float gr, gi;  // gradient of sum((lr rr - li ri) or + (lr ri + li rr) oi)
    if (grad_lhs) {
      gr = rr * orr + ri * oi;
      gi = -ri * orr + rr * oi;
    } else {
      gr = lr * rr - li * ri;
      gi = lr * ri + li * rr;
    }
    const float er = Ee[k] + sgn * gr, ei = Ee[d + k] + sgn * gi;
    const float flr = FL[k], fli = FL[d + k], frr = FR[k], fri = FR[d + k], forr = FO[k], foi = FO[d + k];
    orig += (flr * frr - fli * fri) * forr + (flr * fri + fli * frr) * foi;
    if (swap_lhs) pert += (er * frr - ei * fri) * forr + (er * fri + ei * frr) * foi;
    else pert += (flr * frr - fli * fri) * er + (flr * fri + fli * frr) * ei;
  }
  for (int o = 16; o > 0; o >>= 1) {
    orig += __shfl_xor_sync(0xffffffffu, orig, o);
    pert += __shfl_xor_sync(0xffffffffu, pert, o);
  }
  if (lane == 0) out[j] = sufficient ? -orig + lambd * pert : orig - lambd * pert;
}

}  // namespace

extern "C" int kp_dp_relevance(kp_ctx* ctx, int32_t n_jobs, const int32_t* preds, const int32_t* facts, const int32_t* entity,
                               float epsilon, float lambd, int32_t sufficient, float* out, void* stream) {
  if (!ctx) return KP_EINVAL;
  if (n_jobs < 0 || (n_jobs > 0 && (!preds || !facts || !entity || !out))) KP_FAIL(ctx, KP_EINVAL, "bad data-poisoning arguments");
  if (ctx->kind != KP_COMPLEX)
    KP_FAIL(ctx, KP_EUNSUPPORTED, "the data-poisoning engine needs Model.score_embeddings, which only ComplEx defines among TransE / ComplEx / ConvE");
  if (n_jobs == 0) return KP_OK;
  KP_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dp_kernel<<<(n_jobs + 7) / 8, 256, 0, st>>>(n_jobs, ctx->D, ctx->ent, ctx->rel, preds, facts, entity, epsilon, lambd, sufficient, out);
  KP_LAUNCHED(ctx, 1);
  return KP_OK;
}
