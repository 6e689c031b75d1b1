// Device builder of the known-facts CSR: replaces the Python walk over Dataset.to_filter
// (dataset.py:131-139: to_filter[(h, r)] += [t], to_filter[(t, r + R)] += [h] for every train / valid / test triple)
// that kp_filter_upload needs.  Input: filter facts (entity, relation, id) = "id is a known answer of (entity, relation)",
// in any order, duplicates allowed (the dict holds multiset lists; masking twice = masking once).  Output: the resident
// CSR of the context -- keys entity * R2 + relation ascending, ids ascending and distinct per key -- bit-identical to
// what runtime.filter_csr builds from the dict.
//   pack    (key << 32 | id) per fact                        one thread per fact
//   sort    cub::DeviceRadixSort over the used bits
//   unique  cub::DeviceSelect::Unique                        -> distinct (key, id) pairs, sorted
//   heads   flag the first pair of every key, exclusive scan -> key index of every pair; scatter keys / offsets / ids
#include <cub/cub.cuh>

#include "kp_internal.h"

namespace {

__global__ void flt_pack(int64_t n, const int32_t* __restrict__ facts, int64_t R2, unsigned long long* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = (unsigned long long)facts[3 * i] * (unsigned long long)R2 + (unsigned long long)facts[3 * i + 1];
  out[i] = (key << 32) | (unsigned long long)(uint32_t)facts[3 * i + 2];
}

__global__ void flt_heads(int64_t n, const unsigned long long* __restrict__ pairs, int32_t* __restrict__ head) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = (i == 0 || (pairs[i] >> 32) != (pairs[i - 1] >> 32)) ? 1 : 0;
}

// rank[i] = number of heads before pair i (exclusive scan of head): a head at i opens key number rank[i]
__global__ void flt_scatter(int64_t n, const unsigned long long* __restrict__ pairs, const int32_t* __restrict__ head,
                            const int32_t* __restrict__ rank, int64_t* __restrict__ keys, int64_t* __restrict__ off, int32_t* __restrict__ ids,
                            int64_t n_keys) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i == 0) off[n_keys] = n;
  if (i >= n) return;
  ids[i] = (int32_t)(uint32_t)(pairs[i] & 0xffffffffull);
  if (head[i]) {
    keys[rank[i]] = (int64_t)(pairs[i] >> 32);
    off[rank[i]] = i;
  }
}

}  // namespace

extern "C" int kp_filter_build(kp_ctx* ctx, int64_t n_facts, const int32_t* facts, void* stream) {
  if (!ctx) return KP_EINVAL;
  if (n_facts < 0 || (n_facts > 0 && !facts)) KP_FAIL(ctx, KP_EINVAL, "bad filter facts");
  if ((unsigned long long)ctx->N * (unsigned long long)ctx->R2 >= (1ull << 31))
    KP_FAIL(ctx, KP_EUNSUPPORTED, "kp_filter_build packs entity * R2 + relation into 31 bits (N * R2 = %llu)",
            (unsigned long long)ctx->N * (unsigned long long)ctx->R2);
  KP_CUDA(ctx, cudaSetDevice(ctx->device));
  kp_filter_release(ctx);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n = n_facts;
  auto fresh = [&](void** p, size_t bytes) -> bool {
    if (cudaMalloc(p, bytes + 16) != cudaSuccess) return false;
    ctx->owned.push_back(*p);
    return true;
  };
  if (n == 0) {
    void* o = nullptr;
    if (!fresh(&o, sizeof(int64_t))) KP_FAIL(ctx, KP_ENOMEM, "cannot allocate the filter CSR");
    KP_CUDA(ctx, cudaMemsetAsync(o, 0, sizeof(int64_t), st));
    ctx->f_off = (int64_t*)o;
    ctx->f_keys = (int64_t*)o;
    ctx->f_ids = (int32_t*)o;
    ctx->n_keys = 0;
    return KP_OK;
  }
  // scratch (freed before returning): facts copy when on the host, packed pairs x2, flags, ranks, cub temp
  int32_t* d_facts = nullptr;
  unsigned long long *a = nullptr, *b = nullptr;
  int32_t *head = nullptr, *rank = nullptr;
  int64_t* d_count = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0, t1 = 0, t2 = 0, t3 = 0;
  cudaPointerAttributes attr;
  const bool on_device = cudaPointerGetAttributes(&attr, facts) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  int end_bit = 32;
  for (unsigned long long k = (unsigned long long)ctx->N * (unsigned long long)ctx->R2; k > 0; k >>= 1) ++end_bit;
  cub::DeviceRadixSort::SortKeys(nullptr, t1, a, b, n, 0, end_bit, st);
  cub::DeviceSelect::Unique(nullptr, t2, b, a, d_count, n, st);
  cub::DeviceScan::ExclusiveSum(nullptr, t3, head, rank, n, st);
  tmp_bytes = t1 > t2 ? t1 : t2;
  tmp_bytes = tmp_bytes > t3 ? tmp_bytes : t3;
  int rc = KP_OK;
  int64_t m = 0, n_keys = 0;
  int32_t last_head = 0, last_rank = 0;
  void *keys = nullptr, *off = nullptr, *ids = nullptr;
  const int threads = 256;
  auto blocks = [&](int64_t k) { return (unsigned)((k + threads - 1) / threads); };
  bool ok = cudaMalloc(&a, n * 8) == cudaSuccess && cudaMalloc(&b, n * 8) == cudaSuccess && cudaMalloc(&head, n * 4) == cudaSuccess &&
            cudaMalloc(&rank, n * 4) == cudaSuccess && cudaMalloc(&d_count, 8) == cudaSuccess && cudaMalloc(&tmp, tmp_bytes + 16) == cudaSuccess &&
            (on_device || cudaMalloc(&d_facts, n * 12) == cudaSuccess);
  if (!ok) {
    kp_set_error(ctx, "kp_filter_build: cannot allocate scratch");
    rc = KP_ENOMEM;
    goto done;
  }
  if (!on_device && cudaMemcpyAsync(d_facts, facts, n * 12, cudaMemcpyHostToDevice, st) != cudaSuccess) {
    kp_set_error(ctx, "kp_filter_build: cannot copy the facts");
    rc = KP_ECUDA;
    goto done;
  }
  flt_pack<<<blocks(n), threads, 0, st>>>(n, on_device ? facts : d_facts, ctx->R2, a);
  cub::DeviceRadixSort::SortKeys(tmp, t1, a, b, n, 0, end_bit, st);
  cub::DeviceSelect::Unique(tmp, t2, b, a, d_count, n, st);
  if (cudaMemcpyAsync(&m, d_count, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
    kp_set_error(ctx, "kp_filter_build: sort / unique failed");
    rc = KP_ECUDA;
    goto done;
  }
  flt_heads<<<blocks(m), threads, 0, st>>>(m, a, head);
  cub::DeviceScan::ExclusiveSum(tmp, t3, head, rank, m, st);
  if (cudaMemcpyAsync(&last_head, head + m - 1, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaMemcpyAsync(&last_rank, rank + m - 1, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
    kp_set_error(ctx, "kp_filter_build: scan failed");
    rc = KP_ECUDA;
    goto done;
  }
  n_keys = (int64_t)last_rank + last_head;
  if (!fresh(&keys, n_keys * 8) || !fresh(&off, (n_keys + 1) * 8) || !fresh(&ids, m * 4)) {
    kp_set_error(ctx, "kp_filter_build: cannot allocate the filter CSR");
    rc = KP_ENOMEM;
    goto done;
  }
  flt_scatter<<<blocks(m), threads, 0, st>>>(m, a, head, rank, (int64_t*)keys, (int64_t*)off, (int32_t*)ids, n_keys);
  ctx->launches += 6;
  if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    kp_set_error(ctx, "kp_filter_build: scatter failed");
    rc = KP_ECUDA;
    goto done;
  }
  ctx->f_keys = (int64_t*)keys;
  ctx->f_off = (int64_t*)off;
  ctx->f_ids = (int32_t*)ids;
  ctx->n_keys = n_keys;
done:
  cudaFree(a);
  cudaFree(b);
  cudaFree(head);
  cudaFree(rank);
  cudaFree(d_count);
  cudaFree(tmp);
  cudaFree(d_facts);
  return rc;
}

// Read the resident CSR back (tests / hosts that want to inspect it): sizes first (NULL arrays), then the arrays (host pointers).
extern "C" int kp_filter_download(kp_ctx* ctx, int64_t* n_keys, int64_t* n_ids, int64_t* keys, int64_t* offsets, int32_t* ids) {
  if (!ctx || !n_keys || !n_ids) return KP_EINVAL;
  if (!ctx->f_off) KP_FAIL(ctx, KP_ESTATE, "no filter CSR is resident");
  KP_CUDA(ctx, cudaSetDevice(ctx->device));
  int64_t total = 0;
  KP_CUDA(ctx, cudaMemcpy(&total, ctx->f_off + ctx->n_keys, 8, cudaMemcpyDeviceToHost));
  *n_keys = ctx->n_keys;
  *n_ids = total;
  if (keys && ctx->n_keys) KP_CUDA(ctx, cudaMemcpy(keys, ctx->f_keys, ctx->n_keys * 8, cudaMemcpyDeviceToHost));
  if (offsets) KP_CUDA(ctx, cudaMemcpy(offsets, ctx->f_off, (ctx->n_keys + 1) * 8, cudaMemcpyDeviceToHost));
  if (ids && total) KP_CUDA(ctx, cudaMemcpy(ids, ctx->f_ids, total * 4, cudaMemcpyDeviceToHost));
  return KP_OK;
}
