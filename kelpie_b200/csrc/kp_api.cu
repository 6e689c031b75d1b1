// C-ABI entry points of libkelpie_b200.so: context, tables, resident filter CSR, dispatch.
#include <string.h>

#include <mutex>

#include <stdlib.h>

#include "kp_internal.h"

// message of a failed kp_ctx_create, read back by the calling thread through kp_last_error(NULL)
static thread_local std::string g_create_error;

void kp_set_error(kp_ctx* ctx, const char* msg) {
  if (ctx)
    ctx->err = msg;
  else
    g_create_error = msg;
}

extern "C" const char* kp_last_error(const kp_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int kp_abi_version(void) { return KP_ABI_VERSION; }

extern "C" int64_t kp_launch_count(const kp_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int kp_set_option(kp_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return KP_EINVAL;
  if (!strcmp(name, "force_simt")) {
    ctx->force_simt = value;
    return KP_OK;
  }
  if (!strcmp(name, "umma_min_rows")) {
    ctx->umma_min_rows = value < 1 ? 1 : value;
    return KP_OK;
  }
  if (!strcmp(name, "skinny_fc")) {
    ctx->skinny_fc = value;
    return KP_OK;
  }
  if (!strcmp(name, "cx_rowgrad")) {
    ctx->cx_rowgrad = value;
    return KP_OK;
  }
  if (!strcmp(name, "cx_merge")) {
    ctx->cx_merge = value;
    return KP_OK;
  }
  if (!strcmp(name, "umma_2sm")) {
    ctx->umma_2sm = value;
    return KP_OK;
  }
  if (!strcmp(name, "umma_x4")) {
    ctx->umma_x4 = value;
    return KP_OK;
  }
  if (!strcmp(name, "umma_max_tps")) {
    if (value < 0) KP_FAIL(ctx, KP_EINVAL, "umma_max_tps must be >= 0");
    ctx->umma_max_tps = value;
    return KP_OK;
  }
  if (!strcmp(name, "gemm_ksplit")) {
    ctx->gemm_ksplit = value;
    return KP_OK;
  }
  if (!strcmp(name, "umma_fc_min_rows")) {
    ctx->umma_fc_min_rows = value < 1 ? 1 : value;
    return KP_OK;
  }
  if (!strcmp(name, "conv_split")) {
    ctx->conv_split = value;
    return KP_OK;
  }
  if (!strcmp(name, "umma_qres")) {
    ctx->umma_qres = value;
    return KP_OK;
  }
  if (!strcmp(name, "gemm_wide")) {
    ctx->gemm_wide = value;
    return KP_OK;
  }
  if (!strcmp(name, "sv_dbg")) {
    ctx->sv_dbg = value;
    return KP_OK;
  }
  if (!strcmp(name, "umma_fc")) {
    ctx->umma_fc = value;
    return KP_OK;
  }
  if (!strcmp(name, "umma_rank")) {
    ctx->umma_rank = value;
    return KP_OK;
  }
  if (!strcmp(name, "umma_rotate")) {
    ctx->umma_rotate = value;
    return KP_OK;
  }
  if (!strcmp(name, "umma_prof")) {
    if (value && !ctx->umma_prof) {
      void* d = nullptr;
      if (cudaMalloc(&d, 16 * sizeof(unsigned long long)) != cudaSuccess) KP_FAIL(ctx, KP_ENOMEM, "umma_prof counters");
      cudaMemset(d, 0, 16 * sizeof(unsigned long long));
      ctx->owned.push_back(d);
      ctx->umma_prof = static_cast<unsigned long long*>(d);
    }
    if (!value) ctx->umma_prof = nullptr;
    return KP_OK;
  }
  if (!strcmp(name, "force_tile")) {
    ctx->force_tile = value;
    return KP_OK;
  }
  if (!strcmp(name, "umma_cq")) {
    if (value != 0 && value != 1 && value != 2 && value != 4) KP_FAIL(ctx, KP_EINVAL, "umma_cq must be 0, 1, 2 or 4");
    ctx->umma_cq = value;
    return KP_OK;
  }
  if (!strcmp(name, "timing")) {
    ctx->timing = value;
    return KP_OK;
  }
  KP_FAIL(ctx, KP_EINVAL, "unknown option '%s'", name);
}

static int drain_timers(kp_ctx* ctx) {
  for (auto& t : ctx->timed) {
    KP_CUDA(ctx, cudaEventSynchronize(t.b));
    float ms = 0.f;
    KP_CUDA(ctx, cudaEventElapsedTime(&ms, t.a, t.b));
    ctx->t_ms[t.cat] += ms;
    ctx->t_n[t.cat] += 1;
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  ctx->timed.clear();
  return KP_OK;
}

extern "C" int kp_stat(kp_ctx* ctx, const char* name, double* out) {
  if (!ctx || !name || !out) return KP_EINVAL;
  static const char* cats[] = {"pass", "flash", "transe_train", "update", "conv"};
  int rc = drain_timers(ctx);
  if (rc != KP_OK) return rc;
  if (!strcmp(name, "reset")) {
    for (int i = 0; i < kp_ctx::T_NCAT; ++i) ctx->t_ms[i] = 0, ctx->t_n[i] = 0;
    *out = 0;
    return KP_OK;
  }
  for (int i = 0; i < kp_ctx::T_NCAT; ++i) {
    if (!strncmp(name, "ms_", 3) && !strcmp(name + 3, cats[i])) { *out = ctx->t_ms[i]; return KP_OK; }
    if (!strncmp(name, "n_", 2) && !strcmp(name + 2, cats[i])) { *out = (double)ctx->t_n[i]; return KP_OK; }
  }
  if (!strcmp(name, "rank_rechecks")) {
    unsigned long long v = 0;
    if (ctx->rank_recheck_total) KP_CUDA(ctx, cudaMemcpy(&v, ctx->rank_recheck_total, sizeof(v), cudaMemcpyDeviceToHost));
    *out = (double)v;
    return KP_OK;
  }
  if (!strncmp(name, "umma_prof_", 10) && ctx->umma_prof) {  // slot / own / for / total (MMA-thread cycles, summed over pairs)
    static const char* w[] = {"slot", "own", "for", "total", "send", "wpin", "whdr", "wsfull", "soft"};
    if (name[10] >= '0' && name[10] <= '9') {  // raw counter by index (kp_flash_umma_sv.cu: S slot, S dep, S total, V slot, V dep, V total)
      const int i = atoi(name + 10);
      if (i < 0 || i > 15) KP_FAIL(ctx, KP_EINVAL, "umma_prof index out of range");
      unsigned long long v = 0;
      KP_CUDA(ctx, cudaMemcpy(&v, ctx->umma_prof + i, sizeof(v), cudaMemcpyDeviceToHost));
      *out = (double)v;
      return KP_OK;
    }
    for (int i = 0; i < 9; ++i)
      if (!strcmp(name + 10, w[i])) {
        unsigned long long v = 0;
        KP_CUDA(ctx, cudaMemcpy(&v, ctx->umma_prof + i, sizeof(v), cudaMemcpyDeviceToHost));
        *out = (double)v;
        return KP_OK;
      }
  }
  KP_FAIL(ctx, KP_EINVAL, "unknown stat '%s'", name);
}

static bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Copy `bytes` from a host-or-device pointer into a fresh device allocation owned by ctx.
template <typename T>
static int adopt(kp_ctx* ctx, const T* src, size_t count, T** out) {
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, count * sizeof(T) + 16);
  if (e != cudaSuccess) KP_FAIL(ctx, KP_ENOMEM, "cudaMalloc(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e));
  ctx->owned.push_back(d);
  if (count) KP_CUDA(ctx, cudaMemcpy(d, src, count * sizeof(T), cudaMemcpyDefault));
  *out = static_cast<T*>(d);
  return KP_OK;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  });
  return fn;
}

int kp_encode_2d_f32(kp_ctx* ctx, CUtensorMap* map, const float* base, int64_t rows, int64_t cols,
                     int64_t ld_floats, int box_rows, int box_cols, bool swizzle128) {
  return kp_encode_2d(ctx, map, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows, cols, ld_floats, box_rows, box_cols,
                      swizzle128);
}

int kp_encode_2d(kp_ctx* ctx, CUtensorMap* map, const void* base, CUtensorMapDataType dtype, int elem_bytes,
                 int64_t rows, int64_t cols, int64_t ld_floats, int box_rows, int box_cols, bool swizzle128) {
  const kp_ctx::TmapKey key((const void*)base, (int)dtype * 16 + elem_bytes, rows, cols, ld_floats, box_rows, box_cols, swizzle128);
  if (ctx) {
    auto it = ctx->tmaps.find(key);
    if (it != ctx->tmaps.end()) {
      *map = it->second;
      return KP_OK;
    }
  }
  encode_tiled_fn enc = get_encode();
  if (!enc) KP_FAIL(ctx, KP_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_floats * (cuuint64_t)elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    KP_FAIL(ctx, KP_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
            (long long)rows, (long long)cols, (long long)ld_floats);
  if (ctx) {
    if (ctx->tmaps.size() >= 512) ctx->tmaps.clear();
    ctx->tmaps.emplace(key, *map);
  }
  return KP_OK;
}

int kp_ws_reserve(kp_ctx* ctx, size_t bytes, int arena) {
  if (bytes <= ctx->ws_arena_bytes[arena]) return KP_OK;
  // grow-only; the old block is kept until destroy (in-flight kernels may still read it)
  size_t want = bytes + (bytes >> 2) + (1 << 20);
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, want);
  if (e != cudaSuccess) KP_FAIL(ctx, KP_ENOMEM, "workspace cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  ctx->owned.push_back(d);
  ctx->ws_arena[arena] = static_cast<char*>(d);
  ctx->ws_arena_bytes[arena] = want;
  if (arena == 0) {
    ctx->ws = ctx->ws_arena[0];
    ctx->ws_bytes = want;
  }
  return KP_OK;
}

extern "C" int kp_ctx_create(int device, int model_kind, int64_t n_entities, int64_t n_relations2,
                             int32_t dim, int32_t norm, const float* ent, const float* rel,
                             const kp_conve_weights* conve, kp_ctx** out) {
  if (!out) return KP_EINVAL;
  *out = nullptr;
  if (model_kind < KP_TRANSE || model_kind > KP_CONVE) KP_FAIL(nullptr, KP_EINVAL, "bad model kind %d", model_kind);
  if (n_entities <= 0 || n_entities >= (int64_t(1) << 31) - 2 || n_relations2 <= 0 || dim <= 0)
    KP_FAIL(nullptr, KP_EINVAL, "bad sizes N=%lld R2=%lld D=%d", (long long)n_entities, (long long)n_relations2, dim);
  if (dim % 4 != 0 || dim > 2048) KP_FAIL(nullptr, KP_EUNSUPPORTED, "dim %d must be a multiple of 4 and <= 2048", dim);
  if (!ent || !rel) KP_FAIL(nullptr, KP_EINVAL, "null table pointer");
  if (model_kind == KP_TRANSE && norm != 1 && norm != 2) KP_FAIL(nullptr, KP_EUNSUPPORTED, "TransE norm %d (1 or 2)", norm);
  if (model_kind == KP_COMPLEX && dim % 8 != 0) KP_FAIL(nullptr, KP_EUNSUPPORTED, "ComplEx row width %d must be a multiple of 8", dim);
  if (model_kind == KP_CONVE && (!conve || dim % 20 != 0 || dim / 20 < 3))
    KP_FAIL(nullptr, KP_EINVAL, "ConvE needs weights and dim = 20*h, h >= 3 (dim=%d)", dim);

  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    KP_FAIL(nullptr, KP_ECUDA, "no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
  }
  if (device < 0 || device >= n_dev) KP_FAIL(nullptr, KP_EINVAL, "device %d out of range (%d devices)", device, n_dev);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    KP_FAIL(nullptr, KP_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    KP_FAIL(nullptr, KP_EUNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  if ((e = cudaSetDevice(device)) != cudaSuccess) KP_FAIL(nullptr, KP_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));

  kp_ctx* ctx = new kp_ctx();
  ctx->device = device;
  ctx->kind = model_kind;
  ctx->N = n_entities;
  ctx->R2 = n_relations2;
  ctx->D = dim;
  ctx->norm = norm;
  ctx->sm_count = prop.multiProcessorCount;

  int rc = KP_OK;
  auto fail = [&](int code) {
    g_create_error = ctx->err;
    kp_ctx_destroy(ctx);
    return code;
  };
  // tables: borrow 16B-aligned device pointers, copy anything else
  if (is_device_ptr(ent) && (reinterpret_cast<uintptr_t>(ent) & 15) == 0) {
    ctx->ent = ent;
  } else {
    float* d = nullptr;
    if ((rc = adopt(ctx, ent, (size_t)n_entities * dim, &d)) != KP_OK) return fail(rc);
    ctx->ent = d;
    ctx->own_ent = true;
  }
  if (is_device_ptr(rel) && (reinterpret_cast<uintptr_t>(rel) & 15) == 0) {
    ctx->rel = rel;
  } else {
    float* d = nullptr;
    if ((rc = adopt(ctx, rel, (size_t)n_relations2 * dim, &d)) != KP_OK) return fail(rc);
    ctx->rel = d;
    ctx->own_rel = true;
  }
  if ((rc = kp_encode_2d_f32(ctx, &ctx->ent_map, ctx->ent, n_entities, dim, dim, 128, 32, true)) != KP_OK)
    return fail(rc);
  if (model_kind == KP_CONVE && (rc = kp_conve_setup(ctx, conve)) != KP_OK) return fail(rc);
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) {
    ctx->err = std::string("context setup failed: ") + cudaGetErrorString(e);
    return fail(KP_ECUDA);
  }
  *out = ctx;
  return KP_OK;
}

extern "C" int kp_ctx_destroy(kp_ctx* ctx) {
  if (!ctx) return KP_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (void* p : ctx->owned) cudaFree(p);
  delete ctx;
  return KP_OK;
}

// The resident CSR is replaced, not accumulated: the previous arrays are freed once no kernel can still read them.
void kp_filter_release(kp_ctx* ctx) {
  void* old[3] = {ctx->f_keys, ctx->f_off, ctx->f_ids};
  if (!old[0] && !old[1] && !old[2]) return;
  cudaDeviceSynchronize();
  for (int i = 0; i < 3; ++i) {
    void* p = old[i];
    bool dup = !p;
    for (int j = 0; j < i; ++j) dup = dup || old[j] == p;  // the empty CSR aliases one buffer
    if (dup) continue;
    for (size_t k = 0; k < ctx->owned.size(); ++k)
      if (ctx->owned[k] == p) {
        ctx->owned.erase(ctx->owned.begin() + k);
        cudaFree(p);
        break;
      }
  }
  ctx->f_keys = nullptr;
  ctx->f_off = nullptr;
  ctx->f_ids = nullptr;
  ctx->n_keys = 0;
}

extern "C" int kp_filter_upload(kp_ctx* ctx, int64_t n_keys, const int64_t* keys, const int64_t* offsets,
                                const int32_t* objs) {
  if (!ctx) return KP_EINVAL;
  if (n_keys < 0 || (n_keys > 0 && (!keys || !offsets))) KP_FAIL(ctx, KP_EINVAL, "bad filter CSR arguments");
  KP_CUDA(ctx, cudaSetDevice(ctx->device));
  kp_filter_release(ctx);
  int64_t total = 0;
  if (n_keys > 0) KP_CUDA(ctx, cudaMemcpy(&total, offsets + n_keys, sizeof(int64_t), cudaMemcpyDefault));
  if (total < 0 || (total > 0 && !objs)) KP_FAIL(ctx, KP_EINVAL, "bad filter CSR offsets");
  int rc;
  if ((rc = adopt(ctx, keys, (size_t)n_keys, &ctx->f_keys)) != KP_OK) return rc;
  if ((rc = adopt(ctx, offsets, (size_t)n_keys + 1, &ctx->f_off)) != KP_OK) return rc;
  if ((rc = adopt(ctx, objs, (size_t)total, &ctx->f_ids)) != KP_OK) return rc;
  if (n_keys == 0) KP_CUDA(ctx, cudaMemset(ctx->f_off, 0, sizeof(int64_t)));
  ctx->n_keys = n_keys;
  return KP_OK;
}

extern "C" int kp_all_scores(kp_ctx* ctx, int32_t n_queries, const int32_t* triples, const float* mimic_rows,
                             float* out, int64_t out_ld, void* stream) {
  if (!ctx) return KP_EINVAL;
  if (n_queries < 0 || (n_queries > 0 && (!triples || !out))) KP_FAIL(ctx, KP_EINVAL, "bad all_scores arguments");
  int64_t cols = ctx->N + (mimic_rows ? 1 : 0);
  if (out_ld < cols) KP_FAIL(ctx, KP_EINVAL, "out_ld %lld < %lld columns", (long long)out_ld, (long long)cols);
  if (n_queries == 0) return KP_OK;
  KP_CUDA(ctx, cudaSetDevice(ctx->device));
  return kp_score_impl(ctx, n_queries, triples, mimic_rows, out, out_ld, nullptr, nullptr, 0, nullptr, nullptr,
                       nullptr, nullptr, false, static_cast<cudaStream_t>(stream));
}

extern "C" int kp_score_triples(kp_ctx* ctx, int32_t n_queries, const int32_t* triples, const float* mimic_rows, float* out,
                                void* stream) {
  if (!ctx) return KP_EINVAL;
  if (n_queries < 0 || (n_queries > 0 && (!triples || !out))) KP_FAIL(ctx, KP_EINVAL, "bad score_triples arguments");
  if (n_queries == 0) return KP_OK;
  KP_CUDA(ctx, cudaSetDevice(ctx->device));
  return kp_score_rows_impl(ctx, n_queries, triples, mimic_rows, out, static_cast<cudaStream_t>(stream));
}

extern "C" int kp_filtered_rank(kp_ctx* ctx, int32_t n_queries, const int32_t* triples, const float* mimic_rows,
                                const int64_t* flt_off, const int32_t* flt_ids, int32_t rank_mode,
                                float* target_score, float* best_score, int64_t* rank, int32_t* counters,
                                void* stream) {
  if (!ctx) return KP_EINVAL;
  if (n_queries < 0 || (n_queries > 0 && !triples)) KP_FAIL(ctx, KP_EINVAL, "bad filtered_rank arguments");
  if (rank_mode < KP_RANK_ENGINE_MIN || rank_mode > KP_RANK_CONVE_SORT) KP_FAIL(ctx, KP_EINVAL, "bad rank mode %d", rank_mode);
  if (!flt_off && !ctx->f_off) KP_FAIL(ctx, KP_ESTATE, "no per-query filter given and kp_filter_upload was not called");
  if (n_queries == 0) return KP_OK;
  KP_CUDA(ctx, cudaSetDevice(ctx->device));
  return kp_score_impl(ctx, n_queries, triples, mimic_rows, nullptr, 0, flt_off, flt_ids, rank_mode, target_score,
                       best_score, rank, counters, true, static_cast<cudaStream_t>(stream));
}

extern "C" int kp_post_train_batch(kp_ctx* ctx, const kp_pt_batch* b, const kp_hp* hp, void* stream) {
  if (!ctx) return KP_EINVAL;
  if (!b || !hp) KP_FAIL(ctx, KP_EINVAL, "null batch / hyper-parameters");
  if (b->n_candidates < 0 || hp->epochs < 0 || hp->batch_size <= 0) KP_FAIL(ctx, KP_EINVAL, "bad batch sizes");
  if (b->n_candidates == 0) return KP_OK;
  if (!b->row_off || !b->rows_per_epoch || !b->init_rows || !b->out_rows) KP_FAIL(ctx, KP_EINVAL, "null batch arrays");
  KP_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (ctx->kind) {
    case KP_TRANSE:
      return kp_transe_post_train(ctx, b, hp, st);
    case KP_COMPLEX:
      return kp_complex_post_train(ctx, b, hp, st);
    default:
      return kp_conve_post_train(ctx, b, hp, st);
  }
}
