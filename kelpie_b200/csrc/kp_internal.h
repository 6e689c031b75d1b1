// Internal declarations shared by the translation units of libkelpie_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/kelpie_b200.h"

// B operand of the tcgen05 fp32 GEMM (kp_gemm_umma.cu): split bf16 tables of a frozen [N, K] matrix
struct kp_umma_b {
  CUtensorMap hi64, lo64;  // box {64 k, 64 rows}
  int N = 0, K = 0, Kpad = 0;
  bool ready = false;
};

struct kp_ctx {
  int device = 0;
  int kind = 0;
  int64_t N = 0;   // entities (the mimic id is N)
  int64_t R2 = 0;  // relations incl. inverses
  int D = 0;       // floats per embedding row
  int norm = 2;
  int sm_count = 148;
  long long total_mem = 0;  // device memory in bytes (queried on first use)
  const float* ent = nullptr;  // [N, D]
  const float* rel = nullptr;  // [R2, D]
  bool own_ent = false, own_rel = false;
  CUtensorMap ent_map;  // TMA view of ent: box {32 floats, 128 rows}, 128B swizzle

  // resident filter CSR (kp_filter_upload)
  int64_t n_keys = 0;
  int64_t* f_keys = nullptr;
  int64_t* f_off = nullptr;
  int32_t* f_ids = nullptr;

  // ConvE frozen network (device copies) -- see kp_conve.cu
  struct {
    float *conv_w = nullptr, *conv_b = nullptr, *fc_w = nullptr, *fc_b = nullptr;
    float *bn1 = nullptr, *bn2 = nullptr, *bn3 = nullptr;
    float* ent_colsum = nullptr;  // [D] column sums of the entity table
    int n_filters = 0, hidden = 0, H = 0;  // H = D / 20
    float drop_in = 0, drop_fm = 0, drop_hid = 0;
    kp_umma_b fc_fwd, fc_bwd;  // Linear layer as B operands: W [D, hidden] (forward) and W^T [hidden, D] (backward)
  } cv;

  // split bf16 entity tables for the tcgen05 passes (built on first use, kp_flash_umma.cu)
  struct {
    bool ready = false;
    void *ent_hi = nullptr, *ent_lo = nullptr;
    CUtensorMap eh_map, el_map;      // box {64 bf16, 128 rows}
    CUtensorMap eh64_map, el64_map;  // box {64 bf16, 64 rows} (half entity tile, cta_group::2 pass)
    float* enorm = nullptr;          // [Npad] L2 norm of every entity row
    int Dpad = 0;                    // padded row width of the split tables
  } um;

  // grow-only device workspace arenas (0: drivers' scratch, 1: the tcgen05 pass's own scratch)
  std::vector<void*> owned;
  char* ws = nullptr;  // == ws_arena[0]
  size_t ws_bytes = 0;
  char* ws_arena[2] = {nullptr, nullptr};
  size_t ws_arena_bytes[2] = {0, 0};

  // tensor maps depend only on (base, type, shape, pitch, box): encoded once per distinct key (kp_encode_2d)
  typedef std::tuple<const void*, int, int64_t, int64_t, int64_t, int, int, bool> TmapKey;
  std::map<TmapKey, CUtensorMap> tmaps;

  int64_t launches = 0;
  unsigned long long* rank_recheck_total = nullptr;  // device counter: pairs the tensor-core rank pass handed to the exact re-check
  int64_t force_simt = 0;
  int64_t skinny_fc = 1;      // ConvE Linear forward of < 128 rows on the skinny kernel (else the tiled CUDA-core GEMM)
  int64_t cx_rowgrad = 1;     // ComplEx: per-row stage (strip merge + row gradient, one CTA per row) ahead of the per-candidate update
  int64_t cx_merge = 1;       // ComplEx: strips merged per row by cx_merge_strips before the per-candidate update
  int64_t umma_min_rows = 1;  // fused pass on tcgen05 from this many rows on (one 128-row tile costs the same for 1..128 rows:
                              // 43 us vs 164 us for the CUDA-core pass at 24 620 x 400); 32 = the earlier threshold
  int64_t force_tile = 0;  // route few-query passes through the 64-query tile kernel (tests)
  int64_t umma_2sm = 1;  // use the cta_group::2 pass when there are >= 2 query tiles
  int64_t umma_x4 = 2;   // rows wider than 256 floats: clusters of two pairs that compute S once; 2 = one pair scores, the other
                         // contracts (kp_flash_umma_sv.cu), 1 = both alternate (kp_flash_umma4.cu), 0 = independent pairs
  int64_t sv_dbg = 0;    // debug switches of kp_flash_umma_sv.cu
  int64_t umma_max_tps = 256;  // entity tiles per strip of the tcgen05 fused pass at most (0 = no bound).  The tensor core's fp32
                               // accumulation truncates: a running sum loses ~2^-24 of itself per accumulation, i.e. 1.9e-3 over
                               // the 31 000 K-steps of a 500 000-entity strip -- enough, through Adagrad's scale-invariant update,
                               // to move post-trained rows by 1e-3 (measured: the same candidate in batches of 2 / 64 / 300 / 1200).
                               // 256 tiles = 2048 K-steps bound it at 1.2e-4 of O; the strips are merged in fp32 (round to nearest).
  int64_t umma_qres = 1;  // pair kernel (rows of <= 256 floats): query tile resident in shared memory (kp_flash_umma2.cu)
  int64_t conv_split = 1;  // ConvE conv kernel emits the Linear GEMM's bf16 hi / lo operand directly (kp_conve.cu)
  int64_t gemm_ksplit = 1;   // forward Linear layer of few rows: K cut over CTA pairs, partial sums added in a fixed order (kp_gemm_umma.cu)
  int64_t umma_fc_min_rows = 128;  // ConvE Linear layer on tcgen05 from this many pairs on; below, the fp32 FMA chains of kp_gemm.cu (explain-sized
                                   // batches: 3 % slower than tcgen05 from 16 pairs on, but ranks among saturated sigmoid scores follow the reference more closely)
  int64_t gemm_wide = 1;  // ConvE Linear layer GEMMs with accumulator tiles of up to 256 columns (kp_gemm_umma.cu)
  int64_t umma_fc = 1;      // ConvE Linear layer on the tensor cores (kp_gemm_umma.cu) from 128 rows on
  int64_t umma_rank = 1;    // filtered rank of >= 128 DOT queries on the tensor cores with an exact fp32 re-check (kp_rank_umma.cu)
  int64_t umma_rotate = 1;  // rotating start of the entity walk (clusters share the table pass through L2)
  int* umma_cursor = nullptr;  // device [64]
  unsigned long long* umma_prof = nullptr;  // device [4], option "umma_prof" = 1: wait-cycle counters of the cluster-4 pass
  int64_t umma_cq = 0;  // query tiles per cluster of the tcgen05 pass (0 = automatic)

  // optional per-category kernel timing (kp_set_option("timing", 1); read with kp_stat)
  enum { T_PASS = 0, T_FLASH = 1, T_TRANSE_TRAIN = 2, T_UPDATE = 3, T_CONV = 4, T_NCAT = 5 };
  int64_t timing = 0;
  struct Timed { cudaEvent_t a, b; int cat; };
  std::vector<Timed> timed;
  double t_ms[T_NCAT] = {0, 0, 0, 0, 0};
  int64_t t_n[T_NCAT] = {0, 0, 0, 0, 0};
  std::string err;
};

#define KP_FAIL(ctx, code, ...)                          \
  do {                                                   \
    char _b[512];                                        \
    snprintf(_b, sizeof(_b), __VA_ARGS__);               \
    kp_set_error((ctx), _b);                             \
    return (code);                                       \
  } while (0)

#define KP_CUDA(ctx, expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      KP_FAIL(ctx, KP_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
              __FILE__, __LINE__);                                                      \
  } while (0)

#define KP_LAUNCHED(ctx, n)                                                             \
  do {                                                                                  \
    (ctx)->launches += (n);                                                             \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess)                                                              \
      KP_FAIL(ctx, KP_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
              __FILE__, __LINE__);                                                      \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per device and kernel; safe when several host threads (one per
// context) reach the same launch site at once (the attribute call is idempotent, the flag is atomic)
#define KP_SMEM_ONCE(ctx, func, bytes)                                                                          \
  do {                                                                                                          \
    static std::atomic<unsigned long long> _kp_done{0};                                                         \
    const unsigned long long _kp_bit = 1ull << ((ctx)->device & 63);                                            \
    if (!(_kp_done.load(std::memory_order_acquire) & _kp_bit)) {                                                \
      KP_CUDA(ctx, cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
      _kp_done.fetch_or(_kp_bit, std::memory_order_release);                                                    \
    }                                                                                                           \
  } while (0)

void kp_set_error(kp_ctx* ctx, const char* msg);
// CUDA-event bracket around a kernel launch on its own stream (no-op unless timing is on)
struct KpTimer {
  kp_ctx* ctx;
  cudaStream_t st;
  bool on;
  kp_ctx::Timed t;
  KpTimer(kp_ctx* c, int cat, cudaStream_t s) : ctx(c), st(s), on(c->timing != 0) {
    if (!on) return;
    t.cat = cat;
    if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) { on = false; return; }
    cudaEventRecord(t.a, st);
  }
  ~KpTimer() {
    if (!on) return;
    cudaEventRecord(t.b, st);
    ctx->timed.push_back(t);
  }
};
// workspace: returns a 1024B-aligned device pointer valid until the next kp_ws_reset
int kp_ws_reserve(kp_ctx* ctx, size_t bytes, int arena = 0);
int kp_encode_2d(kp_ctx* ctx, CUtensorMap* map, const void* base, CUtensorMapDataType dtype, int elem_bytes,
                 int64_t rows, int64_t cols, int64_t ld_elems, int box_rows, int box_cols, bool swizzle128);
int kp_encode_2d_f32(kp_ctx* ctx, CUtensorMap* map, const float* base, int64_t rows, int64_t cols,
                     int64_t ld_floats, int box_rows, int box_cols, bool swizzle128);

struct WsCursor {
  char* p;
  char* end;
  template <typename T>
  T* take(size_t n) {
    size_t b = (n * sizeof(T) + 1023) & ~size_t(1023);
    T* r = reinterpret_cast<T*>(p);
    p += b;
    return r;
  }
  static size_t need(size_t n, size_t sz) { return (n * sz + 1023) & ~size_t(1023); }
};

// ---- kp_pass.cu : all-entity scoring pass (CUDA-core tile kernel) -------------------------
enum { KP_OP_DOT = 0, KP_OP_L2 = 1, KP_OP_L1 = 2 };
enum { KP_ACT_NONE = 0, KP_ACT_SIGMOID = 1 };

struct kp_pass_args {
  const float* qmat;  // [Qn padded to a multiple of 64, D] prepared query vectors
  int Qn;
  int op, act;
  // STORE epilogue
  float* out;
  int64_t out_ld;
  // RANK epilogue
  bool rank;
  bool minimize;
  const float* target;     // [Qn]
  const int32_t* tgt_ent;  // [Qn]
  const int64_t* flt_beg;  // [Qn] ranges into flt_ids
  const int64_t* flt_end;
  const int32_t* flt_ids;
  int32_t* cnt;     // [Qn,4] accumulated with atomics: strict, tie, tie_lo, -
  uint32_t* best;   // [Qn] order-preserving uint encoding of the best other score
};
int kp_pass_launch(kp_ctx* ctx, const kp_pass_args& a, cudaStream_t st);
// kp_stream.cu: same contract for Q <= 8 queries (HBM-streaming, no smem staging)
bool kp_stream_usable(kp_ctx* ctx, int Qn);
int kp_stream_launch(kp_ctx* ctx, const kp_pass_args& a, cudaStream_t st);

// ---- kp_score.cu : query preparation, rank finalisation ----------------------------------
int kp_score_impl(kp_ctx* ctx, int Q, const int32_t* triples, const float* mimic, float* out,
                  int64_t out_ld, const int64_t* flt_off, const int32_t* flt_ids, int mode,
                  float* target_score, float* best_score, int64_t* rank, int32_t* counters,
                  bool want_rank, cudaStream_t st);

// out[q] = score of (s_q, p_q, o_q): query preparation + one pair score per row (Model.score)
int kp_score_rows_impl(kp_ctx* ctx, int Q, const int32_t* triples, const float* mimic, float* out, cudaStream_t st);
void kp_filter_release(kp_ctx* ctx);

// ---- kp_transe_train.cu / kp_complex_train.cu / kp_conve.cu -------------------------------
int kp_transe_post_train(kp_ctx* ctx, const kp_pt_batch* b, const kp_hp* hp, cudaStream_t st);
int kp_complex_post_train(kp_ctx* ctx, const kp_pt_batch* b, const kp_hp* hp, cudaStream_t st);
int kp_conve_post_train(kp_ctx* ctx, const kp_pt_batch* b, const kp_hp* hp, cudaStream_t st);
int kp_conve_setup(kp_ctx* ctx, const kp_conve_weights* w);
// features x = phi(lhs, rel) for Q (lhs,rel) pairs; lhs id N -> mimic row q
int kp_conve_features(kp_ctx* ctx, int Q, const int32_t* triples, int stride, const float* mimic,
                      float* x_out, cudaStream_t st);
int kp_conve_features_ex(kp_ctx* ctx, int Q, const int32_t* lhs_ids, const int32_t* rel_ids, int stride,
                         const float* mimic, const int32_t* mimic_index, float* x_out, float* feat_out,
                         cudaStream_t st, const int32_t* drop_ids = nullptr, unsigned long long seed = 0,
                         int step = 0);
// k_rows_b (non-transposed B only): rows of B that exist when K was rounded up over a zero-padded A (-1: K);
// split_k: allow cutting a long K over blockIdx.z (atomic reduction: unordered sums, only for the trainers)
// C[M, N] = A[M, K] B[N, K]^T for few rows and a long K: K strided over the threads of a CTA, fixed-order reduction
int kp_sgemm_skinny_nt(kp_ctx* ctx, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                       cudaStream_t st);
int kp_sgemm(kp_ctx* ctx, bool transb, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
             int ldc, cudaStream_t st, int k_rows_b = -1, bool split_k = false);
int kp_flash_umma2_launch(kp_ctx* ctx, const CUtensorMap& qh_map, const CUtensorMap& ql_map, int G, int KBs,
                          int groups_per_chunk, int n_chunks, int n_qt, int n_strips, int tps, int mode, float* part_m,
                          float* part_l, float* part_O, cudaStream_t st);
int kp_flash_umma4_sms(kp_ctx* ctx);
int kp_flash_umma4_launch(kp_ctx* ctx, const CUtensorMap& qh_map, const CUtensorMap& ql_map, int G, int KBs,
                          int groups_per_chunk, int n_qt, int n_strips, int tps, int mode, float* part_m, float* part_l,
                          float* part_O, cudaStream_t st);
int kp_flash_umma_sv_sms(kp_ctx* ctx);
int kp_flash_umma_sv_launch(kp_ctx* ctx, const CUtensorMap& qh_map, const CUtensorMap& ql_map, int G, int KBs, int ngroup,
                            int n_qt, int n_strips, int tps, int mode, float* part_m, float* part_l, float* part_O,
                            cudaStream_t st);
int kp_umma_min_strips(kp_ctx* ctx, long long g);
int kp_umma_tables(kp_ctx* ctx, cudaStream_t st);
int kp_umma_split_rows(kp_ctx* ctx, const float* mat, int G, long long Gpad, CUtensorMap* hi_map, CUtensorMap* lo_map,
                       cudaStream_t st);
// kp_rank_umma.cu: RANK epilogue of the DOT pass on the tensor cores; same counters as kp_pass_launch (bit-identical)
bool kp_rank_umma_usable(kp_ctx* ctx, const kp_pass_args& a);
int kp_rank_umma_launch(kp_ctx* ctx, const kp_pass_args& a, cudaStream_t st);
int kp_umma_b_prepare(kp_ctx* ctx, const float* B, int N, int K, bool transpose, kp_umma_b* out, cudaStream_t st);
int kp_gemm_umma(kp_ctx* ctx, const float* A, long long lda, int M, const kp_umma_b& B, float* C, long long ldc, size_t ws_offset,
                 cudaStream_t st);
size_t kp_gemm_umma_a_bytes(int M, const kp_umma_b& B);
// parts (optional): room for kp_gemm_umma_ksplit(ctx, M, B) x [M, ldc] partial results; when that count is > 1, K is cut over that many
// CTA pairs per tile, C is NOT written and the caller sums parts[z * M * ldc + ...] over z in order
int kp_gemm_umma_ksplit(const kp_ctx* ctx, int M, const kp_umma_b& B);
int kp_gemm_umma_split(kp_ctx* ctx, const void* ah, const void* al, int M, const kp_umma_b& B, float* C, long long ldc, cudaStream_t st,
                       float* parts = nullptr);
// ConvE Linear layer: forward (x = feat W^T) / backward (dfeat = dh W); tcgen05 from 128 rows on, else CUDA cores
int kp_conve_fc(kp_ctx* ctx, bool forward, int M, const float* A, float* C, size_t ws_offset, cudaStream_t st);
// Whether the forward Linear layer of M pairs runs on tcgen05: the conv kernel then writes the feature maps as the GEMM's split
// A operand -- bf16 hi [M rounded up to 256, Kpad] at the start of `feat_out`, lo kp_conve_feat_half_bytes() behind it -- instead
// of fp32 [M, hidden] followed by a separate split pass.
bool kp_conve_fc_umma(const kp_ctx* ctx, int M);
size_t kp_conve_feat_half_bytes(const kp_ctx* ctx, int M);
int kp_conve_feat_kpad(const kp_ctx* ctx);

// Strip count for a launch of `units` co-resident CTA groups per strip on `slots` group slots: the entity (or
// column) range is cut into s strips so that units * s fills whole waves.  Less than a wave per strip: up to
// ~4 waves in total; otherwise at most 8 strips, chosen against wave quantisation.
inline int kp_plan_strips(long long units, long long slots, long long n_tiles) {
  if (units < 1) units = 1;
  if (slots < 1) slots = 1;
  long long cmax = units >= slots ? 8 : 4 * slots / units;
  if (cmax > 64) cmax = 64;
  if (cmax > n_tiles) cmax = n_tiles;
  int s = 1;
  double best = 0.0;
  for (int c = 1; c <= (int)cmax; ++c) {
    const long long waves = (units * c + slots - 1) / slots;
    const double eff = (double)(units * c) / (double)(waves * slots);
    if (eff > best + 0.02) {
      best = eff;
      s = c;
    }
  }
  return s;
}
int kp_gemm_umma_dyn(kp_ctx* ctx, const float* A, long long lda, bool transA, int M, const float* B, long long ldb, bool transB,
                     int N, int K, float* C, long long ldc, size_t ws_offset, cudaStream_t st, bool tf32 = false);
