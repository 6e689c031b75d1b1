/*
 * kelpie_b200.h -- C ABI of the B200-native relevance-engine hot path.
 *
 * Drop-in boundary for rbarile17/kelpie (pure Python + torch; it has no FFI of its own,
 * so each entry point names the Python call-sites it replaces; INTEGRATION.md shows the
 * ctypes stub a maintainer adds on the reference side).
 *
 * Conventions
 *   - every function returns 0 (KP_OK) or a negative KP_E* code; the message is kept per
 *     context and read with kp_last_error(); nothing throws or exits across the ABI;
 *   - all data-plane pointers are DEVICE pointers on the context's device unless the
 *     parameter is documented as "host or device" (detected with cudaPointerGetAttributes);
 *   - the caller (torch) owns every buffer it passes; the library owns only kp_ctx and
 *     frees everything it allocated in kp_ctx_destroy();
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as void*), the
 *     calls are asynchronous unless stated otherwise;
 *   - entity ids are int32 on the device (N < 2^31); ranks are int64 like the reference's;
 *   - there is NO CPU fallback: without a CUDA device every call fails with KP_ECUDA;
 *   - ONE stream per context at a time: a context owns a single grow-only workspace arena that every call carves its
 *     scratch from, so calls on different streams (or host threads) must be serialised by the caller; use one context
 *     per GPU (several GPUs: one process / one context each, kelpie_b200.parallel).  kp_last_error(NULL) (the message
 *     of a failed kp_ctx_create) is thread-local;
 *   - ids are trusted: the scoring / ranking calls answer NaN for a query id outside [0, N] / [0, R2), the training
 *     kernels do not range-check the ids of a batch -- they come from the dataset the context was built from.
 */
#ifndef KELPIE_B200_H
#define KELPIE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KP_ABI_VERSION 3

enum kp_status {
  KP_OK = 0,
  KP_EINVAL = -1,       /* bad argument */
  KP_ECUDA = -2,        /* CUDA runtime / driver error (sticky errors reported as such) */
  KP_ENOMEM = -3,       /* device allocation failed */
  KP_EUNSUPPORTED = -4, /* valid request this build does not implement */
  KP_ESTATE = -5        /* call order (e.g. rank with resident filter before kp_filter_upload) */
};

/* link_prediction/__init__.py:5-9 MODEL_REGISTRY */
enum kp_model_kind { KP_TRANSE = 0, KP_COMPLEX = 1, KP_CONVE = 2 };

/* Filtered-rank semantics.  "better" is < for the minimiser
 * (TransE) and > for the maximisers (ComplEx, ConvE). */
enum kp_rank_mode {
  /* post_training_engine.py:110-114  filtered -> 1e6, target restored, #(score <= target) */
  KP_RANK_ENGINE_MIN = 0,
  /* post_training_engine.py:116-119  filtered -> -1e6 INCLUDING the target when it is in
   * the filter, #(score >= target) counted before the target is restored */
  KP_RANK_ENGINE_MAX = 1,
  /* model.py:50-60  filtered -> +-1e6, target restored, #(<= target) or #(>= target) */
  KP_RANK_MODEL = 2,
  /* conve.py:160-184  filtered -> 0.0, target restored, 1 + position in a descending sort
   * (ties resolved in index order) */
  KP_RANK_CONVE_SORT = 3
};

enum kp_optimizer { KP_OPT_ADAGRAD = 0, KP_OPT_ADAM = 1, KP_OPT_SGD = 2 };

typedef struct kp_ctx kp_ctx;

/* Frozen ConvE network (conve.py:42-52; eval-mode batch-norm, conve.py:214-237).
 * All pointers host or device, fp32, contiguous. */
typedef struct kp_conve_weights {
  const float* conv_w;   /* [n_filters,1,3,3] */
  const float* conv_b;   /* [n_filters] */
  const float* fc_w;     /* [dim, hidden] row-major (torch Linear.weight) */
  const float* fc_b;     /* [dim] */
  const float* bn1;      /* weight,bias,running_mean,running_var : 4 floats */
  const float* bn2;      /* 4 x [n_filters] */
  const float* bn3;      /* 4 x [dim] */
  int32_t n_filters;     /* 32 */
  int32_t hidden;        /* n_filters * (2*20-2) * (dim/20-2) */
  float drop_input, drop_feature, drop_hidden; /* conve.py:34-36 */
} kp_conve_weights;

/* Hyper-parameters of one post-training (the `training` dict of configs/<model>_<dataset>_explanation.json as the
 * reference's Kelpie*Optimizer consumes it). */
typedef enum kp_regularizer { KP_REG_N3 = 0, KP_REG_N2 = 1 } kp_regularizer;

typedef struct kp_hp {
  int32_t epochs;
  int32_t batch_size;
  int32_t optimizer;      /* kp_optimizer; TransE/ConvE always Adam */
  float lr;               /* ConvE: the caller passes 1e-3, see bce_optimizer.py:165 */
  float beta1, beta2, eps;
  float margin;           /* TransE pairwise_ranking_optimizer.py:44 */
  float reg_weight;       /* TransE L2 / ComplEx N3 or N2 weight */
  float label_smoothing;  /* ConvE bce_optimizer.py:108-110 */
  int32_t regularizer;    /* ABI 3, ComplEx: a kp_regularizer value -- multiclass_nll_optimizer.py:46-49, regularizers.py:25-46 */
} kp_hp;

/* One batch of C independent mimic post-trainings (one mimic row per candidate).
 * Replaces Kelpie{PairwiseRanking,MultiClassNLL,BCE}Optimizer.train for C candidates
 * (pairwise_ranking_optimizer.py:160-203, multiclass_nll_optimizer.py:138-164,
 * bce_optimizer.py:161-208).  All index arrays are device int32 / int64; id N denotes
 * the candidate's own mimic row.
 *
 * TransE / ComplEx: candidate c owns rows [row_off[c], row_off[c+1]) of `pos` (and `neg`),
 * laid out epoch-major: epochs * rows_per_epoch[c] rows, in the order the reference
 * visits them (after its shuffle / permutation); a step covers `batch_size` consecutive
 * rows of one epoch.  TransE: neg[i] is the corrupted version of pos[i].
 * If `static_epochs` != 0 the candidate stores ONE epoch of rows which is reused for
 * every epoch (valid when an epoch is a single step: the loss is a mean over the batch,
 * so the order inside the batch is irrelevant).
 *
 * ConvE: candidate c owns pairs [row_off[c], row_off[c+1]) of `pos` (columns: lhs, rel,
 * unused) in er_vocab order (bce_optimizer.py:92-96); pair i's positives are
 * pos_ids[pos_off[i] .. pos_off[i+1]).  There is no shuffling (bce_optimizer.py:167-176).
 */
typedef struct kp_pt_batch {
  int32_t n_candidates;
  int32_t static_epochs;
  int32_t max_rows_per_epoch;    /* max over candidates of rows_per_epoch (host-known) */
  int32_t reserved;
  int64_t total_rows;            /* rows in `pos` = row_off[C] (host-known) */
  const int64_t* row_off;        /* [C+1] */
  const int32_t* rows_per_epoch; /* [C] (ConvE: pairs per epoch) */
  const int32_t* pos;            /* [rows,3] */
  const int32_t* neg;            /* [rows,3] TransE only */
  const int64_t* pos_off;        /* ConvE only, [pairs+1] */
  const int32_t* pos_ids;        /* ConvE only */
  const float* init_rows;        /* [C, D] */
  float* out_rows;               /* [C, D] */
  uint64_t dropout_seed;         /* ConvE with dropout > 0 */
  /* TransE, optional (ABI 2): compact index tables, 6 instead of 24 bytes per training row over PCIe.  When pos_idx
   * is non-NULL, pos / neg are ignored: candidate c's distinct rows ("triples + inverse triples",
   * pairwise_ranking_optimizer.py:64-65) are facts[fact_off[c] .. fact_off[c+1]) and training row i (same row_off /
   * rows_per_epoch / epoch-major layout) is the positive facts[fact_off[c] + pos_idx[i]] with, as its negative
   * (:171-195), the head (bit 31 of neg_code[i] set) or the tail (clear) replaced by entity neg_code[i] & 0x7fffffff. */
  const int64_t* fact_off;       /* [C+1] */
  const int32_t* facts;          /* [fact_off[C], 3] */
  const uint16_t* pos_idx;       /* [rows] */
  const int32_t* neg_code;       /* [rows] */
} kp_pt_batch;

/* Queries of one scoring / ranking call: Q triples (s,p,o) int32 [Q,3] on the device.
 * mimic_rows (nullable, [Q, D]): when given, entity id N means "row q of mimic_rows" and
 * the score matrix has N+1 columns (KelpieModel.all_scores, model.py:107-108). */

/* ---- context ------------------------------------------------------------------------- */

/* Replaces the per-candidate clone()+cat of the tables (transe.py:86-99,
 * complex.py:146-160, conve.py:206-212): tables are uploaded once (or borrowed when `ent`
 * / `rel` already are device pointers) and the mimic row lives beside them.
 * n_entities = N, n_relations2 = 2 * num_relations, dim = floats per row (2*d for
 * ComplEx), norm = TransE p (1 or 2; ignored otherwise).  ent/rel: host or device. */
int kp_ctx_create(int device, int model_kind, int64_t n_entities, int64_t n_relations2,
                  int32_t dim, int32_t norm, const float* ent, const float* rel,
                  const kp_conve_weights* conve, kp_ctx** out);
int kp_ctx_destroy(kp_ctx* ctx);
const char* kp_last_error(const kp_ctx* ctx); /* ctx may be NULL: last create error */
int kp_abi_version(void);

/* Device-resident CSR of the known facts, replacing the Python dict `Dataset.to_filter`
 * (dataset.py:136-139): keys[i] = entity * n_relations2 + relation, strictly ascending;
 * objs[offsets[i]..offsets[i+1]) = the DISTINCT ids to mask for that key, ascending.
 * Direct keys (s,p) list objects, inverse keys (o,p+R) list subjects.  host or device. */
int kp_filter_upload(kp_ctx* ctx, int64_t n_keys, const int64_t* keys, const int64_t* offsets,
                     const int32_t* objs);

/* ---- scoring / ranking --------------------------------------------------------------- */

/* Device builder of the same CSR: replaces the host walk over the dict Dataset.to_filter
 * (dataset.py:131-139).  facts: [n_facts, 3] int32 rows (entity, relation, id) = "id is a known answer of
 * (entity, relation)", host or device, any order, duplicates allowed (the dict holds multiset lists).  Sort + unique +
 * segment on the device; the resident CSR is identical to what kp_filter_upload receives from the dict walk.
 * Needs N * R2 < 2^31.  kp_filter_download reads the resident CSR back (call with NULL arrays for the sizes). */
int kp_filter_build(kp_ctx* ctx, int64_t n_facts, const int32_t* facts, void* stream);
int kp_filter_download(kp_ctx* ctx, int64_t* n_keys, int64_t* n_ids, int64_t* keys, int64_t* offsets, int32_t* ids);

/* Model.score (transe.py:38-46, complex.py:41-56, conve.py:68-75): out[q] = score of (s_q, p_q, o_q) -- Q rows of work,
 * no pass over the entity table.  Ids equal to N denote the query's own mimic row (mimic_rows [Q, D], nullable). ABI 3. */
int kp_score_triples(kp_ctx* ctx, int32_t n_queries, const int32_t* triples, const float* mimic_rows, float* out,
                     void* stream);

/* Model.all_scores (transe.py:48-65, complex.py:88-113, conve.py:133-158):
 * out[q, j] = score of (s_q, p_q, j), row stride out_ld floats, N (+1) columns. */
int kp_all_scores(kp_ctx* ctx, int32_t n_queries, const int32_t* triples,
                  const float* mimic_rows, float* out, int64_t out_ld, void* stream);

/* Filtered rank of the target o_q among all entities (post_training_engine.py:101-125,
 * model.py:42-68, conve.py:160-184, engine.py:94-124), scores never materialised.
 * Filter of query q: ids flt_ids[flt_off[q]..flt_off[q+1]) (ascending, distinct) when
 * flt_off != NULL, else the resident CSR entry of key (s_q, p_q).
 * Outputs (any may be NULL): target_score[Q], best_score[Q], rank[Q] (int64), and
 * counters[Q,4] = {#strictly better, #ties, #ties with id < o, target-in-filter} over the
 * unfiltered entities other than the target (what select_entities_to_convert needs). */
int kp_filtered_rank(kp_ctx* ctx, int32_t n_queries, const int32_t* triples,
                     const float* mimic_rows, const int64_t* flt_off, const int32_t* flt_ids,
                     int32_t rank_mode, float* target_score, float* best_score, int64_t* rank,
                     int32_t* counters, void* stream);

/* ---- post-training --------------------------------------------------------------------- */

int kp_post_train_batch(kp_ctx* ctx, const kp_pt_batch* batch, const kp_hp* hp, void* stream);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t kp_launch_count(const kp_ctx* ctx);

/* Set a tuning / debugging knob ("force_simt" = 1 routes GEMM-shaped passes through the
 * CUDA-core kernels, used by the tests to cross-check the tcgen05 path; "timing" = 1 brackets
 * the library's kernels with CUDA events, see kp_stat). */
int kp_set_option(kp_ctx* ctx, const char* name, int64_t value);

/* Per-category device time of the library's own kernels, measured with CUDA events on the
 * launching stream while option "timing" is 1.  name = "ms_<cat>" | "n_<cat>" | "reset" with
 * cat in {pass, flash, transe_train, update, conv}.  Synchronises the recorded events. */
int kp_stat(kp_ctx* ctx, const char* name, double* out);

/* ---- full-model TransE training (verify_explanations retrains from scratch) ----------
 * Replaces PairwiseRankingOptimizer.step_on_batch / optim.Adam (pairwise_ranking_optimizer.py:139-157,
 * :46) over the whole entity and relation tables, which are DEVICE tensors updated in place.  The host
 * draws every epoch's shuffle and corruptions in the reference's order (:100-118) and passes them as
 * index tables: step k trains on rows [step_off[k], step_off[k+1]) of pos / neg ([rows, 3] int32, device).
 * The Adam state and step counter live in the handle across calls (one call per epoch or per run). */
typedef struct kp_fit kp_fit;
int kp_transe_fit_create(int device, int64_t n_entities, int64_t n_relations2, int32_t dim, int32_t norm,
                         float lr, float margin, float reg_weight, float* ent, float* rel, kp_fit** out);
int kp_transe_fit_steps(kp_fit* fit, int64_t n_steps, const int64_t* step_off, const int32_t* pos,
                        const int32_t* neg, float* loss_out, void* stream);
int kp_transe_fit_destroy(kp_fit* fit);
const char* kp_transe_fit_error(const kp_fit* fit); /* fit may be NULL: last create error */
int64_t kp_transe_fit_launches(const kp_fit* fit);

/* Full-model ComplEx training: MultiClassNLLOptimizer.step_on_batch (multiclass_nll_optimizer.py:123-135) with
 * ComplEx.forward (complex.py:58-86), CrossEntropyLoss(mean) and optim.Adagrad / Adam / SGD (:41-48) over both
 * tables (device, updated in place).  optimizer: 0 Adagrad, 1 Adam (betas = decay1, decay2), 2 SGD.  reg_weight must
 * be 0 (as in every shipped config).  rows: [total, 3] int32 (device), already permuted by the host in the
 * reference's order (torch.randperm per epoch, :110-111); step k = rows [step_off[k], step_off[k+1]). */
typedef struct kp_cfit kp_cfit;
int kp_complex_fit_create(int device, int64_t n_entities, int64_t n_relations2, int32_t dim, int32_t optimizer,
                          float lr, float beta1, float beta2, float reg_weight, int32_t max_batch, float* ent,
                          float* rel, kp_cfit** out);
int kp_complex_fit_steps(kp_cfit* fit, int64_t n_steps, const int64_t* step_off, const int32_t* rows,
                         float* loss_out, void* stream);
int kp_complex_fit_destroy(kp_cfit* fit);
const char* kp_complex_fit_error(const kp_cfit* fit); /* fit may be NULL: last create error */
int64_t kp_complex_fit_launches(const kp_cfit* fit);

/* Full-model ConvE training: BCEOptimizer.step_on_batch (bce_optimizer.py:137-158) with ConvE.forward = all_scores
 * (conve.py:133-158: batch-norm -> 3x3 conv -> batch-norm -> relu -> Linear -> batch-norm -> relu -> 1-vs-all sigmoid
 * against the whole entity table), BCELoss(mean) with label smoothing (:104-110) and optim.Adam over EVERY parameter
 * (:36).  All tensors are DEVICE fp32, contiguous, updated in place (parameters and the batch-norm running statistics).
 * Batch-norm runs in train mode (batch statistics; running statistics momentum 0.1, unbiased variance) except on a
 * step of ONE pair, which the reference runs in eval mode (:140-156).  Dropout rates > 0 use counter-based masks
 * (seed, pair id, step), not torch's Philox stream. */
typedef struct kp_conve_params {
  float* ent;      /* [N, dim] */
  float* rel;      /* [R2, dim] */
  float* conv_w;   /* [n_filters, 1, 3, 3] */
  float* conv_b;   /* [n_filters] */
  float* fc_w;     /* [dim, hidden] */
  float* fc_b;     /* [dim] */
  float *bn1_w, *bn1_b, *bn1_mean, *bn1_var; /* [1] each */
  float *bn2_w, *bn2_b, *bn2_mean, *bn2_var; /* [n_filters] each */
  float *bn3_w, *bn3_b, *bn3_mean, *bn3_var; /* [dim] each */
  int32_t n_filters; /* 32 */
  int32_t hidden;    /* n_filters * 38 * (dim / 20 - 2) */
  float drop_input, drop_feature, drop_hidden;
} kp_conve_params;
typedef struct kp_vfit kp_vfit;
/* pairs: [n_pairs, 2] int32 (lhs, rel) in er_vocab order (:92-96); pos_off [n_pairs + 1] int64 / pos_ids int32: the
 * distinct objects of every pair (all three on the DEVICE, kept by the caller for the life of the handle). */
int kp_conve_fit_create(int device, int64_t n_entities, int64_t n_relations2, int32_t dim, const kp_conve_params* params,
                        float label_smoothing, int32_t max_batch, int64_t n_pairs, const int32_t* pairs,
                        const int64_t* pos_off, const int32_t* pos_ids, uint64_t dropout_seed, kp_vfit** out);
/* n_steps consecutive steps of one epoch at learning rate lr (the caller applies ExponentialLR between epochs, :125-126);
 * step k trains on pairs order[step_off[k] .. step_off[k+1]) -- order: [n_pairs] int32 (device), this epoch's shuffle;
 * step_off: host.  loss_out (device, nullable): [n_steps] mean BCE of every step. */
int kp_conve_fit_steps(kp_vfit* fit, int64_t n_steps, const int64_t* step_off, const int32_t* order, float lr,
                       float* loss_out, void* stream);
int kp_conve_fit_destroy(kp_vfit* fit);
const char* kp_conve_fit_error(const kp_vfit* fit); /* fit may be NULL: last create error */
int64_t kp_conve_fit_launches(const kp_vfit* fit);

/* Data-poisoning baseline (data_poisoning_engine.py:21-141) for ComplEx: per job (prediction, training fact,
 * perspective entity in {pred.s, pred.o}) the relevance of the fact = change of its score when the entity's embedding
 * moves by epsilon along -/+ the gradient of the prediction's score (necessary / sufficient).  preds / facts: [n, 3]
 * int32, entity: [n] int32, out: [n] fp32 -- all on the device.  Other model kinds: KP_EUNSUPPORTED (the reference's
 * engine calls Model.score_embeddings, which TransE and ConvE do not define). */
int kp_dp_relevance(kp_ctx* ctx, int32_t n_jobs, const int32_t* preds, const int32_t* facts, const int32_t* entity,
                    float epsilon, float lambd, int32_t sufficient, float* out, void* stream);

/* Host-side replay of the reference's per-epoch random draws (no device work; kp_host_rng.cu).  Both host generators
 * the Kelpie optimizers consume are 32-bit Mersenne Twisters -- torch's default CPU generator and numpy's legacy global
 * RandomState -- and each draw is a fixed function of consecutive output words.  key: the 624 state words, *pos: index of
 * the next word (624 = regenerate first); both are advanced in place exactly as the reference's calls would.
 *   kp_mt19937_words: the next `count` tempered words (out == NULL only advances; torch.randperm(n) consumes n - 1).
 *   kp_replay_transe_corruptions (pairwise_ranking_optimizer.py:171-172,187-195): per epoch torch.randint(high, (drawn,))
 *     then torch.randint(2, (drawn,)), of which the first `used` are kept:
 *     neg_code[e * used + i] = corrupting entity | head-corrupted << 31 (the compact table of kp_pt_batch), high <= 2^31;
 *     element = word % high, which is what torch (2.11) computes for high < 2^28 (two words per element above).
 *   kp_replay_numpy_shuffles (:167-168): `epochs` cumulative np.random.shuffle calls on one index vector of length n;
 *     perm[e * n + i] = its content after shuffle e (Fisher-Yates from the back, masked rejection sampling). */
int kp_mt19937_words(uint32_t* key, int32_t* pos, int64_t count, uint32_t* out);
int kp_replay_transe_corruptions(uint32_t* key, int32_t* pos, int32_t epochs, int64_t drawn, int64_t used, uint32_t high,
                                 int32_t* neg_code);
int kp_replay_numpy_shuffles(uint32_t* key, int32_t* pos, int32_t epochs, int32_t n, int32_t* perm);
/* Both of the above for one TransE job in one call (pairwise_ranking_optimizer.py:165-195): torch's generator travels as the
 * 5056-byte buffer of torch.get_rng_state() and is updated in place; pos_idx[e * n + i] = index of the positive of training
 * row i of epoch e (= shuffled row i / ratio), neg_code as kp_replay_transe_corruptions with drawn = ratio * n, used = n. */
int kp_replay_transe_job(uint32_t* np_key, int32_t* np_pos, uint8_t* torch_state, int64_t torch_state_bytes, int32_t epochs,
                         int32_t n, int32_t ratio, uint32_t high, int32_t* pos_idx, int32_t* neg_code);

/* Diagnostic: the fused score -> softmax (mode 0) / sigmoid (mode 1) -> contract pass alone, for
 * n_rows query vectors [n_rows, D] (device) against the resident entity table:
 *   out_m[g] = max_j z_gj (softmax: the reference max used, >= true max - 8),  out_l[g] = sum_j p_gj,
 *   out_O[g, :] = sum_j p_gj * E[j, :]   with p = exp(z - out_m) or sigmoid(z),  z = q_g . E[j].
 * This is the inner contraction of the ComplEx / ConvE post-training steps (SURVEY 9.3/9.4:
 * logsumexp and delta-a = sum_j G_ij E[j]); the tests check it against an fp64 restatement. */
int kp_debug_contract(kp_ctx* ctx, int32_t n_rows, const float* queries, int32_t mode, float* out_m,
                      float* out_l, float* out_O, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KELPIE_B200_H */
