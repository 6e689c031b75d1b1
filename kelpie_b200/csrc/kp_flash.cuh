// Interface of the fused score->normalise->contract pass and the strip-merge helpers.
#pragma once
#include <cuda_runtime.h>

struct kp_ctx;
enum { KP_FLASH_SOFTMAX = 0, KP_FLASH_SIGMOID = 1 };

// Number of strips the pass will produce for G rows (partials are [n_strips, G(, D)]).
int kp_flash_plan(kp_ctx* ctx, int G, int* n_strips);
int kp_flash_simt(kp_ctx* ctx, const float* qmat, int G, int mode, float* part_m, float* part_l, float* part_O,
                  cudaStream_t st);

// tcgen05 version (kp_flash_umma.cu): same contract, used for >= 32 rows and D <= 512
bool kp_flash_umma_usable(kp_ctx* ctx, int G);
int kp_flash_umma_plan(kp_ctx* ctx, int G, int* n_strips);
int kp_flash_umma(kp_ctx* ctx, const float* qmat, int G, int mode, float* part_m, float* part_l, float* part_O,
                  cudaStream_t st);
// Dispatcher: runs the pass on the tensor cores when usable, else on CUDA cores; returns the
// number of strips written.  kp_flash_max_strips bounds it for buffer sizing.
int kp_flash_run(kp_ctx* ctx, const float* qmat, int G, int mode, float* part_m, float* part_l, float* part_O,
                 cudaStream_t st, int* n_strips);
int kp_flash_max_strips(kp_ctx* ctx);
size_t kp_flash_part_rows(kp_ctx* ctx, int G);

#ifdef __CUDACC__
// Merge the per-strip softmax statistics of row g: returns M = max_s m_s and L = sum_s l_s e^{m_s-M}.
__device__ __forceinline__ void kp_flash_merge_stats(const float* part_m, const float* part_l, int n_strips, int G,
                                                     int g, float& M, float& L) {
  M = -INFINITY;
  for (int s = 0; s < n_strips; ++s) M = fmaxf(M, part_m[(size_t)s * G + g]);
  L = 0.f;
  for (int s = 0; s < n_strips; ++s) {
    const float m = part_m[(size_t)s * G + g];
    if (m != -INFINITY) L += part_l[(size_t)s * G + g] * expf(m - M);
  }
}
#endif
