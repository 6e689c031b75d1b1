// ComplEx mimic post-training (placeholder until the softmax pass lands).
#include "kp_internal.h"

int kp_complex_post_train(kp_ctx* ctx, const kp_pt_batch*, const kp_hp*, cudaStream_t) {
  KP_FAIL(ctx, KP_EUNSUPPORTED, "ComplEx post-training is not built yet");
}
