// Tensor-core variant of fdm_self_attention for bf16 activations (placeholder until the mma path lands:
// reports "not handled" so the caller uses the exact-fp32 kernel).
#include "common.cuh"

int fdm_attention_mma_try(const fdm_attn_args& a, cudaStream_t stream, bool* handled) {
  (void)a; (void)stream;
  *handled = false;
  return 0;
}
