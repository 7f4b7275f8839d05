// convnext_internal.h — launchers shared between the inference forward (convnext.cu) and the training
// forward/backward (convnext_train.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace vdk {

// mode 0: y = LayerNorm_C(dwconv7(x) + bias) (rstd_out optional: 1/sigma per pixel for the backward)
// mode 1: y = dwconv7(x) with the taps `w49` as given (+ addend): the backward-data pass uses reversed taps
int launch_dwconv7(int mode, const __nv_bfloat16* x, int batch, int H, int W, int C, const float* w49, const float* bias,
                   const float* ln_w, const float* ln_b, float eps, __nv_bfloat16* y, float* rstd_out,
                   const __nv_bfloat16* addend, cudaStream_t s);

// out = LayerNorm_C(x) per pixel, patch == 2: regrouped into 2x2/s2 patch rows (kh, kw, c); rstd_out optional
int launch_ln_patchify(const __nv_bfloat16* x, int B, int H, int W, int C, const float* ln_w, const float* ln_b, float eps,
                       int patch, __nv_bfloat16* out, float* rstd_out, cudaStream_t s);

// embeddings[row] = bias + sum of split-K slabs (fixed order) [, canonically L2-normalised]
int launch_neck_finalize(const float* slabs, int n_slabs, size_t slab_stride, int B, int F, const float* bias, int l2norm,
                         float* out, cudaStream_t s);

}  // namespace vdk

namespace vdk {
// ---- training-side launchers (train_ops.cu) ----
int launch_col_sum(const __nv_bfloat16* x, int64_t M, int C, int ld, float* out, cudaStream_t s);
int launch_ln_bwd(const __nv_bfloat16* dy, const __nv_bfloat16* y, const float* rstd, int B, int H, int W, int C,
                  const float* ln_w, const float* ln_b, int patch, __nv_bfloat16* dx, const __nv_bfloat16* addend,
                  float* dgamma, float* dbeta, cudaStream_t s);
int launch_dwconv7_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dconv, int B, int H, int W, int C, float* dw49,
                         float* dbias, cudaStream_t s);
int launch_permute021(const float* in, int A, int Bd, int Cd, const float* row_scale, __nv_bfloat16* out_bf16,
                      float* out_f32, int accumulate, cudaStream_t s);
int launch_cast_bf16(const float* in, int64_t n, __nv_bfloat16* out, cudaStream_t s);
int launch_layerscale_finalize(const float* G, const float* W2, const float* b2, const float* gamma, const float* sdo, int C,
                               int K4, float* dW2, float* dgamma, float* db2, cudaStream_t s);
int launch_bn_fwd_bf16(const __nv_bfloat16* x, int R, int C, const float* w, const float* b, float eps, float momentum,
                       __nv_bfloat16* y, float* save_mean, float* save_rstd, float* run_mean, float* run_var, cudaStream_t s);
int launch_bn_fwd_f32(const float* x, int R, int C, const float* w, const float* b, float eps, float momentum, float* y,
                      float* save_mean, float* save_rstd, float* run_mean, float* run_var, cudaStream_t s);
int launch_bn_bwd_bf16(const __nv_bfloat16* dy, const __nv_bfloat16* x, int R, int C, const float* w, const float* save_mean,
                       const float* save_rstd, __nv_bfloat16* dx, float* dweight, float* dbias, cudaStream_t s);
int launch_bn_bwd_f32(const float* dy, const float* x, int R, int C, const float* w, const float* save_mean,
                      const float* save_rstd, float* dx, float* dweight, float* dbias, cudaStream_t s);
int launch_add_f32(float* dst, const float* src, int64_t n, cudaStream_t s);
}  // namespace vdk
