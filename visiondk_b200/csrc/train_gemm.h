// train_gemm.h — the GEMM forms of the training paths (convnext_train.cu, vit_train.cu) on top of gemm_run.
#pragma once
#include <algorithm>

#include "vdk_host.h"

namespace vdk {

// dst[i] (+)= sum over slabs, fixed order (convnext_train.cu)
int launch_slab_reduce(const float* slabs, int n_slabs, size_t stride, int64_t n4, float* dst, int accumulate, cudaStream_t s);

// split count for a weight-gradient GEMM (few output tiles, very long contraction): at least two, so that vdk_gemm
// takes its raw-partials output mode; every split stores its own fp32 slab, which a reduction kernel then adds in a
// fixed order (deterministic, and no atomics on the few hot output addresses).
inline int wgrad_splits(int M, int N, size_t K) {
  const int tiles = ((M + 127) / 128) * ((N + 255) / 256);
  const int want = std::max(2, (2 * sm_count()) / std::max(1, tiles));
  return vdk_gemm_effective_splits(static_cast<int>(K), want);
}
inline size_t wgrad_slab_bytes(int M, int N, size_t K) {
  return static_cast<size_t>(std::max(2, wgrad_splits(M, N, K))) * M * N * 4;
}

struct Gemm {
  cudaStream_t s;
  int run(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd, int epi, const float* bias,
          const float* gamma, const void* res, int ldr, int out_dtype, int split, long long stride, int ta, int tb,
          void* aux_out = nullptr) const {
    vdk_gemm_desc g{};
    g.A = A; g.B = B; g.D = D;
    g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldd = ldd;
    g.in_dtype = VDK_DTYPE_BF16; g.out_dtype = out_dtype; g.epilogue = epi;
    g.bias = bias; g.gamma = gamma; g.residual = res; g.ldr = ldr;
    g.ln_eps = 1e-6f; g.split_k = split; g.split_stride = stride; g.trans_a = ta; g.trans_b = tb; g.aux_out = aux_out;
    return gemm_run(g, s);
  }
  // weight gradient D[M,N] (+)= A^T B over a long K (A stored [K,M], B stored [K,N]): split-K partial slabs + fixed-order reduction
  int wgrad(const void* A, const void* B, float* D, int M, int N, int K, int lda, int ldb, float* slabs, bool accumulate) const {
    VDK_REQUIRE((static_cast<size_t>(M) * N) % 4 == 0, "wgrad: M*N must be a multiple of 4");
    const int split = wgrad_splits(M, N, static_cast<size_t>(K));
    const size_t stride = static_cast<size_t>(M) * N;
    int rc = run(A, B, slabs, M, N, K, lda, ldb, N, VDK_EPI_NONE, nullptr, nullptr, nullptr, 0, VDK_DTYPE_FP32, std::max(2, split),
                 static_cast<long long>(stride), 1, 1);
    if (rc != VDK_OK) return rc;
    return launch_slab_reduce(slabs, split, stride, static_cast<int64_t>(stride / 4), D, accumulate ? 1 : 0, s);
  }
};

#define RC(expr)                   \
  do {                             \
    int _rc = (expr);              \
    if (_rc != VDK_OK) return _rc; \
  } while (0)

}  // namespace vdk
