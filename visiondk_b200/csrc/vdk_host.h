// vdk_host.h — host-side helpers shared by the translation units of libvdk_b200.so:
// error reporting (vdk_last_error_string), TMA descriptor encoding, device properties.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/vdk_b200.h"

namespace vdk {

// Thread-local last error text; every C-ABI entry point returns a VDK_* code and records why.
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

// Checks a CUDA runtime call inside a C-ABI entry point.
#define VDK_CUDA_OK(expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::vdk::fail(VDK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define VDK_REQUIRE(cond, ...)                                       \
  do {                                                               \
    if (!(cond)) return ::vdk::fail(VDK_ERR_INVALID, __VA_ARGS__);   \
  } while (0)

// Encodes a 2-D row-major tensor map for 16-bit elements: `rows` x `cols`, row pitch `ld` elements,
// box = box_rows x box_cols, 128-byte swizzle (box_cols * 2 bytes must be 128).
// Returns VDK_OK or an error code (driver entry point missing, bad alignment...).
int make_tma_2d_16bit(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);

// 4-D tile map over an NHWC 16-bit tensor: box = [1, box_h, box_w, box_c], no swizzle, out-of-bounds reads give 0
// (which is exactly the zero padding of a convolution).
int make_tma_nhwc_16bit(CUtensorMap* map, const void* base, int B, int H, int W, int C, int box_h, int box_w, int box_c);

// 3-D map over a 16-bit tensor [d2][d1][d0] (d0 contiguous; pitches in elements): box = [1][box_rows][64 elements], 128-byte
// swizzle, out-of-bounds rows read as zero.  The attention kernel's view of the qkv buffer: [batch][tokens][3*heads*64].
int make_tma_3d_16bit(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t pitch1, uint64_t pitch2,
                      uint32_t box_rows);

int sm_count();

// Live kernel timing inside a real step (bench.py's roofline): while a profile is open (vdk_prof_begin), every launch wrapped
// in a ProfScope is bracketed by two CUDA events on its own stream; vdk_prof_end sums launch count, milliseconds and the
// algorithmic FLOPs / bytes per category.  Closed (the default), a ProfScope costs one predictable branch.
enum ProfCategory { kProfGemm = 0, kProfDepthwise = 1, kProfAttention = 2, kProfScoreFilter = 3, kProfOther = 4, kProfCategories = 5 };
struct ProfScope {
  ProfScope(int category, double flops, double bytes, cudaStream_t stream);
  ~ProfScope();
  int slot;
  cudaStream_t stream;
};

}  // namespace vdk

namespace vdk {
// Internal form of vdk_gemm used by the composite entry points (convnext forward, heads).
int gemm_run(const vdk_gemm_desc& g, cudaStream_t stream);
}  // namespace vdk
