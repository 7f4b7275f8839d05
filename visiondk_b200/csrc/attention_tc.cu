// attention_tc.cu — softmax(Q K^T / sqrt(d)) V on the 5th-generation tensor cores (tcgen05), head_dim 64, bf16.
//
// Replaces timm's Attention.forward inside every ViT block of the reference's backbone
// (models/faceX/backbone/timm_wrapper.py:52 -> timm VisionTransformer blocks): scores are never
// written to memory.  Input is the qkv Linear's output as stored, bf16 [B, N, 3, H, 64]; output bf16 [B, N, H*64]
// (+ optionally the log2-domain log-sum-exp per row, which the training backward consumes).
//
// One CTA = one (image, head) x TWO 128-query tiles (ViT-B/16's 197 tokens are exactly two), 320 threads:
//   warps 0-3   softmax group of query tile 0 (TMEM lane = query row: one thread owns one row, no shuffles)
//   warps 4-7   softmax group of query tile 1
//   warp  8     TMA producer: Q tiles once, then K/V tiles (128 keys) through a three-deep mbarrier ring
//   warp  9     tcgen05.mma issuer: S_t = Q_t K_j^T (128x128x64, both operands K-major) into TMEM, O_t += P_t V_j (128x64x128:
//               P from shared memory, written by the softmax group in the swizzle-128B K-major layout; V as stored =
//               MN-major operand)
// TMEM: S_0, S_1 (128 fp32 columns each), O_0, O_1 (64 each) = 384 of 512 columns.  The two groups ping-pong: while one
// exponentiates tile j, the tensor core computes the other group's P V and next S.  CTAs are PERSISTENT (one per SM, items
// strided over the grid): TMEM is allocated once, barrier phases run across items, and the producer loads the next item's Q
// and K/V while the current one is still in its softmax — the first version (one CTA per item) spent 3/4 of every CTA's life
// in allocation, barrier set-up and the first TMA round trip (152 TFLOP/s at 197 tokens).
// Softmax is the online recurrence in the log2 domain with a LAZY rescale: the running reference maximum only moves (and
// O / l are only rescaled, TMEM -> registers -> TMEM) when a row's maximum grows by more than 2^8; P = exp2(s - m_ref) then
// stays below 256, exact in bf16's range, and the final O / l cancels the stale reference.
#include "vdk_host.h"
#include "vdk_ptx.cuh"

#include <algorithm>
#include <cmath>

namespace vdk {

constexpr int kAtD = 64;          // head dim
constexpr int kAtQM = 128;        // query rows per tile (TMEM lanes)
constexpr int kAtKV = 128;        // keys per tile
constexpr int kAtStages = 3;
constexpr int kAtThreads = 320;
constexpr int kAtTile = kAtQM * kAtD * 2;  // 16 KB: a 128 x 64 bf16 tile
constexpr int kAtSmem = 2 * kAtTile + kAtStages * 2 * kAtTile + 2 * 2 * kAtTile + 16 * 8 + 16 + 1024;
static_assert(kAtSmem <= 227 * 1024, "attention shared memory budget");

struct AttParams {
  int B, N, H;
  float scale_log2e;
  __nv_bfloat16* out;
  float* lse2;
  int n_qtiles, n_kvtiles;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__global__ void __launch_bounds__(kAtThreads, 1)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const AttParams p) {
  extern __shared__ uint8_t att_smem_raw[];
  uint8_t* smem = att_smem_raw + ((1024u - (smem_u32(att_smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_q = smem;                                   // [2][16 KB]
  uint8_t* smem_kv = smem + 2 * kAtTile;                    // [stages][K 16 KB | V 16 KB]
  uint8_t* smem_p = smem_kv + kAtStages * 2 * kAtTile;      // [2 groups][2 K-blocks of 16 KB]
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem_p + 2 * 2 * kAtTile);
  uint64_t* q_empty = q_full + 1;
  uint64_t* kv_full = q_empty + 1;
  uint64_t* kv_empty = kv_full + kAtStages;
  uint64_t* s_full = kv_empty + kAtStages;   // [2]
  uint64_t* p_full = s_full + 2;             // [2]
  uint64_t* p_empty = p_full + 2;            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(p_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int J = p.n_kvtiles;
  const int n_pairs = (p.n_qtiles + 1) / 2;
  const int n_items = n_pairs * p.H * p.B;  // persistent: this CTA takes items blockIdx.x, + gridDim.x, ...

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_qkv);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < kAtStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);   // one arrival per warp of the group
      mbar_init(&p_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  auto decode = [&](int item, int& pair, int& h, int& b) {
    pair = item % n_pairs;
    const int r = item / n_pairs;
    h = r % p.H;
    b = r / p.H;
  };

  if (warp == 8) {
    // ===================== TMA producer: Q of the next item and its K/V tiles run ahead of the tensor core =====================
    if (lane == 0) {
      int qc = 0, kvc = 0;  // items / K-V tiles loaded so far (barrier phases run across items)
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++qc) {
        int pair, h, b;
        decode(item, pair, h, b);
        const int qt0 = pair * 2, n_t = min(2, p.n_qtiles - qt0);
        if (qc > 0) mbar_wait_relaxed(q_empty, (qc - 1) & 1);  // every S = Q K^T of the previous item has retired
        mbar_arrive_expect_tx(q_full, n_t * kAtTile);
        for (int t = 0; t < n_t; ++t) tma_load_3d(smem_q + t * kAtTile, &map_qkv, q_full, h * kAtD, (qt0 + t) * kAtQM, b);
        for (int j = 0; j < J; ++j, ++kvc) {
          const int st = kvc % kAtStages;
          if (kvc >= kAtStages) mbar_wait_relaxed(&kv_empty[st], ((kvc / kAtStages) - 1) & 1);
          mbar_arrive_expect_tx(&kv_full[st], 2 * kAtTile);
          tma_load_3d(smem_kv + st * 2 * kAtTile, &map_qkv, &kv_full[st], (p.H + h) * kAtD, j * kAtKV, b);
          tma_load_3d(smem_kv + st * 2 * kAtTile + kAtTile, &map_qkv, &kv_full[st], (2 * p.H + h) * kAtD, j * kAtKV, b);
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16<true>(kAtQM, kAtKV);        // Q, K both K-major
      constexpr uint32_t idesc_o = umma_idesc_f16<true>(kAtQM, kAtD, 0u, 1u);  // P K-major, V MN-major (d contiguous)
      int qc = 0, kvc = 0;
      int gt[2] = {0, 0};  // tiles of group t handed to the softmax warps so far (phases of s_full / p_full / p_empty)
      auto issue_s = [&](int t, int stage) {
        const uint64_t da = umma_desc_k_sw128(smem_u32(smem_q + t * kAtTile));
        const uint64_t db = umma_desc_k_sw128(smem_u32(smem_kv + stage * 2 * kAtTile));
        const uint32_t d = tmem_base + t * kAtKV;
#pragma unroll
        for (int k = 0; k < kAtD / 16; ++k) umma_f16_ss(d, da + 2 * k, db + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int stage, bool first) {
        const uint32_t pa = smem_u32(smem_p + t * 2 * kAtTile);
        const uint64_t db = umma_desc_mn_sw128(smem_u32(smem_kv + stage * 2 * kAtTile + kAtTile), 8192);
        const uint32_t d = tmem_base + 2 * kAtKV + t * kAtD;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t da = umma_desc_k_sw128(pa + kb * kAtTile);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(d, da + 2 * k, db + 128u * (kb * 4 + k), idesc_o, (!first || kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&p_empty[t]);
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++qc) {
        int pair, h, b;
        decode(item, pair, h, b);
        const int n_t = min(2, p.n_qtiles - pair * 2);
        mbar_wait(q_full, qc & 1);
        mbar_wait(&kv_full[kvc % kAtStages], (kvc / kAtStages) & 1);
        tc_fence_after();
        // S of the item's first key tile.  The previous item's last P of group t was consumed before its p_full arrival, which
        // the loop below waited for, so S_t may be overwritten.
        for (int t = 0; t < n_t; ++t) issue_s(t, kvc % kAtStages);
        if (J == 1) umma_commit(q_empty);
        for (int j = 0; j < J; ++j, ++kvc) {
          const int st = kvc % kAtStages;
          for (int t = 0; t < n_t; ++t) {
            mbar_wait(&p_full[t], gt[t] & 1);  // P_t of this tile is in shared memory; S_t has been consumed as well
            tc_fence_after();
            issue_pv(t, st, j == 0);
            ++gt[t];
            if (j + 1 < J) {
              if (t == 0) {
                mbar_wait(&kv_full[(kvc + 1) % kAtStages], ((kvc + 1) / kAtStages) & 1);
                tc_fence_after();
              }
              issue_s(t, (kvc + 1) % kAtStages);
              if (j + 2 == J && t == n_t - 1) umma_commit(q_empty);  // the item's last S MMAs are in flight: Q is free once they retire
            }
          }
          umma_commit(&kv_empty[st]);  // K_j / V_j free once every MMA issued so far has retired
        }
      }
    }
  } else {
    // ===================== softmax group t: one query row per thread =====================
    const int t = warp >> 2;
    const int lane_base = (warp & 3) * 32;
    const int r = lane_base + lane;                // row inside the tile
    const uint32_t s_addr = tmem_base + (static_cast<uint32_t>(lane_base) << 16) + t * kAtKV;
    const uint32_t o_addr = tmem_base + (static_cast<uint32_t>(lane_base) << 16) + 2 * kAtKV + t * kAtD;
    uint8_t* prow = smem_p + t * 2 * kAtTile + r * 128;
    const int rsw = r & 7;
    int gt = 0;  // tiles this group has processed (phases)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int pair, h, b;
      decode(item, pair, h, b);
      const int qt0 = pair * 2;
      if (t >= min(2, p.n_qtiles - qt0)) continue;   // odd tile count: the pair's second group idles for this item
      const int row = (qt0 + t) * kAtQM + r;         // token index of this query
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < J; ++j, ++gt) {
        mbar_wait(&s_full[t], gt & 1);
        tc_fence_after();
        const int valid = min(kAtKV, p.N - j * kAtKV);  // keys of this tile that exist
        // ---- pass A over TMEM: row maximum ----
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < kAtKV / 32; ++c) {
          if (c * 32 >= valid) break;
          uint32_t v[32];
          tmem_ld_32x32b_x32(s_addr + c * 32, v);
          tmem_ld_wait();
          if (c * 32 + 32 <= valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
        mx *= p.scale_log2e;
        if (j == 0) {
          m_ref = mx;
        } else {
          const float m_new = fmaxf(m_ref, mx);
          if (__any_sync(0xffffffffu, m_new - m_ref > 8.0f)) {
            // rescale O and l of the rows of this warp (alpha = 1 for rows whose maximum did not move)
            const float alpha = ex2_approx(m_ref - m_new);
            m_ref = m_new;
            l *= alpha;
            mbar_wait(&p_empty[t], (gt - 1) & 1);  // P V of the previous tile has retired: O is stable
            tc_fence_after();
#pragma unroll
            for (int hc = 0; hc < kAtD / 32; ++hc) {
              uint32_t o[32];
              tmem_ld_32x32b_x32(o_addr + hc * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32b_x32(o_addr + hc * 32, o);
            }
            tmem_st_wait();
          }
        }
        if (gt > 0) mbar_wait(&p_empty[t], (gt - 1) & 1);  // the P buffer is free (already passed if the rescale / epilogue waited)
        // ---- pass B: P = exp2(s * c - m_ref) -> bf16 -> shared memory (K-major, 128-byte swizzle), row sum in fp32 ----
        const float nm = -m_ref;
#pragma unroll 1
        for (int c = 0; c < kAtKV / 32; ++c) {
          uint32_t pk[16];
          if (c * 32 < valid) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(s_addr + c * 32, v);
            tmem_ld_wait();
            float e[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              e[i] = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2e, nm));
              if (c * 32 + 32 > valid && c * 32 + i >= valid) e[i] = 0.f;
            }
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              s0 += e[i]; s1 += e[i + 1]; s2 += e[i + 2]; s3 += e[i + 3];
            }
            l += (s0 + s1) + (s2 + s3);
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;  // keys beyond N: P must be exactly zero (never NaN garbage)
          }
          uint8_t* blk = prow + (c >> 1) * kAtTile;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = (c & 1) * 4 + q;  // 16-byte chunk inside the 128-byte row of this K-block
            *reinterpret_cast<uint4*>(blk + ((chunk ^ rsw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();  // generic-proxy stores of P -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      // ---- epilogue: O / l -> bf16 row ----
      mbar_wait(&p_empty[t], (gt - 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / l;
      // tcgen05.ld is .sync.aligned: EVERY lane of the warp executes it (rows beyond N included); only the stores are predicated
      __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * p.N + (row < p.N ? row : 0)) * (static_cast<size_t>(p.H) * kAtD) + h * kAtD;
#pragma unroll
      for (int hc = 0; hc < kAtD / 32; ++hc) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(o_addr + hc * 32, o);
        tmem_ld_wait();
        if (row < p.N) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv);
            w.y = pack_bf16x2(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv);
            w.z = pack_bf16x2(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv);
            w.w = pack_bf16x2(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + hc * 32 + q * 8) = w;
          }
        }
      }
      if (row < p.N && p.lse2) p.lse2[(static_cast<size_t>(b) * p.H + h) * p.N + row] = m_ref + log2f(l);
      tc_fence_before();  // this item's O reads are ordered before the p_full arrival that lets the next item's P V overwrite O
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// qkv bf16 [B, N, 3, H, 64] -> out bf16 [B, N, H*64]; lse2 (optional) fp32 [B, H, N] in the log2 domain
int launch_attention_tc(const __nv_bfloat16* qkv, int B, int N, int H, __nv_bfloat16* out, float* lse2, cudaStream_t s) {
  VDK_REQUIRE(B > 0 && N > 0 && H > 0 && H <= 65535 && B <= 65535, "attention: bad shape");
  static bool attr = false;
  if (!attr) {
    VDK_CUDA_OK(cudaFuncSetAttribute(attention_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    attr = true;
  }
  CUtensorMap map;
  const uint64_t pitch = 3ull * H * kAtD;
  int rc = make_tma_3d_16bit(&map, qkv, pitch, static_cast<uint64_t>(N), static_cast<uint64_t>(B), pitch, pitch * N, kAtQM);
  if (rc != VDK_OK) return rc;
  AttParams p{};
  p.B = B; p.N = N; p.H = H;
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(kAtD));
  p.out = out;
  p.lse2 = lse2;
  p.n_qtiles = (N + kAtQM - 1) / kAtQM;
  p.n_kvtiles = (N + kAtKV - 1) / kAtKV;
  const long long n_items = static_cast<long long>((p.n_qtiles + 1) / 2) * H * B;
  VDK_REQUIRE(n_items < (1ll << 31), "attention: too many (image, head, tile pair) items");
  const int grid = static_cast<int>(std::min<long long>(n_items, sm_count()));  // persistent: one CTA per SM
  attention_fwd_tc_kernel<<<grid, kAtThreads, kAtSmem, s>>>(map, p);
  VDK_CUDA_OK(cudaGetLastError());
  return VDK_OK;
}

}  // namespace vdk
