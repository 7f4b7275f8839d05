// retrieval.cu — exact inner-product top-k for the CBIR path: L2-normalise -> score -> select -> re-rank.
//
// Replaces F.normalize (models/faceX/face_model.py:139), faiss GpuIndexFlat add/search
// (engine/cbir/evaluation.py:155-168,190-195; cbir_eval.py:82-95,113-118).
//
// Pipeline (all on device, nothing leaves HBM):
//   rows_prepare   fp32 rows -> canonical unit rows (fp32) + fp16 copy + per-row rounding-error norm
//   score_filter   fp16 tcgen05 GEMM of a 128-query tile (resident in smem) against streamed gallery tiles;
//                  the epilogue never writes the score matrix: each thread owns one query row in TMEM and
//                  appends only scores >= tau[row] to that row's candidate list
//   select         per query: k-th largest approximate score A_k (radix select), keep a >= A_k - 2*eps
//                  (eps bounds |approx - canonical|, so the true top-k survive), tighten tau for the next
//                  gallery range; on the last range re-score survivors canonically (fp64, fixed order) and
//                  sort by (score desc, id asc)
// The gallery is scanned in geometrically growing ranges so that tau is tight when most of it streams by.
#include "vdk_host.h"
#include "vdk_ptx.cuh"

#include <cfloat>
#include <cmath>

namespace vdk {

// ------------------------------------------------------------------------------------------------
// canonical arithmetic (restated in oracle/retrieval.py; the two must agree bit for bit)
// ------------------------------------------------------------------------------------------------
// Fixed-order fp64 dot: lane l accumulates elements l, l+32, ... in order, then a 16/8/4/2/1 xor butterfly.
// Products of two fp32 values are exact in fp64, so fma(a,b,acc) and acc + a*b round identically.
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__device__ __forceinline__ uint32_t ord_u32(float f) {  // order-preserving float -> uint32
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord_u32(uint32_t o) {
  const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

// ------------------------------------------------------------------------------------------------
// rows_prepare: one warp per row
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rows_prepare_kernel(const float* __restrict__ x, int64_t n, int dim,
                                                           int normalize, float* __restrict__ xn,
                                                           __half* __restrict__ xh, float* __restrict__ row_norm,
                                                           float* __restrict__ row_err) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  const float* xr = x + row * dim;
  float denom = 1.0f;
  if (normalize) {
    double ss = 0.0;
    for (int i = lane; i < dim; i += 32) {
      const double v = static_cast<double>(xr[i]);
      ss = fma(v, v, ss);
    }
    ss = warp_sum_f64(ss);
    const float nrm = static_cast<float>(sqrt(ss));
    denom = fmaxf(nrm, 1e-12f);  // F.normalize eps
  }
  double s2 = 0.0, e2 = 0.0;
  for (int i = lane; i < dim; i += 32) {
    const float v = normalize ? __fdiv_rn(xr[i], denom) : xr[i];
    const __half h = __float2half_rn(v);
    const float d = v - __half2float(h);
    s2 = fma(static_cast<double>(v), static_cast<double>(v), s2);
    e2 = fma(static_cast<double>(d), static_cast<double>(d), e2);
    if (xn) xn[row * dim + i] = v;
    xh[row * dim + i] = h;
  }
  s2 = warp_sum_f64(s2);
  e2 = warp_sum_f64(e2);
  if (lane == 0) {
    // round the bounds up: they are used as upper bounds on ||xn|| and ||xn - xh||
    if (row_norm) row_norm[row] = __double2float_ru(sqrt(s2)) * 1.000001f;
    row_err[row] = __double2float_ru(sqrt(e2)) * 1.000001f + 1e-30f;
  }
}

__global__ void reduce_max_kernel(const float* __restrict__ x, int64_t n, float* out) {
  float m = -FLT_MAX;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    m = fmaxf(m, x[i]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  // values are non-negative norms: the int ordering of their bit patterns equals the float ordering
  if ((threadIdx.x & 31) == 0 && m >= 0.f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}

// ------------------------------------------------------------------------------------------------
// score_filter: fp16 tcgen05 GEMM with a threshold-filter epilogue
// ------------------------------------------------------------------------------------------------
constexpr int kQM = 128;          // queries per tile (TMEM lanes)
constexpr int kGN = 256;          // gallery rows per tile (UMMA N)
constexpr int kSBK = 64;          // K per smem block (128-byte swizzle row of fp16)
constexpr int kQBlockBytes = kQM * kSBK * 2;   // 16 KB
constexpr int kGStageBytes = kGN * kSBK * 2;   // 32 KB
constexpr int kGStages = 3;
constexpr int kScoreThreads = 192;
constexpr int kMaxKB = 8;         // dim <= 512

struct ScoreParams {
  int n_query;
  int num_kb;  // dim / 64
  int64_t g_lo, g_hi;
  int n_qtiles, n_splits, tiles_per_split, n_tiles;
  const float* tau;  // per-query admission threshold (sparse mode)
  uint2* cand;       // [n_query][cap] {score bits, gallery row}
  int cap;
  unsigned* counts;  // [n_query]
};

static int score_smem_bytes(int num_kb) {
  return num_kb * kQBlockBytes + kGStages * kGStageBytes + (2 * kGStages + 6) * 8 + 16 + 1024;
}

__device__ __noinline__ void cand_append(uint2* cand, unsigned* counts, int cap, int row, float a, uint32_t gidx) {
  const unsigned pos = atomicAdd(&counts[row], 1u);
  if (pos < static_cast<unsigned>(cap))
    cand[static_cast<size_t>(row) * cap + pos] = make_uint2(__float_as_uint(a), gidx);
}

template <bool kDense>
__global__ void __launch_bounds__(kScoreThreads, 1)
score_filter_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_g,
                    const ScoreParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_g = smem + p.num_kb * kQBlockBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_g + kGStages * kGStageBytes);
  uint64_t* empty_bar = full_bar + kGStages;
  uint64_t* tmem_full = empty_bar + kGStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* q_full = tmem_empty + 2;
  uint64_t* q_empty = q_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(q_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_units = p.n_qtiles * p.n_splits;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_q);
    prefetch_tensormap(&map_g);
    for (int i = 0; i < kGStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, uphase = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int qt = u % p.n_qtiles, sp = u / p.n_qtiles;
        const int t0 = sp * p.tiles_per_split;
        const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
        if (t0 >= t1) continue;
        // the query tile stays resident for the whole unit
        mbar_wait(q_empty, uphase ^ 1);
        mbar_arrive_expect_tx(q_full, p.num_kb * kQBlockBytes);
        for (int kb = 0; kb < p.num_kb; ++kb)
          tma_load_2d(smem_q + kb * kQBlockBytes, &map_q, q_full, kb * kSBK, qt * kQM, kEvictLast);
        uphase ^= 1;
        for (int t = t0; t < t1; ++t) {
          const int grow = static_cast<int>(p.g_lo) + t * kGN;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], kGStageBytes);
            tma_load_2d(smem_g + stage * kGStageBytes, &map_g, &full_bar[stage], kb * kSBK, grow, kEvictNormal);
            if (++stage == kGStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16<false>(kQM, kGN);
      int stage = 0, it = 0;
      uint32_t phase = 0, uphase = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int sp = u / p.n_qtiles;
        const int t0 = sp * p.tiles_per_split;
        const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
        if (t0 >= t1) continue;
        mbar_wait(q_full, uphase);
        uphase ^= 1;
        tc_fence_after();
        for (int t = t0; t < t1; ++t, ++it) {
          const int acc = it & 1;
          mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kGN;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t da = umma_desc_k_sw128(smem_u32(smem_q + kb * kQBlockBytes));
            const uint64_t db = umma_desc_k_sw128(smem_u32(smem_g + stage * kGStageBytes));
#pragma unroll
            for (int k = 0; k < kSBK / 16; ++k)
              umma_f16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (++stage == kGStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit(&tmem_full[acc]);
        }
        umma_commit(q_empty);  // query tile may be overwritten once every MMA of this unit has retired
      }
    }
  } else {
    // ===================== epilogue: one thread = one query row =====================
    const int lane_base = (warp & 3) * 32;
    int it = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int qt = u % p.n_qtiles, sp = u / p.n_qtiles;
      const int t0 = sp * p.tiles_per_split;
      const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
      if (t0 >= t1) continue;
      const int row = qt * kQM + lane_base + lane;
      const bool row_ok = row < p.n_query;
      float tau = INFINITY;
      if (!kDense && row_ok) tau = p.tau[row];
      for (int t = t0; t < t1; ++t, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_full[acc], (it >> 1) & 1);
        tc_fence_after();
        const int64_t gbase = p.g_lo + static_cast<int64_t>(t) * kGN;
#pragma unroll 1
        for (int c = 0; c < kGN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(lane_base) << 16) + acc * kGN + c * 32, r);
          tmem_ld_wait();
          const int64_t g0 = gbase + c * 32;
          if (kDense) {
            if (row_ok && g0 < p.g_hi) {
              uint2* dst = p.cand + static_cast<size_t>(row) * p.cap + (g0 - p.g_lo);
              const int64_t rem = p.g_hi - g0;
              const int nv = rem < 32 ? static_cast<int>(rem) : 32;
              if (nv == 32) {
#pragma unroll
                for (int j = 0; j < 32; j += 2)
                  *reinterpret_cast<uint4*>(dst + j) = make_uint4(r[j], static_cast<uint32_t>(g0 + j), r[j + 1],
                                                                  static_cast<uint32_t>(g0 + j + 1));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < nv) dst[j] = make_uint2(r[j], static_cast<uint32_t>(g0 + j));
              }
            }
          } else {
            float m = __uint_as_float(r[0]);
#pragma unroll
            for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
            if (m >= tau) {  // rare once tau is tight; rows beyond n_query carry tau = +inf
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float a = __uint_as_float(r[j]);
                if (a >= tau && g0 + j < p.g_hi)
                  cand_append(p.cand, p.counts, p.cap, row, a, static_cast<uint32_t>(g0 + j));
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[acc]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// select: one CTA per query
// ------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 256;
constexpr int kSortMax = 2048;  // survivors the final sort can hold

struct SelectParams {
  int n_query, dim, k, cap;
  uint2* cand;
  unsigned* counts;
  float* tau;
  const float* eps;  // per-query bound on |approx - canonical|
  int dense_n;       // > 0: counts[] are implied (dense stage wrote dense_n entries per row)
  int final_stage;
  const float* q32;
  const float* g32;
  int64_t id_offset;
  float* out_scores;
  int64_t* out_ids;
  int32_t* status;  // {overflow_rows, max_candidates, max_survivors, reserved}
};

// k-th largest of keys[0..n) (order-preserving uint32), n >= k >= 1.  All threads must call.
__device__ uint32_t block_kth_largest(const uint32_t* keys, int n, int k, unsigned* hist /*[256]*/,
                                      unsigned* bcast /*[2]*/) {
  uint32_t prefix = 0, mask = 0;
  int k_rem = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = keys[i];
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0, b = 255;
      for (; b > 0; --b) {
        if (acc + static_cast<int>(hist[b]) >= k_rem) break;
        acc += hist[b];
      }
      bcast[0] = static_cast<unsigned>(b);
      bcast[1] = static_cast<unsigned>(k_rem - acc);
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    mask |= 255u << shift;
    k_rem = static_cast<int>(bcast[1]);
    __syncthreads();
  }
  return prefix;
}

__global__ void __launch_bounds__(kSelThreads) select_kernel(const SelectParams p) {
  extern __shared__ uint8_t sel_smem[];
  uint32_t* s_key = reinterpret_cast<uint32_t*>(sel_smem);          // [cap]
  uint32_t* s_idx = s_key + p.cap;                                   // [cap]
  unsigned long long* s_sort = reinterpret_cast<unsigned long long*>(s_idx + p.cap);  // [kSortMax] (final only)
  __shared__ unsigned hist[256];
  __shared__ unsigned bcast[2];
  __shared__ unsigned s_m;

  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  const unsigned cnt = p.dense_n > 0 ? static_cast<unsigned>(p.dense_n) : p.counts[row];
  const int n = static_cast<int>(min(cnt, static_cast<unsigned>(p.cap)));
  uint2* rc = p.cand + static_cast<size_t>(row) * p.cap;
  if (tid == 0) {
    if (cnt > static_cast<unsigned>(p.cap)) atomicAdd(&p.status[0], 1);
    atomicMax(&p.status[1], static_cast<int>(min(cnt, 0x7fffffffu)));
    s_m = 0;
  }
  for (int i = tid; i < n; i += blockDim.x) {
    const uint2 e = rc[i];
    s_key[i] = ord_u32(__uint_as_float(e.x));
    s_idx[i] = e.y;
  }
  __syncthreads();

  // admission bound: everything within 2*eps below the k-th largest approximate score may be a true top-k member
  float tau_use = -INFINITY;
  if (n >= p.k) {
    const uint32_t kth = block_kth_largest(s_key, n, p.k, hist, bcast);
    tau_use = unord_u32(kth) - 2.0f * p.eps[row];
  }
  const uint32_t tau_key = ord_u32(tau_use);

  if (!p.final_stage) {
    // compact survivors to the front of the row's list; later ranges append behind them
    for (int i = tid; i < n; i += blockDim.x) {
      if (s_key[i] >= tau_key) {
        const unsigned pos = atomicAdd(&s_m, 1u);
        rc[pos] = make_uint2(__float_as_uint(unord_u32(s_key[i])), s_idx[i]);
      }
    }
    __syncthreads();
    if (tid == 0) {
      p.counts[row] = s_m;
      p.tau[row] = tau_use;
    }
    return;
  }

  // ---- final range: canonical re-score of the survivors, then sort by (score desc, id asc) ----
  // survivors are gathered into s_sort as (idx) first, then overwritten with sortable 64-bit keys
  for (int i = tid; i < n; i += blockDim.x) {
    if (s_key[i] >= tau_key) {
      const unsigned pos = atomicAdd(&s_m, 1u);
      if (pos < kSortMax) s_sort[pos] = s_idx[i];
    }
  }
  __syncthreads();
  const unsigned m_all = s_m;
  const int m = static_cast<int>(min(m_all, static_cast<unsigned>(kSortMax)));
  if (tid == 0) {
    if (m_all > kSortMax) atomicAdd(&p.status[0], 1);
    atomicMax(&p.status[2], static_cast<int>(m_all));
  }
  {
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const float* q = p.q32 + static_cast<size_t>(row) * p.dim;
    for (int c = warp; c < m; c += nwarps) {
      const uint32_t gi = static_cast<uint32_t>(s_sort[c]);
      const float* g = p.g32 + static_cast<size_t>(gi) * p.dim;
      double acc = 0.0;
      for (int i = lane; i < p.dim; i += 32) acc = fma(static_cast<double>(q[i]), static_cast<double>(g[i]), acc);
      acc = warp_sum_f64(acc);
      __syncwarp();
      if (lane == 0)
        s_sort[c] = (static_cast<unsigned long long>(ord_u32(static_cast<float>(acc))) << 32) |
                    static_cast<unsigned long long>(~gi);
    }
  }
  int m2 = 1;
  while (m2 < m) m2 <<= 1;
  for (int i = m + tid; i < m2; i += blockDim.x) s_sort[i] = 0ull;  // below every real key
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= m2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (m2 >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = s_sort[lo], b = s_sort[hi];
        if ((a < b) == desc) {
          s_sort[lo] = b;
          s_sort[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < p.k; j += blockDim.x) {
    float sc = -FLT_MAX;  // faiss pads inner-product results with lowest() and id -1
    int64_t id = -1;
    if (j < m) {
      const unsigned long long key = s_sort[j];
      sc = unord_u32(static_cast<uint32_t>(key >> 32));
      id = static_cast<int64_t>(~static_cast<uint32_t>(key & 0xffffffffull)) + p.id_offset;
    }
    p.out_scores[static_cast<size_t>(row) * p.k + j] = sc;
    p.out_ids[static_cast<size_t>(row) * p.k + j] = id;
  }
}

__global__ void eps_kernel(const float* __restrict__ q_norm, const float* __restrict__ q_err,
                           const float* __restrict__ g_norm_max, const float* __restrict__ g_err_max, int n,
                           float* __restrict__ eps, float* __restrict__ tau, unsigned* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gn = *g_norm_max, ge = *g_err_max;
  const float qn = q_norm[i] + q_err[i];
  // |approx - canonical| <= |dq.g| + |qh.dg| + tensor-core accumulation error ("error bound")
  const float e = q_err[i] * gn + qn * ge + 1.220703125e-4f /*2^-13*/ * qn * (gn + ge);
  eps[i] = e * 1.0001f + 1e-30f;
  tau[i] = -INFINITY;
  counts[i] = 0;
}

// ------------------------------------------------------------------------------------------------
// merge of per-shard lists, and brute-force pair scores for verification
// ------------------------------------------------------------------------------------------------
__global__ void topk_merge_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids, int n_lists,
                                  int64_t n_query, int k, float* __restrict__ out_scores,
                                  int64_t* __restrict__ out_ids) {
  // one warp per query; lists are individually ordered, so a k-step tournament over n_lists heads suffices
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_query) return;
  int head = 0;  // lane l owns list l (n_lists <= 32)
  const size_t list_stride = static_cast<size_t>(n_query) * k;
  for (int j = 0; j < k; ++j) {
    float sc = -FLT_MAX;
    int64_t id = -1;
    if (lane < n_lists && head < k) {
      sc = scores[lane * list_stride + row * k + head];
      id = ids[lane * list_stride + row * k + head];
    }
    // best = max score, then smallest non-negative id
    float bs = sc;
    int64_t bid = id;
    int bl = lane;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, off);
      const int64_t oid = __shfl_xor_sync(0xffffffffu, bid, off);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, off);
      const bool o_valid = oid >= 0, b_valid = bid >= 0;
      bool take;
      if (o_valid != b_valid) take = o_valid;
      else if (os != bs) take = os > bs;
      else if (oid != bid) take = oid < bid;
      else take = ol < bl;
      if (take) {
        bs = os;
        bid = oid;
        bl = ol;
      }
    }
    if (lane == 0) {
      out_scores[row * k + j] = bid >= 0 ? bs : -FLT_MAX;
      out_ids[row * k + j] = bid;
    }
    if (lane == bl && bid >= 0) ++head;
  }
}

__global__ void exact_pairs_kernel(const float* __restrict__ q32, const float* __restrict__ g32, int dim,
                                   const int64_t* __restrict__ qi, const int64_t* __restrict__ gi, int64_t n,
                                   float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const float* q = q32 + qi[w] * dim;
  const float* g = g32 + gi[w] * dim;
  double acc = 0.0;
  for (int i = lane; i < dim; i += 32) acc = fma(static_cast<double>(q[i]), static_cast<double>(g[i]), acc);
  acc = warp_sum_f64(acc);
  if (lane == 0) out[w] = static_cast<float>(acc);
}

static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// One gallery range [lo, hi) of the scan: the launch both vdk_ip_topk and vdk_score_range use.
static int launch_score_range(const CUtensorMap& mq, const CUtensorMap& mg, int nq, int dim, int64_t lo, int64_t hi,
                              bool dense, const float* tau, uint2* cand, int cap, unsigned* counts, cudaStream_t s) {
  static bool score_attr = false;
  if (!score_attr) {
    VDK_CUDA_OK(cudaFuncSetAttribute(score_filter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     score_smem_bytes(kMaxKB)));
    VDK_CUDA_OK(cudaFuncSetAttribute(score_filter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     score_smem_bytes(kMaxKB)));
    score_attr = true;
  }
  VDK_REQUIRE(lo % kGN == 0, "score range must start on a multiple of %d", kGN);
  if (dense) VDK_REQUIRE(hi - lo <= cap, "dense first range exceeds candidate capacity");
  const int sms = sm_count();
  ScoreParams p{};
  p.n_query = nq;
  p.num_kb = dim / kSBK;
  p.g_lo = lo;
  p.g_hi = hi;
  p.n_qtiles = (nq + kQM - 1) / kQM;
  p.n_tiles = static_cast<int>((hi - lo + kGN - 1) / kGN);
  int splits = (4 * sms + p.n_qtiles - 1) / p.n_qtiles;
  splits = std::max(1, std::min(splits, p.n_tiles));
  p.tiles_per_split = (p.n_tiles + splits - 1) / splits;
  p.n_splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.tau = tau;
  p.cand = cand;
  p.cap = cap;
  p.counts = counts;
  const int units = p.n_qtiles * p.n_splits;
  const int grid = std::min(units, sms);
  const int smem = score_smem_bytes(p.num_kb);
  if (dense)
    score_filter_kernel<true><<<grid, kScoreThreads, smem, s>>>(mq, mg, p);
  else
    score_filter_kernel<false><<<grid, kScoreThreads, smem, s>>>(mq, mg, p);
  VDK_CUDA_OK(cudaGetLastError());
  return VDK_OK;
}

struct TopkWorkspace {
  uint2* cand;
  unsigned* counts;
  float* tau;
  float* eps;
};
static size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
static TopkWorkspace carve_workspace(void* workspace, int64_t nq, int cap) {
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  TopkWorkspace w;
  w.cand = reinterpret_cast<uint2*>(ws);
  ws += align256(static_cast<size_t>(nq) * cap * sizeof(uint2));
  w.counts = reinterpret_cast<unsigned*>(ws);
  ws += align256(static_cast<size_t>(nq) * sizeof(float));
  w.tau = reinterpret_cast<float*>(ws);
  ws += align256(static_cast<size_t>(nq) * sizeof(float));
  w.eps = reinterpret_cast<float*>(ws);
  return w;
}

}  // namespace vdk

using namespace vdk;

extern "C" int vdk_rows_prepare(const float* x, int64_t n, int dim, int normalize, float* xn, void* xh,
                                float* row_norm, float* row_err, void* stream) {
  VDK_REQUIRE(x && xh && row_err, "vdk_rows_prepare: x, xh and row_err are required");
  VDK_REQUIRE(n >= 0 && dim > 0, "vdk_rows_prepare: bad shape n=%lld dim=%d", (long long)n, dim);
  if (n == 0) return VDK_OK;
  const int64_t warps_per_block = 256 / 32;
  const int64_t blocks = (n + warps_per_block - 1) / warps_per_block;
  VDK_REQUIRE(blocks < (1ll << 31), "vdk_rows_prepare: too many rows");
  rows_prepare_kernel<<<static_cast<unsigned>(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, n, dim, normalize, xn, reinterpret_cast<__half*>(xh), row_norm, row_err);
  VDK_CUDA_OK(cudaGetLastError());
  return VDK_OK;
}

extern "C" int vdk_reduce_max(const float* x, int64_t n, float* out, void* stream) {
  VDK_REQUIRE(x && out && n >= 0, "vdk_reduce_max: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  VDK_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float), s));
  if (n == 0) return VDK_OK;
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 1184));
  reduce_max_kernel<<<blocks, 256, 0, s>>>(x, n, out);
  VDK_CUDA_OK(cudaGetLastError());
  return VDK_OK;
}

extern "C" int vdk_topk_plan_default(vdk_topk_plan* plan, int64_t n_query, int64_t n_gallery, int dim, int k) {
  VDK_REQUIRE(plan, "vdk_topk_plan_default: null plan");
  VDK_REQUIRE(n_query >= 0 && n_gallery >= 0 && n_gallery < (1ll << 31), "vdk_topk_plan_default: bad sizes");
  VDK_REQUIRE(dim > 0 && dim % 64 == 0 && dim <= 64 * kMaxKB, "vdk_topk_plan_default: dim must be a multiple of 64, <= 512 (got %d)", dim);
  VDK_REQUIRE(k >= 1 && k <= 1024, "vdk_topk_plan_default: k must be in [1,1024] (got %d)", k);
  plan->n_query = n_query;
  plan->n_gallery = n_gallery;
  plan->dim = dim;
  plan->k = k;
  plan->cand_capacity = pow2_ceil(std::max(8192, 16 * k));
  // first range is scored densely (no threshold yet); each later range is 8x the prefix before it, so the
  // expected number of admitted candidates per range stays near 7k.
  int64_t end = std::min<int64_t>(n_gallery, std::max(4096, 4 * k));
  end = (end + kGN - 1) / kGN * kGN;
  if (end > plan->cand_capacity) end = plan->cand_capacity;
  int s = 0;
  for (; s < 8; ++s) {
    if (end >= n_gallery || s == 7) {
      plan->stage_end[s] = n_gallery;
      ++s;
      break;
    }
    plan->stage_end[s] = end;
    end *= 8;
  }
  plan->n_stages = s;
  for (int i = s; i < 8; ++i) plan->stage_end[i] = n_gallery;
  return VDK_OK;
}

extern "C" size_t vdk_topk_workspace_bytes(const vdk_topk_plan* plan) {
  if (!plan) return 0;
  const size_t nq = static_cast<size_t>(plan->n_query);
  return align256(nq * plan->cand_capacity * sizeof(uint2)) + 3 * align256(nq * sizeof(float)) + 256;
}

extern "C" int vdk_ip_topk(const vdk_topk_plan* plan, const float* q32, const void* qh, const float* q_norm,
                           const float* q_err, const float* g32, const void* gh, const float* g_norm_max,
                           const float* g_err_max, int64_t id_offset, float* out_scores, int64_t* out_ids,
                           int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  VDK_REQUIRE(plan && out_scores && out_ids && status, "vdk_ip_topk: null plan/output");
  const int64_t nq = plan->n_query, ng = plan->n_gallery;
  const int dim = plan->dim, k = plan->k, cap = plan->cand_capacity;
  VDK_REQUIRE(dim > 0 && dim % 64 == 0 && dim <= 64 * kMaxKB, "vdk_ip_topk: unsupported dim %d", dim);
  VDK_REQUIRE(k >= 1 && k <= 1024 && cap >= 2 * k && (cap & (cap - 1)) == 0, "vdk_ip_topk: bad k/capacity");
  VDK_REQUIRE(plan->n_stages >= 1 && plan->n_stages <= 8 && plan->stage_end[plan->n_stages - 1] == ng,
              "vdk_ip_topk: stage table must end at n_gallery");
  VDK_REQUIRE(nq < (1ll << 31) / kQM * kQM && ng < (1ll << 31), "vdk_ip_topk: sizes exceed 32-bit tiling");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  VDK_CUDA_OK(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), s));
  if (nq == 0) return VDK_OK;
  VDK_REQUIRE(q32 && qh && q_norm && q_err, "vdk_ip_topk: null query operand");
  VDK_REQUIRE(workspace && workspace_bytes >= vdk_topk_workspace_bytes(plan), "vdk_ip_topk: workspace too small");
  VDK_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "vdk_ip_topk: workspace must be 256-byte aligned");

  const TopkWorkspace w = carve_workspace(workspace, nq, cap);
  uint2* cand = w.cand;
  unsigned* counts = w.counts;
  float* tau = w.tau;
  float* eps = w.eps;

  SelectParams sp{};
  sp.n_query = static_cast<int>(nq);
  sp.dim = dim;
  sp.k = k;
  sp.cap = cap;
  sp.cand = cand;
  sp.counts = counts;
  sp.tau = tau;
  sp.eps = eps;
  sp.q32 = q32;
  sp.g32 = g32;
  sp.id_offset = id_offset;
  sp.out_scores = out_scores;
  sp.out_ids = out_ids;
  sp.status = status;
  const int sel_smem = cap * 8 + kSortMax * 8;
  static bool sel_attr = false;
  if (!sel_attr) {
    VDK_CUDA_OK(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    sel_attr = true;
  }
  VDK_REQUIRE(sel_smem <= 200 * 1024, "vdk_ip_topk: candidate capacity too large for the select kernel");

  if (ng == 0) {  // empty gallery: all results are padding
    VDK_CUDA_OK(cudaMemsetAsync(counts, 0, nq * sizeof(unsigned), s));
    sp.dense_n = 0;
    sp.final_stage = 1;
    static const float zero = 0.f;
    (void)zero;
    eps_kernel<<<(static_cast<int>(nq) + 255) / 256, 256, 0, s>>>(q_norm, q_err, q_err, q_err, static_cast<int>(nq),
                                                                   eps, tau, counts);
    select_kernel<<<static_cast<unsigned>(nq), kSelThreads, sel_smem, s>>>(sp);
    VDK_CUDA_OK(cudaGetLastError());
    return VDK_OK;
  }
  VDK_REQUIRE(g32 && gh && g_norm_max && g_err_max, "vdk_ip_topk: null gallery operand");

  eps_kernel<<<(static_cast<int>(nq) + 255) / 256, 256, 0, s>>>(q_norm, q_err, g_norm_max, g_err_max,
                                                                 static_cast<int>(nq), eps, tau, counts);
  VDK_CUDA_OK(cudaGetLastError());

  CUtensorMap mq, mg;
  int rc = make_tma_2d_16bit(&mq, qh, static_cast<uint64_t>(nq), dim, dim, kQM, kSBK);
  if (rc != VDK_OK) return rc;
  rc = make_tma_2d_16bit(&mg, gh, static_cast<uint64_t>(ng), dim, dim, kGN, kSBK);
  if (rc != VDK_OK) return rc;

  int64_t lo = 0;
  for (int st = 0; st < plan->n_stages; ++st) {
    const int64_t hi = plan->stage_end[st];
    VDK_REQUIRE(hi > lo || (hi == lo && st > 0), "vdk_ip_topk: stage table must be increasing");
    const bool dense = (st == 0);
    const bool last = (st == plan->n_stages - 1);
    if (hi > lo) {
      rc = launch_score_range(mq, mg, static_cast<int>(nq), dim, lo, hi, dense, tau, cand, cap, counts, s);
      if (rc != VDK_OK) return rc;
    }
    sp.dense_n = dense ? static_cast<int>(hi - lo) : 0;
    sp.final_stage = last ? 1 : 0;
    select_kernel<<<static_cast<unsigned>(nq), kSelThreads, sel_smem, s>>>(sp);
    VDK_CUDA_OK(cudaGetLastError());
    lo = hi;
  }
  return VDK_OK;
}

extern "C" int vdk_topk_merge(const float* scores, const int64_t* ids, int n_lists, int64_t n_query, int k,
                              float* out_scores, int64_t* out_ids, void* stream) {
  VDK_REQUIRE(scores && ids && out_scores && out_ids, "vdk_topk_merge: null operand");
  VDK_REQUIRE(n_lists >= 1 && n_lists <= 32 && k >= 1 && n_query >= 0, "vdk_topk_merge: n_lists must be in [1,32]");
  if (n_query == 0) return VDK_OK;
  const int64_t blocks = (n_query + 7) / 8;
  topk_merge_kernel<<<static_cast<unsigned>(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      scores, ids, n_lists, n_query, k, out_scores, out_ids);
  VDK_CUDA_OK(cudaGetLastError());
  return VDK_OK;
}

extern "C" int vdk_ip_exact_pairs(const float* q32, const float* g32, int dim, const int64_t* qi, const int64_t* gi,
                                  int64_t n, float* out, void* stream) {
  VDK_REQUIRE(q32 && g32 && qi && gi && out && dim > 0 && n >= 0, "vdk_ip_exact_pairs: bad arguments");
  if (n == 0) return VDK_OK;
  const int64_t blocks = (n + 7) / 8;
  exact_pairs_kernel<<<static_cast<unsigned>(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      q32, g32, dim, qi, gi, n, out);
  VDK_CUDA_OK(cudaGetLastError());
  return VDK_OK;
}

extern "C" int vdk_score_range(const vdk_topk_plan* plan, const void* qh, const void* gh, int64_t lo, int64_t hi,
                               int dense, void* workspace, size_t workspace_bytes, void* stream) {
  VDK_REQUIRE(plan && qh && gh && workspace, "vdk_score_range: null operand");
  VDK_REQUIRE(workspace_bytes >= vdk_topk_workspace_bytes(plan), "vdk_score_range: workspace too small");
  VDK_REQUIRE(lo >= 0 && hi > lo && hi <= plan->n_gallery, "vdk_score_range: bad range");
  const TopkWorkspace w = carve_workspace(workspace, plan->n_query, plan->cand_capacity);
  CUtensorMap mq, mg;
  int rc = make_tma_2d_16bit(&mq, qh, static_cast<uint64_t>(plan->n_query), plan->dim, plan->dim, kQM, kSBK);
  if (rc != VDK_OK) return rc;
  rc = make_tma_2d_16bit(&mg, gh, static_cast<uint64_t>(plan->n_gallery), plan->dim, plan->dim, kGN, kSBK);
  if (rc != VDK_OK) return rc;
  return launch_score_range(mq, mg, static_cast<int>(plan->n_query), plan->dim, lo, hi, dense != 0, w.tau, w.cand,
                            plan->cand_capacity, w.counts, reinterpret_cast<cudaStream_t>(stream));
}
