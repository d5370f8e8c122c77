// Ranked-list extraction and multi-ground-truth ranking for the result writers:
//   laff_topk_dense       top-K (K <= 2048) of every row of a dense fp32 score matrix, ordered by the tie rule
//                         (score desc, index desc) == np.argsort(kind='stable')[::-1][:K]
//                         (predictor.py:53-88 txt2video_write_to_file: `inds[index][::-1][0:TopK]`)
//   laff_rank_multi_gt    0-based rank of every ground-truth column of every row (CSR lists), same tie rule
//                         (predictor.py:262-270: video -> text, several captions per video)
//   laff_multi_gt_metrics evaluation.eval (evaluation.py:92-109) from those ranks: first-GT rank + AP per row, R@K/MedR/...
// All three are HBM-bound integer/compare work: rows are streamed with coalesced loads, no tensor cores involved.
#include <cstdint>

#include "host_util.cuh"

namespace laff {

constexpr int kSelThreads = 1024;
constexpr int kSelMaxK = LAFF_MAX_TOPK_DENSE;      // 2048
constexpr int kSortMaxCols = 16384;               // rows up to this many candidates are sorted whole in shared memory
constexpr int kRadixBits = 11;
constexpr int kRadixBins = 1 << kRadixBits;       // 2048 = 2 bins per thread

// Monotone map fp32 -> u32 (ascending): negative floats flip all bits, others flip the sign bit.
__device__ __forceinline__ uint32_t f2key(float f) {
  uint32_t u = __float_as_uint(f);
  if (u == 0x80000000u) u = 0u;  // -0.0 == +0.0 for the float comparisons numpy sorts by: a tie, not an order
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  const uint32_t u = k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu);
  return __uint_as_float(u);
}

// Descending bitonic sort of n (power of two) 64-bit keys in shared memory by the whole block.
__device__ void bitonic_sort_desc(unsigned long long* keys, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a < b) == desc) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void write_topk(const unsigned long long* keys, int n_valid, int k, float scale, float* out_val,
                                           int32_t* out_idx) {
  for (int t = threadIdx.x; t < k; t += blockDim.x) {
    if (t < n_valid) {
      out_val[t] = scale * key2f(static_cast<uint32_t>(keys[t] >> 32));
      out_idx[t] = static_cast<int32_t>(keys[t] & 0xFFFFFFFFull);
    } else {
      out_val[t] = -INFINITY;
      out_idx[t] = -1;
    }
  }
}

// Whole-row sort: cols <= kSortMaxCols.  idx_in (optional) supplies the index each candidate carries (merging shards).
__global__ void __launch_bounds__(kSelThreads) topk_sort_kernel(const float* __restrict__ scores, long long ld,
                                                                 const int32_t* __restrict__ idx_in, long long ld_idx, int cols,
                                                                 int n_pad, int k, float scale, float* __restrict__ out_val,
                                                                 int32_t* __restrict__ out_idx) {
  extern __shared__ unsigned long long s_keys[];
  const long long row = blockIdx.x;
  const float* s = scores + row * ld;
  int n_valid = 0;
  for (int j = threadIdx.x; j < n_pad; j += blockDim.x) {
    unsigned long long key = 0ull;  // sorts below every real candidate: real keys have a non-zero upper half unless -NaN
    if (j < cols) {
      const int32_t id = idx_in ? idx_in[row * ld_idx + j] : j;
      if (id >= 0) key = (static_cast<unsigned long long>(f2key(s[j])) << 32) | static_cast<uint32_t>(id);
    }
    s_keys[j] = key;
  }
  // number of real candidates (entries with idx -1 are padding from shards smaller than k)
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int local = 0;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) local += (idx_in ? idx_in[row * ld_idx + j] >= 0 : 1);
  if (local) atomicAdd(&s_cnt, local);
  bitonic_sort_desc(s_keys, n_pad);
  n_valid = s_cnt;
  write_topk(s_keys, n_valid, k, scale, out_val + row * k, out_idx + row * k);
}

// Long rows: 3-pass radix select (11 + 11 + 10 bits) of the k-th largest score, ordered collection of the winners
// (ties at the threshold taken by descending column), then a k-wide sort.
__global__ void __launch_bounds__(kSelThreads) topk_select_kernel(const float* __restrict__ scores, long long ld, long long cols,
                                                                   int k, int n_pad, float scale, float* __restrict__ out_val,
                                                                   int32_t* __restrict__ out_idx) {
  __shared__ unsigned long long s_keys[kSelMaxK];
  __shared__ int s_hist[kRadixBins];
  __shared__ int s_warp[kSelThreads / 32];
  __shared__ uint32_t s_prefix;
  __shared__ int s_need, s_taken_gt;
  const long long row = blockIdx.x;
  const float* s = scores + row * ld;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  if (tid == 0) {
    s_prefix = 0;
    s_need = k;
    s_taken_gt = 0;
  }
  // ---- radix select: after the passes s_prefix = key of the k-th largest score, s_need = how many of the elements
  //      equal to it belong to the top k ----
  const int shifts[3] = {21, 10, 0};
  const int widths[3] = {11, 11, 10};
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = shifts[pass], bins = 1 << widths[pass];
    for (int b = tid; b < kRadixBins; b += kSelThreads) s_hist[b] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t hi_mask = pass == 0 ? 0u : ~((1u << (shift + widths[pass])) - 1u);
    for (long long j = tid; j < cols; j += kSelThreads) {
      // cosine scores crowd a few exponent bins: aggregate equal digits inside the warp before touching the histogram
      const uint32_t key = f2key(s[j]);
      const bool valid = (key & hi_mask) == prefix;
      const uint32_t digit = valid ? ((key >> shift) & (bins - 1)) : 0xFFFFFFFFu;
      const unsigned peers = __match_any_sync(__activemask(), digit);
      if (valid && lane == __ffs(peers) - 1) atomicAdd(&s_hist[digit], __popc(peers));
    }
    __syncthreads();
    // suffix scan over the bins from the largest digit down: thread t owns descending positions 2t, 2t+1
    const int p0 = 2 * tid, p1 = 2 * tid + 1;
    const int h0 = p0 < bins ? s_hist[bins - 1 - p0] : 0;
    const int h1 = p1 < bins ? s_hist[bins - 1 - p1] : 0;
    int v = h0 + h1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    if (lane == 31) s_warp[wid] = v;
    __syncthreads();
    if (wid == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int incl1 = v + (wid ? s_warp[wid - 1] : 0);  // elements in positions <= p1
    const int incl0 = incl1 - h1;                        // ... <= p0
    const int excl0 = incl0 - h0;
    const int need = s_need;
    __syncthreads();
    if (h0 > 0 && excl0 < need && incl0 >= need) {
      s_prefix = prefix | (static_cast<uint32_t>(bins - 1 - p0) << shift);
      s_need = need - excl0;
    } else if (h1 > 0 && incl0 < need && incl1 >= need) {
      s_prefix = prefix | (static_cast<uint32_t>(bins - 1 - p1) << shift);
      s_need = need - incl0;
    }
    __syncthreads();
  }
  const uint32_t T = s_prefix;
  const int need_eq = s_need;
  const int n_gt = k - need_eq;
  // ---- collection, walking the row from its last column so that ties at T are met in descending column order ----
  for (int j = tid; j < n_pad; j += kSelThreads) s_keys[j] = 0ull;
  __syncthreads();
  int eq_base = 0;
  for (long long base = 0; base < cols; base += kSelThreads) {
    const long long j = cols - 1 - base - tid;
    uint32_t key = 0;
    bool gt = false, eq = false;
    if (j >= 0) {
      key = f2key(s[j]);
      gt = key > T;
      eq = key == T;
    }
    if (gt) {
      const int slot = atomicAdd(&s_taken_gt, 1);
      s_keys[slot] = (static_cast<unsigned long long>(key) << 32) | static_cast<uint32_t>(j);
    }
    const bool want_eq = eq && eq_base < need_eq;
    const int tile_eq = __syncthreads_count(want_eq);
    if (tile_eq) {
      const unsigned bal = __ballot_sync(0xffffffffu, want_eq);
      if (lane == 0) s_warp[wid] = __popc(bal);
      __syncthreads();
      int before = __popc(bal & ((1u << lane) - 1u));
      for (int w = 0; w < wid; ++w) before += s_warp[w];
      const int slot = eq_base + before;
      if (want_eq && slot < need_eq) s_keys[n_gt + slot] = (static_cast<unsigned long long>(key) << 32) | static_cast<uint32_t>(j);
      eq_base += tile_eq;
      __syncthreads();
    }
  }
  bitonic_sort_desc(s_keys, n_pad);
  write_topk(s_keys, k, k, scale, out_val + row * k, out_idx + row * k);
}

// rank0 of every ground-truth column of a row; gts processed 16 at a time with per-thread counters.
constexpr int kMultiGtChunk = 16;
__global__ void __launch_bounds__(256) rank_multi_gt_kernel(const float* __restrict__ scores, long long ld, long long cols,
                                                            const long long* __restrict__ gt_offsets,
                                                            const int32_t* __restrict__ gt_cols, int32_t* __restrict__ rank0) {
  __shared__ float s_sg[kMultiGtChunk];
  __shared__ int s_col[kMultiGtChunk];
  __shared__ int s_cnt[kMultiGtChunk];
  const long long row = blockIdx.x;
  const float* s = scores + row * ld;
  const long long g0 = gt_offsets[row], g1 = gt_offsets[row + 1];
  for (long long gb = g0; gb < g1; gb += kMultiGtChunk) {
    const int n = static_cast<int>(g1 - gb < kMultiGtChunk ? g1 - gb : kMultiGtChunk);
    __syncthreads();
    if (threadIdx.x < kMultiGtChunk) {
      const int c = threadIdx.x < n ? gt_cols[gb + threadIdx.x] : -1;
      s_col[threadIdx.x] = c;
      s_sg[threadIdx.x] = (c >= 0 && c < cols) ? s[c] : INFINITY;
      s_cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    int cnt[kMultiGtChunk];
#pragma unroll
    for (int g = 0; g < kMultiGtChunk; ++g) cnt[g] = 0;
    for (long long j = threadIdx.x; j < cols; j += blockDim.x) {
      const float v = s[j];
#pragma unroll
      for (int g = 0; g < kMultiGtChunk; ++g) cnt[g] += (v > s_sg[g]) || (v == s_sg[g] && j > s_col[g]);
    }
#pragma unroll
    for (int g = 0; g < kMultiGtChunk; ++g) {
      int c = cnt[g];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt[g], c);
    }
    __syncthreads();
    if (threadIdx.x < n) rank0[gb + threadIdx.x] = (s_col[threadIdx.x] >= 0 && s_col[threadIdx.x] < cols) ? s_cnt[threadIdx.x] : -1;
  }
}

// evaluation.eval per row from the ranks of its ground truths: first = min rank, AP = mean_i (i + 1) / (r_(i) + 1) with
// r_(i) the i-th smallest rank (ranks of distinct columns are distinct, so i = #{smaller ranks}).
__global__ void multi_gt_row_kernel(const int32_t* __restrict__ rank0, const long long* __restrict__ gt_offsets, int rows,
                                    int32_t* __restrict__ first, double* __restrict__ ap) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const long long g0 = gt_offsets[row], g1 = gt_offsets[row + 1];
  int best = -1;
  double acc = 0.0;
  long long n = 0;
  for (long long a = g0; a < g1; ++a) {
    const int ra = rank0[a];
    if (ra < 0) continue;
    int smaller = 0;
    for (long long b = g0; b < g1; ++b) smaller += (rank0[b] >= 0 && rank0[b] < ra);
    acc += static_cast<double>(smaller + 1) / static_cast<double>(ra + 1);
    best = (best < 0 || ra < best) ? ra : best;
    ++n;
  }
  first[row] = best;
  ap[row] = n ? acc / static_cast<double>(n) : 0.0;
}

__global__ void mean_into_kernel(const double* __restrict__ x, int n, double* __restrict__ dst) {
  __shared__ double s[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += x[i];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *dst = n ? s[0] / n : 0.0;
}

static int next_pow2(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace laff

using namespace laff;

extern "C" int laff_topk_dense(const float* scores, long long ld, const int32_t* idx_in, long long ld_idx, int rows,
                               long long cols, int k, float scale, float* out_val, int32_t* out_idx, void* stream) {
  LAFF_REQUIRE(scores && out_val && out_idx && rows >= 0 && cols >= 0 && ld >= cols, LAFF_EINVAL, "laff_topk_dense: bad arguments");
  LAFF_REQUIRE(k >= 1 && k <= kSelMaxK, LAFF_ENOTSUP, "laff_topk_dense: k must be in [1, %d] (got %d)", kSelMaxK, k);
  LAFF_REQUIRE(cols < (1LL << 31), LAFF_ENOTSUP, "laff_topk_dense: rows longer than 2^31 - 1");
  LAFF_REQUIRE(!idx_in || (cols <= kSortMaxCols && ld_idx >= cols), LAFF_ENOTSUP,
               "laff_topk_dense: candidate lists with explicit indices are limited to %d entries", kSortMaxCols);
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if (rows == 0) return LAFF_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cols <= kSortMaxCols) {
    const int n_pad = next_pow2(static_cast<int>(cols));
    const size_t smem = static_cast<size_t>(n_pad) * sizeof(unsigned long long);
    static bool configured = false;
    if (!configured) {
      LAFF_CUDA(cudaFuncSetAttribute(topk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kSortMaxCols * static_cast<int>(sizeof(unsigned long long))));
      configured = true;
    }
    topk_sort_kernel<<<rows, kSelThreads, smem, st>>>(scores, ld, idx_in, ld_idx, static_cast<int>(cols), n_pad, k, scale,
                                                       out_val, out_idx);
  } else {
    topk_select_kernel<<<rows, kSelThreads, 0, st>>>(scores, ld, cols, k, next_pow2(k), scale, out_val, out_idx);
  }
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_rank_multi_gt(const float* scores, long long ld, int rows, long long cols, const long long* gt_offsets,
                                  const int32_t* gt_cols, int32_t* rank0, void* stream) {
  LAFF_REQUIRE(scores && gt_offsets && rank0 && rows >= 0 && cols >= 0 && ld >= cols, LAFF_EINVAL,
               "laff_rank_multi_gt: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if (rows == 0) return LAFF_OK;
  rank_multi_gt_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(scores, ld, cols, gt_offsets, gt_cols, rank0);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_multi_gt_metrics(const int32_t* rank0, const long long* gt_offsets, int rows, int32_t* first, double* ap,
                                     double* out8, void* stream) {
  LAFF_REQUIRE(rank0 && gt_offsets && first && ap && out8 && rows > 0, LAFF_EINVAL, "laff_multi_gt_metrics: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  multi_gt_row_kernel<<<(rows + 127) / 128, 128, 0, st>>>(rank0, gt_offsets, rows, first, ap);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  rc = laff_rank_metrics(first, rows, out8, stream);
  if (rc) return rc;
  mean_into_kernel<<<1, 256, 0, st>>>(ap, rows, out8 + 6);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}
