// Similarity + ranking kernels.
//   laff_sim_dense       : Q x V fp32 score matrix (small-config predict() path)            model/model.py:1003-1016
//   laff_sim_gt_scores   : s_i,gt by the same MMA sequence as the sweep                     predictor.py:239-244
//   laff_sim_rank_topk   : similarity GEMM fused with exact rank counting + streaming top-k predictor.py:232, evaluation.py:64-79
//   laff_topk_merge      : merge per-chunk / per-shard ordered lists
//   laff_rank_from_scores: the same tie rule applied to a materialised score matrix
//   laff_rank_metrics    : R@1/5/10, MedR, MeanR, MIR on device                             evaluation.py:81-89, :105-109
#include <cfloat>
#include <climits>
#include <cstdlib>
#include <cstring>

#include "gemm_engine.cuh"
#include "host_util.cuh"

namespace laff {

// ------------------------------------------------------------------------------------------------------------
// Epilogues
// ------------------------------------------------------------------------------------------------------------
struct EpiDense {
  struct Params {
    float* out;
    long long ld;
    int M, N;
    float scale;
  };
  static constexpr int kSmemBytes = 0;
  Params p;
  __device__ EpiDense(const Params& p_, uint8_t*, int) : p(p_) {}
  __device__ __forceinline__ void unit_begin(int, const Unit&) {}
  __device__ __forceinline__ void unit_end(int, const Unit&) {}
  __device__ __forceinline__ void chunk(const uint32_t (&r)[32], int row, int col0) {
    if (row >= p.M || col0 >= p.N) return;
    float* dst = p.out + static_cast<long long>(row) * p.ld + col0;
    if (col0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 v = make_float4(__uint_as_float(r[j]) * p.scale, __uint_as_float(r[j + 1]) * p.scale,
                               __uint_as_float(r[j + 2]) * p.scale, __uint_as_float(r[j + 3]) * p.scale);
        *reinterpret_cast<float4*>(dst + j) = v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) dst[j] = __uint_as_float(r[j]) * p.scale;
    }
  }
};

// Diagnostics only: drains TMEM (mode 1) or skips the loads entirely (mode 0) so the mainloop can be timed alone.
struct EpiNull {
  struct Params {
    float* sink;
    int mode;
  };
  static constexpr int kSmemBytes = 0;
  Params p;
  float acc;
  __device__ EpiNull(const Params& p_, uint8_t*, int) : p(p_), acc(0.f) {}
  __device__ __forceinline__ void unit_begin(int, const Unit&) {}
  __device__ __forceinline__ void unit_end(int row, const Unit&) {
    if (acc == 123.456f) p.sink[row & 1023] = acc;
  }
  __device__ __forceinline__ void chunk(const uint32_t (&r)[32], int, int) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc = fmaxf(acc, __uint_as_float(r[j]));
  }
};

// D[r, r] of A x gather(B)[r]: the raw accumulator the sweep would produce for (query r, its ground-truth video).
struct EpiDiag {
  struct Params {
    float* sgt;
    const int32_t* gt_local;
    int M;
  };
  static constexpr int kSmemBytes = 0;
  Params p;
  __device__ EpiDiag(const Params& p_, uint8_t*, int) : p(p_) {}
  __device__ __forceinline__ void unit_begin(int, const Unit&) {}
  __device__ __forceinline__ void unit_end(int, const Unit&) {}
  __device__ __forceinline__ void chunk(const uint32_t (&r)[32], int row, int col0) {
    if (row >= p.M || row < col0 || row >= col0 + 32) return;
    const int j = row - col0;
    uint32_t v = 0;
#pragma unroll
    for (int t = 0; t < 32; ++t)
      if (t == j) v = r[t];
    p.sgt[row] = p.gt_local[row] >= 0 ? __uint_as_float(v) : 0.0f;
  }
};

// total order used everywhere: (score desc, index desc)
__device__ __forceinline__ bool better(float va, int ia, float vb, int ib) {
  return va > vb || (va == vb && ia > ib);
}

// Monotone float <-> int key so that a per-query score threshold can be raised with atomicMax.
__device__ __forceinline__ int float_key(float v) {
  const int b = __float_as_int(v);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// (score, index) packed so that "better" under the tie rule (score desc, index desc) is a plain signed 64-bit ">".
__device__ __forceinline__ long long pack_entry(float v, int idx) {
  return (static_cast<long long>(float_key(v)) << 32) | static_cast<long long>(static_cast<unsigned>(idx));
}
constexpr long long kEmptySlot = static_cast<long long>(0x8000000000000000ull);  // below every real entry

// Out-of-line slow path of the rank epilogue: offer entry e to the query's k global slots (an unordered set that
// converges to the exact top-k).  Lock-free: replace the current minimum slot by CAS while e beats it.  Slots only
// ever improve, so any value read here is a valid lower bound of the final k-th best.  Returns the new threshold.
static __device__ __noinline__ float topk_offer(long long* slots, int32_t* thr_key, long long e, int k) {
  for (;;) {
    long long mn = 0x7fffffffffffffffLL, mn2 = 0x7fffffffffffffffLL;
    int at = 0;
    for (int i = 0; i < k; ++i) {
      const long long s = __ldcg(slots + i);
      if (s < mn) {
        mn2 = mn;
        mn = s;
        at = i;
      } else if (s < mn2) {
        mn2 = s;
      }
    }
    if (e <= mn) return key_float(static_cast<int>(mn >> 32));  // not (or no longer) among the k best
    const long long old = atomicCAS(reinterpret_cast<unsigned long long*>(slots + at),
                                    static_cast<unsigned long long>(mn), static_cast<unsigned long long>(e));
    if (old == mn) {
      const long long new_min = e < mn2 ? e : mn2;  // k == 1: mn2 stays at +max, so new_min = e
      const int key = static_cast<int>(new_min >> 32);
      if (new_min != kEmptySlot) atomicMax(thr_key, key);
      return new_min == kEmptySlot ? -INFINITY : key_float(key);
    }
  }
}

// Similarity sweep epilogue: exact rank counting + streaming exact top-k.  One epilogue thread owns (query row,
// 128-column half of each tile) for all tiles of a unit.
//   fast path per score: compare-and-count against s_gt, and two compares OR-ed into one "look closer" predicate;
//   one branch per 32 scores;
//   slow path (a score reaches the query's threshold, or ties s_gt exactly): tie rule + topk_offer.
// The threshold of a query is a published lower bound of its current k-th best over ALL gallery columns any CTA has
// swept so far, so the slow path dies out like k/columns_seen.
template <int KMAX>
struct EpiRank {
  struct Params {
    const float* sgt;     // raw accumulator of (i, gt_i)
    const int32_t* gt;    // global gallery index of the ground truth
    int32_t* count;       // [M] += local rank contribution
    int32_t* thr_key;     // [M] float_key of a lower bound of the k-th best (init: float_key(-inf))
    long long* slots;     // [M, KMAX] packed entries (init: kEmptySlot)
    int M, N;             // queries, local gallery size
    int col_offset;       // global index of local column 0
    int k;                // top-k size (<= KMAX); 0: rank only
  };
  static constexpr int kSmemBytes = 0;
  Params p;
  int cnt;
  float sg;    // s_gt
  int g;       // gt (global index)
  float thr;   // current threshold (lower bound of the k-th best)
  long long* my_slots;
  int32_t* my_thr;

  __device__ EpiRank(const Params& p_, uint8_t*, int) : p(p_) {}

  __device__ __forceinline__ void unit_begin(int row, const Unit&) {
    cnt = 0;
    sg = INFINITY;  // rows beyond M: nothing counts, nothing is offered
    g = -1;
    thr = INFINITY;
    my_slots = nullptr;
    my_thr = nullptr;
    if (row < p.M) {
      sg = p.sgt[row];
      g = p.gt[row];
      if (p.k > 0) {
        my_slots = p.slots + static_cast<long long>(row) * KMAX;
        my_thr = p.thr_key + row;
        thr = key_float(__ldcg(my_thr));
      }
    }
  }

  __device__ __forceinline__ void slow_chunk(const uint32_t (&r)[32], int gcol0) {
    if (my_thr != nullptr) thr = fmaxf(thr, key_float(__ldcg(my_thr)));  // pick up other CTAs' progress
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float v = __uint_as_float(r[j]);
      if (v >= thr || v == sg) {
        const int gcol = gcol0 + j;
        if (v == sg && gcol > g) ++cnt;  // tie with the ground truth: higher index ranks first
        if (v >= thr && v > -INFINITY) thr = fmaxf(thr, topk_offer(my_slots, my_thr, pack_entry(v, gcol), p.k));
      }
    }
  }

  __device__ __forceinline__ void chunk(uint32_t (&r)[32], int row, int col0) {
    if (col0 + 32 > p.N) {  // last, partial tile: columns past the gallery read as zero -> make them -inf
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j >= p.N) r[j] = 0xff800000u;
    }
    bool look = false;
    int c = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float v = __uint_as_float(r[j]);
      c += (v > sg) ? 1 : 0;
      look = look || (v >= thr) || (v == sg);
    }
    cnt += c;
    if (look) slow_chunk(r, col0 + p.col_offset);
  }

  __device__ __forceinline__ void unit_end(int row, const Unit&) {
    if (row < p.M && cnt) atomicAdd(p.count + row, cnt);
  }
};

// N1 (writer lists, k up to 2048) without the dense matrix: every score >= the query's threshold is appended, unordered,
// to the query's candidate list.  With a threshold a little below the k-th best (estimated from a sample of the
// gallery) ~2k of the V scores survive, so the Q x V matrix is never written and the exact top-k is a sort of the
// survivors.  count may end above cap (list truncated) or below k: the caller checks both and falls back.
// Optionally (sgt != nullptr) the same pass also counts the ground truth's rank exactly as EpiRank does (compare-and-count
// against the raw score of the ground truth + the tie rule), so a caller that wants the long lists AND the ranks
// (predictor.py:232-259: metrics and t2v.pkl of one query set) sweeps the gallery once instead of twice.
struct EpiCollect {
  struct Params {
    const float* thr;     // [M] threshold on the scaled score
    int32_t* count;       // [M] number of scores >= thr seen so far
    float* cand_val;      // [M, cap]
    int32_t* cand_idx;    // [M, cap] global gallery index
    int cap;
    int M, N;
    int col_offset;
    float scale;
    const float* sgt;     // optional [M]: raw accumulator of (i, gt_i)
    const int32_t* gt;    // optional [M]: global gallery index of the ground truth
    int32_t* rank_count;  // optional [M] += number of local videos ranked above the ground truth
  };
  static constexpr int kSmemBytes = 0;
  Params p;
  float t;
  float sg;
  int g, cnt;
  __device__ EpiCollect(const Params& p_, uint8_t*, int) : p(p_) {}
  __device__ __forceinline__ void unit_begin(int row, const Unit&) {
    t = row < p.M ? __ldg(p.thr + row) : INFINITY;
    cnt = 0;
    sg = INFINITY;
    g = -1;
    if (p.sgt != nullptr && row < p.M) {
      sg = __ldg(p.sgt + row);
      g = __ldg(p.gt + row);
    }
  }
  __device__ __forceinline__ void unit_end(int row, const Unit&) {
    if (p.sgt != nullptr && row < p.M && cnt) atomicAdd(p.rank_count + row, cnt);
  }
  __device__ __forceinline__ void chunk(const uint32_t (&r)[32], int row, int col0) {
    uint32_t hit = 0;
    const uint32_t valid = (col0 + 32 > p.N) ? ((col0 >= p.N) ? 0u : (0xffffffffu >> (32 - (p.N - col0)))) : 0xffffffffu;
    if (p.sgt != nullptr) {   // rank of the ground truth: #{v > s_gt} + #{v == s_gt, index > gt} (columns past the gallery excluded)
      uint32_t above = 0, tie = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float v = __uint_as_float(r[j]);
        above |= (v > sg) ? (1u << j) : 0u;
        tie |= (v == sg) ? (1u << j) : 0u;
      }
      cnt += __popc(above & valid);
      tie &= valid;
      while (tie) {
        const int j = __ffs(tie) - 1;
        tie &= tie - 1;
        if (col0 + j + p.col_offset > g) ++cnt;
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) hit |= (__uint_as_float(r[j]) * p.scale >= t) ? (1u << j) : 0u;
    hit &= valid;  // columns past the gallery read as zero
    if (hit == 0) return;
    const int n = __popc(hit);
    const int base = atomicAdd(p.count + row, n);
    float* cv = p.cand_val + static_cast<long long>(row) * p.cap;
    int32_t* ci = p.cand_idx + static_cast<long long>(row) * p.cap;
    int o = base;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if ((hit >> j) & 1u) {
        if (o < p.cap) {
          cv[o] = __uint_as_float(r[j]) * p.scale;
          ci[o] = col0 + j + p.col_offset;
        }
        ++o;
      }
    }
  }
};

__global__ void collect_init_kernel(int32_t* count, float* cand_val, int32_t* cand_idx, int Q, long long total) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long j = i; j < Q; j += stride) count[j] = 0;   // the grid is capped: Q may exceed it
  for (long long j = i; j < total; j += stride) {
    cand_val[j] = -INFINITY;
    cand_idx[j] = -1;
  }
}

__global__ void rank_init_kernel(int32_t* count, int32_t* thr_key, long long* slots, int Q, int kmax) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Q) {
    count[i] = 0;
    thr_key[i] = float_key(-INFINITY);
  }
  if (i < Q * kmax) slots[i] = kEmptySlot;
}

// Order each query's k slots by the tie rule and unpack: one thread per query, k <= 16.
__global__ void topk_finalize_kernel(const long long* __restrict__ slots, int Q, int kmax, int k, float scale,
                                     float* __restrict__ out_val, int32_t* __restrict__ out_idx) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  long long e[LAFF_MAX_TOPK];
#pragma unroll
  for (int i = 0; i < LAFF_MAX_TOPK; ++i) e[i] = i < k ? slots[static_cast<long long>(q) * kmax + i] : kEmptySlot;
#pragma unroll
  for (int i = 1; i < LAFF_MAX_TOPK; ++i) {  // insertion sort, descending, fully unrolled (registers)
#pragma unroll
    for (int j = i; j > 0; --j) {
      const long long a = e[j - 1], b = e[j];
      e[j - 1] = a > b ? a : b;
      e[j] = a > b ? b : a;
    }
  }
#pragma unroll
  for (int i = 0; i < LAFF_MAX_TOPK; ++i) {
    if (i < k) {
      const bool empty = e[i] == kEmptySlot;
      out_val[static_cast<long long>(q) * k + i] = empty ? -INFINITY : key_float(static_cast<int>(e[i] >> 32)) * scale;
      out_idx[static_cast<long long>(q) * k + i] = empty ? -1 : static_cast<int>(e[i] & 0xffffffffLL);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Launch helper
// ------------------------------------------------------------------------------------------------------------
static uint64_t g_hint_override[2] = {0, 0};

// Gallery column tiles per sweep launch.  The CTA pairs that share a gallery tile drift apart as a launch goes on (their
// epilogue work is data dependent), the shared tile then misses L2 for the late ones, and on a power-capped kernel the
// extra HBM traffic costs SM clock.  A kernel boundary re-aligns them: cutting a 1 M-video pass into 8 launches of ~480
// column tiles measured 69.3 / 68.6 ms -> 60.8 ms per 10 k x 1 M sweep on the same box (2: 62.1, 4: 60.9, 16: 61.3;
// profiles/r01_sched_experiments.md).  Launches stay >= 480 tiles (~65 waves of work units) so their tails remain small.
// LAFF_SWEEP_COLSPLIT=n (read once) forces n launches per pass (diagnostics).
static int sweep_tiles_per_launch(int tiles) {
  static const int forced = [] { const char* e = getenv("LAFF_SWEEP_COLSPLIT"); return e ? atoi(e) : 0; }();
  int split = forced > 0 ? forced : tiles / 480;
  if (split < 1) split = 1;
  return (tiles + split - 1) / split;
}

template <int CG, class Epi>
static int launch_gemm_cg(const CUtensorMap& tmA, const CUtensorMap& tmB, int num_kb, uint32_t idesc, const Sched& s,
                          const typename Epi::Params& ep, int sms, cudaStream_t st) {
  // queries (A) are re-read by every gallery tile: keep them in L2; the gallery (B) streams through once per m-group
  uint64_t hintA = ptx::kEvictLast, hintB = ptx::kEvictNormal;
  if (g_hint_override[0]) hintA = g_hint_override[0];
  if (g_hint_override[1]) hintB = g_hint_override[1];
  LAFF_CUDA((launch_gemm_kernel<CG, Epi>(tmA, tmB, num_kb, idesc, s, ep, hintA, hintB, sms, st)));
  return LAFF_OK;
}

struct GemmOperands {
  CUtensorMap tmA, tmB;
  int num_kb;
  uint32_t idesc;
  int cg;
  int sms;
};

// A [M, K] pitch lda, B [N, K] pitch ldb
static int prepare_operands(GemmOperands* op, const void* A, const void* B, long long M, long long N, int K,
                            long long lda, long long ldb, int dtype, int cg) {
  LAFF_REQUIRE(is16(dtype), LAFF_EINVAL, "operand dtype must be LAFF_F16 or LAFF_BF16, got %d", dtype);
  LAFF_REQUIRE(M > 0 && N > 0 && K > 0, LAFF_EINVAL, "empty GEMM operand (M=%lld N=%lld K=%d)", M, N, K);
  LAFF_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && lda >= K && ldb >= K, LAFF_EINVAL,
               "K (%d) and operand pitches (%lld, %lld) must be multiples of 8 and pitches >= K", K, lda, ldb);
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  op->cg = cg;
  op->sms = di.sms;
  op->num_kb = (K + kBlockK - 1) / kBlockK;
  op->idesc = make_idesc_f16(dtype, kBlockM * cg, kBlockN);
  rc = make_tmap_2d(&op->tmA, A, dtype, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), kBlockM);
  if (rc) return rc;
  rc = make_tmap_2d(&op->tmB, B, dtype, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldb),
                    static_cast<uint32_t>(kBlockN / cg));
  return rc;
}

template <class Epi>
static int launch_gemm(const GemmOperands& op, const Sched& s, const typename Epi::Params& ep, cudaStream_t st) {
  if (op.cg == 2) return launch_gemm_cg<2, Epi>(op.tmA, op.tmB, op.num_kb, op.idesc, s, ep, op.sms, st);
  return launch_gemm_cg<1, Epi>(op.tmA, op.tmB, op.num_kb, op.idesc, s, ep, op.sms, st);
}

// ------------------------------------------------------------------------------------------------------------
// Small CUDA-core kernels around the GEMM
// ------------------------------------------------------------------------------------------------------------
__global__ void gather_rows16_kernel(const uint4* __restrict__ g, long long ldg16, const int32_t* __restrict__ gt_local,
                                     int Q, int vec_per_row, uint4* __restrict__ out) {
  const long long total = static_cast<long long>(Q) * vec_per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int row = static_cast<int>(i / vec_per_row);
    const int v = static_cast<int>(i - static_cast<long long>(row) * vec_per_row);
    int src = gt_local[row];
    if (src < 0) src = 0;
    out[i] = g[static_cast<long long>(src) * ldg16 + v];
  }
}

// One warp per query: k_out rounds of "best candidate strictly worse than the previous pick".
__global__ void topk_merge_kernel(const float* __restrict__ vals, const int32_t* __restrict__ idx, int n_lists, int Q,
                                  int k_in, long long list_stride, int row_stride, int k_out, float in_scale,
                                  float* __restrict__ out_val, int32_t* __restrict__ out_idx) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= Q) return;
  const int n_cand = n_lists * k_in;
  float last_v = INFINITY;
  int last_i = INT_MAX;
  for (int r = 0; r < k_out; ++r) {
    float bv = -INFINITY;
    int bi = -1;
    for (int c = lane; c < n_cand; c += 32) {
      const int l = c / k_in, e = c - l * k_in;
      const long long o = static_cast<long long>(l) * list_stride + static_cast<long long>(warp) * row_stride + e;
      const float v = vals[o];
      const int i = idx[o];
      if (i >= 0 && better(last_v, last_i, v, i) && better(v, i, bv, bi)) {
        bv = v;
        bi = i;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (better(ov, oi, bv, bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      out_val[static_cast<long long>(warp) * k_out + r] = bi >= 0 ? bv * in_scale : -INFINITY;
      out_idx[static_cast<long long>(warp) * k_out + r] = bi;
    }
    if (bi < 0) {
      // nothing left: fill the tail
      for (int t = r + 1 + lane; t < k_out; t += 32) {
        out_val[static_cast<long long>(warp) * k_out + t] = -INFINITY;
        out_idx[static_cast<long long>(warp) * k_out + t] = -1;
      }
      break;
    }
    last_v = bv;
    last_i = bi;
  }
}

// One block per query on a materialised score row.
__global__ void rank_from_scores_kernel(const float* __restrict__ scores, int V, long long ld,
                                        const int32_t* __restrict__ gt, int k, int32_t* __restrict__ rank0,
                                        float* __restrict__ topk_val, int32_t* __restrict__ topk_idx) {
  __shared__ int s_cnt[32];
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  __shared__ float s_lastv;
  __shared__ int s_lasti;
  const int q = blockIdx.x;
  const float* row = scores + static_cast<long long>(q) * ld;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (rank0 != nullptr) {
    const int g = gt[q];
    const float sg = row[g];
    int cnt = 0;
    for (int j = threadIdx.x; j < V; j += blockDim.x) {
      const float v = row[j];
      cnt += (j != g && better(v, j, sg, g)) ? 1 : 0;
    }
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    if (lane == 0) s_cnt[w] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int i = 0; i < nw; ++i) t += s_cnt[i];
      rank0[q] = t;
    }
    __syncthreads();
  }
  if (k <= 0) return;
  if (threadIdx.x == 0) {
    s_lastv = INFINITY;
    s_lasti = INT_MAX;
  }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    const float lv = s_lastv;
    const int li = s_lasti;
    float bv = -INFINITY;
    int bi = -1;
    for (int j = threadIdx.x; j < V; j += blockDim.x) {
      const float v = row[j];
      if (better(lv, li, v, j) && better(v, j, bv, bi)) {
        bv = v;
        bi = j;
      }
    }
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (better(ov, oi, bv, bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      s_v[w] = bv;
      s_i[w] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float fv = -INFINITY;
      int fi = -1;
      for (int i = 0; i < nw; ++i)
        if (better(s_v[i], s_i[i], fv, fi)) {
          fv = s_v[i];
          fi = s_i[i];
        }
      topk_val[static_cast<long long>(q) * k + r] = fi >= 0 ? fv : -INFINITY;
      topk_idx[static_cast<long long>(q) * k + r] = fi;
      s_lastv = fi >= 0 ? fv : -INFINITY;
      s_lasti = fi;
    }
    __syncthreads();
  }
}

// Single block.  Two order statistics at once by a 31-step bitwise descent (ranks are non-negative int32); each thread
// keeps up to kMetricCache of its elements in registers so a pass costs two warp reductions and one barrier pair.
constexpr int kMetricCache = 16;
__device__ void block_two_order_stats(const int32_t* __restrict__ x, int n, int kth_a, int kth_b, int* s_red, int& out_a,
                                      int& out_b) {
  int cache[kMetricCache];
  const bool cached = n <= kMetricCache * static_cast<int>(blockDim.x);
#pragma unroll
  for (int j = 0; j < kMetricCache; ++j) {
    const int i = threadIdx.x + j * blockDim.x;
    cache[j] = (cached && i < n) ? x[i] : -1;  // -1 never matches a prefix of non-negative values
  }
  int pa = 0, pb = 0, ra = kth_a, rb = kth_b;
  const int nw = blockDim.x >> 5;
  for (int bit = 30; bit >= 0; --bit) {
    const int mask_hi = static_cast<int>(~((1u << (bit + 1)) - 1u));  // bits above `bit`
    int ca = 0, cb = 0;
    if (cached) {
#pragma unroll
      for (int j = 0; j < kMetricCache; ++j) {
        const int v = cache[j];
        const bool zero = v >= 0 && !((v >> bit) & 1);
        ca += (zero && (v & mask_hi) == (pa & mask_hi)) ? 1 : 0;
        cb += (zero && (v & mask_hi) == (pb & mask_hi)) ? 1 : 0;
      }
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int v = x[i];
        const bool zero = !((v >> bit) & 1);
        ca += (zero && (v & mask_hi) == (pa & mask_hi)) ? 1 : 0;
        cb += (zero && (v & mask_hi) == (pb & mask_hi)) ? 1 : 0;
      }
    }
    for (int off = 16; off > 0; off >>= 1) {
      ca += __shfl_xor_sync(0xffffffffu, ca, off);
      cb += __shfl_xor_sync(0xffffffffu, cb, off);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      s_red[threadIdx.x >> 5] = ca;
      s_red[32 + (threadIdx.x >> 5)] = cb;
    }
    __syncthreads();
    int ta = 0, tb = 0;
    for (int i = 0; i < nw; ++i) {
      ta += s_red[i];
      tb += s_red[32 + i];
    }
    if (ra >= ta) { ra -= ta; pa |= (1 << bit); }
    if (rb >= tb) { rb -= tb; pb |= (1 << bit); }
  }
  out_a = pa;
  out_b = pb;
}

__global__ void rank_metrics_kernel(const int32_t* __restrict__ rank0, int Q, double* __restrict__ out) {
  __shared__ int s_red[64];
  __shared__ double s_d[32][4];
  int c1 = 0, c5 = 0, c10 = 0;
  double sum_r = 0.0, sum_inv = 0.0;
  for (int i = threadIdx.x; i < Q; i += blockDim.x) {
    const int r = rank0[i];
    c1 += r < 1;
    c5 += r < 5;
    c10 += r < 10;
    sum_r += static_cast<double>(r);
    sum_inv += 1.0 / (static_cast<double>(r) + 1.0);
  }
  double v[5] = {static_cast<double>(c1), static_cast<double>(c5), static_cast<double>(c10), sum_r, sum_inv};
  double tot[5];
  for (int t = 0; t < 5; ++t) {
    double x = v[t];
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_d[threadIdx.x >> 5][0] = x;
    __syncthreads();
    double a = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) a += s_d[i][0];
    tot[t] = a;
  }
  // np.median: middle element (odd Q) or the mean of the two middle elements (even Q)
  int lo, hi;
  block_two_order_stats(rank0, Q, (Q - 1) / 2, Q / 2, s_red, lo, hi);
  if (threadIdx.x == 0) {
    const double q = static_cast<double>(Q);
    const double med = 0.5 * (static_cast<double>(lo) + static_cast<double>(hi));
    out[0] = 100.0 * tot[0] / q;
    out[1] = 100.0 * tot[1] / q;
    out[2] = 100.0 * tot[2] / q;
    out[3] = floor(med) + 1.0;
    out[4] = tot[3] / q + 1.0;
    out[5] = tot[4] / q;
    out[6] = tot[4] / q;
    out[7] = q;
  }
}

// evaluation.eval (evaluation.py:92-109) on a 0/1 label matrix whose column p is the p-th ranked item: one warp per
// query finds the 1-based rank of the first ground truth and AP = mean_j (j + 1) / rank_j over its ground truths.
__global__ void label_metrics_kernel(const uint8_t* __restrict__ label, int Q, int V, long long ld,
                                     int32_t* __restrict__ rank0, double* __restrict__ ap) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= Q) return;
  const uint8_t* row = label + static_cast<long long>(warp) * ld;
  int first = -1;
  int seen = 0;
  double acc = 0.0;
  for (int base = 0; base < V; base += 32) {
    const int p = base + lane;
    const bool hit = p < V && row[p] == 1;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) {
      if (first < 0) first = base + __ffs(m) - 1;
      if (hit) {
        const int j = seen + __popc(m & ((1u << lane) - 1u));  // index of this ground truth among the row's
        acc += static_cast<double>(j + 1) / static_cast<double>(p + 1);
      }
      seen += __popc(m);
    }
  }
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    rank0[warp] = first;  // -1: no ground truth in the row (the reference would raise IndexError)
    ap[warp] = seen > 0 ? acc / static_cast<double>(seen) : 0.0;
  }
}

__global__ void mean_double_kernel(const double* __restrict__ x, int n, double* __restrict__ out) {
  __shared__ double s[32];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += x[i];
  for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += s[i];
    *out = t / static_cast<double>(n);
  }
}

}  // namespace laff

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
using namespace laff;

extern "C" {

int laff_sim_dense(const void* q, const void* g, int Q, int V, int D, long long ldq, long long ldg, int dtype,
                   float scale, float* out, long long ld_out, void* stream) {
  LAFF_REQUIRE(q && g && out, LAFF_EINVAL, "laff_sim_dense: null pointer");
  LAFF_REQUIRE(ld_out >= V, LAFF_EINVAL, "laff_sim_dense: ld_out %lld < V %d", ld_out, V);
  const Tuning t = get_tuning();
  GemmOperands op;
  int rc = prepare_operands(&op, q, g, Q, V, D, ldq, ldg, dtype, t.cta_group);
  if (rc) return rc;
  const Sched s = make_sched(Q, V, op.cg, t.chunk_tiles, t.m_group, 0);
  EpiDense::Params ep{out, ld_out, Q, V, scale};
  return launch_gemm<EpiDense>(op, s, ep, static_cast<cudaStream_t>(stream));
}

int laff_debug_gemm(const void* q, const void* g, int Q, int V, int D, long long ldq, long long ldg, int dtype,
                    int mode, int hint_a, int hint_b, float* sink, void* stream) {
  static const uint64_t hints[4] = {0, ptx::kEvictNormal, ptx::kEvictFirst, ptx::kEvictLast};
  g_hint_override[0] = hints[hint_a & 3];
  g_hint_override[1] = hints[hint_b & 3];
  if (mode < 0) return LAFF_OK;  // only set the hint override
  const Tuning t = get_tuning();
  GemmOperands op;
  int rc = prepare_operands(&op, q, g, Q, V, D, ldq, ldg, dtype, t.cta_group);
  if (rc) return rc;
  const Sched s = make_sched(Q, V, op.cg, t.chunk_tiles, t.m_group, 0);
  EpiNull::Params ep{sink, mode};
  return launch_gemm<EpiNull>(op, s, ep, static_cast<cudaStream_t>(stream));
}

int laff_sim_collect(const void* q, const void* g, int Q, int V, int D, long long ldq, long long ldg, int dtype,
                     float scale, const float* thr, int col_offset, int cap, int32_t* count, float* cand_val,
                     int32_t* cand_idx, void* stream) {
  return laff_sim_collect_rank(q, g, Q, V, D, ldq, ldg, dtype, scale, thr, col_offset, cap, count, cand_val, cand_idx, nullptr,
                               nullptr, nullptr, stream);
}

int laff_sim_collect_rank(const void* q, const void* g, int Q, int V, int D, long long ldq, long long ldg, int dtype,
                          float scale, const float* thr, int col_offset, int cap, int32_t* count, float* cand_val,
                          int32_t* cand_idx, const float* sgt_raw, const int32_t* gt_global, int32_t* rank_count, void* stream) {
  LAFF_REQUIRE(q && g && thr && count && cand_val && cand_idx, LAFF_EINVAL, "laff_sim_collect: null pointer");
  LAFF_REQUIRE(cap > 0, LAFF_EINVAL, "laff_sim_collect: cap must be positive (got %d)", cap);
  LAFF_REQUIRE((sgt_raw != nullptr) == (gt_global != nullptr) && (sgt_raw != nullptr) == (rank_count != nullptr), LAFF_EINVAL,
               "laff_sim_collect_rank: sgt_raw, gt_global and rank_count go together");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rank_count) LAFF_CUDA(cudaMemsetAsync(rank_count, 0, sizeof(int32_t) * static_cast<size_t>(Q), st));
  const Tuning t = get_tuning();
  GemmOperands op;
  int rc = prepare_operands(&op, q, g, Q, V, D, ldq, ldg, dtype, t.cta_group);
  if (rc) return rc;
  const long long total = static_cast<long long>(Q) * cap;
  int blocks = static_cast<int>((total + 255) / 256 < op.sms * 8 ? (total + 255) / 256 : op.sms * 8);
  collect_init_kernel<<<blocks, 256, 0, st>>>(count, cand_val, cand_idx, Q, total); laff::count_launch();
  LAFF_CUDA(cudaGetLastError());
  const int tiles = (V + kBlockN - 1) / kBlockN;
  const int tiles_per = sweep_tiles_per_launch(tiles);   // see there: launches of ~480 column tiles keep the CTAs aligned
  for (int c0 = 0; c0 < V; c0 += tiles_per * kBlockN) {
    const int cols = V - c0 < tiles_per * kBlockN ? V - c0 : tiles_per * kBlockN;
    GemmOperands oc = op;
    if (c0 > 0 || cols < V) {
      rc = prepare_operands(&oc, q, static_cast<const uint16_t*>(g) + static_cast<long long>(c0) * ldg, Q, cols, D, ldq, ldg, dtype,
                            t.cta_group);
      if (rc) return rc;
    }
    const Sched s = make_sched(Q, cols, oc.cg, t.chunk_tiles, t.m_group, 0);
    EpiCollect::Params ep{thr, count, cand_val, cand_idx, cap, Q, cols, col_offset + c0, scale, sgt_raw, gt_global, rank_count};
    rc = launch_gemm<EpiCollect>(oc, s, ep, st);
    if (rc) return rc;
  }
  return LAFF_OK;
}

size_t laff_sim_gt_workspace_bytes(int Q, int D) {
  if (Q <= 0 || D <= 0) return 0;
  return static_cast<size_t>(Q) * static_cast<size_t>(D) * 2 + 256;
}

int laff_sim_gt_scores(const void* q, const void* g, int Q, int V, int D, long long ldq, long long ldg, int dtype,
                       const int32_t* gt_local, float* sgt_raw, void* workspace, size_t workspace_bytes, void* stream) {
  LAFF_REQUIRE(q && g && gt_local && sgt_raw && workspace, LAFF_EINVAL, "laff_sim_gt_scores: null pointer");
  LAFF_REQUIRE(workspace_bytes >= laff_sim_gt_workspace_bytes(Q, D), LAFF_EWORKSPACE,
               "laff_sim_gt_scores: workspace too small (%zu < %zu)", workspace_bytes, laff_sim_gt_workspace_bytes(Q, D));
  LAFF_REQUIRE(D % 8 == 0 && ldg % 8 == 0, LAFF_EINVAL, "laff_sim_gt_scores: D and ldg must be multiples of 8");
  LAFF_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, LAFF_EINVAL, "gallery not 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 256-byte aligned gather buffer inside the workspace
  uintptr_t wp = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255);
  void* ggt = reinterpret_cast<void*>(wp);
  const Tuning t = get_tuning();
  GemmOperands op;
  int rc = prepare_operands(&op, q, ggt, Q, Q, D, ldq, D, dtype, t.cta_group);
  if (rc) return rc;
  const int vec_per_row = D / 8;
  const long long total = static_cast<long long>(Q) * vec_per_row;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > op.sms * 8) blocks = op.sms * 8;
  gather_rows16_kernel<<<blocks, 256, 0, st>>>(static_cast<const uint4*>(g), ldg / 8, gt_local, Q, vec_per_row,
                                               static_cast<uint4*>(ggt)); laff::count_launch();
  LAFF_CUDA(cudaGetLastError());
  const Sched s = make_sched(Q, Q, op.cg, 1, 1, 1);
  EpiDiag::Params ep{sgt_raw, gt_local, Q};
  (void)V;
  return launch_gemm<EpiDiag>(op, s, ep, st);
}

size_t laff_sim_rank_workspace_bytes(int Q, int V, int D) {
  (void)D;
  if (Q <= 0 || V <= 0) return 0;
  // per query: LAFF_MAX_TOPK packed slots + one threshold word
  return static_cast<size_t>(Q) * (LAFF_MAX_TOPK * 8 + 4) + 512;
}

int laff_sim_rank_topk(const void* q, const void* g, int Q, int V, int D, long long ldq, long long ldg, int dtype,
                       float scale, const float* sgt_raw, const int32_t* gt_global, int col_offset, int k,
                       int32_t* count, float* topk_val, int32_t* topk_idx, void* workspace, size_t workspace_bytes,
                       void* stream) {
  LAFF_REQUIRE(q && g && sgt_raw && gt_global && count && workspace, LAFF_EINVAL, "laff_sim_rank_topk: null pointer");
  LAFF_REQUIRE(k >= 0 && k <= LAFF_MAX_TOPK, LAFF_ENOTSUP, "laff_sim_rank_topk: k=%d outside [0, %d]", k, LAFF_MAX_TOPK);
  LAFF_REQUIRE(k == 0 || (topk_val && topk_idx), LAFF_EINVAL, "laff_sim_rank_topk: top-k outputs missing");
  LAFF_REQUIRE(workspace_bytes >= laff_sim_rank_workspace_bytes(Q, V, D), LAFF_EWORKSPACE,
               "laff_sim_rank_topk: workspace too small (%zu < %zu)", workspace_bytes,
               laff_sim_rank_workspace_bytes(Q, V, D));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Tuning t = get_tuning();
  GemmOperands op;
  int rc = prepare_operands(&op, q, g, Q, V, D, ldq, ldg, dtype, t.cta_group);
  if (rc) return rc;
  const Sched s = make_sched(Q, V, op.cg, t.chunk_tiles, t.m_group, 0);
  uintptr_t wp = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255);
  long long* slots = reinterpret_cast<long long*>(wp);
  int32_t* thr_key = reinterpret_cast<int32_t*>(wp + static_cast<size_t>(Q) * LAFF_MAX_TOPK * 8);
  const int init_n = Q * LAFF_MAX_TOPK;
  rank_init_kernel<<<(init_n + 255) / 256, 256, 0, st>>>(count, thr_key, slots, Q, LAFF_MAX_TOPK); laff::count_launch();
  LAFF_CUDA(cudaGetLastError());
  // One launch per group of m_group row tiles (2560 queries): the groups are independent passes over the gallery anyway,
  // and separate launches measured 4-5 % faster than one launch walking the groups back to back (10 000 x 1 M:
  // 67.0 / 69.7 / 67.7 ms as one launch, 62.7 / 65.6 / 65.5 ms as four; profiles/r01_sched_experiments.md).
  static const bool split_groups = [] { const char* e = getenv("LAFF_SWEEP_SPLIT"); return !e || atoi(e) != 0; }();
  const int rows_per_launch = split_groups ? s.m_group * s.rows_per_mtile : Q;
  // Experiment kept behind LAFF_SWEEP_TAIL_CG1=1 (off): sweep a trailing row tile that is at most half full (10 000 queries
  // = 39 tiles of 256 + 16 rows) with the cta_group::1 kernel (128-row tiles on single CTAs) instead of a full 256-row
  // pair MMA.  On paper 1.2 % of the sweep; measured (three A/B pairs on one box) 64.38 vs 64.13 ms, i.e. nothing: the
  // sweep is power-bound, and an MMA over mostly-zero rows draws little power while eight extra launches are not free.
  static const bool split_tail = [] { const char* e = getenv("LAFF_SWEEP_TAIL_CG1"); return e && atoi(e) != 0; }();
  auto sweep_rows = [&](int r0, int rows, int cg) -> int {
    const int tiles = (V + kBlockN - 1) / kBlockN;
    const int tiles_per = sweep_tiles_per_launch(tiles);
    for (int c0 = 0; c0 < V; c0 += tiles_per * kBlockN) {
      const int cols = V - c0 < tiles_per * kBlockN ? V - c0 : tiles_per * kBlockN;
      GemmOperands oc;
      int rc2 = prepare_operands(&oc, static_cast<const uint16_t*>(q) + static_cast<long long>(r0) * ldq,
                                 static_cast<const uint16_t*>(g) + static_cast<long long>(c0) * ldg, rows, cols, D, ldq, ldg, dtype, cg);
      if (rc2) return rc2;
      const Sched ss = make_sched(rows, cols, oc.cg, t.chunk_tiles, t.m_group, 0);
      EpiRank<LAFF_MAX_TOPK>::Params ep{sgt_raw + r0, gt_global + r0, count + r0, thr_key + r0,
                                       slots + static_cast<long long>(r0) * LAFF_MAX_TOPK, rows, cols, col_offset + c0, k};
      rc2 = launch_gemm<EpiRank<LAFF_MAX_TOPK>>(oc, ss, ep, st);
      if (rc2) return rc2;
    }
    return LAFF_OK;
  };
  for (int r0 = 0; r0 < Q; r0 += rows_per_launch) {
    int rows = Q - r0 < rows_per_launch ? Q - r0 : rows_per_launch;
    const int rem = rows % s.rows_per_mtile;
    const bool tail = split_tail && op.cg == 2 && rem > 0 && rem <= kBlockM && rows > rem;
    if (tail) rows -= rem;
    rc = sweep_rows(r0, rows, op.cg);
    if (rc) return rc;
    if (tail) {
      rc = sweep_rows(r0 + rows, rem, 1);
      if (rc) return rc;
    }
  }
  if (k > 0) {
    topk_finalize_kernel<<<(Q + 127) / 128, 128, 0, st>>>(slots, Q, LAFF_MAX_TOPK, k, scale, topk_val, topk_idx); laff::count_launch();
    LAFF_CUDA(cudaGetLastError());
  }
  return LAFF_OK;
}

int laff_topk_merge(const float* vals, const int32_t* idx, int n_lists, int Q, int k_in, long long list_stride,
                    int k_out, float in_scale, float* out_val, int32_t* out_idx, void* stream) {
  LAFF_REQUIRE(vals && idx && out_val && out_idx, LAFF_EINVAL, "laff_topk_merge: null pointer");
  LAFF_REQUIRE(n_lists > 0 && Q > 0 && k_in > 0 && k_out > 0, LAFF_EINVAL, "laff_topk_merge: bad sizes");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const int threads = 128;
  const int blocks = (Q * 32 + threads - 1) / threads;
  topk_merge_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(vals, idx, n_lists, Q, k_in, list_stride,
                                                                              k_in, k_out, in_scale, out_val, out_idx); laff::count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

int laff_rank_from_scores(const float* scores, int Q, int V, long long ld, const int32_t* gt, int k, int32_t* rank0,
                          float* topk_val, int32_t* topk_idx, void* stream) {
  LAFF_REQUIRE(scores && Q > 0 && V > 0 && ld >= V, LAFF_EINVAL, "laff_rank_from_scores: bad arguments");
  LAFF_REQUIRE(rank0 == nullptr || gt != nullptr, LAFF_EINVAL, "laff_rank_from_scores: rank0 needs gt");
  LAFF_REQUIRE(k == 0 || (topk_val && topk_idx), LAFF_EINVAL, "laff_rank_from_scores: top-k outputs missing");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  rank_from_scores_kernel<<<Q, 256, 0, static_cast<cudaStream_t>(stream)>>>(scores, V, ld, gt, k, rank0, topk_val,
                                                                           topk_idx); laff::count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

int laff_label_metrics(const uint8_t* label, int Q, int V, long long ld, int32_t* rank0, double* ap, double* out8,
                       void* stream) {
  LAFF_REQUIRE(label && rank0 && ap && out8 && Q > 0 && V > 0 && ld >= V, LAFF_EINVAL, "laff_label_metrics: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 128;
  label_metrics_kernel<<<(Q * 32 + threads - 1) / threads, threads, 0, st>>>(label, Q, V, ld, rank0, ap); laff::count_launch();
  rank_metrics_kernel<<<1, 1024, 0, st>>>(rank0, Q, out8); laff::count_launch();
  mean_double_kernel<<<1, 1024, 0, st>>>(ap, Q, out8 + 6); laff::count_launch();  // mAP over the per-query APs
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

int laff_rank_metrics(const int32_t* rank0, int Q, double* out8, void* stream) {
  LAFF_REQUIRE(rank0 && out8 && Q > 0, LAFF_EINVAL, "laff_rank_metrics: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  rank_metrics_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(rank0, Q, out8); laff::count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

}  // extern "C"
