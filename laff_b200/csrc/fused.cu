// laff_fuse_forward — the whole LAFF fusion of one net in ONE kernel:
//   per-feature FC projection (tcgen05 GEMM, TMA-fed) -> bias -> activation -> eval-BN -> per-head attention logit ->
//   softmax over the features -> weighted sum -> L2 normalise, never writing the projected features to HBM.
//   model/model.py:257-276 (TransformNet), :1807-1876 / :1663-1705 (nets), model/Attention.py:78-105, :508-531.
//
// Two variants of one kernel (template parameter CG = tcgen05 cta_group):
//   CG = 1  work unit = (128-row tile, head), cluster of 2 CTAs; CTA c computes columns [256c, 256c+256) of the head's
//           512 with cta_group::1 MMAs (128 x 256 x K_l).  All 148 SMs run, but each SM pulls 48 KB of operands per
//           k-block (96 B/clk) and ends up L2->SM bound.
//   CG = 2  work unit = (256-row tile, head), cluster of 4 CTAs = two MMA pairs; pair q computes columns
//           [256q, 256q+256) with cta_group::2 MMAs (256 x 256 x K_l), each CTA holding 128 of the rows and loading
//           half of the pair's W tile: 32 KB per k-block per SM (64 B/clk, as in the similarity sweep).  Only 33
//           4-CTA clusters fit the 148 SMs (132 SMs busy), still the faster variant for large row counts.
// Per FC feature the GEMM lands in one of two TMEM accumulator buffers, so the epilogue of feature l overlaps the
// MMAs of feature l+1.  Epilogue threads own (row, 128 columns):
//   pass A  y = BN(act(acc + b)), partial logit  sum_c w_h[c] y[c]           (TMEM -> registers)
//   exchange the 4 partial logits of a row (2 column halves x the 2 CTAs that hold the row) through shared memory /
//   DSMEM (st.async) + an mbarrier
//   pass B  re-read TMEM, recompute y, online-softmax update of the running weighted sum g[128] held in registers
// The softmax denominator cancels under the final L2 normalisation (with_ave = False), so only exp(e - max) weights are
// kept.  "No-transform" features (raw 512-d vector tiled over the heads + BN, model/model.py:1822-1823) take the same
// two passes reading the raw feature from global memory.  Register budget: setmaxnreg moves registers from the
// producer/MMA warpgroup to the two epilogue warpgroups (g[128] + a 32-column chunk per thread).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <atomic>
#include <cstring>

#include "gemm_engine.cuh"
#include "host_util.cuh"

namespace laff {

constexpr int kFuseBarBytes = 256;
constexpr int kFuseParamBufs = 3;                          // see the staging protocol in the epilogue
constexpr int kFuseParamFloats = kFuseParamBufs * 4 * kBlockN;  // per buffer: {bias, scale, shift, w_h} x 256 columns
constexpr int kFuseXchgFloats = 2 * 4 * kBlockM;            // [parity][partial][row]
// Warp-private 32-row x 64-byte transposition tiles (one per epilogue warp).  An epilogue thread owns a ROW, so its
// natural global accesses are 16 bytes at 32 different rows per warp instruction (32 half-used sectors).  Through the
// tile a warp instruction moves 8 rows x 64 contiguous bytes instead: 4x fewer lines, every sector fully used.
constexpr int kFuseTileRowBytes = 64;
constexpr int kFuseTileWarpBytes = 32 * kFuseTileRowBytes;
constexpr int kFuseTileBytes = kEpiWarps * kFuseTileWarpBytes;
template <int CG>
struct FuseCfg {
  static constexpr int kStages = (CG == 2) ? 6 : 4;
  static constexpr int kABytes = kBlockM * kBlockK * 2;          // 16 KB: this CTA's 128 rows of x
  static constexpr int kBRows = kBlockN / CG;                    // rows of W (= output columns) this CTA loads
  static constexpr int kBBytes = kBRows * kBlockK * 2;           // 32 KB (CG 1) / 16 KB (CG 2)
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTxBytes = kStageBytes * CG;              // bytes landing per stage on the MMA issuer's barrier
  static constexpr int kCluster = 2 * CG;
  static constexpr int kSmem = kStages * kStageBytes + kFuseBarBytes + (kFuseParamFloats + kFuseXchgFloats) * 4 +
                               kFuseTileBytes + 1024;
  static_assert(kSmem <= 227 * 1024, "fused kernel shared memory budget");
};

struct FuseTmaps {
  CUtensorMap x[LAFF_FUSE_MAX_FC];
  CUtensorMap w[LAFF_FUSE_MAX_FC];
};

struct FuseParams {
  int n_fc, n_tiled, heads;
  long long rows;
  int num_kb[LAFF_FUSE_MAX_FC];
  int act[LAFF_FUSE_MAX_FC];
  const float* bias[LAFF_FUSE_MAX_FC];
  const float* bn_scale[LAFF_FUSE_MAX_FC];
  const float* bn_shift[LAFF_FUSE_MAX_FC];
  const float* tiled_x[LAFF_FUSE_MAX_TILED];
  long long tiled_ld[LAFF_FUSE_MAX_TILED];
  int tiled_in_dim[LAFF_FUSE_MAX_TILED];
  const float* tiled_scale[LAFF_FUSE_MAX_TILED];
  const float* tiled_shift[LAFF_FUSE_MAX_TILED];
  const float* att_w;
  const float* att_b;
  float* out;
  long long ld_out;
  void* out16;
  int out16_dtype;
  long long ld_out16;
  float norm_eps;
  uint32_t idesc;
  int total_units;
};

#ifdef LAFF_FUSE_PROFILE
// Phase timers of the epilogue (cycles, one sampling thread: CTA 0, warp 4, lane 0); read by laff_debug_fuse_profile.
__device__ unsigned long long g_fuse_prof[8];
#define LAFF_PROF_T(var) const long long var = clock64()
#define LAFF_PROF_ADD(slot, a, b) \
  if (blockIdx.x == 0 && threadIdx.x == 128) atomicAdd(&g_fuse_prof[slot], static_cast<unsigned long long>((b) - (a)))
#else
#define LAFF_PROF_T(var)
#define LAFF_PROF_ADD(slot, a, b)
#endif

namespace fptx {
__device__ __forceinline__ void setmaxnreg_dec56() { asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory"); }
__device__ __forceinline__ void setmaxnreg_inc224() { asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory"); }
// Remote (peer CTA) shared-memory store that reports its 4 bytes to an mbarrier in the same peer CTA: the receiver
// needs no fence and the sender no separate arrive.
__device__ __forceinline__ void st_async_f32(uint32_t cluster_addr, float v, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(cluster_addr),
               "r"(__float_as_uint(v)), "r"(cluster_mbar)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > LAFF_WATCHDOG_CYCLES) ptx::watchdog_fire(bar, parity, tag);
  }
}
__device__ __forceinline__ void mbar_arrive_release_cluster_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
}  // namespace fptx

// Activations in the epilogue, with every per-column constant folded on the host side of the loop (load_params):
//   tanh(z) = 2 / (1 + 2^(-2 z log2e)) - 1,  sigmoid(z) = 1 / (1 + 2^(-z log2e))     ("sigm" family: a / (1 + 2^(n z)) + c)
//   y = BN(act(acc + b)) = A' * rcp(1 + ex2(acc * n + n b)) + C'   with  A' = a * scale,  C' = c * scale + shift
//   relu / none: y = max(acc + b, floor) * scale + shift
// ex2.approx / rcp.approx: |error| < 3e-7 absolute on y.
struct ActCoef {
  float neg_s;   // n = -s * log2(e)
  float a, c;
  float floor;   // piecewise-linear family: -inf for 'none', 0 for relu
  bool sigm;
};
__device__ __forceinline__ ActCoef make_act(int act) {
  ActCoef k;
  k.sigm = act == LAFF_ACT_TANH || act == LAFF_ACT_SIGMOID;
  const float l2e = 1.4426950408889634f;
  k.neg_s = act == LAFF_ACT_TANH ? -2.0f * l2e : -l2e;
  k.a = act == LAFF_ACT_TANH ? 2.0f : 1.0f;
  k.c = act == LAFF_ACT_TANH ? -1.0f : 0.0f;
  k.floor = act == LAFF_ACT_RELU ? 0.0f : -INFINITY;
  return k;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}
// One 16-column chunk of pass A: r holds the accumulators on entry and y on exit; returns the running partial logit.
// p0/p1/p2 are the folded per-column constants described above, pw the attention weights w_h.
constexpr int kFuseChunk = 16;
__device__ __forceinline__ float pass_a_chunk(uint32_t (&r)[kFuseChunk], const float* p0, const float* p1, const float* p2,
                                              const float* pw, const ActCoef& ak, float part) {
  auto act1 = [&](uint32_t acc, float n, float a, float c) -> float {
    return fmaf(a, rcp_approx(1.0f + ex2_approx(fmaf(__uint_as_float(acc), ak.neg_s, n))), c);
  };
  if (ak.sigm) {
#pragma unroll
    for (int j = 0; j < kFuseChunk; j += 4) {
      const float4 n4 = *reinterpret_cast<const float4*>(p0 + j);
      const float4 a4 = *reinterpret_cast<const float4*>(p1 + j);
      const float4 c4 = *reinterpret_cast<const float4*>(p2 + j);
      const float4 w4 = *reinterpret_cast<const float4*>(pw + j);
      const float y0 = act1(r[j], n4.x, a4.x, c4.x);
      const float y1 = act1(r[j + 1], n4.y, a4.y, c4.y);
      const float y2 = act1(r[j + 2], n4.z, a4.z, c4.z);
      const float y3 = act1(r[j + 3], n4.w, a4.w, c4.w);
      part = fmaf(w4.x, y0, part);
      part = fmaf(w4.y, y1, part);
      part = fmaf(w4.z, y2, part);
      part = fmaf(w4.w, y3, part);
      r[j] = __float_as_uint(y0); r[j + 1] = __float_as_uint(y1); r[j + 2] = __float_as_uint(y2); r[j + 3] = __float_as_uint(y3);
    }
  } else {
#pragma unroll
    for (int j = 0; j < kFuseChunk; j += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(p0 + j);
      const float4 a4 = *reinterpret_cast<const float4*>(p1 + j);
      const float4 c4 = *reinterpret_cast<const float4*>(p2 + j);
      const float4 w4 = *reinterpret_cast<const float4*>(pw + j);
      const float y0 = fmaf(fmaxf(__uint_as_float(r[j]) + b4.x, ak.floor), a4.x, c4.x);
      const float y1 = fmaf(fmaxf(__uint_as_float(r[j + 1]) + b4.y, ak.floor), a4.y, c4.y);
      const float y2 = fmaf(fmaxf(__uint_as_float(r[j + 2]) + b4.z, ak.floor), a4.z, c4.z);
      const float y3 = fmaf(fmaxf(__uint_as_float(r[j + 3]) + b4.w, ak.floor), a4.w, c4.w);
      part = fmaf(w4.x, y0, part);
      part = fmaf(w4.y, y1, part);
      part = fmaf(w4.z, y2, part);
      part = fmaf(w4.w, y3, part);
      r[j] = __float_as_uint(y0); r[j + 1] = __float_as_uint(y1); r[j + 2] = __float_as_uint(y2); r[j + 3] = __float_as_uint(y3);
    }
  }
  return part;
}
__device__ __forceinline__ uint16_t fuse_to16(float v, int dtype) {
  if (dtype == LAFF_BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  return __half_as_ushort(__float2half_rn(v));
}

template <int CG>
__global__ void __launch_bounds__(kNumThreads, 1)
    laff_fuse_kernel(const __grid_constant__ FuseTmaps tm, const __grid_constant__ FuseParams p) {
  using Cfg = FuseCfg<CG>;
  constexpr int kFuseStages = Cfg::kStages;
  constexpr int kFuseABytes = Cfg::kABytes;
  constexpr int kFuseBBytes = Cfg::kBBytes;
  constexpr int kFuseStageBytes = Cfg::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);

  const uint32_t sA = base;
  const uint32_t sB = base + kFuseStages * kFuseABytes;
  const uint32_t bar0 = base + kFuseStages * kFuseStageBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kFuseStages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kFuseStages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kFuseStages + 2 + a); };
  auto xchg_bar = [&](int a) { return bar0 + 8u * (2 * kFuseStages + 4 + a); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * kFuseStages + 6);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + kFuseStages * kFuseStageBytes + 8 * (2 * kFuseStages + 6));
  float* s_param = reinterpret_cast<float*>(smem + kFuseStages * kFuseStageBytes + kFuseBarBytes);
  float* s_xchg = s_param + kFuseParamFloats;
  uint8_t* s_tiles = reinterpret_cast<uint8_t*>(s_xchg + kFuseXchgFloats);
  const uint32_t s_xchg_addr = bar0 + kFuseBarBytes + kFuseParamFloats * 4;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(ptx::lane_id());
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const int cluster_id = static_cast<int>(blockIdx.x) / Cfg::kCluster;
  const int num_clusters = static_cast<int>(gridDim.x) / Cfg::kCluster;
  // Which 256 columns of the head / which 128 rows of the unit this CTA holds, who issues its MMAs, and which CTA
  // holds the other 256 columns of the same rows (the logit-exchange partner).
  const int col_half = (CG == 2) ? static_cast<int>(cta_rank >> 1) : static_cast<int>(cta_rank);
  const int row_sub = (CG == 2) ? static_cast<int>(cta_rank & 1u) : 0;
  const uint32_t leader = (CG == 2) ? (cta_rank & ~1u) : cta_rank;
  const bool is_leader = leader == cta_rank;
  const uint32_t peer = (CG == 2) ? (cta_rank ^ 2u) : (cta_rank ^ 1u);
  constexpr int kUnitRows = kBlockM * CG;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < p.n_fc; ++l) {
      ptx::prefetch_tensormap(&tm.x[l]);
      ptx::prefetch_tensormap(&tm.w[l]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kFuseStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), kEpiWarps * CG);  // every epilogue warp of the CTAs sharing the accumulator
      ptx::mbar_init(xchg_bar(a), kEpiWarps);  // local epilogue warps; the peer's partials arrive as transaction bytes
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<CG>(tmem_slot, kTmemCols);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tcgen05_fence_before();
  ptx::cluster_sync_all();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    fptx::setmaxnreg_dec56();
    if (warp == 0) {
      // ========================================= TMA producer =========================================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int u = cluster_id; u < p.total_units; u += num_clusters) {
          const int row_tile = u / p.heads, head = u - row_tile * p.heads;
          const int m0 = row_tile * kUnitRows + row_sub * kBlockM;
          const int n0 = head * 2 * kBlockN + col_half * kBlockN + row_sub * Cfg::kBRows;
          for (int l = 0; l < p.n_fc; ++l) {
            for (int kb = 0; kb < p.num_kb[l]; ++kb) {
              ptx::mbar_wait_nocall(empty_bar(stage), phase ^ 1u);
              if (is_leader) ptx::mbar_arrive_expect_tx(full_bar(stage), Cfg::kTxBytes);
              if constexpr (CG == 1) {
                ptx::tma_load_2d(sA + stage * kFuseABytes, &tm.x[l], full_bar(stage), kb * kBlockK, m0, ptx::kEvictNormal);
                ptx::tma_load_2d(sB + stage * kFuseBBytes, &tm.w[l], full_bar(stage), kb * kBlockK, n0, ptx::kEvictLast);
              } else {
                const uint32_t leader_full = ptx::mapa(full_bar(stage), leader);
                ptx::tma_load_2d_pair(sA + stage * kFuseABytes, &tm.x[l], leader_full, kb * kBlockK, m0, ptx::kEvictNormal);
                ptx::tma_load_2d_pair(sB + stage * kFuseBBytes, &tm.w[l], leader_full, kb * kBlockK, n0, ptx::kEvictLast);
              }
              if (++stage == kFuseStages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ========================================== MMA issuer ==========================================
      if (is_leader && lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint16_t pair_mask = static_cast<uint16_t>(3u << leader);
        auto commit = [&](uint32_t bar) {
          if constexpr (CG == 1) ptx::umma_commit<1>(bar);
          else ptx::umma_commit_pair_mask(bar, pair_mask);
        };
        for (int u = cluster_id; u < p.total_units; u += num_clusters) {
          for (int l = 0; l < p.n_fc; ++l) {
            ptx::mbar_wait_nocall(tempty_bar(acc), acc_phase ^ 1u);
            ptx::tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kBlockN);
            for (int kb = 0; kb < p.num_kb[l]; ++kb) {
              ptx::mbar_wait_nocall(full_bar(stage), phase);
              ptx::tcgen05_fence_after();
              const uint64_t da = ptx::make_smem_desc_sw128(sA + stage * kFuseABytes);
              const uint64_t db = ptx::make_smem_desc_sw128(sB + stage * kFuseBBytes);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k)
                ptx::umma_f16<CG>(d_tmem, da + 2u * k, db + 2u * k, p.idesc, static_cast<uint32_t>((kb | k) != 0));
              commit(empty_bar(stage));
              if (++stage == kFuseStages) { stage = 0; phase ^= 1u; }
            }
            commit(tfull_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ============================================ epilogue ============================================
    fptx::setmaxnreg_inc224();
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row_in_cta = quad * 32 + lane;
    const int epi_tid = half * kBlockM + row_in_cta;       // 0..255
    const int my_part = half + 2 * col_half;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t xchg_count = 0;  // exchanges done so far: parity = count & 1, phase = (count >> 1) & 1
    // transposition tile of this warp: 16-byte piece c (0..3) of row r sits at r * 64 + 16 * (c ^ ((r >> 1) & 3)), which
    // makes both the row-per-thread accesses and the 4-lanes-per-row accesses bank-conflict free
    uint8_t* tile = s_tiles + (warp - 4) * kFuseTileWarpBytes;
    auto tile_at = [&](int r, int c) -> uint4* { return reinterpret_cast<uint4*>(tile + r * kFuseTileRowBytes + 16 * (c ^ ((r >> 1) & 3))); };
    const int t_row = lane >> 2, t_piece = lane & 3;   // coalesced side: lane handles piece t_piece of rows t_row + 8 i

    // All-to-all of one float per (row, partial) among the 4 owners of a row (2 column halves x 2 CTAs); returns the
    // sum in a fixed order, identical in all four threads.  post() publishes, collect() waits: independent work placed
    // between the two hides the cluster round trip.  Exchange n uses slot / barrier n & 1; a thread can only post n + 2
    // after collecting n + 1, which needs every thread's post of n + 1, which they issue after reading slot n.
    auto post = [&](float v) {
      const int par = static_cast<int>(xchg_count & 1u);
      const int off = (par * 4 + my_part) * kBlockM + row_in_cta;
      s_xchg[off] = v;
      fptx::st_async_f32(ptx::mapa(s_xchg_addr + 4u * off, peer), v, ptx::mapa(xchg_bar(par), peer));
      __syncwarp();
      if (lane == 0) {
        if (warp == 4) ptx::mbar_arrive_expect_tx(xchg_bar(par), 2 * kBlockM * 4);  // the peer's 2 x 128 partials
        else ptx::mbar_arrive(xchg_bar(par));
      }
    };
    auto collect = [&]() -> float {
      const int par = static_cast<int>(xchg_count & 1u);
      const uint32_t ph = (xchg_count >> 1) & 1u;
      fptx::mbar_wait_cluster(xchg_bar(par), ph, 14);
      const float* sx = s_xchg + par * 4 * kBlockM + row_in_cta;
      const float r = ((sx[0] + sx[kBlockM]) + sx[2 * kBlockM]) + sx[3 * kBlockM];
      ++xchg_count;
      return r;
    };
    // Per-column parameters {bias, BN scale, BN shift, w_h} of (feature f of head `hd`), one column per epilogue thread:
    // loaded into registers before pass A of the current feature and stored to buffer (cur + 1) % 3 after it, i.e.
    // before this thread's post(); collect() then orders every thread's stores before anybody's reads.  Three buffers:
    // a slower warp may still read buffer cur - 1 (pass B of the previous feature) while a faster one writes cur + 1.
    auto load_params = [&](int f, int hd, float (&v)[4]) {
      const bool tl = f < p.n_tiled;
      const int l = tl ? f : f - p.n_tiled;
      const int col = hd * 2 * kBlockN + col_half * kBlockN + epi_tid;
      const float* bs = tl ? nullptr : p.bias[l];
      const float* sc = tl ? p.tiled_scale[l] : p.bn_scale[l];
      const float* sh = tl ? p.tiled_shift[l] : p.bn_shift[l];
      const float b = bs ? __ldg(bs + col) : 0.f;
      const float scale = sc ? __ldg(sc + col) : 1.f;
      const float shift = sh ? __ldg(sh + col) : 0.f;
      const ActCoef k = make_act(tl ? LAFF_ACT_NONE : p.act[l]);
      v[0] = k.sigm ? k.neg_s * b : b;                    // sigm family: n * b, else the bias itself
      v[1] = k.sigm ? k.a * scale : scale;                // A'
      v[2] = k.sigm ? fmaf(k.c, scale, shift) : shift;    // C'
      v[3] = __ldg(p.att_w + col);
    };
    auto store_params = [&](int buf, const float (&v)[4]) {
      float* sp = s_param + buf * 4 * kBlockN;
      sp[epi_tid] = v[0];
      sp[kBlockN + epi_tid] = v[1];
      sp[2 * kBlockN + epi_tid] = v[2];
      sp[3 * kBlockN + epi_tid] = v[3];
    };

    int stage_buf = 0;  // buffer holding the parameters of the feature about to be processed
    if (cluster_id < p.total_units) {
      float v[4];
      load_params(0, cluster_id % p.heads, v);
      store_params(0, v);
    }
    fptx::epi_bar_sync();
    for (int u = cluster_id; u < p.total_units; u += num_clusters) {
      const int row_tile = u / p.heads, head = u - row_tile * p.heads;
      const long long row = static_cast<long long>(row_tile) * kUnitRows + row_sub * kBlockM + row_in_cta;
      const bool row_ok = row < p.rows;
      const int colbase = head * 2 * kBlockN + col_half * kBlockN;                   // first global column of this CTA
      const int mycol = half * kEpiCols;                                             // first column (in CTA) of this thread

      float g[kEpiCols];
#pragma unroll
      for (int j = 0; j < kEpiCols; ++j) g[j] = 0.f;
      float m_ref = -INFINITY;  // reference logit of the row: weights are exp(e - m_ref)
      // Weight of a feature with logit e.  The reference moves (and g is rescaled) only when exp() could overflow --
      // always on the first feature (m_ref = -inf), practically never afterwards -- so the common case costs one FMA
      // per element in pass B.  All four owners of a row see the same e and share their warp's row set, so they take
      // the same decision and the row keeps one common scale.
      auto rebase = [&](float e) -> float {
        if (__any_sync(0xffffffffu, e - m_ref > 60.0f)) {
          const float nm = fmaxf(m_ref, e);
          const float corr = __expf(m_ref - nm);
          m_ref = nm;
#pragma unroll
          for (int j = 0; j < kEpiCols; ++j) g[j] *= corr;
        }
        return __expf(e - m_ref);
      };

      const int n_feat = p.n_tiled + p.n_fc;
      for (int f = 0; f < n_feat; ++f) {
        const bool tiled = f < p.n_tiled;
        const int l = tiled ? f : f - p.n_tiled;
        const float* sp = s_param + stage_buf * 4 * kBlockN;
        float nextp[4];
        const bool has_next = (f + 1 < n_feat) || (u + num_clusters < p.total_units);
        if (f + 1 < n_feat) load_params(f + 1, head, nextp);
        else if (has_next) load_params(0, (u + num_clusters) % p.heads, nextp);
        const float* p0 = sp + mycol;                  // see load_params for the meaning of the four rows
        const float* p1 = sp + kBlockN + mycol;
        const float* p2 = sp + 2 * kBlockN + mycol;
        const float* pw = sp + 3 * kBlockN + mycol;
        const int next_buf = stage_buf == kFuseParamBufs - 1 ? 0 : stage_buf + 1;

        if (tiled) {
          const float* xrow = p.tiled_x[l] + (row_ok ? row : 0) * p.tiled_ld[l];
          const int xoff = (colbase + mycol) % p.tiled_in_dim[l];  // in_dim is a multiple of 128: 128 consecutive columns never wrap
          float part = 0.f;
          LAFF_PROF_T(pt0);
          if (f == 0) {
            // First feature of the row: its weight is exp(0) = 1 whatever its logit turns out to be, so y goes
            // straight into g and no second pass is needed.
            // 16-column chunks: 4 coalesced loads per lane (8 rows x 64 bytes per warp instruction).  All 8 chunks are
            // requested before the first one is used -- g is still empty, so the 128 registers are free -- and the
            // global-load latency is paid once per unit instead of once per chunk.
            const long long row0 = static_cast<long long>(row_tile) * kUnitRows + row_sub * kBlockM + quad * 32;
            float4 xv[kEpiCols / 16][4];
#pragma unroll
            for (int k = 0; k < kEpiCols / 16; ++k) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const long long rr = row0 + t_row + 8 * i;
                xv[k][i] = rr < p.rows ? __ldg(reinterpret_cast<const float4*>(p.tiled_x[l] + rr * p.tiled_ld[l] + xoff + 16 * k + 4 * t_piece))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
#pragma unroll
            for (int k = 0; k < kEpiCols / 16; ++k) {
#pragma unroll
              for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(tile_at(t_row + 8 * i, t_piece)) = xv[k][i];
              __syncwarp();
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int cc = 16 * k + 4 * c;
                const float4 v = *reinterpret_cast<const float4*>(tile_at(lane, c));
                const float4 sc4 = *reinterpret_cast<const float4*>(p1 + cc);
                const float4 sh4 = *reinterpret_cast<const float4*>(p2 + cc);
                const float4 w4 = *reinterpret_cast<const float4*>(pw + cc);
                g[cc] = fmaf(v.x, sc4.x, sh4.x);
                g[cc + 1] = fmaf(v.y, sc4.y, sh4.y);
                g[cc + 2] = fmaf(v.z, sc4.z, sh4.z);
                g[cc + 3] = fmaf(v.w, sc4.w, sh4.w);
                part = fmaf(w4.x, g[cc], part);
                part = fmaf(w4.y, g[cc + 1], part);
                part = fmaf(w4.z, g[cc + 2], part);
                part = fmaf(w4.w, g[cc + 3], part);
              }
              __syncwarp();
            }
            LAFF_PROF_T(ptm);
            if (has_next) store_params(next_buf, nextp);
            post(part);
            m_ref = collect() + __ldg(p.att_b + head);
            stage_buf = next_buf;
            LAFF_PROF_T(pt1);
            LAFF_PROF_ADD(5, pt0, ptm);   // loads + math of the first tiled feature
            LAFF_PROF_ADD(2, ptm, pt1);   // its exchange goes to the 'post' slot (debug only)
            continue;
          }
#pragma unroll 1
          for (int c = 0; c < kEpiCols / 32; ++c) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int cc = c * 32 + j;
              const float4 v = row_ok ? __ldg(reinterpret_cast<const float4*>(xrow + xoff + cc)) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float4 sc4 = *reinterpret_cast<const float4*>(p1 + cc);
              const float4 sh4 = *reinterpret_cast<const float4*>(p2 + cc);
              const float4 w4 = *reinterpret_cast<const float4*>(pw + cc);
              part = fmaf(w4.x, fmaf(v.x, sc4.x, sh4.x), part);
              part = fmaf(w4.y, fmaf(v.y, sc4.y, sh4.y), part);
              part = fmaf(w4.z, fmaf(v.z, sc4.z, sh4.z), part);
              part = fmaf(w4.w, fmaf(v.w, sc4.w, sh4.w), part);
            }
          }
          if (has_next) store_params(next_buf, nextp);
          post(part);
          const float pe = rebase(collect() + __ldg(p.att_b + head));
#pragma unroll
          for (int cc = 0; cc < kEpiCols; cc += 4) {
            const float4 v = row_ok ? __ldg(reinterpret_cast<const float4*>(xrow + xoff + cc)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 sc4 = *reinterpret_cast<const float4*>(p1 + cc);
            const float4 sh4 = *reinterpret_cast<const float4*>(p2 + cc);
            g[cc] = fmaf(pe, fmaf(v.x, sc4.x, sh4.x), g[cc]);
            g[cc + 1] = fmaf(pe, fmaf(v.y, sc4.y, sh4.y), g[cc + 1]);
            g[cc + 2] = fmaf(pe, fmaf(v.z, sc4.z, sh4.z), g[cc + 2]);
            g[cc + 3] = fmaf(pe, fmaf(v.w, sc4.w, sh4.w), g[cc + 3]);
          }
          stage_buf = next_buf;
          continue;
        }

        // ---------------- projected feature: the accumulator of its GEMM is in TMEM buffer `acc` ----------------
        const ActCoef ak = make_act(p.act[l]);
        LAFF_PROF_T(q0);
        ptx::mbar_wait(tfull_bar(acc), acc_phase, 15);
        ptx::tcgen05_fence_after();
        LAFF_PROF_T(q1);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * kBlockN + mycol);
        // ---- pass A: y = BN(act(acc + b)) written back to TMEM in place, partial logit over my 128 columns.  The
        //      TMEM load of chunk c + 1 is in flight while chunk c is computed. ----
        float part = 0.f;
        constexpr int kChunks = kEpiCols / kFuseChunk;
        {
          uint32_t ra[kFuseChunk], rb[kFuseChunk];
          ptx::tmem_ld_32x32b_x16(taddr, ra);
#pragma unroll 1
          for (int c = 0; c < kChunks; c += 2) {
            const int o0 = c * kFuseChunk, o1 = o0 + kFuseChunk;
            ptx::tmem_ld_wait();
            ptx::tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(o1), rb);
            part = pass_a_chunk(ra, p0 + o0, p1 + o0, p2 + o0, pw + o0, ak, part);
            ptx::tmem_st_32x32b_x16(taddr + static_cast<uint32_t>(o0), ra);  // pass B re-reads y, not the accumulator
            ptx::tmem_ld_wait();
            // (tcgen05.st consumed ra when it issued: reloading ra needs no wait::st)
            if (c + 2 < kChunks) ptx::tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(o1 + kFuseChunk), ra);
            part = pass_a_chunk(rb, p0 + o1, p1 + o1, p2 + o1, pw + o1, ak, part);
            ptx::tmem_st_32x32b_x16(taddr + static_cast<uint32_t>(o1), rb);
          }
          ptx::tmem_st_wait();
        }
        LAFF_PROF_T(q2);
        if (has_next) store_params(next_buf, nextp);
        post(part);
        LAFF_PROF_T(q3);
        // ---- pass B: g += exp(e - m_ref) * y (the softmax denominator cancels under the final L2 norm) ----
        const float pe = rebase(collect() + __ldg(p.att_b + head));
        LAFF_PROF_T(q4);
        //      (Measured: requesting 48 columns at a time -- x32 + x16 loads in flight -- is slower than this x16 ping-pong,
        //      2.6 k against 2.0 k cycles per feature in the pair variant, where the MMAs keep the TMEM port busy.)
        {
          uint32_t ra[kFuseChunk], rb[kFuseChunk];
          ptx::tmem_ld_32x32b_x16(taddr, ra);
#pragma unroll
          for (int c = 0; c < kChunks; c += 2) {
            const int o0 = c * kFuseChunk, o1 = o0 + kFuseChunk;
            ptx::tmem_ld_wait();
            ptx::tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(o1), rb);
#pragma unroll
            for (int j = 0; j < kFuseChunk; ++j) g[o0 + j] = fmaf(pe, __uint_as_float(ra[j]), g[o0 + j]);
            ptx::tmem_ld_wait();
            if (c + 2 < kChunks) ptx::tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(o1 + kFuseChunk), ra);
#pragma unroll
            for (int j = 0; j < kFuseChunk; ++j) g[o1 + j] = fmaf(pe, __uint_as_float(rb[j]), g[o1 + j]);
          }
        }
        ptx::tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 1) ptx::mbar_arrive(tempty_bar(acc));
          else ptx::mbar_arrive_remote(tempty_bar(acc), leader);
        }
        LAFF_PROF_T(q5);
        LAFF_PROF_ADD(0, q0, q1);
        LAFF_PROF_ADD(1, q1, q2);
        LAFF_PROF_ADD(2, q2, q3);
        LAFF_PROF_ADD(3, q3, q4);
        LAFF_PROF_ADD(4, q4, q5);
        LAFF_PROF_ADD(7, 0, 1);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        stage_buf = next_buf;
      }
      LAFF_PROF_T(fin0);
      // ---- L2 normalise over the whole head (4 partial sums of squares per row) and write ----
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < kEpiCols; ++j) ss = fmaf(g[j], g[j], ss);
      LAFF_PROF_T(fin_a);
      post(ss);
      const float den = sqrtf(collect()) + p.norm_eps;
      LAFF_PROF_T(fin_b);
      LAFF_PROF_ADD(3, fin_a, fin_b);     // exchange of the final normalisation goes to the 'collect' slot (debug only)
      const float inv = 1.0f / den;  // one IEEE division per row; x * (1/den) is within 1 ulp of x / den
#pragma unroll
      for (int j = 0; j < kEpiCols; ++j) g[j] *= inv;
      {
        const long long c0 = colbase + mycol;
        const long long row0 = static_cast<long long>(row_tile) * kUnitRows + row_sub * kBlockM + quad * 32;
        if (p.out) {  // fp32: 16 columns (64 bytes) of every row per round
#pragma unroll
          for (int k = 0; k < kEpiCols / 16; ++k) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              *reinterpret_cast<float4*>(tile_at(lane, c)) = make_float4(g[16 * k + 4 * c], g[16 * k + 4 * c + 1], g[16 * k + 4 * c + 2],
                                                                         g[16 * k + 4 * c + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const long long rr = row0 + t_row + 8 * i;
              if (rr < p.rows) *reinterpret_cast<uint4*>(p.out + rr * p.ld_out + c0 + 16 * k + 4 * t_piece) = *tile_at(t_row + 8 * i, t_piece);
            }
            __syncwarp();
          }
        }
        if (p.out16) {  // 16-bit: 32 columns (64 bytes) of every row per round
          uint16_t* o16 = static_cast<uint16_t*>(p.out16);
#pragma unroll
          for (int k = 0; k < kEpiCols / 32; ++k) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int j = 32 * k + 8 * c;
              uint4 pk;
              if (p.out16_dtype == LAFF_BF16) {
                __nv_bfloat162 a0 = __floats2bfloat162_rn(g[j], g[j + 1]), a1 = __floats2bfloat162_rn(g[j + 2], g[j + 3]);
                __nv_bfloat162 a2 = __floats2bfloat162_rn(g[j + 4], g[j + 5]), a3 = __floats2bfloat162_rn(g[j + 6], g[j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&a0); pk.y = *reinterpret_cast<uint32_t*>(&a1);
                pk.z = *reinterpret_cast<uint32_t*>(&a2); pk.w = *reinterpret_cast<uint32_t*>(&a3);
              } else {
                __half2 a0 = __floats2half2_rn(g[j], g[j + 1]), a1 = __floats2half2_rn(g[j + 2], g[j + 3]);
                __half2 a2 = __floats2half2_rn(g[j + 4], g[j + 5]), a3 = __floats2half2_rn(g[j + 6], g[j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&a0); pk.y = *reinterpret_cast<uint32_t*>(&a1);
                pk.z = *reinterpret_cast<uint32_t*>(&a2); pk.w = *reinterpret_cast<uint32_t*>(&a3);
              }
              *tile_at(lane, c) = pk;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const long long rr = row0 + t_row + 8 * i;
              if (rr < p.rows) *reinterpret_cast<uint4*>(o16 + rr * p.ld_out16 + c0 + 32 * k + 8 * t_piece) = *tile_at(t_row + 8 * i, t_piece);
            }
            __syncwarp();
          }
        }
      }
      LAFF_PROF_T(fin1);
      LAFF_PROF_ADD(6, fin_b, fin1);      // scale + stores
      LAFF_PROF_ADD(4, fin0, fin_a);      // sum of squares goes to the 'pass B' slot (debug only)
    }
  }

  ptx::tcgen05_fence_before();
  ptx::cluster_sync_all();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

}  // namespace laff

using namespace laff;

namespace {

std::atomic<int> g_fuse_variant{0};  // 0 = choose per call, 1 / 2 = force the cta_group::1 / ::2 kernel

// Co-resident clusters of the given variant (4-CTA clusters cannot straddle a GPC: 33 fit on a 148-SM B200).
template <int CG>
int fuse_max_clusters(int sms, int* out) {
  static int cached = 0;
  if (!cached) {
    auto kern = laff_fuse_kernel<CG>;
    LAFF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FuseCfg<CG>::kSmem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(sms / FuseCfg<CG>::kCluster * FuseCfg<CG>::kCluster));
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = FuseCfg<CG>::kSmem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = FuseCfg<CG>::kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    LAFF_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    LAFF_REQUIRE(n > 0, LAFF_ENODEV, "laff_fuse_forward: no %d-CTA cluster of the fused kernel fits this device", FuseCfg<CG>::kCluster);
    cached = n;
  }
  *out = cached;
  return LAFF_OK;
}

template <int CG>
int fuse_launch(const FuseTmaps& tm, const FuseParams& p, int clusters, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * FuseCfg<CG>::kCluster));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = FuseCfg<CG>::kSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = FuseCfg<CG>::kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LAFF_CUDA(cudaLaunchKernelEx(&cfg, laff_fuse_kernel<CG>, tm, p));
  count_launch();
  return LAFF_OK;
}

}  // namespace

// Debug: phase cycle counters of the fused kernel's epilogue (all zero unless the library was built with
// -DLAFF_FUSE_PROFILE): [wait accumulator, pass A, post, collect, pass B, first tiled feature, normalise + store, #features].
extern "C" int laff_debug_fuse_profile(unsigned long long* out8, int reset) {
  LAFF_REQUIRE(out8, LAFF_EINVAL, "laff_debug_fuse_profile: bad arguments");
#ifdef LAFF_FUSE_PROFILE
  LAFF_CUDA(cudaMemcpyFromSymbol(out8, g_fuse_prof, sizeof(unsigned long long) * 8));
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    LAFF_CUDA(cudaMemcpyToSymbol(g_fuse_prof, z, sizeof(z)));
  }
#else
  (void)reset;
  for (int i = 0; i < 8; ++i) out8[i] = 0;
#endif
  return LAFF_OK;
}

extern "C" int laff_set_fuse_variant(int cta_group) {
  LAFF_REQUIRE(cta_group >= 0 && cta_group <= 2, LAFF_EINVAL, "laff_set_fuse_variant: cta_group must be 0 (auto), 1 or 2");
  g_fuse_variant.store(cta_group);
  return LAFF_OK;
}

extern "C" int laff_get_fuse_variant(void) { return g_fuse_variant.load(); }

extern "C" int laff_fuse_forward(const laff_fuse_desc* d, long long rows, float* out, long long ld_out, void* out16,
                                 int out16_dtype, long long ld_out16, void* stream) {
  LAFF_REQUIRE(d && rows > 0, LAFF_EINVAL, "laff_fuse_forward: bad arguments");
  LAFF_REQUIRE(out || out16, LAFF_EINVAL, "laff_fuse_forward: no output buffer");
  LAFF_REQUIRE(d->head_dim == 2 * kBlockN, LAFF_ENOTSUP, "laff_fuse_forward: head_dim must be 512 (got %d); use "
               "laff_project + laff_attention_pool for other shapes", d->head_dim);
  LAFF_REQUIRE(d->heads > 0 && d->n_fc >= 1 && d->n_fc <= LAFF_FUSE_MAX_FC && d->n_tiled >= 0 &&
                   d->n_tiled <= LAFF_FUSE_MAX_TILED, LAFF_ENOTSUP, "laff_fuse_forward: n_fc=%d n_tiled=%d unsupported",
               d->n_fc, d->n_tiled);
  LAFF_REQUIRE(is16(d->dtype) && d->att_weight && d->att_bias, LAFF_EINVAL, "laff_fuse_forward: bad descriptor");
  LAFF_REQUIRE(rows < (1LL << 31) - kBlockM, LAFF_ENOTSUP, "laff_fuse_forward: too many rows");
  const int D = d->heads * d->head_dim;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  LAFF_REQUIRE(!out || (ld_out >= D && ld_out % 4 == 0 && al16(out)), LAFF_EINVAL, "laff_fuse_forward: out pitch/alignment");
  LAFF_REQUIRE(!out16 || (ld_out16 >= D && ld_out16 % 8 == 0 && al16(out16) && is16(out16_dtype)), LAFF_EINVAL,
               "laff_fuse_forward: out16 pitch/alignment");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  // Variant: wall time ~ waves x unit time; a cta_group::2 unit covers twice the rows in ~4/3 of the time, but only
  // the clusters that fit the GPCs run at once.
  int max1 = 0, max2 = 0;
  rc = fuse_max_clusters<1>(di.sms_total, &max1);
  if (rc) return rc;
  rc = fuse_max_clusters<2>(di.sms_total, &max2);
  if (rc) return rc;
  // an SM budget (laff_set_sm_limit) caps the number of co-resident clusters; below four SMs only pairs fit
  const bool budget = di.sms < di.sms_total;
  if (budget) {
    max1 = max1 < di.sms / 2 ? max1 : di.sms / 2;
    max2 = max2 < di.sms / 4 ? max2 : di.sms / 4;
    LAFF_REQUIRE(max1 >= 1, LAFF_EINVAL, "laff_fuse_forward: an SM budget of %d leaves no room for a 2-CTA cluster", di.sms);
  }
  const long long units1 = (rows + kBlockM - 1) / kBlockM * d->heads;
  const long long units2 = (rows + 2 * kBlockM - 1) / (2 * kBlockM) * d->heads;
  int cg = g_fuse_variant.load();
  if (max2 < 1) cg = 1;
  if (cg == 0) {
    // Measured (profiles/README.md): the pair variant only wins when a wide feature (K >= 3072, the 3981-word BoW)
    // makes the kernel operand-feed bound; otherwise running on all 148 SMs is worth more than the cheaper feed.
    int kmax = 0;
    for (int l = 0; l < d->n_fc; ++l) kmax = d->fc[l].K > kmax ? d->fc[l].K : kmax;
    const long long waves1 = (units1 + max1 - 1) / max1;
    const long long waves2 = (units2 + max2 - 1) / max2;
    cg = (kmax >= 3072 && waves2 * 5 <= waves1 * 6) ? 2 : 1;  // a pair wave takes ~0.83 of a cta_group::1 wave there
  }
  const int b_rows = kBlockN / cg;
  FuseTmaps tm;
  FuseParams p;
  memset(&tm, 0, sizeof(tm));
  memset(&p, 0, sizeof(p));
  p.n_fc = d->n_fc;
  p.n_tiled = d->n_tiled;
  p.heads = d->heads;
  p.rows = rows;
  // Processing order of the projected features: widest first (stable).  The pooled sum does not depend on the order
  // beyond rounding, but the pipeline does: the MMAs of feature l + 2 can only start once the epilogue has released
  // the TMEM buffer of feature l, so a wide GEMM placed last in a unit leaves the epilogue waiting for it.
  int order[LAFF_FUSE_MAX_FC];
  for (int l = 0; l < d->n_fc; ++l) order[l] = l;
  for (int a = 1; a < d->n_fc; ++a)
    for (int b = a; b > 0 && d->fc[order[b]].K > d->fc[order[b - 1]].K; --b) { const int t = order[b]; order[b] = order[b - 1]; order[b - 1] = t; }
  for (int l = 0; l < d->n_fc; ++l) {
    const auto& f = d->fc[order[l]];
    LAFF_REQUIRE(f.x16 && f.w16 && f.K > 0 && f.K % 8 == 0 && f.ldx % 8 == 0 && f.ldw % 8 == 0 && f.ldx >= f.K && f.ldw >= f.K,
                 LAFF_EINVAL, "laff_fuse_forward: fc feature %d: K and pitches must be multiples of 8", l);
    LAFF_REQUIRE((f.bn_scale == nullptr) == (f.bn_shift == nullptr) && f.activation >= 0 && f.activation <= 3, LAFF_EINVAL,
                 "laff_fuse_forward: fc feature %d: bad bn/activation", l);
    rc = make_tmap_2d(&tm.x[l], f.x16, d->dtype, static_cast<uint64_t>(rows), static_cast<uint64_t>(f.K), static_cast<uint64_t>(f.ldx), kBlockM);
    if (rc) return rc;
    rc = make_tmap_2d(&tm.w[l], f.w16, d->dtype, static_cast<uint64_t>(D), static_cast<uint64_t>(f.K), static_cast<uint64_t>(f.ldw), static_cast<uint32_t>(b_rows));
    if (rc) return rc;
    p.num_kb[l] = (f.K + kBlockK - 1) / kBlockK;
    p.act[l] = f.activation;
    p.bias[l] = f.bias;
    p.bn_scale[l] = f.bn_scale;
    p.bn_shift[l] = f.bn_shift;
  }
  for (int l = 0; l < d->n_tiled; ++l) {
    const auto& t = d->tiled[l];
    LAFF_REQUIRE(t.x && t.in_dim > 0 && t.in_dim % kEpiCols == 0 && D % t.in_dim == 0 && t.ld >= t.in_dim && t.ld % 4 == 0 && al16(t.x),
                 LAFF_ENOTSUP, "laff_fuse_forward: tiled feature %d: in_dim must be a multiple of 128 dividing D, 16-byte aligned rows", l);
    LAFF_REQUIRE((t.bn_scale == nullptr) == (t.bn_shift == nullptr), LAFF_EINVAL, "laff_fuse_forward: tiled feature %d: bn mismatch", l);
    p.tiled_x[l] = t.x;
    p.tiled_ld[l] = t.ld;
    p.tiled_in_dim[l] = t.in_dim;
    p.tiled_scale[l] = t.bn_scale;
    p.tiled_shift[l] = t.bn_shift;
  }
  p.att_w = d->att_weight;
  p.att_b = d->att_bias;
  p.out = out;
  p.ld_out = ld_out;
  p.out16 = out16;
  p.out16_dtype = out16_dtype;
  p.ld_out16 = ld_out16;
  p.norm_eps = static_cast<float>(d->norm_eps);
  p.idesc = make_idesc_f16(d->dtype, kBlockM * cg, kBlockN);
  const long long units = cg == 2 ? units2 : units1;
  LAFF_REQUIRE(units < (1LL << 31), LAFF_ENOTSUP, "laff_fuse_forward: too many work units");
  p.total_units = static_cast<int>(units);
  int clusters = cg == 2 ? max2 : max1;
  if (clusters > p.total_units) clusters = p.total_units;
  if (cg == 2) return fuse_launch<2>(tm, p, clusters, static_cast<cudaStream_t>(stream));
  return fuse_launch<1>(tm, p, clusters, static_cast<cudaStream_t>(stream));
}
