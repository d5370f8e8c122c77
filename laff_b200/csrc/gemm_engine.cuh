// Persistent warp-specialised tcgen05 GEMM engine for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T, 16-bit operands (fp16 or
// bf16, both K-major = row-major with K contiguous), fp32 accumulation in TMEM.
//
//   warp 0      TMA producer      (one lane)  global -> smem ring, SWIZZLE_128B tiles of 64 K-elements
//   warp 1      MMA issuer        (one lane, leader CTA of the pair only)  tcgen05.mma, commits to mbarriers
//   warp 2      TMEM allocator
//   warps 4..11 epilogue          tcgen05.ld: warp w reads TMEM lanes 32*(w%4)..+31 (one accumulator row per thread);
//                                 warps 4..7 take columns 0..127 of a tile, warps 8..11 columns 128..255
//
// The accumulator is double buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile
// i+1.  CG (cta_group) = 1: one CTA owns a 128 x 256 tile.  CG = 2: a CTA pair (cluster of 2) owns a 256 x 256 tile;
// each CTA loads its 128 rows of A and its 128 rows (=columns of D) of B, the leader issues cta_group::2 MMAs which
// read both CTAs' shared memory and write both CTAs' TMEM.
//
// Work is cut into units = (m_tile, chunk of consecutive n_tiles).  A cluster walks units u = cluster_id,
// cluster_id + num_clusters, ... ; all three roles decode the same static schedule independently.  The epilogue
// functor keeps per-row state across the n_tiles of a unit (that is what the streaming rank / top-k needs).
#pragma once
#include "ptx.cuh"

#define LAFF_COUNT_LAUNCH_DECLARED
namespace laff { void count_launch(int n = 1); }

namespace laff {

constexpr int kBlockM = 128;  // accumulator rows per CTA (TMEM lanes)
constexpr int kBlockN = 256;  // accumulator columns per tile (MMA N)
constexpr int kBlockK = 64;   // 16-bit K elements per pipeline stage = one 128-byte swizzle span
constexpr int kUmmaK = 16;    // K per tcgen05.mma for 16-bit operands
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;
constexpr int kNumThreads = 384;
constexpr int kEpiWarps = 8;                  // two per TMEM lane quadrant
constexpr int kEpiThreads = kEpiWarps * 32;   // epilogue threads per CTA: (row, column half)
constexpr int kEpiCols = kBlockN / 2;         // columns of a tile one epilogue thread walks

template <int CG>
struct EngineCfg {
  static constexpr int kStages = (CG == 2) ? 6 : 4;
  static constexpr int kBRows = kBlockN / CG;  // rows of B each CTA loads per stage
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTxBytes = kStageBytes * CG;  // bytes landing per stage across the CTAs of one MMA
  static constexpr int kBarBytes = 256;
  static constexpr int kPipeBytes = kStages * kStageBytes + kBarBytes;
  // + per-epilogue scratch (Epi::kSmemBytes) + alignment slack
  static constexpr int smem_bytes(int epi_bytes) { return kPipeBytes + epi_bytes + 1024; }
};

struct Sched {
  int m_tiles;      // row tiles of 128*CG rows
  int n_tiles;      // column tiles of 256
  int chunk_tiles;  // n_tiles per unit
  int n_chunks;     // ceil(n_tiles / chunk_tiles)
  int m_group;      // m_tiles that sweep the N range together (keeps their A rows hot in L2)
  int diag;         // 1: unit u = the single tile that holds D[r, r] for the rows of m_tile u
  int total_units;
  int rows_per_mtile;
};

struct Unit {
  int m_tile, chunk, n_begin, n_end;
};

__device__ __forceinline__ Unit decode_unit(const Sched& s, int u) {
  Unit r;
  if (s.diag) {
    r.m_tile = u;
    r.chunk = 0;
    r.n_begin = (u * s.rows_per_mtile) / kBlockN;
    r.n_end = r.n_begin + 1;
    return r;
  }
  const int per_group = s.m_group * s.n_chunks;
  const int full_groups = s.m_tiles / s.m_group;
  int g = u / per_group;
  int gsize = s.m_group;
  int rem = u - g * per_group;
  if (g >= full_groups) {
    g = full_groups;
    gsize = s.m_tiles - full_groups * s.m_group;
    rem = u - full_groups * per_group;
  }
  r.chunk = rem / gsize;
  r.m_tile = g * s.m_group + (rem - r.chunk * gsize);
  r.n_begin = r.chunk * s.chunk_tiles;
  r.n_end = min(r.n_begin + s.chunk_tiles, s.n_tiles);
  return r;
}

inline Sched make_sched(int M, int N, int cg, int chunk_tiles, int m_group, int diag) {
  Sched s;
  s.rows_per_mtile = kBlockM * cg;
  s.m_tiles = (M + s.rows_per_mtile - 1) / s.rows_per_mtile;
  s.n_tiles = (N + kBlockN - 1) / kBlockN;
  if (chunk_tiles < 1) chunk_tiles = 1;
  if (chunk_tiles > s.n_tiles) chunk_tiles = s.n_tiles > 0 ? s.n_tiles : 1;
  s.chunk_tiles = chunk_tiles;
  s.n_chunks = (s.n_tiles + chunk_tiles - 1) / chunk_tiles;
  if (m_group < 1) m_group = 1;
  if (m_group > s.m_tiles) m_group = s.m_tiles > 0 ? s.m_tiles : 1;
  s.m_group = m_group;
  s.diag = diag;
  s.total_units = diag ? s.m_tiles : s.m_tiles * s.n_chunks;
  return s;
}

template <int CG, class Epi>
__global__ void __launch_bounds__(kNumThreads, 1)
    gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int num_kb,
                uint32_t idesc, Sched sched, uint64_t hintA, uint64_t hintB, typename Epi::Params ep) {
  using Cfg = EngineCfg<CG>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);

  const uint32_t sA = base;
  const uint32_t sB = base + Cfg::kStages * Cfg::kABytes;
  const uint32_t bar0 = base + Cfg::kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * Cfg::kStages + kAccStages + a); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * Cfg::kStages + 2 * kAccStages);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem + Cfg::kStages * Cfg::kStageBytes + 8 * (2 * Cfg::kStages + 2 * kAccStages));

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = static_cast<int>(ptx::lane_id());
  const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
  const int cluster_id = static_cast<int>(blockIdx.x) / CG;
  const int num_clusters = static_cast<int>(gridDim.x) / CG;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);   // leader's arrive.expect_tx; TMA bytes of all CTAs complete_tx here
      ptx::mbar_init(empty_bar(s), 1);  // one tcgen05.commit (multicast to both CTAs when CG == 2)
    }
    for (int a = 0; a < kAccStages; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);        // one tcgen05.commit per tile
      ptx::mbar_init(tempty_bar(a), kEpiWarps * CG);  // one arrive per epilogue warp of every CTA of the pair
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<CG>(tmem_slot, kTmemCols);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tcgen05_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = cluster_id; u < sched.total_units; u += num_clusters) {
        const Unit un = decode_unit(sched, u);
        const int m0 = un.m_tile * (kBlockM * CG) + static_cast<int>(cta_rank) * kBlockM;
        for (int n = un.n_begin; n < un.n_end; ++n) {
          const int n0 = n * kBlockN + static_cast<int>(cta_rank) * Cfg::kBRows;
          for (int kb = 0; kb < num_kb; ++kb) {
            ptx::mbar_wait(empty_bar(stage), phase ^ 1u, 1);
            const uint32_t dstA = sA + stage * Cfg::kABytes;
            const uint32_t dstB = sB + stage * Cfg::kBBytes;
            if constexpr (CG == 1) {
              ptx::mbar_arrive_expect_tx(full_bar(stage), Cfg::kTxBytes);
              ptx::tma_load_2d(dstA, &tmA, full_bar(stage), kb * kBlockK, m0, hintA);
              ptx::tma_load_2d(dstB, &tmB, full_bar(stage), kb * kBlockK, n0, hintB);
            } else {
              if (cta_rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), Cfg::kTxBytes);
              const uint32_t leader_full = ptx::mapa(full_bar(stage), 0);
              ptx::tma_load_2d_pair(dstA, &tmA, leader_full, kb * kBlockK, m0, hintA);
              ptx::tma_load_2d_pair(dstB, &tmB, leader_full, kb * kBlockK, n0, hintB);
            }
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================================ MMA issuer ============================================
    if (cta_rank == 0 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = cluster_id; u < sched.total_units; u += num_clusters) {
        const Unit un = decode_unit(sched, u);
        for (int n = un.n_begin; n < un.n_end; ++n) {
          ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 2);  // epilogue has drained this accumulator buffer
          ptx::tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kBlockN);
          for (int kb = 0; kb < num_kb; ++kb) {
            ptx::mbar_wait(full_bar(stage), phase, 3);  // TMA bytes of this stage have landed (both CTAs)
            ptx::tcgen05_fence_after();
            const uint64_t da = ptx::make_smem_desc_sw128(sA + stage * Cfg::kABytes);
            const uint64_t db = ptx::make_smem_desc_sw128(sB + stage * Cfg::kBBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              // +32 bytes per UMMA_K inside the 128-byte swizzle span = +2 in the (addr >> 4) field
              ptx::umma_f16<CG>(d_tmem, da + 2u * k, db + 2u * k, idesc, static_cast<uint32_t>((kb | k) != 0));
            }
            ptx::umma_commit<CG>(empty_bar(stage));  // frees the smem slot once these MMAs retire
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
          }
          ptx::umma_commit<CG>(tfull_bar(acc));  // accumulator complete -> epilogue
          if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ============================================= epilogue =============================================
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;   // which 128 columns of each tile
    const int row_in_cta = quad * 32 + lane;
    Epi epi(ep, smem + Cfg::kPipeBytes, half * kBlockM + row_in_cta);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = cluster_id; u < sched.total_units; u += num_clusters) {
      const Unit un = decode_unit(sched, u);
      const int row = un.m_tile * (kBlockM * CG) + static_cast<int>(cta_rank) * kBlockM + row_in_cta;
      epi.unit_begin(row, un);
      for (int n = un.n_begin; n < un.n_end; ++n) {
        ptx::mbar_wait(tfull_bar(acc), acc_phase, 4);
        ptx::tcgen05_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                               static_cast<uint32_t>(acc * kBlockN + half * kEpiCols);
#pragma unroll 1
        for (int c = 0; c < kEpiCols / 32; ++c) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c * 32), r);
          ptx::tmem_ld_wait();
          epi.chunk(r, row, n * kBlockN + half * kEpiCols + c * 32);
        }
        ptx::tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 1) ptx::mbar_arrive(tempty_bar(acc));
          else ptx::mbar_arrive_remote(tempty_bar(acc), 0);
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
      }
      epi.unit_end(row, un);
    }
  }

  // ============================================== teardown ==============================================
  ptx::tcgen05_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Host-side launch (one persistent CTA, or CTA pair, per SM)
// ------------------------------------------------------------------------------------------------------------
template <int CG, class Epi>
inline cudaError_t launch_gemm_kernel(const CUtensorMap& tmA, const CUtensorMap& tmB, int num_kb, uint32_t idesc,
                                      const Sched& s, const typename Epi::Params& ep, uint64_t hintA, uint64_t hintB,
                                      int sms, cudaStream_t st) {
  using Cfg = EngineCfg<CG>;
  constexpr int kSmem = Cfg::smem_bytes(Epi::kSmemBytes);
  static_assert(kSmem <= 227 * 1024, "shared memory budget exceeded");
  auto kern = gemm_kernel<CG, Epi>;
  static bool configured = false;  // per instantiation; one device per process
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (s.total_units <= 0) return cudaSuccess;
  int clusters = sms / CG;
  if (clusters > s.total_units) clusters = s.total_units;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * CG));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  count_launch();
  return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, num_kb, idesc, s, hintA, hintB, ep);
}

}  // namespace laff
