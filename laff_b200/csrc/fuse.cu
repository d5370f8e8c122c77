// Multi-feature fusion kernels.
//   laff_l2norm_quantize / laff_cast_pad_16 / laff_split3_16 : operand preparation for the tensor-core GEMMs
//   laff_bn_fold        : eval-mode BatchNorm1d as an affine map                       model/model.py:232, :273-274
//   laff_project        : y = BN(act(x W^T + b)), tcgen05 GEMM with fused epilogue      model/model.py:257-276
//   laff_attention_pool : per-head LAFF block over L features                          model/Attention.py:78-105, :508-531
//   laff_frame_pool     : frame-level LAFF block                                       model/model.py:2160-2173
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gemm_engine.cuh"
#include "host_util.cuh"

namespace laff {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

// fp16 operands (the default: 8x finer rounding than bf16 at the same MMA rate) have a narrow range: a raw
// feature beyond +-65504 saturates instead of turning into an inf that would poison a whole output row (NaN stays NaN).
__device__ __forceinline__ uint16_t to16(float v, int dtype) {
  if (dtype == LAFF_BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  return __half_as_ushort(__float2half_rn(v != v ? v : fminf(fmaxf(v, -65504.f), 65504.f)));
}
__device__ __forceinline__ float from16(uint16_t b, int dtype) {
  if (dtype == LAFF_BF16) return __bfloat162float(__ushort_as_bfloat16(b));
  return __half2float(__ushort_as_half(b));
}

// ------------------------------------------------------------------------------------------------------------
// S1: per-head L2 normalisation + rounding.  One warp per (row, head).
// ------------------------------------------------------------------------------------------------------------
__global__ void l2norm_quantize_kernel(const float* __restrict__ x, long long rows, int heads, int dh, long long ldx,
                                       float eps, int normalise, int out_dtype, void* __restrict__ out, long long ld_out) {
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = rows * heads;
  if (warp >= total) return;
  const long long row = warp / heads;
  const int h = static_cast<int>(warp - row * heads);
  const float* src = x + row * ldx + static_cast<long long>(h) * dh;
  float den = 1.0f;
  if (normalise) {
    float ss = 0.f;
    for (int d = lane; d < dh; d += 32) {
      const float v = src[d];
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    den = sqrtf(ss) + eps;  // loss.py:11  norm = sqrt(sum(x^2)) + eps + 1e-14
  }
  const long long o = row * ld_out + static_cast<long long>(h) * dh;
  for (int d = lane; d < dh; d += 32) {
    const float v = normalise ? src[d] / den : src[d];  // loss.py:12  torch.div(X, norm)
    if (out_dtype == LAFF_F32)
      static_cast<float*>(out)[o + d] = v;
    else
      static_cast<uint16_t*>(out)[o + d] = to16(v, out_dtype);
  }
}

// Vector form for head sizes that are multiples of 128 with 16-byte aligned rows (every shipped setting: d_h = 512): the
// head is read ONCE with 16-byte loads (NV float4 per lane, all in flight together), kept in registers across the
// reduction, and written with 16-byte (fp32) / 8-byte (16-bit) stores -- the scalar kernel above reads it twice with
// 4-byte accesses (0.78 of the measured copy bandwidth).
template <int NV>
__global__ void __launch_bounds__(256) l2norm_quantize_vec_kernel(const float* __restrict__ x, long long rows, int heads, long long ldx,
                                                                  float eps, int normalise, int out_dtype, void* __restrict__ out,
                                                                  long long ld_out) {
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = rows * heads;
  if (warp >= total) return;
  const long long row = warp / heads;
  const int h = static_cast<int>(warp - row * heads);
  constexpr int dh = NV * 128;
  const float4* src = reinterpret_cast<const float4*>(x + row * ldx + static_cast<long long>(h) * dh);
  float4 v[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) v[q] = __ldcs(src + lane + 32 * q);   // streamed: read once
  float inv = 1.0f;
  float den = 1.0f;
  if (normalise) {
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      ss = fmaf(v[q].x, v[q].x, ss);
      ss = fmaf(v[q].y, v[q].y, ss);
      ss = fmaf(v[q].z, v[q].z, ss);
      ss = fmaf(v[q].w, v[q].w, ss);
    }
    ss = warp_sum(ss);
    den = sqrtf(ss) + eps;  // loss.py:11
  }
  (void)inv;
  const long long o = row * ld_out + static_cast<long long>(h) * dh;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    float4 r = v[q];
    if (normalise) { r.x = r.x / den; r.y = r.y / den; r.z = r.z / den; r.w = r.w / den; }  // loss.py:12 torch.div
    const int c = 4 * (lane + 32 * q);
    if (out_dtype == LAFF_F32) {
      *reinterpret_cast<float4*>(static_cast<float*>(out) + o + c) = r;
    } else {
      uint2 pk;
      pk.x = static_cast<uint32_t>(to16(r.x, out_dtype)) | (static_cast<uint32_t>(to16(r.y, out_dtype)) << 16);
      pk.y = static_cast<uint32_t>(to16(r.z, out_dtype)) | (static_cast<uint32_t>(to16(r.w, out_dtype)) << 16);
      *reinterpret_cast<uint2*>(static_cast<uint16_t*>(out) + o + c) = pk;
    }
  }
}

// fp32 -> 16-bit operand copies.  HBM-bound: one thread moves 8 consecutive columns (two 16-byte loads, one 16-byte store
// per output plane); VEC = false is the element-wise path for pitches / pointers that are not 16-byte aligned.
__device__ __forceinline__ void load8(const float* __restrict__ src, int valid, bool vec, float (&v)[8]) {
  if (vec && valid >= 8) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = j < valid ? __ldg(src + j) : 0.f;   // columns [cols, cols_pad) are zero
  }
}
__device__ __forceinline__ uint4 pack8(const uint16_t (&h)[8]) {
  uint4 o;
  o.x = h[0] | (static_cast<uint32_t>(h[1]) << 16);
  o.y = h[2] | (static_cast<uint32_t>(h[3]) << 16);
  o.z = h[4] | (static_cast<uint32_t>(h[5]) << 16);
  o.w = h[6] | (static_cast<uint32_t>(h[7]) << 16);
  return o;
}
__device__ __forceinline__ void store8(uint16_t* __restrict__ dst, const uint16_t (&h)[8], bool vec) {
  if (vec) {
    *reinterpret_cast<uint4*>(dst) = pack8(h);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = h[j];
  }
}

// cols_pad % 8 == 0 (the callers pad to TMA's 16-byte pitch granularity)
__global__ void cast_pad_kernel(const float* __restrict__ x, long long rows, int cols, long long ldx, int out_dtype,
                                uint16_t* __restrict__ out, int cols_pad, long long ld_out, int vec_in, int vec_out) {
  const int gpr = cols_pad >> 3;
  const long long total = rows * gpr;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / gpr;
    const int c = static_cast<int>(i - r * gpr) << 3;
    float v[8];
    load8(x + r * ldx + c, cols - c, vec_in != 0, v);
    uint16_t h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = to16(v[j], out_dtype);
    store8(out + r * ld_out + c, h, vec_out != 0);
  }
}

__global__ void split3_kernel(const float* __restrict__ x, long long rows, int cols, long long ldx, int side,
                              int out_dtype, uint16_t* __restrict__ out, int cols_pad, long long ld_out, int vec_in,
                              int vec_out) {
  const int gpr = cols_pad >> 3;
  const long long total = rows * gpr;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / gpr;
    const int c = static_cast<int>(i - r * gpr) << 3;
    float v[8];
    load8(x + r * ldx + c, cols - c, vec_in != 0, v);
    uint16_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hi[j] = to16(v[j], out_dtype);
      lo[j] = to16(v[j] - from16(hi[j], out_dtype), out_dtype);
    }
    uint16_t* o = out + r * ld_out + c;
    // left: [hi | lo | hi]   right: [hi | hi | lo]   =>  left . right = hi.hi + lo.hi + hi.lo
    store8(o, hi, vec_out != 0);
    store8(o + cols_pad, side == 0 ? lo : hi, vec_out != 0);
    store8(o + 2 * static_cast<long long>(cols_pad), side == 0 ? hi : lo, vec_out != 0);
  }
}

__global__ void bn_fold_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, int D, float* __restrict__ scale,
                               float* __restrict__ shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  const float invstd = 1.0f / sqrtf(var[i] + eps);
  const float s = (w ? w[i] : 1.0f) * invstd;
  scale[i] = s;
  shift[i] = (b ? b[i] : 0.0f) - mean[i] * s;
}

// ------------------------------------------------------------------------------------------------------------
// F1: projection epilogue  y = BN(act(acc + bias))
// ------------------------------------------------------------------------------------------------------------
// tanh(z) = 1 - 2 / (exp(2z) + 1) with the hardware ex2 / rcp approximations: |error| < 3e-7 absolute over the whole
// range (the reference's tanh output feeds a unit-norm embedding compared at 2e-6), 5 instructions instead of ~40.
__device__ __forceinline__ float tanh_fast(float z) {
  const float e = __expf(2.0f * z);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__device__ __forceinline__ float activate(float z, int act) {
  switch (act) {
    case LAFF_ACT_TANH: return tanh_fast(z);
    case LAFF_ACT_RELU: return fmaxf(z, 0.f);
    case LAFF_ACT_SIGMOID: return __fdividef(1.0f, 1.0f + __expf(-z));
    default: return z;
  }
}

// Projection epilogue.  The accumulator arrives one row per thread (TMEM lane = row); each warp transposes its
// 32 x 32 chunk through a padded shared-memory tile so that a lane owns one output column: the per-column bias / BN
// parameters become three registers per lane and every global store is a fully coalesced 128-byte row segment.
struct EpiProject {
  struct Params {
    float* y;
    long long ldy;
    long long M;
    int N;
    const float* bias;
    const float* bn_scale;
    const float* bn_shift;
    int act;
  };
  static constexpr int kTile = 32 * 33;
  static constexpr int kSmemBytes = kEpiWarps * kTile * 4;
  Params p;
  float* tile;
  int lane;
  __device__ EpiProject(const Params& p_, uint8_t* smem, int epi_tid) : p(p_) {
    tile = reinterpret_cast<float*>(smem) + (epi_tid >> 5) * kTile;
    lane = epi_tid & 31;
  }
  __device__ __forceinline__ void unit_begin(int, const Unit&) {}
  __device__ __forceinline__ void unit_end(int, const Unit&) {}
  __device__ __forceinline__ void chunk(const uint32_t (&r)[32], int row, int col0) {
#pragma unroll
    for (int j = 0; j < 32; ++j) tile[j * 33 + lane] = __uint_as_float(r[j]);  // tile[col][row]
    __syncwarp();
    const int col = col0 + lane;
    const bool cok = col < p.N;
    const float b = (cok && p.bias) ? __ldg(p.bias + col) : 0.f;
    const float sc = (cok && p.bn_scale) ? __ldg(p.bn_scale + col) : 1.f;
    const float sh = (cok && p.bn_shift) ? __ldg(p.bn_shift + col) : 0.f;
    const int row0 = row - lane;
    float* dst = p.y + static_cast<long long>(row0) * p.ldy + col;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      float z = tile[lane * 33 + i] + b;
      z = activate(z, p.act);
      z = fmaf(z, sc, sh);
      if (cok && row0 + i < p.M) dst[static_cast<long long>(i) * p.ldy] = z;
    }
    __syncwarp();
  }
};

// ------------------------------------------------------------------------------------------------------------
// F5/F6: LAFF block.  One warp per (row, head).  A lane owns VPL elements of the head:
//   VEC  : 4 consecutive elements at 4*lane + 128*q (128-bit loads / stores; head_dim % 128 == 0, 16-byte aligned rows)
//   !VEC : elements lane + 32*t (any head_dim % 32 == 0)
// All L feature slices of the head are held in registers (LMAX * VPL values), so every input byte is read once.
// ------------------------------------------------------------------------------------------------------------
template <int VPL, bool VEC>
__device__ __forceinline__ int pool_elem(int lane, int t) {
  return VEC ? (4 * lane + 128 * (t >> 2) + (t & 3)) : (lane + 32 * t);
}

template <int VPL, bool VEC>
__device__ __forceinline__ void pool_load(const float* __restrict__ p, int lane, float (&v)[VPL]) {
  if constexpr (VEC) {
#pragma unroll
    for (int q = 0; q < VPL / 4; ++q) {
      const float4 x = *reinterpret_cast<const float4*>(p + 4 * lane + 128 * q);
      v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int t = 0; t < VPL; ++t) v[t] = p[lane + 32 * t];
  }
}

// PLAIN = neither with_ave nor mul (the shipped setting): the mean-pooled vector is never needed, which frees VPL registers
// per lane and lets a fifth block of the d_h = 512, L <= 4 instantiation fit an SM (a warp holds its L heads in registers
// from the loads to the final normalisation, so occupancy is what keeps enough loads in flight).
template <int VPL, int LMAX, bool VEC, bool PLAIN>
__global__ void __launch_bounds__(128, (PLAIN && VPL == 16 && LMAX == 4) ? 5 : 1) attention_pool_kernel(laff_pool_desc d, long long rows, float* __restrict__ out,
                                                            long long ld_out, void* __restrict__ out16, int out16_dtype,
                                                            long long ld_out16, float* __restrict__ att) {
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = rows * d.heads;
  if (warp >= total) return;
  const long long row = warp / d.heads;
  const int h = static_cast<int>(warp - row * d.heads);
  const int dh = d.head_dim;
  const int L = d.n_features;

  float y[LMAX][VPL];
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    if (l < L) {
      const laff_pool_source& s = d.src[l];
      if (s.kind == 0) {
        pool_load<VPL, VEC>(s.src + row * s.ld + static_cast<long long>(h) * dh, lane, y[l]);
      } else {
        // "no-transform": x tiled along D (x.repeat(1, heads)), then BatchNorm1d(D)   model/model.py:1822-1823
        const float* p = s.src + row * s.ld;
        if constexpr (VEC) {
#pragma unroll
          for (int q = 0; q < VPL / 4; ++q) {
            const int col = h * dh + 4 * lane + 128 * q;
            const float4 x = *reinterpret_cast<const float4*>(p + col % s.in_dim);  // in_dim % 4 == 0: never wraps
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s.bn_scale) {
              sc = __ldg(reinterpret_cast<const float4*>(s.bn_scale + col));
              sh = __ldg(reinterpret_cast<const float4*>(s.bn_shift + col));
            }
            y[l][4 * q] = fmaf(x.x, sc.x, sh.x); y[l][4 * q + 1] = fmaf(x.y, sc.y, sh.y);
            y[l][4 * q + 2] = fmaf(x.z, sc.z, sh.z); y[l][4 * q + 3] = fmaf(x.w, sc.w, sh.w);
          }
        } else {
#pragma unroll
          for (int t = 0; t < VPL; ++t) {
            const int col = h * dh + lane + 32 * t;
            float v = p[col % s.in_dim];
            if (s.bn_scale) v = fmaf(v, __ldg(s.bn_scale + col), __ldg(s.bn_shift + col));
            y[l][t] = v;
          }
        }
      }
    } else {
#pragma unroll
      for (int t = 0; t < VPL; ++t) y[l][t] = 0.f;
    }
  }
  float w[VPL];
  pool_load<VPL, VEC>(d.att_weight + static_cast<long long>(h) * dh, lane, w);

  // raw_global_emb = mean over features (Attention.py:81)
  float mean[PLAIN ? 1 : VPL];
  if constexpr (!PLAIN) {
    const float invL = 1.0f / static_cast<float>(L);
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      float s = 0.f;
#pragma unroll
      for (int l = 0; l < LMAX; ++l)
        if (l < L) s += y[l][t];
      mean[t] = s * invL;
    }
  }

  // logits e_l = w_h . common_l + c_h  (Attention.py:88), common = local (* mean if mul, Attention.py:83-86)
  float e[LMAX];
  const float cb = __ldg(d.att_bias + h);
  float emax = -INFINITY;
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    if (l < L) {
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < VPL; ++t) {
        if constexpr (PLAIN) s = fmaf(w[t], y[l][t], s);
        else s = fmaf(w[t], d.mul ? y[l][t] * mean[t] : y[l][t], s);
      }
      e[l] = warp_sum(s) + cb;
      emax = fmaxf(emax, e[l]);
    } else {
      e[l] = -INFINITY;
    }
  }
  // softmax over features (Attention.py:89)
  float z = 0.f;
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    if (l < L) {
      e[l] = expf(e[l] - emax);
      z += e[l];
    }
  }
  const float invz = 1.0f / z;
  // weighted sum (+ omega * mean-pool when with_ave, Attention.py:93-101)
  float g[VPL];
  float ss = 0.f;
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
      if (l < L) s = fmaf(e[l] * invz, y[l][t], s);
    if constexpr (!PLAIN) {
      if (d.with_ave) s = fmaf(d.omega, mean[t] * static_cast<float>(L), s);  // sum_l omega * raw_global_emb
    }
    g[t] = s;
    ss = fmaf(s, s, ss);
  }
  ss = warp_sum(ss);
  const float den = sqrtf(ss) + static_cast<float>(d.norm_eps);  // l2norm(eps=0): + 0 + 1e-14  (Attention.py:103)
#pragma unroll
  for (int t = 0; t < VPL; ++t) g[t] = g[t] / den;
  const long long c0 = static_cast<long long>(h) * dh;
  if constexpr (VEC) {
#pragma unroll
    for (int q = 0; q < VPL / 4; ++q) {
      const long long c = c0 + 4 * lane + 128 * q;
      if (out) *reinterpret_cast<float4*>(out + row * ld_out + c) = make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
      if (out16) {
        uint2 pk;
        pk.x = static_cast<uint32_t>(to16(g[4 * q], out16_dtype)) | (static_cast<uint32_t>(to16(g[4 * q + 1], out16_dtype)) << 16);
        pk.y = static_cast<uint32_t>(to16(g[4 * q + 2], out16_dtype)) | (static_cast<uint32_t>(to16(g[4 * q + 3], out16_dtype)) << 16);
        *reinterpret_cast<uint2*>(static_cast<uint16_t*>(out16) + row * ld_out16 + c) = pk;
      }
    }
  } else {
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      const long long c = c0 + lane + 32 * t;
      if (out) out[row * ld_out + c] = g[t];
      if (out16) static_cast<uint16_t*>(out16)[row * ld_out16 + c] = to16(g[t], out16_dtype);
    }
  }
  if (att && lane == 0) {
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
      if (l < L) {
        float a = e[l] * invz;
        if (d.with_ave) a += d.omega / static_cast<float>(L);  // Attention.py:97
        att[(row * d.heads + h) * L + l] = a;
      }
  }
}

// ------------------------------------------------------------------------------------------------------------
// F7: frame-level LAFF block, one warp per video, online softmax over frames.
// ------------------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(128) frame_pool_kernel(const float* __restrict__ frames, long long B, int F, int dim,
                                                        const float* __restrict__ att_w, float att_b, int with_ave,
                                                        int mul, float omega, float norm_eps, float* __restrict__ out,
                                                        long long ld_out) {
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* base = frames + warp * static_cast<long long>(F) * dim;
  float w[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) w[t] = __ldg(att_w + lane + 32 * t);

  float mean[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) mean[t] = 0.f;
  if (mul || with_ave) {
    for (int f = 0; f < F; ++f) {
      const float* p = base + static_cast<long long>(f) * dim;
#pragma unroll
      for (int t = 0; t < VPL; ++t) mean[t] += p[lane + 32 * t];
    }
    const float invF = 1.0f / static_cast<float>(F);
#pragma unroll
    for (int t = 0; t < VPL; ++t) mean[t] *= invF;
  }

  float acc[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) acc[t] = 0.f;
  float m = -INFINITY, z = 0.f;
  for (int f = 0; f < F; ++f) {
    const float* p = base + static_cast<long long>(f) * dim;
    float x[VPL];
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < VPL; ++t) {
      x[t] = p[lane + 32 * t];
      s = fmaf(w[t], mul ? x[t] * mean[t] : x[t], s);
    }
    const float e = warp_sum(s) + att_b;
    const float mn = fmaxf(m, e);
    const float corr = expf(m - mn);  // 0 on the first frame (m = -inf)
    const float pe = expf(e - mn);
    z = z * corr + pe;
#pragma unroll
    for (int t = 0; t < VPL; ++t) acc[t] = fmaf(acc[t], corr, pe * x[t]);
    m = mn;
  }
  const float invz = 1.0f / z;
  float ss = 0.f;
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    float g = acc[t] * invz;
    if (with_ave) g = fmaf(omega, mean[t] * static_cast<float>(F), g);
    acc[t] = g;
    ss = fmaf(g, g, ss);
  }
  ss = warp_sum(ss);
  const float den = sqrtf(ss) + norm_eps;
#pragma unroll
  for (int t = 0; t < VPL; ++t) out[warp * ld_out + lane + 32 * t] = acc[t] / den;
}

}  // namespace laff

using namespace laff;

static int grid_for(long long total, int block, int sms) {
  long long b = (total + block - 1) / block;
  const long long cap = static_cast<long long>(sms) * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

extern "C" {

int laff_l2norm_quantize(const float* x, long long rows, int heads, int head_dim, long long ldx, double eps,
                         int out_dtype, void* out, long long ld_out, void* stream) {
  LAFF_REQUIRE(x && out && rows > 0 && heads > 0 && head_dim > 0, LAFF_EINVAL, "laff_l2norm_quantize: bad arguments");
  LAFF_REQUIRE(out_dtype == LAFF_F16 || out_dtype == LAFF_BF16 || out_dtype == LAFF_F32, LAFF_EINVAL,
               "laff_l2norm_quantize: bad out_dtype %d", out_dtype);
  LAFF_REQUIRE(ldx >= static_cast<long long>(heads) * head_dim && ld_out >= static_cast<long long>(heads) * head_dim,
               LAFF_EINVAL, "laff_l2norm_quantize: pitch smaller than heads*head_dim");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const long long warps = rows * heads;
  const int block = 256;
  const long long blocks = (warps * 32 + block - 1) / block;
  LAFF_REQUIRE(blocks < (1LL << 31), LAFF_ENOTSUP, "laff_l2norm_quantize: too many rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float feps = static_cast<float>(eps < 0 ? 0.0 : eps);
  const int norm = eps >= 0 ? 1 : 0;
  const size_t esz = out_dtype == LAFF_F32 ? 4 : 2;
  const bool vec = head_dim % 128 == 0 && head_dim <= 1024 && (head_dim & (head_dim - 1)) == 0 && ldx % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                   (ld_out * esz) % 16 == 0;
  if (vec) {
    const unsigned nb = static_cast<unsigned>(blocks);
    switch (head_dim / 128) {
      case 1: l2norm_quantize_vec_kernel<1><<<nb, block, 0, st>>>(x, rows, heads, ldx, feps, norm, out_dtype, out, ld_out); break;
      case 2: l2norm_quantize_vec_kernel<2><<<nb, block, 0, st>>>(x, rows, heads, ldx, feps, norm, out_dtype, out, ld_out); break;
      case 4: l2norm_quantize_vec_kernel<4><<<nb, block, 0, st>>>(x, rows, heads, ldx, feps, norm, out_dtype, out, ld_out); break;
      default: l2norm_quantize_vec_kernel<8><<<nb, block, 0, st>>>(x, rows, heads, ldx, feps, norm, out_dtype, out, ld_out); break;
    }
  } else {
    l2norm_quantize_kernel<<<static_cast<unsigned>(blocks), block, 0, st>>>(x, rows, heads, head_dim, ldx, feps, norm, out_dtype, out,
                                                                           ld_out);
  }
  laff::count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

int laff_cast_pad_16(const float* x, long long rows, int cols, long long ldx, int out_dtype, void* out, int cols_pad,
                     long long ld_out, void* stream) {
  LAFF_REQUIRE(x && out && rows > 0 && cols > 0 && cols_pad >= cols && ld_out >= cols_pad && ldx >= cols, LAFF_EINVAL,
               "laff_cast_pad_16: bad arguments");
  LAFF_REQUIRE(is16(out_dtype), LAFF_EINVAL, "laff_cast_pad_16: out_dtype must be 16-bit");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  LAFF_REQUIRE(cols_pad % 8 == 0, LAFF_EINVAL, "laff_cast_pad_16: cols_pad (%d) must be a multiple of 8", cols_pad);
  const int vec_in = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && ldx % 4 == 0;
  const int vec_out = (reinterpret_cast<uintptr_t>(out) & 15) == 0 && ld_out % 8 == 0;
  cast_pad_kernel<<<grid_for(rows * (cols_pad / 8), 256, di.sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, rows, cols, ldx, out_dtype, static_cast<uint16_t*>(out), cols_pad, ld_out, vec_in, vec_out); laff::count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

int laff_split3_16(const float* x, long long rows, int cols, long long ldx, int side, int out_dtype, void* out,
                   int cols_pad, long long ld_out, void* stream) {
  LAFF_REQUIRE(x && out && rows > 0 && cols > 0 && cols_pad >= cols && ld_out >= 3LL * cols_pad && ldx >= cols,
               LAFF_EINVAL, "laff_split3_16: bad arguments");
  LAFF_REQUIRE(is16(out_dtype) && (side == 0 || side == 1), LAFF_EINVAL, "laff_split3_16: bad dtype/side");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  LAFF_REQUIRE(cols_pad % 8 == 0, LAFF_EINVAL, "laff_split3_16: cols_pad (%d) must be a multiple of 8", cols_pad);
  const int vec_in = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && ldx % 4 == 0;
  const int vec_out = (reinterpret_cast<uintptr_t>(out) & 15) == 0 && ld_out % 8 == 0;
  split3_kernel<<<grid_for(rows * (cols_pad / 8), 256, di.sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, rows, cols, ldx, side, out_dtype, static_cast<uint16_t*>(out), cols_pad, ld_out, vec_in, vec_out); laff::count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

int laff_bn_fold(const float* weight, const float* bias, const float* running_mean, const float* running_var,
                 double eps, int D, float* scale, float* shift, void* stream) {
  LAFF_REQUIRE(running_mean && running_var && scale && shift && D > 0, LAFF_EINVAL, "laff_bn_fold: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  bn_fold_kernel<<<(D + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(weight, bias, running_mean, running_var,
                                                                                static_cast<float>(eps), D, scale, shift); laff::count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

int laff_project(const void* x16, const void* w16, long long rows, int K, int D, long long ldx, long long ldw,
                 int dtype, const float* bias, int activation, const float* bn_scale, const float* bn_shift,
                 float* y, long long ldy, void* stream) {
  LAFF_REQUIRE(x16 && w16 && y, LAFF_EINVAL, "laff_project: null pointer");
  LAFF_REQUIRE(ldy >= D, LAFF_EINVAL, "laff_project: ldy %lld < D %d", ldy, D);
  LAFF_REQUIRE((bn_scale == nullptr) == (bn_shift == nullptr), LAFF_EINVAL, "laff_project: bn_scale/bn_shift mismatch");
  LAFF_REQUIRE(activation >= 0 && activation <= 3, LAFF_EINVAL, "laff_project: bad activation %d", activation);
  LAFF_REQUIRE(rows < (1LL << 31), LAFF_ENOTSUP, "laff_project: too many rows");
  LAFF_REQUIRE(is16(dtype), LAFF_EINVAL, "laff_project: dtype must be LAFF_F16/LAFF_BF16");
  LAFF_REQUIRE(rows > 0 && D > 0 && K > 0 && K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0 && ldx >= K && ldw >= K,
               LAFF_EINVAL, "laff_project: K and pitches must be multiples of 8 (K=%d ldx=%lld ldw=%lld)", K, ldx, ldw);
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const Tuning t = get_tuning();
  const int cg = t.cta_group;
  CUtensorMap tmA, tmB;
  rc = make_tmap_2d(&tmA, x16, dtype, static_cast<uint64_t>(rows), static_cast<uint64_t>(K), static_cast<uint64_t>(ldx), kBlockM);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, w16, dtype, static_cast<uint64_t>(D), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw),
                    static_cast<uint32_t>(kBlockN / cg));
  if (rc) return rc;
  // W (<= 4096 x K 16-bit) always stays in L2.  Many row tiles: one unit = one row tile sweeping all of W, so x is read
  // from HBM once and re-read from L2.  Few row tiles: one unit per output tile to fill the SMs.
  const int m_tiles = static_cast<int>((rows + kBlockM * cg - 1) / (kBlockM * cg));
  const bool many = m_tiles >= 2 * (di.sms / cg);
  const Sched s = make_sched(static_cast<int>(rows), D, cg, many ? (1 << 20) : 1, 1 << 20, 0);
  EpiProject::Params ep{y, ldy, rows, D, bias, bn_scale, bn_shift, activation};
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  const uint32_t idesc = make_idesc_f16(dtype, kBlockM * cg, kBlockN);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint64_t hintA = ptx::kEvictNormal, hintB = ptx::kEvictLast;
  if (cg == 2)
    LAFF_CUDA((launch_gemm_kernel<2, EpiProject>(tmA, tmB, num_kb, idesc, s, ep, hintA, hintB, di.sms, st)));
  else
    LAFF_CUDA((launch_gemm_kernel<1, EpiProject>(tmA, tmB, num_kb, idesc, s, ep, hintA, hintB, di.sms, st)));
  return LAFF_OK;
}

int laff_attention_pool(const laff_pool_desc* desc, long long rows, float* out, long long ld_out, void* out16,
                        int out16_dtype, long long ld_out16, float* att, void* stream) {
  LAFF_REQUIRE(desc && rows > 0, LAFF_EINVAL, "laff_attention_pool: bad arguments");
  LAFF_REQUIRE(out || out16, LAFF_EINVAL, "laff_attention_pool: no output buffer");
  LAFF_REQUIRE(desc->n_features >= 1 && desc->n_features <= LAFF_MAX_FEATURES, LAFF_ENOTSUP,
               "laff_attention_pool: n_features=%d outside [1, %d]", desc->n_features, LAFF_MAX_FEATURES);
  LAFF_REQUIRE(desc->heads > 0 && desc->att_weight && desc->att_bias, LAFF_EINVAL, "laff_attention_pool: bad desc");
  const int D = desc->heads * desc->head_dim;
  LAFF_REQUIRE(out == nullptr || ld_out >= D, LAFF_EINVAL, "laff_attention_pool: ld_out < D");
  LAFF_REQUIRE(out16 == nullptr || (ld_out16 >= D && is16(out16_dtype)), LAFF_EINVAL, "laff_attention_pool: bad out16");
  for (int l = 0; l < desc->n_features; ++l) {
    const laff_pool_source& s = desc->src[l];
    LAFF_REQUIRE(s.src != nullptr, LAFF_EINVAL, "laff_attention_pool: feature %d has no source", l);
    if (s.kind == 0) {
      LAFF_REQUIRE(s.ld >= D, LAFF_EINVAL, "laff_attention_pool: feature %d pitch < D", l);
    } else {
      LAFF_REQUIRE(s.kind == 1 && s.in_dim > 0 && D % s.in_dim == 0 && s.ld >= s.in_dim, LAFF_EINVAL,
                   "laff_attention_pool: tiled feature %d: in_dim %d must divide D %d", l, s.in_dim, D);
      LAFF_REQUIRE((s.bn_scale == nullptr) == (s.bn_shift == nullptr), LAFF_EINVAL, "feature %d: bn mismatch", l);
    }
  }
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const long long warps = rows * desc->heads;
  const int block = 128;
  const long long blocks = (warps * 32 + block - 1) / block;
  LAFF_REQUIRE(blocks < (1LL << 31), LAFF_ENOTSUP, "laff_attention_pool: too many rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dh = desc->head_dim;
  LAFF_REQUIRE(dh % 32 == 0, LAFF_ENOTSUP, "laff_attention_pool: head_dim %d must be a multiple of 32", dh);
  // 128-bit path: head_dim % 128 == 0 and every pointer / pitch 16-byte aligned
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  bool vec = dh % 128 == 0 && al16(desc->att_weight) && (!out || (al16(out) && ld_out % 4 == 0)) &&
             (!out16 || ((reinterpret_cast<uintptr_t>(out16) & 7) == 0 && ld_out16 % 4 == 0));
  for (int l = 0; l < desc->n_features && vec; ++l) {
    const laff_pool_source& s = desc->src[l];
    vec = al16(s.src) && s.ld % 4 == 0;
    if (s.kind == 1) vec = vec && s.in_dim % 4 == 0 && (!s.bn_scale || (al16(s.bn_scale) && al16(s.bn_shift)));
  }
  const int L = desc->n_features;
  const bool plain = !desc->with_ave && !desc->mul;
#define LAFF_POOL_LAUNCH(V, LM, VE)                                                                                    \
  do {                                                                                                                 \
    if (plain) attention_pool_kernel<V, LM, VE, true><<<static_cast<unsigned>(blocks), block, 0, st>>>(                 \
        *desc, rows, out, ld_out, out16, out16_dtype, ld_out16, att);                                                  \
    else attention_pool_kernel<V, LM, VE, false><<<static_cast<unsigned>(blocks), block, 0, st>>>(                      \
        *desc, rows, out, ld_out, out16, out16_dtype, ld_out16, att);                                                  \
  } while (0)
#define LAFF_POOL_CASE(V)                                                                                              \
  case V:                                                                                                              \
    if (vec && V >= 4) {                                                                                               \
      if (L <= 4) LAFF_POOL_LAUNCH((V >= 4 ? V : 4), 4, true);                                                         \
      else LAFF_POOL_LAUNCH((V >= 4 ? V : 4), 8, true);                                                                \
    } else {                                                                                                           \
      LAFF_POOL_LAUNCH(V, 8, false);                                                                                   \
    }                                                                                                                  \
    break;
  switch (dh / 32) {
    LAFF_POOL_CASE(1)
    LAFF_POOL_CASE(2)
    LAFF_POOL_CASE(4)
    LAFF_POOL_CASE(8)
    LAFF_POOL_CASE(16)
    default:
      LAFF_REQUIRE(false, LAFF_ENOTSUP, "laff_attention_pool: head_dim %d not in {32,64,128,256,512}", dh);
  }
#undef LAFF_POOL_LAUNCH
#undef LAFF_POOL_CASE
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

int laff_frame_pool(const float* frames, long long B, int F, int dim, const float* att_weight, float att_bias,
                    int with_ave, int mul, float omega, double norm_eps, float* out, long long ld_out, void* stream) {
  LAFF_REQUIRE(frames && att_weight && out && B > 0 && F > 0 && dim > 0 && ld_out >= dim, LAFF_EINVAL,
               "laff_frame_pool: bad arguments");
  LAFF_REQUIRE(dim % 32 == 0, LAFF_ENOTSUP, "laff_frame_pool: dim %d must be a multiple of 32", dim);
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const int block = 128;
  const long long blocks = (B * 32 + block - 1) / block;
  LAFF_REQUIRE(blocks < (1LL << 31), LAFF_ENOTSUP, "laff_frame_pool: too many videos");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LAFF_FRAME_CASE(V)                                                                                           \
  case V:                                                                                                            \
    frame_pool_kernel<V><<<static_cast<unsigned>(blocks), block, 0, st>>>(frames, B, F, dim, att_weight, att_bias,    \
                                                                          with_ave, mul, omega,                      \
                                                                          static_cast<float>(norm_eps), out, ld_out); laff::count_launch(); \
    break;
  switch (dim / 32) {
    LAFF_FRAME_CASE(1)
    LAFF_FRAME_CASE(2)
    LAFF_FRAME_CASE(4)
    LAFF_FRAME_CASE(8)
    LAFF_FRAME_CASE(16)
    LAFF_FRAME_CASE(32)
    default:
      LAFF_REQUIRE(false, LAFF_ENOTSUP, "laff_frame_pool: dim %d not in {32,...,1024} powers of two", dim);
  }
#undef LAFF_FRAME_CASE
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

}  // extern "C"
