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

__device__ __forceinline__ uint16_t to16(float v, int dtype) {
  if (dtype == LAFF_BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  return __half_as_ushort(__float2half_rn(v));
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

__global__ void cast_pad_kernel(const float* __restrict__ x, long long rows, int cols, long long ldx, int out_dtype,
                                uint16_t* __restrict__ out, int cols_pad, long long ld_out) {
  const long long total = rows * cols_pad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols_pad;
    const int c = static_cast<int>(i - r * cols_pad);
    const float v = c < cols ? x[r * ldx + c] : 0.f;
    out[r * ld_out + c] = to16(v, out_dtype);
  }
}

__global__ void split3_kernel(const float* __restrict__ x, long long rows, int cols, long long ldx, int side,
                              int out_dtype, uint16_t* __restrict__ out, int cols_pad, long long ld_out) {
  const long long total = rows * cols_pad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols_pad;
    const int c = static_cast<int>(i - r * cols_pad);
    const float v = c < cols ? x[r * ldx + c] : 0.f;
    const uint16_t hi = to16(v, out_dtype);
    const uint16_t lo = to16(v - from16(hi, out_dtype), out_dtype);
    uint16_t* o = out + r * ld_out + c;
    // left: [hi | lo | hi]   right: [hi | hi | lo]   =>  left . right = hi.hi + lo.hi + hi.lo
    o[0] = hi;
    o[cols_pad] = side == 0 ? lo : hi;
    o[2 * static_cast<long long>(cols_pad)] = side == 0 ? hi : lo;
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
__device__ __forceinline__ float activate(float z, int act) {
  switch (act) {
    case LAFF_ACT_TANH: return tanhf(z);
    case LAFF_ACT_RELU: return fmaxf(z, 0.f);
    case LAFF_ACT_SIGMOID: return 1.0f / (1.0f + expf(-z));
    default: return z;
  }
}

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
  static constexpr int kSmemBytes = 0;
  Params p;
  __device__ EpiProject(const Params& p_, uint8_t*, int) : p(p_) {}
  __device__ __forceinline__ void unit_begin(int, const Unit&) {}
  __device__ __forceinline__ void unit_end(int, const Unit&) {}
  __device__ __forceinline__ void chunk(const uint32_t (&r)[32], int row, int col0) {
    if (row >= p.M || col0 >= p.N) return;
    float* dst = p.y + static_cast<long long>(row) * p.ldy + col0;
    const bool full = (col0 + 32 <= p.N) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int c = min(col0 + j, p.N - 1);
      float z = __uint_as_float(r[j]);
      if (p.bias) z += __ldg(p.bias + c);
      z = activate(z, p.act);
      if (p.bn_scale) z = fmaf(z, __ldg(p.bn_scale + c), __ldg(p.bn_shift + c));
      o[j] = z;
    }
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) dst[j] = o[j];
    }
  }
};

// ------------------------------------------------------------------------------------------------------------
// F5/F6: LAFF block.  One warp per (row, head); lane owns elements d = lane + 32*t of the head.
// ------------------------------------------------------------------------------------------------------------
template <int VPL>  // values per lane = head_dim / 32
__global__ void __launch_bounds__(128) attention_pool_kernel(laff_pool_desc d, long long rows, float* __restrict__ out,
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

  float y[LAFF_MAX_FEATURES][VPL];
  float w[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) w[t] = __ldg(d.att_weight + static_cast<long long>(h) * dh + lane + 32 * t);

#pragma unroll
  for (int l = 0; l < LAFF_MAX_FEATURES; ++l) {
    if (l < L) {
      const laff_pool_source& s = d.src[l];
      if (s.kind == 0) {
        const float* p = s.src + row * s.ld + static_cast<long long>(h) * dh;
#pragma unroll
        for (int t = 0; t < VPL; ++t) y[l][t] = p[lane + 32 * t];
      } else {
        // "no-transform": x tiled along D (x.repeat(1, heads)), then BatchNorm1d(D)   model/model.py:1822-1823
        const float* p = s.src + row * s.ld;
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
          const int col = h * dh + lane + 32 * t;
          float v = p[col % s.in_dim];
          if (s.bn_scale) v = fmaf(v, __ldg(s.bn_scale + col), __ldg(s.bn_shift + col));
          y[l][t] = v;
        }
      }
    } else {
#pragma unroll
      for (int t = 0; t < VPL; ++t) y[l][t] = 0.f;
    }
  }

  // raw_global_emb = mean over features (Attention.py:81)
  float mean[VPL];
  const float invL = 1.0f / static_cast<float>(L);
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < LAFF_MAX_FEATURES; ++l)
      if (l < L) s += y[l][t];
    mean[t] = s * invL;
  }

  // logits e_l = w_h . common_l + c_h  (Attention.py:88), common = local (* mean if mul, Attention.py:83-86)
  float e[LAFF_MAX_FEATURES];
  const float cb = __ldg(d.att_bias + h);
  float emax = -INFINITY;
#pragma unroll
  for (int l = 0; l < LAFF_MAX_FEATURES; ++l) {
    if (l < L) {
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < VPL; ++t) s = fmaf(w[t], d.mul ? y[l][t] * mean[t] : y[l][t], s);
      e[l] = warp_sum(s) + cb;
      emax = fmaxf(emax, e[l]);
    } else {
      e[l] = -INFINITY;
    }
  }
  // softmax over features (Attention.py:89)
  float z = 0.f;
#pragma unroll
  for (int l = 0; l < LAFF_MAX_FEATURES; ++l) {
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
    for (int l = 0; l < LAFF_MAX_FEATURES; ++l)
      if (l < L) s = fmaf(e[l] * invz, y[l][t], s);
    if (d.with_ave) s = fmaf(d.omega, mean[t] * static_cast<float>(L), s);  // sum_l omega * raw_global_emb
    g[t] = s;
    ss = fmaf(s, s, ss);
  }
  ss = warp_sum(ss);
  const float den = sqrtf(ss) + static_cast<float>(d.norm_eps);  // l2norm(eps=0): + 0 + 1e-14  (Attention.py:103)
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    const float v = g[t] / den;
    const long long c = static_cast<long long>(h) * dh + lane + 32 * t;
    if (out) out[row * ld_out + c] = v;
    if (out16) static_cast<uint16_t*>(out16)[row * ld_out16 + c] = to16(v, out16_dtype);
  }
  if (att && lane == 0) {
#pragma unroll
    for (int l = 0; l < LAFF_MAX_FEATURES; ++l)
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
  l2norm_quantize_kernel<<<static_cast<unsigned>(blocks), block, 0, static_cast<cudaStream_t>(stream)>>>(
      x, rows, heads, head_dim, ldx, static_cast<float>(eps < 0 ? 0.0 : eps), eps >= 0 ? 1 : 0, out_dtype, out, ld_out); laff::count_launch();
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
  cast_pad_kernel<<<grid_for(rows * cols_pad, 256, di.sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, rows, cols, ldx, out_dtype, static_cast<uint16_t*>(out), cols_pad, ld_out); laff::count_launch();
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
  split3_kernel<<<grid_for(rows * cols_pad, 256, di.sms), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, rows, cols, ldx, side, out_dtype, static_cast<uint16_t*>(out), cols_pad, ld_out); laff::count_launch();
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
#define LAFF_POOL_CASE(V)                                                                                              \
  case V:                                                                                                              \
    attention_pool_kernel<V><<<static_cast<unsigned>(blocks), block, 0, st>>>(*desc, rows, out, ld_out, out16,          \
                                                                              out16_dtype, ld_out16, att); laff::count_launch();             \
    break;
  LAFF_REQUIRE(dh % 32 == 0, LAFF_ENOTSUP, "laff_attention_pool: head_dim %d must be a multiple of 32", dh);
  switch (dh / 32) {
    LAFF_POOL_CASE(1)
    LAFF_POOL_CASE(2)
    LAFF_POOL_CASE(4)
    LAFF_POOL_CASE(8)
    LAFF_POOL_CASE(16)
    default:
      LAFF_REQUIRE(false, LAFF_ENOTSUP, "laff_attention_pool: head_dim %d not in {32,64,128,256,512}", dh);
  }
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
