// Training step of the LAFF fusion nets: the train-mode stages around the projection GEMMs
// and their backward, the backward of the LAFF pooling block, and the optimizer step.
//   reference: TransformNet.forward in train mode (model/model.py:257-276: FC -> activation -> dropout -> BatchNorm1d with
//   batch statistics), Attention_1 / Multi_head_MyApply_Attention (model/Attention.py:78-105, :508-531) under autograd,
//   W2VVPP.forward (model/model.py:964-1001: backward, clip_grad_norm_(params, 2), RMSprop / Adam step).
// At the reference's batch size (128) every stage is latency-bound: the kernels below are column- or row-parallel
// CUDA-core code; the matrix products (x W^T forward, dZ^T x for the weight gradients) go through the tcgen05 engine.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstdint>

#include "host_util.cuh"

namespace laff {

// Counter-based uniform in [0, 1): a 64-bit mix of (seed, element index).  Same value in forward and backward.
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float act_grad_from_output(float a, int act) {
  switch (act) {
    case LAFF_ACT_TANH: return 1.0f - a * a;
    case LAFF_ACT_SIGMOID: return a * (1.0f - a);
    case LAFF_ACT_RELU: return a > 0.f ? 1.0f : 0.f;
    default: return 1.0f;
  }
}

// Column statistics need every row of a column; rows are the short dimension (B = 128 in the shipped configs), so a
// block owns 32 columns and its 32 x kRowLanes threads stride over the rows: coalesced 128-byte row segments, then a
// shared-memory tree over the row lanes.  Sums are kept in double: the batch statistics feed rsqrt and differences.
constexpr int kRowLanes = 16;

__device__ __forceinline__ double col_reduce(double v, double (*sh)[33]) {
  __syncthreads();  // previous use of sh is over
  sh[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  for (int o = kRowLanes / 2; o > 0; o >>= 1) {
    if (threadIdx.y < o) sh[threadIdx.y][threadIdx.x] += sh[threadIdx.y + o][threadIdx.x];
    __syncthreads();
  }
  return sh[0][threadIdx.x];
}

// src is either the activated projection a [B, D] (src_cols == D) or a raw "no-transform" feature [B, in_dim] tiled over
// the heads (src_cols == in_dim, model/model.py:1822-1823).
__global__ void __launch_bounds__(32 * kRowLanes)
    transform_train_fwd_kernel(const float* __restrict__ src, long long ld_src, int src_cols, int B, int D, float p_drop,
                               unsigned long long seed_base, const unsigned long long* __restrict__ seed_dev,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps,
                               int use_bn, float* __restrict__ y, long long ld_y, uint8_t* __restrict__ mask,
                               float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  __shared__ double sh[kRowLanes][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < D;
  const int sc = ok ? c % src_cols : 0;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  // seed_dev: a step counter kept on the device, so that a captured CUDA graph draws a fresh mask on every replay
  const unsigned long long seed = seed_base + (seed_dev ? *seed_dev * 0x632BE59BD9B4E019ull : 0ull);
  double sum = 0.0;
  if (ok) {
    for (int r = threadIdx.y; r < B; r += kRowLanes) {
      float v = src[r * ld_src + sc];
      if (p_drop > 0.f) {
        const bool keep = uniform01(seed, static_cast<unsigned long long>(r) * D + c) >= p_drop;
        mask[static_cast<long long>(r) * D + c] = keep ? 1 : 0;
        v = keep ? v * keep_scale : 0.f;
      }
      y[r * ld_y + c] = v;
      sum += v;
    }
  }
  if (!use_bn) return;
  const double mean = col_reduce(sum, sh) / B;
  double ss = 0.0;
  if (ok) {
    for (int r = threadIdx.y; r < B; r += kRowLanes) {
      const double d = static_cast<double>(y[r * ld_y + c]) - mean;
      ss += d * d;
    }
  }
  ss = col_reduce(ss, sh);
  if (!ok) return;
  const double var = ss / B;  // biased: what normalises the batch
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
  const float meanf = static_cast<float>(mean);
  for (int r = threadIdx.y; r < B; r += kRowLanes) y[r * ld_y + c] = (y[r * ld_y + c] - meanf) * invstd * g + b;
  if (threadIdx.y == 0) {
    save_mean[c] = meanf;
    save_invstd[c] = invstd;
    if (running_mean) {
      const double unbiased = B > 1 ? ss / (B - 1) : var;
      running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * meanf;
      running_var[c] = (1.0f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
    }
  }
}

// Backward of the stage above: dy -> dz (gradient at the GEMM output, before the activation), d gamma, d beta, d bias
// (= column sum of dz).  `a` is the activated projection saved by the forward (NULL for a tiled feature, whose input is a
// leaf: only the BatchNorm parameters get gradients).
__global__ void __launch_bounds__(32 * kRowLanes)
    transform_train_bwd_kernel(const float* __restrict__ dy, long long ld_dy, const float* __restrict__ a, long long ld_a,
                               const float* __restrict__ tiled_x, long long ld_x, int in_dim, const uint8_t* __restrict__ mask,
                               float p_drop, int act, int use_bn, const float* __restrict__ gamma,
                               const float* __restrict__ save_mean, const float* __restrict__ save_invstd, int B, int D,
                               float* __restrict__ dz, long long ld_dz, float* __restrict__ dgamma, float* __restrict__ dbeta,
                               float* __restrict__ dbias) {
  __shared__ double sh[kRowLanes][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < D;
  const float keep_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;
  auto dropped = [&](int r) -> float {  // value that entered BatchNorm
    float v = a ? a[r * ld_a + c] : tiled_x[r * ld_x + (c % in_dim)];
    if (p_drop > 0.f) v = mask[static_cast<long long>(r) * D + c] ? v * keep_scale : 0.f;
    return v;
  };
  float g = 1.f, mean = 0.f, invstd = 1.f;
  double sb = 0.0, sg = 0.0;
  if (use_bn) {
    if (ok) {
      g = gamma ? gamma[c] : 1.0f;
      mean = save_mean[c];
      invstd = save_invstd[c];
      for (int r = threadIdx.y; r < B; r += kRowLanes) {
        const float d = dy[r * ld_dy + c];
        sb += d;
        sg += static_cast<double>(d) * ((dropped(r) - mean) * invstd);
      }
    }
    sb = col_reduce(sb, sh);
    sg = col_reduce(sg, sh);
    if (ok && threadIdx.y == 0) {
      if (dgamma) dgamma[c] = static_cast<float>(sg);
      if (dbeta) dbeta[c] = static_cast<float>(sb);
    }
  }
  if (!dz) return;
  const float mb = static_cast<float>(sb / B), mg = static_cast<float>(sg / B);
  double sbias = 0.0;
  if (ok) {
    for (int r = threadIdx.y; r < B; r += kRowLanes) {
      float d = dy[r * ld_dy + c];
      if (use_bn) d = g * invstd * (d - mb - ((dropped(r) - mean) * invstd) * mg);
      if (p_drop > 0.f) d = mask[static_cast<long long>(r) * D + c] ? d * keep_scale : 0.f;
      if (a) d *= act_grad_from_output(a[r * ld_a + c], act);
      dz[r * ld_dz + c] = d;
      sbias += d;
    }
  }
  sbias = col_reduce(sbias, sh);
  if (ok && threadIdx.y == 0 && dbias) dbias[c] = static_cast<float>(sbias);
}

// Backward of the LAFF block (Attention_1, model/Attention.py:78-105):
//   r = mean_l y_l,  common_l = mul ? y_l * r : y_l,  e_l = w_h . common_l + c_h,  p = softmax_l(e),
//   g = sum_l (p_l + omega') y_l  (omega' = with_ave ? omega : 0; omega is read with .item() by the reference, so it
//   gets no gradient),  out = g / (|g| + eps).   The shipped setting is with_ave = mul = 0.
// One warp per (row, head); lanes own d_h / 32 columns.  dW / dc are written per (row, head) and reduced afterwards
// (deterministic).
struct PoolBwdArgs {  // passed by value: nothing to upload, and a captured graph keeps its own copy
  const float* ys[LAFF_MAX_FEATURES];
  float* dys[LAFF_MAX_FEATURES];
  long long lds[LAFF_MAX_FEATURES];
};

template <int VPL, int LMAX>
__global__ void __launch_bounds__(128) pool_train_bwd_kernel(const __grid_constant__ PoolBwdArgs args, int L,
                                                             const float* __restrict__ att_w, const float* __restrict__ att_b,
                                                             const float* __restrict__ dout, long long ld_dout, long long rows, int heads,
                                                             float norm_eps, int with_ave, int mul, float omega,
                                                             float* __restrict__ dw_part, float* __restrict__ dc_part) {
  const float* const* ys = args.ys;
  float* const* dys = args.dys;
  const long long* lds = args.lds;
  const int lane = threadIdx.x & 31;
  const long long wid = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= rows * heads) return;
  const long long row = wid / heads;
  const int h = static_cast<int>(wid - row * heads);
  const int dh = VPL * 32;
  const long long col0 = static_cast<long long>(h) * dh;
  float w[VPL], y[LMAX][VPL], e[LMAX];
#pragma unroll
  for (int v = 0; v < VPL; ++v) w[v] = att_w[col0 + v * 32 + lane];
  auto wsum = [&](float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
  };
  float emax = -INFINITY;
  float rmean[VPL];  // r = mean over the features (only used by the product variant)
#pragma unroll
  for (int v = 0; v < VPL; ++v) rmean[v] = 0.f;
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    if (l < L) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        y[l][v] = ys[l][row * lds[l] + col0 + v * 32 + lane];
        rmean[v] += y[l][v];
      }
    }
  }
  const float invL = 1.0f / static_cast<float>(L);
#pragma unroll
  for (int v = 0; v < VPL; ++v) rmean[v] *= invL;
  const float wave = with_ave ? omega : 0.f;
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    if (l < L) {
      float part = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) part = fmaf(w[v], mul ? y[l][v] * rmean[v] : y[l][v], part);
      e[l] = wsum(part) + att_b[h];
      emax = fmaxf(emax, e[l]);
    }
  }
  float den = 0.f;
#pragma unroll
  for (int l = 0; l < LMAX; ++l)
    if (l < L) {
      e[l] = expf(e[l] - emax);
      den += e[l];
    }
  float g[VPL], dg[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) g[v] = 0.f;
#pragma unroll
  for (int l = 0; l < LMAX; ++l)
    if (l < L) {
      e[l] /= den;  // p_l
#pragma unroll
      for (int v = 0; v < VPL; ++v) g[v] = fmaf(e[l] + wave, y[l][v], g[v]);
    }
  float ss = 0.f, dot = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    dg[v] = dout[row * ld_dout + col0 + v * 32 + lane];
    ss = fmaf(g[v], g[v], ss);
  }
  const float nrm = sqrtf(wsum(ss));
  const float inv = 1.0f / (nrm + norm_eps);
#pragma unroll
  for (int v = 0; v < VPL; ++v) dot = fmaf(g[v] * inv, dg[v], dot);  // out . dout
  dot = wsum(dot);
  // d g = (dout - out * (out . dout) * |g| / (|g| + eps)) / (|g| + eps)
  const float shrink = nrm > 0.f ? dot * nrm * inv : 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) dg[v] = (dg[v] - (g[v] * inv) * shrink) * inv;
  float dp[LMAX], mix = 0.f;
#pragma unroll
  for (int l = 0; l < LMAX; ++l)
    if (l < L) {
      float part = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) part = fmaf(dg[v], y[l][v], part);
      dp[l] = wsum(part);
      mix = fmaf(e[l], dp[l], mix);
    }
  float dwv[VPL], dcv = 0.f, dmean[VPL];  // dmean = (1/L) sum_m de_m * w * y_m: gradient reaching every y_l through r
#pragma unroll
  for (int v = 0; v < VPL; ++v) dwv[v] = dmean[v] = 0.f;
  if (mul) {
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
      if (l < L) {
        const float de = e[l] * (dp[l] - mix);
#pragma unroll
        for (int v = 0; v < VPL; ++v) dmean[v] = fmaf(de * w[v], y[l][v], dmean[v]);
      }
#pragma unroll
    for (int v = 0; v < VPL; ++v) dmean[v] *= invL;
  }
#pragma unroll
  for (int l = 0; l < LMAX; ++l)
    if (l < L) {
      const float de = e[l] * (dp[l] - mix);
      dcv += de;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const float common = mul ? y[l][v] * rmean[v] : y[l][v];
        const float through_e = mul ? de * w[v] * rmean[v] + dmean[v] : de * w[v];
        dys[l][row * lds[l] + col0 + v * 32 + lane] = fmaf(e[l] + wave, dg[v], through_e);
        dwv[v] = fmaf(de, common, dwv[v]);
      }
    }
#pragma unroll
  for (int v = 0; v < VPL; ++v) dw_part[wid * dh + v * 32 + lane] = dwv[v];
  if (lane == 0) dc_part[wid] = dcv;
}

// Gradient w.r.t. a tiled ("no-transform") feature: dx[b, j] = sum_h dz[b, h * in_dim + j]  (x.repeat(1, heads) backward).
__global__ void fold_tiles_kernel(const float* __restrict__ dz, long long ld_dz, int B, int D, int in_dim, float* __restrict__ dx,
                                  long long ld_dx) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * in_dim) return;
  const int b = static_cast<int>(i / in_dim), j = static_cast<int>(i - static_cast<long long>(b) * in_dim);
  float s = 0.f;
  for (int c = j; c < D; c += in_dim) s += dz[b * ld_dz + c];
  dx[b * ld_dx + j] = s;
}

// Backward of the frame-level LAFF block (Attention_1(dim) over the F frames of a video, with_ave = mul = False;
// model/model.py:2167-2173): gradients of the logit weight / bias only (frame features are leaves).  One warp per
// video; softmax weights and d p_f are kept in shared memory (F <= kMaxFrames).
constexpr int kMaxFrames = 128;
template <int VPL>
__global__ void __launch_bounds__(128) frame_pool_bwd_kernel(const float* __restrict__ frames, long long B, int F, int dim,
                                                            const float* __restrict__ att_w, const float* __restrict__ dout,
                                                            long long ld_dout, float norm_eps, float* __restrict__ dw_part,
                                                            float* __restrict__ dc_part) {
  __shared__ float s_p[4][kMaxFrames], s_dp[4][kMaxFrames];
  const long long vid = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  if (vid >= B) return;
  const float* base = frames + vid * static_cast<long long>(F) * dim;
  auto wsum = [&](float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
  };
  float w[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) w[t] = att_w[lane + 32 * t];
  float m = -INFINITY;
  for (int f = 0; f < F; ++f) {
    float sdot = 0.f;
#pragma unroll
    for (int t = 0; t < VPL; ++t) sdot = fmaf(w[t], base[static_cast<long long>(f) * dim + lane + 32 * t], sdot);
    const float e = wsum(sdot);  // the bias shifts every logit alike: it cancels in the softmax
    if (lane == 0) s_p[wl][f] = e;
    m = fmaxf(m, e);
  }
  __syncwarp();
  float z = 0.f;
  for (int f = 0; f < F; ++f) z += expf(s_p[wl][f] - m);
  float g[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) g[t] = 0.f;
  for (int f = 0; f < F; ++f) {
    const float pf = expf(s_p[wl][f] - m) / z;
    __syncwarp();
    if (lane == 0) s_p[wl][f] = pf;
#pragma unroll
    for (int t = 0; t < VPL; ++t) g[t] = fmaf(pf, base[static_cast<long long>(f) * dim + lane + 32 * t], g[t]);
  }
  __syncwarp();
  float ss = 0.f, dot = 0.f, dg[VPL];
#pragma unroll
  for (int t = 0; t < VPL; ++t) {
    dg[t] = dout[vid * ld_dout + lane + 32 * t];
    ss = fmaf(g[t], g[t], ss);
  }
  const float nrm = sqrtf(wsum(ss));
  const float inv = 1.0f / (nrm + norm_eps);
#pragma unroll
  for (int t = 0; t < VPL; ++t) dot = fmaf(g[t] * inv, dg[t], dot);
  dot = wsum(dot);
  const float shrink = nrm > 0.f ? dot * nrm * inv : 0.f;
#pragma unroll
  for (int t = 0; t < VPL; ++t) dg[t] = (dg[t] - (g[t] * inv) * shrink) * inv;
  float mix = 0.f;
  for (int f = 0; f < F; ++f) {
    float part = 0.f;
#pragma unroll
    for (int t = 0; t < VPL; ++t) part = fmaf(dg[t], base[static_cast<long long>(f) * dim + lane + 32 * t], part);
    part = wsum(part);
    if (lane == 0) s_dp[wl][f] = part;
    mix = fmaf(s_p[wl][f], part, mix);
  }
  __syncwarp();
  float dwv[VPL], dcv = 0.f;
#pragma unroll
  for (int t = 0; t < VPL; ++t) dwv[t] = 0.f;
  for (int f = 0; f < F; ++f) {
    const float de = s_p[wl][f] * (s_dp[wl][f] - mix);
    dcv += de;
#pragma unroll
    for (int t = 0; t < VPL; ++t) dwv[t] = fmaf(de, base[static_cast<long long>(f) * dim + lane + 32 * t], dwv[t]);
  }
#pragma unroll
  for (int t = 0; t < VPL; ++t) dw_part[vid * dim + lane + 32 * t] = dwv[t];
  if (lane == 0) dc_part[vid] = dcv;
}

// out[c] = sum_r part[r, c]   (rows x cols, deterministic order)
__global__ void colsum_kernel(const float* __restrict__ part, long long rows, int cols, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double s = 0.0;
  for (long long r = 0; r < rows; ++r) s += part[r * cols + c];
  out[c] = static_cast<float>(s);
}

// fp32 [R, C] -> 16-bit [C, Kpad * terms] with K = R (zero padded to a multiple of 8): the K-major operands of the
// weight-gradient GEMM dW = dZ^T x.  terms == 1: plain rounding; terms == 3: the 3-term split ([hi|lo|hi] for side 0,
// [hi|hi|lo] for side 1) whose product hi*hi + lo*hi + hi*lo is fp32-grade.
__global__ void transpose16_kernel(const float* __restrict__ x, long long ld, int R, int C, int dtype, int terms, int side,
                                   uint16_t* __restrict__ out, long long ld_out, int Kpad) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? x[r * ld + c] : 0.f;
  }
  __syncthreads();
  auto to16 = [&](float v) -> uint16_t {
    return dtype == LAFF_BF16 ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
  };
  auto from16 = [&](uint16_t u) -> float {
    return dtype == LAFF_BF16 ? __bfloat162float(__ushort_as_bfloat16(u)) : __half2float(__ushort_as_half(u));
  };
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c >= C || r >= Kpad) continue;
    const float v = tile[threadIdx.x][i];
    uint16_t* o = out + static_cast<long long>(c) * ld_out + r;
    const uint16_t hi = to16(v);
    if (terms == 1) {
      o[0] = hi;
    } else {
      const uint16_t lo = to16(v - from16(hi));
      o[0] = hi;
      o[Kpad] = side == 0 ? lo : hi;
      o[2 * Kpad] = side == 0 ? hi : lo;
    }
  }
}

// ---- optimizer -------------------------------------------------------------------------------------------------------
constexpr int kOptChunk = 256 * 8;  // elements one block of the optimizer kernels walks

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// partial[blk] = sum of squares of the block's gradient chunk; partial_max[blk] (optional, the loss-scaler variant) = the
// largest |g| of the chunk (NaN-propagating: a non-finite gradient makes it non-finite).
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const laff_opt_tensor* __restrict__ ts, const int* __restrict__ blk_tensor,
                                                          const long long* __restrict__ blk_start, double* __restrict__ partial,
                                                          float* __restrict__ partial_max) {
  const laff_opt_tensor t = ts[blk_tensor[blockIdx.x]];
  const long long s = blk_start[blockIdx.x];
  float part = 0.f;  // 8 squares per thread in fp32, everything above that in fp64
  float mx = 0.f;
  auto amax = [](float m, float v) { const float a = fabsf(v); return (a > m || a != a) ? a : m; };
  if (t.grad) {
    const long long j0 = s + threadIdx.x * 8;
    if (aligned16(t.grad) && j0 + 8 <= t.n) {
      const float4 a = *reinterpret_cast<const float4*>(t.grad + j0), b = *reinterpret_cast<const float4*>(t.grad + j0 + 4);
      part = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
      if (partial_max) mx = amax(amax(amax(amax(amax(amax(amax(amax(0.f, a.x), a.y), a.z), a.w), b.x), b.y), b.z), b.w);
    } else {
      for (long long j = j0; j < j0 + 8 && j < t.n; ++j) {
        part = fmaf(t.grad[j], t.grad[j], part);
        mx = amax(mx, t.grad[j]);
      }
    }
  }
  const double acc = part;
  __shared__ double sh[256];
  __shared__ float shm[256];
  sh[threadIdx.x] = acc;
  shm[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[threadIdx.x] += sh[threadIdx.x + o];
      shm[threadIdx.x] = (shm[threadIdx.x] != shm[threadIdx.x]) ? shm[threadIdx.x] : amax(shm[threadIdx.x], shm[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = sh[0];
    if (partial_max) partial_max[blockIdx.x] = shm[0];
  }
}

__global__ void sqnorm_final_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sqrt(sh[0]);
}

// Loss-scaler variant (torch.cuda.amp.GradScaler as the reference drives it, model/model.py:970-989): the gradients the
// reference clips are the SCALED ones, so clip_grad_norm_ sees S * ||g||; scaler.step() then unscales and skips the
// optimizer step when a gradient is non-finite; scaler.update() halves S after a skipped step and doubles it after
// `growth_interval` good ones.  Gradients here are fp32 and unscaled, so (a) the clip coefficient is
// min(1, max_norm / (S ||g|| + 1e-6)) on the raw gradients (scale and unscale cancel), and (b) the fp16 overflow that
// makes the reference skip is emulated on the parameter gradients, which all pass through fp16 in the reference's
// autocast backward: a step is skipped when S * max|g| would round to inf in fp16 (>= overflow_limit = 65520) or a
// gradient is non-finite.  One thread; writes ctl = {coef, skip flag}, the scaled norm, and advances S / the growth
// tracker / the optimizer step count for the step kernel that follows.
__global__ void scaler_final_kernel(const double* __restrict__ partial, const float* __restrict__ partial_max, int n, float max_norm,
                                    laff_scaler_state* __restrict__ sc, float growth, float backoff, int interval,
                                    float overflow_limit, double* __restrict__ norm_out, float* __restrict__ ctl,
                                    long long* __restrict__ step_dev) {
  __shared__ double sh[256];
  __shared__ float shm[256];
  double acc = 0.0;
  float mx = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    acc += partial[i];
    const float v = partial_max[i];
    if (v != v || v > mx) mx = (mx != mx) ? mx : v;
  }
  sh[threadIdx.x] = acc;
  shm[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[threadIdx.x] += sh[threadIdx.x + o];
      const float a = shm[threadIdx.x], b = shm[threadIdx.x + o];
      shm[threadIdx.x] = (a != a) ? a : ((b != b || b > a) ? b : a);
    }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  const double S = static_cast<double>(sc->scale);
  const double norm = sqrt(sh[0]) * S;   // what clip_grad_norm_ measures on the scaled gradients
  const double gmax = static_cast<double>(shm[0]) * S;
  const bool bad = !(norm == norm) || isinf(norm) || !(gmax == gmax) || gmax >= static_cast<double>(overflow_limit);
  float coef = 1.0f;
  if (max_norm > 0.f && !bad) {
    const float c = static_cast<float>(static_cast<double>(max_norm) / (norm + 1e-6));
    coef = c < 1.0f ? c : 1.0f;
  }
  ctl[0] = coef;
  ctl[1] = bad ? 1.0f : 0.0f;
  *norm_out = norm;
  sc->found_inf = bad ? 1 : 0;
  if (bad) {
    sc->scale = static_cast<float>(S * backoff);
    sc->growth_tracker = 0;
    sc->skipped += 1;
  } else {
    if (step_dev) *step_dev += 1;   // optimizer.step() only runs (and Adam's step only advances) on a good step
    if (++sc->growth_tracker >= interval) {
      sc->scale = static_cast<float>(S * growth);
      sc->growth_tracker = 0;
    }
  }
}

// kind 0: RMSprop (torch defaults: no momentum, not centered), kind 1: Adam.  The clip coefficient is derived on the
// device from the total gradient norm (clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1), so the step
// needs no host round trip.
__global__ void __launch_bounds__(256) opt_step_kernel(const laff_opt_tensor* __restrict__ ts, const int* __restrict__ blk_tensor,
                                                       const long long* __restrict__ blk_start, const double* __restrict__ total_norm,
                                                       float max_norm, int kind, float lr, float alpha_or_beta1, float beta2, float eps,
                                                       float bias_c1, float bias_c2, const long long* __restrict__ step_dev,
                                                       const float* __restrict__ lr_dev, const float* __restrict__ ctl) {
  const laff_opt_tensor t = ts[blk_tensor[blockIdx.x]];
  if (!t.grad) return;
  if (ctl && ctl[1] != 0.f) return;   // loss-scaler variant: overflow step, the optimizer is skipped
  if (lr_dev) lr = *lr_dev;
  if (step_dev && kind == 1) {  // graph replays: the step count lives on the device
    const double st = static_cast<double>(*step_dev);
    bias_c1 = static_cast<float>(1.0 - pow(static_cast<double>(alpha_or_beta1), st));
    bias_c2 = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), st)));
  }
  const long long s = blk_start[blockIdx.x];
  float coef = 1.0f;
  if (ctl) {
    coef = ctl[0];
  } else if (max_norm > 0.f) {
    const float c = static_cast<float>(static_cast<double>(max_norm) / (*total_norm + 1e-6));
    coef = c < 1.0f ? c : 1.0f;
  }
  auto update = [&](float gr, float& p, float& s1, float& s2) -> float {
    const float g = gr * coef;
    if (kind == 0) {
      s1 = alpha_or_beta1 * s1 + (1.0f - alpha_or_beta1) * g * g;
      p -= lr * (g / (sqrtf(s1) + eps));
    } else {
      s1 = alpha_or_beta1 * s1 + (1.0f - alpha_or_beta1) * g;
      s2 = beta2 * s2 + (1.0f - beta2) * g * g;
      const float denom = sqrtf(s2) / bias_c2 + eps;  // bias_c2 = sqrt(1 - beta2^t)
      p -= (lr / bias_c1) * (s1 / denom);             // bias_c1 = 1 - beta1^t
    }
    return g;
  };
  const long long j0 = s + threadIdx.x * 8;  // 8 consecutive elements per thread: two 16-byte accesses per array
  const bool vec = j0 + 8 <= t.n && aligned16(t.param) && aligned16(t.grad) && aligned16(t.state1) &&
                   (kind == 0 || aligned16(t.state2)) && (!t.grad_out || aligned16(t.grad_out));
  if (vec) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long j = j0 + 4 * h;
      float4 g = *reinterpret_cast<const float4*>(t.grad + j);
      float4 p = *reinterpret_cast<float4*>(t.param + j);
      float4 a = *reinterpret_cast<float4*>(t.state1 + j);
      float4 b = kind == 1 ? *reinterpret_cast<float4*>(t.state2 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      g.x = update(g.x, p.x, a.x, b.x);
      g.y = update(g.y, p.y, a.y, b.y);
      g.z = update(g.z, p.z, a.z, b.z);
      g.w = update(g.w, p.w, a.w, b.w);
      *reinterpret_cast<float4*>(t.param + j) = p;
      *reinterpret_cast<float4*>(t.state1 + j) = a;
      if (kind == 1) *reinterpret_cast<float4*>(t.state2 + j) = b;
      if (t.grad_out) *reinterpret_cast<float4*>(t.grad_out + j) = g;
    }
  } else {
    for (long long j = j0; j < j0 + 8 && j < t.n; ++j) {
      float p = t.param[j], a = t.state1[j], b = kind == 1 ? t.state2[j] : 0.f;
      const float g = update(t.grad[j], p, a, b);
      t.param[j] = p;
      t.state1[j] = a;
      if (kind == 1) t.state2[j] = b;
      if (t.grad_out) t.grad_out[j] = g;
    }
  }
}

}  // namespace laff

using namespace laff;

extern "C" int laff_transform_train_forward(const float* src, long long ld_src, int src_cols, int B, int D, float p_drop,
                                            unsigned long long seed, const unsigned long long* seed_dev, const float* gamma,
                                            const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                            int use_bn, float* y, long long ld_y, uint8_t* mask, float* save_mean,
                                            float* save_invstd, void* stream) {
  LAFF_REQUIRE(src && y && B > 0 && D > 0 && src_cols > 0 && src_cols <= D && D % src_cols == 0 && ld_src >= src_cols && ld_y >= D,
               LAFF_EINVAL, "laff_transform_train_forward: bad arguments");
  LAFF_REQUIRE(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || mask), LAFF_EINVAL,
               "laff_transform_train_forward: dropout needs 0 <= p < 1 and a mask buffer");
  LAFF_REQUIRE(!use_bn || (save_mean && save_invstd && (running_mean == nullptr) == (running_var == nullptr)), LAFF_EINVAL,
               "laff_transform_train_forward: BatchNorm needs save_mean / save_invstd");
  LAFF_REQUIRE(!use_bn || B > 1, LAFF_EINVAL, "laff_transform_train_forward: BatchNorm in train mode needs more than 1 row "
               "(torch raises 'Expected more than 1 value per channel')");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  transform_train_fwd_kernel<<<(D + 31) / 32, dim3(32, kRowLanes), 0, static_cast<cudaStream_t>(stream)>>>(
      src, ld_src, src_cols, B, D, p_drop, seed, seed_dev, gamma, beta, running_mean, running_var, momentum, eps, use_bn, y, ld_y,
      mask, save_mean, save_invstd);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_transform_train_backward(const float* dy, long long ld_dy, const float* a, long long ld_a, const float* tiled_x,
                                             long long ld_x, int in_dim, const uint8_t* mask, float p_drop, int activation,
                                             int use_bn, const float* gamma, const float* save_mean, const float* save_invstd, int B,
                                             int D, float* dz, long long ld_dz, float* dgamma, float* dbeta, float* dbias,
                                             void* stream) {
  LAFF_REQUIRE(dy && B > 0 && D > 0 && ld_dy >= D && (a != nullptr) != (tiled_x != nullptr), LAFF_EINVAL,
               "laff_transform_train_backward: exactly one of a / tiled_x must be given");
  LAFF_REQUIRE(!a || ld_a >= D, LAFF_EINVAL, "laff_transform_train_backward: bad pitch of a");
  LAFF_REQUIRE(!tiled_x || (in_dim > 0 && D % in_dim == 0 && ld_x >= in_dim), LAFF_EINVAL,
               "laff_transform_train_backward: tiled feature: in_dim must divide D");
  LAFF_REQUIRE(!dz || ld_dz >= D, LAFF_EINVAL, "laff_transform_train_backward: bad pitch of dz");
  LAFF_REQUIRE(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || mask), LAFF_EINVAL, "laff_transform_train_backward: dropout mask missing");
  LAFF_REQUIRE(!use_bn || (save_mean && save_invstd), LAFF_EINVAL, "laff_transform_train_backward: BatchNorm statistics missing");
  LAFF_REQUIRE(activation >= 0 && activation <= 3, LAFF_EINVAL, "laff_transform_train_backward: bad activation");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  transform_train_bwd_kernel<<<(D + 31) / 32, dim3(32, kRowLanes), 0, static_cast<cudaStream_t>(stream)>>>(
      dy, ld_dy, a, ld_a, tiled_x, ld_x, in_dim, mask, p_drop, activation, use_bn, gamma, save_mean, save_invstd, B, D, dz, ld_dz,
      dgamma, dbeta, dbias);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_attention_pool_backward(const float* const* ys, const long long* lds, int n_features, int heads,
                                            int head_dim, const float* att_weight, const float* att_bias, const float* dout,
                                            long long ld_dout, long long rows, float norm_eps, int with_ave, int mul, float omega,
                                            float* const* dys, float* dw_part, float* dc_part, float* dw, float* dc, void* stream) {
  LAFF_REQUIRE(ys && lds && dys && att_weight && att_bias && dout && dw_part && dc_part && dw && dc && rows > 0,
               LAFF_EINVAL, "laff_attention_pool_backward: bad arguments");
  LAFF_REQUIRE(n_features >= 1 && n_features <= LAFF_MAX_FEATURES && heads > 0 && ld_dout >= static_cast<long long>(heads) * head_dim,
               LAFF_EINVAL, "laff_attention_pool_backward: bad shape");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PoolBwdArgs args = {};
  const long long D = static_cast<long long>(heads) * head_dim;
  for (int l = 0; l < n_features; ++l) {
    LAFF_REQUIRE(ys[l] && dys[l] && lds[l] >= D, LAFF_EINVAL, "laff_attention_pool_backward: feature %d: bad pointer / pitch", l);
    args.ys[l] = ys[l];
    args.dys[l] = dys[l];
    args.lds[l] = lds[l];
  }
  const long long warps = rows * heads;
  const unsigned blocks = static_cast<unsigned>((warps + 3) / 4);
#define LAFF_POOL_BWD(VPL)                                                                                                       \
  pool_train_bwd_kernel<VPL, LAFF_MAX_FEATURES><<<blocks, 128, 0, st>>>(args, n_features, att_weight, att_bias, dout, ld_dout, rows, \
                                                                         heads, norm_eps, with_ave, mul, omega, dw_part, dc_part)
  switch (head_dim) {
    case 32: LAFF_POOL_BWD(1); break;
    case 64: LAFF_POOL_BWD(2); break;
    case 128: LAFF_POOL_BWD(4); break;
    case 256: LAFF_POOL_BWD(8); break;
    case 512: LAFF_POOL_BWD(16); break;
    default:
      set_error("laff_attention_pool_backward: head_dim %d unsupported (32, 64, 128, 256, 512)", head_dim);
      return LAFF_ENOTSUP;
  }
#undef LAFF_POOL_BWD
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  // dw[h, :] = sum_rows dw_part[row, h, :]: view the partials as [rows, heads * head_dim]
  const int cols = heads * head_dim;
  colsum_kernel<<<(cols + 127) / 128, 128, 0, st>>>(dw_part, rows, cols, dw);
  colsum_kernel<<<(heads + 127) / 128, 128, 0, st>>>(dc_part, rows, heads, dc);
  count_launch(2);
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_fold_tiles(const float* dz, long long ld_dz, int B, int D, int in_dim, float* dx, long long ld_dx, void* stream) {
  LAFF_REQUIRE(dz && dx && B > 0 && D > 0 && in_dim > 0 && D % in_dim == 0 && ld_dz >= D && ld_dx >= in_dim, LAFF_EINVAL,
               "laff_fold_tiles: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const long long n = static_cast<long long>(B) * in_dim;
  fold_tiles_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dz, ld_dz, B, D, in_dim, dx,
                                                                                                        ld_dx);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_frame_pool_backward(const float* frames, long long B, int F, int dim, const float* att_weight, const float* dout,
                                        long long ld_dout, double norm_eps, float* dw_part, float* dc_part, float* dw, float* dc,
                                        void* stream) {
  LAFF_REQUIRE(frames && att_weight && dout && dw_part && dc_part && dw && dc && B > 0 && F > 0 && ld_dout >= dim, LAFF_EINVAL,
               "laff_frame_pool_backward: bad arguments");
  LAFF_REQUIRE(F <= kMaxFrames, LAFF_ENOTSUP, "laff_frame_pool_backward: at most %d frames per video (got %d)", kMaxFrames, F);
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((B + 3) / 4);
#define LAFF_FRAME_BWD(V)                                                                                                   \
  case V:                                                                                                                   \
    frame_pool_bwd_kernel<V><<<blocks, 128, 0, st>>>(frames, B, F, dim, att_weight, dout, ld_dout, static_cast<float>(norm_eps), \
                                                     dw_part, dc_part);                                                     \
    break;
  switch (dim % 32 == 0 ? dim / 32 : 0) {
    LAFF_FRAME_BWD(1)
    LAFF_FRAME_BWD(2)
    LAFF_FRAME_BWD(4)
    LAFF_FRAME_BWD(8)
    LAFF_FRAME_BWD(16)
    LAFF_FRAME_BWD(32)
    default:
      set_error("laff_frame_pool_backward: dim %d not in {32, 64, 128, 256, 512, 1024}", dim);
      return LAFF_ENOTSUP;
  }
#undef LAFF_FRAME_BWD
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  colsum_kernel<<<(dim + 127) / 128, 128, 0, st>>>(dw_part, B, dim, dw);
  colsum_kernel<<<1, 128, 0, st>>>(dc_part, B, 1, dc);
  count_launch(2);
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_transpose_16(const float* x, long long ld, int rows, int cols, int dtype, int terms, int side, void* out16,
                                 long long ld_out, void* stream) {
  LAFF_REQUIRE(x && out16 && rows > 0 && cols > 0 && ld >= cols && is16(dtype) && (terms == 1 || terms == 3) && (side == 0 || side == 1),
               LAFF_EINVAL, "laff_transpose_16: bad arguments");
  const int Kpad = (rows + 7) / 8 * 8;
  LAFF_REQUIRE(ld_out >= static_cast<long long>(Kpad) * terms && ld_out % 8 == 0, LAFF_EINVAL,
               "laff_transpose_16: output pitch must hold %d x %d elements and be a multiple of 8", terms, Kpad);
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  dim3 grid((cols + 31) / 32, (Kpad + 31) / 32), block(32, 8);
  transpose16_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(x, ld, rows, cols, dtype, terms, side,
                                                                           static_cast<uint16_t*>(out16), ld_out, Kpad);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_optimizer_blocks(const long long* sizes, int n_tensors, int* blk_tensor, long long* blk_start, int capacity) {
  LAFF_REQUIRE(sizes && n_tensors >= 0, LAFF_EINVAL, "laff_optimizer_blocks: bad arguments");
  long long total = 0;
  for (int t = 0; t < n_tensors; ++t) {
    LAFF_REQUIRE(sizes[t] >= 0, LAFF_EINVAL, "laff_optimizer_blocks: negative size");
    for (long long s = 0; s < sizes[t]; s += kOptChunk) {
      if (blk_tensor && total < capacity) {
        blk_tensor[total] = t;
        blk_start[total] = s;
      }
      ++total;
    }
  }
  LAFF_REQUIRE(total < (1LL << 31), LAFF_ENOTSUP, "laff_optimizer_blocks: too many parameters");
  return static_cast<int>(total);
}

extern "C" int laff_optimizer_step(const laff_opt_tensor* tensors_dev, const int* blk_tensor_dev, const long long* blk_start_dev,
                                   int n_blocks, int kind, float lr, float alpha_or_beta1, float beta2, float eps, long long step,
                                   float max_grad_norm, double* partial_dev, double* total_norm_dev, const long long* step_dev,
                                   const float* lr_dev, void* stream) {
  LAFF_REQUIRE(tensors_dev && blk_tensor_dev && blk_start_dev && partial_dev && total_norm_dev && n_blocks >= 0 &&
                   (kind == 0 || kind == 1) && step >= 1,
               LAFF_EINVAL, "laff_optimizer_step: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if (n_blocks == 0) return LAFF_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  grad_sqnorm_kernel<<<n_blocks, 256, 0, st>>>(tensors_dev, blk_tensor_dev, blk_start_dev, partial_dev, nullptr);
  sqnorm_final_kernel<<<1, 256, 0, st>>>(partial_dev, n_blocks, total_norm_dev);
  float c1 = 1.f, c2 = 1.f;
  if (kind == 1) {
    c1 = static_cast<float>(1.0 - pow(static_cast<double>(alpha_or_beta1), static_cast<double>(step)));
    c2 = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), static_cast<double>(step))));
  }
  opt_step_kernel<<<n_blocks, 256, 0, st>>>(tensors_dev, blk_tensor_dev, blk_start_dev, total_norm_dev, max_grad_norm, kind, lr,
                                            alpha_or_beta1, beta2, eps, c1, c2, step_dev, lr_dev, nullptr);
  count_launch(3);
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_optimizer_step_scaled(const laff_opt_tensor* tensors_dev, const int* blk_tensor_dev,
                                          const long long* blk_start_dev, int n_blocks, int kind, float alpha_or_beta1, float beta2,
                                          float eps, float max_grad_norm, double* partial_dev, float* partial_max_dev,
                                          double* total_norm_dev, long long* step_dev, const float* lr_dev,
                                          laff_scaler_state* scaler_dev, float growth_factor, float backoff_factor,
                                          int growth_interval, float overflow_limit, float* ctl_dev, void* stream) {
  LAFF_REQUIRE(tensors_dev && blk_tensor_dev && blk_start_dev && partial_dev && partial_max_dev && total_norm_dev && step_dev &&
                   lr_dev && scaler_dev && ctl_dev && n_blocks >= 0 && (kind == 0 || kind == 1) && growth_factor >= 1.f &&
                   backoff_factor > 0.f && backoff_factor <= 1.f && growth_interval >= 1,
               LAFF_EINVAL, "laff_optimizer_step_scaled: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if (n_blocks == 0) return LAFF_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  grad_sqnorm_kernel<<<n_blocks, 256, 0, st>>>(tensors_dev, blk_tensor_dev, blk_start_dev, partial_dev, partial_max_dev);
  scaler_final_kernel<<<1, 256, 0, st>>>(partial_dev, partial_max_dev, n_blocks, max_grad_norm, scaler_dev, growth_factor,
                                         backoff_factor, growth_interval, overflow_limit, total_norm_dev, ctl_dev, step_dev);
  opt_step_kernel<<<n_blocks, 256, 0, st>>>(tensors_dev, blk_tensor_dev, blk_start_dev, total_norm_dev, max_grad_norm, kind, 0.f,
                                            alpha_or_beta1, beta2, eps, 1.f, 1.f, step_dev, lr_dev, ctl_dev);
  count_launch(3);
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}
