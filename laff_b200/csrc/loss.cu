// Hardest-negative margin ranking loss, forward + backward.
//   loss.py:95-135   MarginRankingLoss.forward(s, im): scores = cosine_sim(im, s) -> rows = videos, cols = sentences
//   loss.py:161-200  MarginRankingLossWithScore.forward(score)
//   model/model.py:852-862, :2036-2038  sum over heads
// B = 128, H = 8, d_h = 512 is 134 MFLOP: latency-bound, so plain fp32 CUDA-core kernels (exact fp32 products, like
// the reference's fp32 mm) and a handful of launches.
#include <cstring>

#include "host_util.cuh"

namespace laff {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// xhat[r, h, :] = x[r, h, :] / (||x[r, h, :]|| + eps);  den[r, h] = ||x|| + eps;  nrm[r, h] = ||x||   (loss.py:8-13)
__global__ void mrl_normalize_kernel(const float* __restrict__ x, long long items, int dh, float eps,
                                     float* __restrict__ xhat, float* __restrict__ nrm) {
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= items) return;
  const float* p = x + warp * dh;
  float ss = 0.f;
  for (int d = lane; d < dh; d += 32) ss = fmaf(p[d], p[d], ss);
  ss = wsum(ss);
  const float n = sqrtf(ss);
  const float den = n + eps;
  for (int d = lane; d < dh; d += 32) xhat[warp * dh + d] = p[d] / den;
  if (lane == 0) nrm[warp] = n;
}

// S[h][i][j] = vis_hat[i, h, :] . txt_hat[j, h, :]      32x32 tile per block, K chunks of 32 through smem
__global__ void mrl_scores_kernel(const float* __restrict__ vis_hat, const float* __restrict__ txt_hat, int B, int H,
                                  int dh, float* __restrict__ S) {
  __shared__ float sv[32][33];
  __shared__ float st[32][33];
  const int h = blockIdx.z;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < dh; k0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      const int i = i0 + r, j = j0 + r, k = k0 + tx;
      sv[r][tx] = (i < B && k < dh) ? vis_hat[(static_cast<long long>(i) * H + h) * dh + k] : 0.f;
      st[r][tx] = (j < B && k < dh) ? txt_hat[(static_cast<long long>(j) * H + h) * dh + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float t = st[tx][k];
#pragma unroll
      for (int a = 0; a < 4; ++a) acc[a] = fmaf(sv[ty + 8 * a][k], t, acc[a]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty + 8 * a, j = j0 + tx;
    if (i < B && j < B) S[(static_cast<long long>(h) * B + i) * B + j] = acc[a];
  }
}

// One block per head.  Thread t owns column t (t2i: max/sum over rows i) and row t (i2t: max/sum over columns j).
// dS (may be NULL) receives dLoss/dS; it must be zero on entry.  head_loss[h] = this head's loss.
__global__ void mrl_hinge_kernel(const float* __restrict__ S, int B, long long ld, long long head_stride, float margin,
                                 int max_violation, int direction, int cost_mean, float* __restrict__ head_loss,
                                 float* __restrict__ dS) {
  __shared__ float s_red[32];
  const int h = blockIdx.x;
  const float* Sh = S + h * head_stride;
  float* dSh = dS ? dS + h * head_stride : nullptr;
  const float denom = cost_mean ? (max_violation ? static_cast<float>(B) : static_cast<float>(B) * static_cast<float>(B)) : 1.0f;
  const float gscale = 1.0f / denom;
  float local = 0.f;
  for (int t = threadIdx.x; t < B; t += blockDim.x) {
    const float dt = Sh[static_cast<long long>(t) * ld + t];
    if (direction == LAFF_DIR_T2I || direction == LAFF_DIR_BIDIR) {
      // cost_im[i][t] = max(0, margin + S[i][t] - diag[t]), i != t      (loss.py:115-118)
      float best = 0.f;
      int besti = -1;
      float sum = 0.f;
      int nviol = 0;
      for (int i = 0; i < B; ++i) {
        if (i == t) continue;
        const float c = fmaxf(margin + Sh[static_cast<long long>(i) * ld + t] - dt, 0.f);
        if (max_violation) {
          if (c > best) {
            best = c;
            besti = i;
          }
        } else if (c > 0.f) {
          sum += c;
          ++nviol;
          if (dSh) atomicAdd(dSh + static_cast<long long>(i) * ld + t, gscale);
        }
      }
      if (max_violation) {
        local += best;
        if (dSh && besti >= 0) {
          atomicAdd(dSh + static_cast<long long>(besti) * ld + t, gscale);
          atomicAdd(dSh + static_cast<long long>(t) * ld + t, -gscale);
        }
      } else {
        local += sum;
        if (dSh && nviol) atomicAdd(dSh + static_cast<long long>(t) * ld + t, -gscale * static_cast<float>(nviol));
      }
    }
    if (direction == LAFF_DIR_I2T || direction == LAFF_DIR_BIDIR) {
      // cost_s[t][j] = max(0, margin + S[t][j] - diag[t]), j != t       (loss.py:110-113)
      float best = 0.f;
      int bestj = -1;
      float sum = 0.f;
      int nviol = 0;
      for (int j = 0; j < B; ++j) {
        if (j == t) continue;
        const float c = fmaxf(margin + Sh[static_cast<long long>(t) * ld + j] - dt, 0.f);
        if (max_violation) {
          if (c > best) {
            best = c;
            bestj = j;
          }
        } else if (c > 0.f) {
          sum += c;
          ++nviol;
          if (dSh) atomicAdd(dSh + static_cast<long long>(t) * ld + j, gscale);
        }
      }
      if (max_violation) {
        local += best;
        if (dSh && bestj >= 0) {
          atomicAdd(dSh + static_cast<long long>(t) * ld + bestj, gscale);
          atomicAdd(dSh + static_cast<long long>(t) * ld + t, -gscale);
        }
      } else {
        local += sum;
        if (dSh && nviol) atomicAdd(dSh + static_cast<long long>(t) * ld + t, -gscale * static_cast<float>(nviol));
      }
    }
  }
  local = wsum(local);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += s_red[w];
    head_loss[h] = tot * gscale;
  }
}

// DualSoftmaxLoss (loss.py:291-310) on one head's score matrix X [B, B] (rows = videos, columns = sentences):
//   loss = (f(X^T) + f(X)) / 2,   f(A) = -sum_i log_softmax_row(B * A * softmax_col(A / temp))[i, i]
// One block per head, thread t owns column t in the column passes and row t in the row passes.  dX (may be NULL)
// receives dLoss/dX.  All statistics live in shared memory (B <= 1024).
template <bool TR>
__device__ void dsl_term(const float* __restrict__ X, int B, float temp, float* s_cm, float* s_cz, float* s_rm, float* s_rz,
                         float* s_cc, float* s_red, float* loss_acc, float* __restrict__ dX, bool accumulate) {
  auto A = [&](int i, int j) -> float { return TR ? X[static_cast<long long>(j) * B + i] : X[static_cast<long long>(i) * B + j]; };
  const float invT = 1.0f / temp, fB = static_cast<float>(B);
  for (int j = threadIdx.x; j < B; j += blockDim.x) {  // column softmax statistics
    float m = -INFINITY;
    for (int i = 0; i < B; ++i) m = fmaxf(m, A(i, j) * invT);
    float z = 0.f;
    for (int i = 0; i < B; ++i) z += expf(A(i, j) * invT - m);
    s_cm[j] = m;
    s_cz[j] = z;
  }
  __syncthreads();
  auto P0 = [&](int i, int j) -> float { return expf(A(i, j) * invT - s_cm[j]) / s_cz[j]; };
  float local = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {  // row log-softmax of M = B * A * P0
    float m = -INFINITY;
    for (int j = 0; j < B; ++j) m = fmaxf(m, fB * A(i, j) * P0(i, j));
    float z = 0.f;
    for (int j = 0; j < B; ++j) z += expf(fB * A(i, j) * P0(i, j) - m);
    s_rm[i] = m;
    s_rz[i] = z;
    local -= fB * A(i, i) * P0(i, i) - m - logf(z);
  }
  local = wsum(local);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += s_red[w];
    *loss_acc += 0.5f * t;
  }
  if (!dX) {
    __syncthreads();
    return;
  }
  auto dM = [&](int i, int j) -> float {
    const float a = A(i, j), p0 = P0(i, j);
    return expf(fB * a * p0 - s_rm[i]) / s_rz[i] - (i == j ? 1.0f : 0.0f);
  };
  for (int j = threadIdx.x; j < B; j += blockDim.x) {  // c_j = sum_k P0_kj * dP0_kj,  dP0 = B * A * dM
    float c = 0.f;
    for (int k = 0; k < B; ++k) c = fmaf(P0(k, j), fB * A(k, j) * dM(k, j), c);
    s_cc[j] = c;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < B * B; e += blockDim.x) {
    const int i = e / B, j = e - i * B;
    const float a = A(i, j), p0 = P0(i, j), dm = dM(i, j);
    const float d = 0.5f * (fB * p0 * dm + invT * p0 * (fB * a * dm - s_cc[j]));
    float* o = TR ? dX + static_cast<long long>(j) * B + i : dX + static_cast<long long>(i) * B + j;
    *o = accumulate ? *o + d : d;
  }
  __syncthreads();
}

__global__ void dsl_kernel(const float* __restrict__ S, int B, long long head_stride, float temp, float* __restrict__ head_loss,
                           float* __restrict__ dS) {
  extern __shared__ float s_dsl[];  // 5 * B + 32 floats
  float* s_cm = s_dsl;
  float* s_cz = s_cm + B;
  float* s_rm = s_cz + B;
  float* s_rz = s_rm + B;
  float* s_cc = s_rz + B;
  float* s_red = s_cc + B;
  __shared__ float s_loss;
  const int h = blockIdx.x;
  if (threadIdx.x == 0) s_loss = 0.f;
  __syncthreads();
  const float* X = S + h * head_stride;
  float* dX = dS ? dS + h * head_stride : nullptr;
  dsl_term<true>(X, B, temp, s_cm, s_cz, s_rm, s_rz, s_cc, s_red, &s_loss, dX, false);   // f(cosine_sim(s, im)) = f(X^T)
  dsl_term<false>(X, B, temp, s_cm, s_cz, s_rm, s_rz, s_cc, s_red, &s_loss, dX, true);   // f(sim^T) = f(X)
  if (threadIdx.x == 0) head_loss[h] = s_loss;
}

__global__ void mrl_sum_heads_kernel(const float* __restrict__ head_loss, int H, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t = 0.f;
    for (int h = 0; h < H; ++h) t += head_loss[h];  // model/model.py:857-858 accumulates head by head
    *loss = t;
  }
}

// Block per (row r, head h).  g = sum_c coef[c] * other_hat[c, h, :] where coef = dS[h][r][c] (vis rows) or
// dS[h][c][r] (txt rows); then back through x / (||x|| + eps):  dx = g/den - xhat * (xhat . g) / n
__global__ void mrl_grad_kernel(const float* __restrict__ dS, const float* __restrict__ other_hat,
                                const float* __restrict__ self_hat, const float* __restrict__ self_nrm, int B, int H,
                                int dh, float eps, int transpose, float* __restrict__ dx) {
  extern __shared__ float s_g[];  // dh floats + 32
  float* s_red = s_g + dh;
  const int r = blockIdx.x, h = blockIdx.y;
  const float* dSh = dS + static_cast<long long>(h) * B * B;
  for (int d = threadIdx.x; d < dh; d += blockDim.x) s_g[d] = 0.f;
  __syncthreads();
  for (int c = 0; c < B; ++c) {
    const float coef = transpose ? dSh[static_cast<long long>(c) * B + r] : dSh[static_cast<long long>(r) * B + c];
    if (coef != 0.f) {
      const float* o = other_hat + (static_cast<long long>(c) * H + h) * dh;
      for (int d = threadIdx.x; d < dh; d += blockDim.x) s_g[d] = fmaf(coef, o[d], s_g[d]);
    }
  }
  __syncthreads();
  const float* xh = self_hat + (static_cast<long long>(r) * H + h) * dh;
  float dot = 0.f;
  for (int d = threadIdx.x; d < dh; d += blockDim.x) dot = fmaf(xh[d], s_g[d], dot);
  dot = wsum(dot);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dot;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) tot += s_red[w];
  const float n = self_nrm[static_cast<long long>(r) * H + h];
  const float den = n + eps;
  const float k = n > 0.f ? tot / n : 0.f;
  float* o = dx + (static_cast<long long>(r) * H + h) * dh;
  for (int d = threadIdx.x; d < dh; d += blockDim.x) o[d] = s_g[d] / den - xh[d] * k;
}

}  // namespace laff

using namespace laff;

static size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

extern "C" {

size_t laff_mrl_workspace_bytes(int B, int H, int dh) {
  if (B <= 0 || H <= 0 || dh <= 0) return 0;
  const size_t emb = align256(static_cast<size_t>(B) * H * dh * 4);
  const size_t nrm = align256(static_cast<size_t>(B) * H * 4);
  const size_t mat = align256(static_cast<size_t>(H) * B * B * 4);
  return 2 * emb + 2 * nrm + 2 * mat + align256(static_cast<size_t>(H) * 4) + 256;
}

static int embedding_loss(int kind, const float* txt, const float* vis, int B, int H, int dh, float margin, int max_violation,
                          int direction, int cost_mean, float temp, float* loss, float* d_txt, float* d_vis, void* workspace,
                          size_t workspace_bytes, void* stream) {
  LAFF_REQUIRE(txt && vis && loss && workspace && B > 0 && H > 0 && dh > 0, LAFF_EINVAL,
               "laff_mrl_forward_backward / laff_dsl_forward_backward: bad arguments");
  LAFF_REQUIRE(direction >= 0 && direction <= 2, LAFF_EINVAL, "laff_mrl_forward_backward: bad direction %d", direction);
  LAFF_REQUIRE(kind == 0 || (B <= 1024 && temp > 0.f), LAFF_ENOTSUP, "laff_dsl_forward_backward: batch size %d > 1024 or temp <= 0", B);
  LAFF_REQUIRE(workspace_bytes >= laff_mrl_workspace_bytes(B, H, dh), LAFF_EWORKSPACE,
               "laff_mrl_forward_backward: workspace too small");
  LAFF_REQUIRE(dh <= 8192, LAFF_ENOTSUP, "laff_mrl_forward_backward: head dim %d too large", dh);
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t emb = align256(static_cast<size_t>(B) * H * dh * 4);
  const size_t nrm = align256(static_cast<size_t>(B) * H * 4);
  const size_t mat = align256(static_cast<size_t>(H) * B * B * 4);
  uintptr_t p = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~static_cast<uintptr_t>(255);
  float* txt_hat = reinterpret_cast<float*>(p); p += emb;
  float* vis_hat = reinterpret_cast<float*>(p); p += emb;
  float* txt_nrm = reinterpret_cast<float*>(p); p += nrm;
  float* vis_nrm = reinterpret_cast<float*>(p); p += nrm;
  float* S = reinterpret_cast<float*>(p); p += mat;
  float* dS = reinterpret_cast<float*>(p); p += mat;
  float* head_loss = reinterpret_cast<float*>(p);
  const float eps = 1e-13f + 1e-14f;  // loss.py:8, :11 (default eps + 1e-14)
  const long long items = static_cast<long long>(B) * H;
  const int nblk = static_cast<int>((items * 32 + 255) / 256);
  mrl_normalize_kernel<<<nblk, 256, 0, st>>>(txt, items, dh, eps, txt_hat, txt_nrm); laff::count_launch();
  mrl_normalize_kernel<<<nblk, 256, 0, st>>>(vis, items, dh, eps, vis_hat, vis_nrm); laff::count_launch();
  dim3 grid((B + 31) / 32, (B + 31) / 32, H), block(32, 8);
  mrl_scores_kernel<<<grid, block, 0, st>>>(vis_hat, txt_hat, B, H, dh, S); laff::count_launch();
  const bool need_grad = d_txt != nullptr || d_vis != nullptr;
  if (need_grad) LAFF_CUDA(cudaMemsetAsync(dS, 0, static_cast<size_t>(H) * B * B * 4, st));
  const int hthreads = B >= 1024 ? 1024 : ((B + 31) / 32) * 32;
  if (kind == 0) {
    mrl_hinge_kernel<<<H, hthreads, 0, st>>>(S, B, B, static_cast<long long>(B) * B, margin, max_violation, direction,
                                             cost_mean, head_loss, need_grad ? dS : nullptr);
  } else {
    dsl_kernel<<<H, hthreads, static_cast<size_t>(5 * B + 32) * 4, st>>>(S, B, static_cast<long long>(B) * B, temp, head_loss,
                                                                         need_grad ? dS : nullptr);
  }
  laff::count_launch();
  mrl_sum_heads_kernel<<<1, 32, 0, st>>>(head_loss, H, loss); laff::count_launch();
  const size_t gsm = static_cast<size_t>(dh + 32) * 4;
  if (d_vis) mrl_grad_kernel<<<dim3(B, H), 128, gsm, st>>>(dS, txt_hat, vis_hat, vis_nrm, B, H, dh, eps, 0, d_vis); laff::count_launch();
  if (d_txt) mrl_grad_kernel<<<dim3(B, H), 128, gsm, st>>>(dS, vis_hat, txt_hat, txt_nrm, B, H, dh, eps, 1, d_txt); laff::count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

int laff_mrl_forward_backward(const float* txt, const float* vis, int B, int H, int dh, float margin,
                              int max_violation, int direction, int cost_mean, float* loss, float* d_txt,
                              float* d_vis, void* workspace, size_t workspace_bytes, void* stream) {
  return embedding_loss(0, txt, vis, B, H, dh, margin, max_violation, direction, cost_mean, 1.0f, loss, d_txt, d_vis, workspace,
                        workspace_bytes, stream);
}

int laff_dsl_forward_backward(const float* txt, const float* vis, int B, int H, int dh, float temp, float* loss, float* d_txt,
                              float* d_vis, void* workspace, size_t workspace_bytes, void* stream) {
  return embedding_loss(1, txt, vis, B, H, dh, 0.f, 0, 0, 0, temp, loss, d_txt, d_vis, workspace, workspace_bytes, stream);
}

int laff_mrl_score_forward_backward(const float* score, int B, long long ld, float margin, int max_violation,
                                    int direction, int cost_mean, float* loss, float* d_score, void* stream) {
  LAFF_REQUIRE(score && loss && B > 0 && ld >= B, LAFF_EINVAL, "laff_mrl_score_forward_backward: bad arguments");
  LAFF_REQUIRE(direction >= 0 && direction <= 2, LAFF_EINVAL, "laff_mrl_score_forward_backward: bad direction");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_score) LAFF_CUDA(cudaMemset2DAsync(d_score, static_cast<size_t>(ld) * 4, 0, static_cast<size_t>(B) * 4, B, st));
  const int hthreads = B >= 1024 ? 1024 : ((B + 31) / 32) * 32;
  // head_loss[0] is written straight into *loss (H = 1)
  mrl_hinge_kernel<<<1, hthreads, 0, st>>>(score, B, ld, 0, margin, max_violation, direction, cost_mean, loss, d_score); laff::count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

}  // extern "C"
