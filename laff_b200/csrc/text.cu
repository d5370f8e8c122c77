// Text front-end on the device: the numeric half of BoWTxtEncoder / W2VTxtEncoder / GruTxtEncoder
// (model/model.py:322-434, txt2vec.py:49-109).  Tokenising strings stays on the host; everything after the token ids
// runs here.  All of it is small HBM/latency-bound work next to the fusion GEMMs.
//   laff_bow_counts     BowVec._encoding (txt2vec.py:56-63): vec[idx] += 1 per in-vocabulary token
//   laff_gather_mean    W2Vec._encoding (txt2vec.py:97-104): mean of the word vectors, accumulated in float64 in the
//                       order given (numpy: np.array(list of float lists).mean(axis=0) is a sequential fp64 row sum)
//   laff_gather_rows    nn.Embedding lookup (model/model.py:352) for the GRU input
//   laff_gru_cell       one nn.GRU time step (gates r, z, n in PyTorch's order) for packed variable-length sequences,
//                       with the running sum for 'mean' pooling and the last valid state for 'last' pooling (:361-383)
#include <cstdint>

#include "host_util.cuh"

namespace laff {

__global__ void bow_counts_kernel(const long long* __restrict__ offsets, const int32_t* __restrict__ ids, int ndims,
                                  float* __restrict__ out, long long ld) {
  const long long row = blockIdx.x;
  float* o = out + row * ld;
  for (int c = threadIdx.x; c < ndims; c += blockDim.x) o[c] = 0.f;
  __syncthreads();
  const long long a = offsets[row], b = offsets[row + 1];
  for (long long t = a + threadIdx.x; t < b; t += blockDim.x) {
    const int id = ids[t];
    if (id >= 0 && id < ndims) atomicAdd(o + id, 1.0f);  // counts are small integers: exact in fp32 in any order
  }
}

// Sparse BoW projection (BoWTxtEncoder + its TransformNet, model/model.py:399-417 then :257-276): the reference builds the
// dense count vector c (txt2vec.py:56-63: vec[idx] += 1 per in-vocabulary token, ~8 non-zeros of 3981) and multiplies it
// by W [D, vocab]; here y[r, :] = BN(act(row_scale[r] * sum_t Wt[id_t, :] + b)) is a gather-sum over the rows of the
// transposed weight Wt [vocab, D] fp32 -- a repeated token is simply added twice, which IS its count.  One block per
// caption, 4 consecutive columns per thread per pass (16-byte loads of 16 KB rows: coalesced); tokens in caption order,
// fp32 accumulation (the dense tensor-core path rounds W to 16 bits; this one does not).  HBM/L2-bound:
// n_tokens * D * 4 bytes read + D * 4 written per caption; Wt (65 MB at D = 4096, vocab = 3981) stays L2-resident.
__device__ __forceinline__ float bow_act(float z, int act) {
  switch (act) {
    case LAFF_ACT_TANH: {   // same formulation as EpiProject (fuse.cu): 1 - 2 / (exp(2z) + 1), |err| < 3e-7
      const float e = __expf(2.0f * z);
      return 1.0f - __fdividef(2.0f, e + 1.0f);
    }
    case LAFF_ACT_RELU: return fmaxf(z, 0.f);
    case LAFF_ACT_SIGMOID: return __fdividef(1.0f, 1.0f + __expf(-z));
    default: return z;
  }
}

__global__ void __launch_bounds__(256) bow_project_kernel(const long long* __restrict__ offsets, const int32_t* __restrict__ ids,
                                                          long long id_base, int vocab, const float* __restrict__ wt, long long ld_wt,
                                                          int D, const float* __restrict__ bias, int act,
                                                          const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                                                          const float* __restrict__ row_scale, float* __restrict__ y, long long ld_y) {
  const long long row = blockIdx.x;
  const long long a = offsets[row] - id_base, b = offsets[row + 1] - id_base;
  __shared__ int s_ids[256];
  const float rs = row_scale ? row_scale[row] : 1.0f;
  for (int c0 = 0; c0 < D; c0 += 256 * 4) {
    const int c = c0 + threadIdx.x * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long t0 = a; t0 < b; t0 += 256) {
      __syncthreads();
      if (t0 + threadIdx.x < b) s_ids[threadIdx.x] = ids[t0 + threadIdx.x];
      __syncthreads();
      const int n = b - t0 < 256 ? static_cast<int>(b - t0) : 256;
      if (c < D) {
        for (int t = 0; t < n; ++t) {
          const int id = s_ids[t];
          if (id < 0 || id >= vocab) continue;   // out-of-vocabulary marker: contributes nothing (vocab.find() < 0)
          const float4 w = __ldg(reinterpret_cast<const float4*>(wt + static_cast<long long>(id) * ld_wt + c));
          acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
        }
      }
    }
    if (c < D) {
      float v[4] = {acc.x * rs, acc.y * rs, acc.z * rs, acc.w * rs};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float z = v[j] + (bias ? __ldg(bias + c + j) : 0.f);
        z = bow_act(z, act);
        if (bn_scale) z = fmaf(z, __ldg(bn_scale + c + j), __ldg(bn_shift + c + j));
        v[j] = z;
      }
      *reinterpret_cast<float4*>(y + row * ld_y + c) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

__global__ void gather_mean_kernel(const float* __restrict__ table, long long ld_t, long long n_table,
                                   const long long* __restrict__ offsets, const int32_t* __restrict__ ids, int dim,
                                   float* __restrict__ out, long long ld) {
  const long long row = blockIdx.x;
  const long long a = offsets[row], b = offsets[row + 1];
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    double acc = 0.0;
    long long n = 0;
    for (long long t = a; t < b; ++t) {
      const long long id = ids[t];
      if (id < 0 || id >= n_table) continue;
      acc += static_cast<double>(table[id * ld_t + c]);
      ++n;
    }
    out[row * ld + c] = n ? static_cast<float>(acc / static_cast<double>(n)) : 0.f;
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ table, long long ld_t, long long n_table,
                                   const int32_t* __restrict__ ids, long long n, int dim, float* __restrict__ out,
                                   long long ld) {
  const long long row = blockIdx.x;
  if (row >= n) return;
  const long long id = ids[row];
  const bool ok = id >= 0 && id < n_table;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) out[row * ld + c] = ok ? table[id * ld_t + c] : 0.f;
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// gi: input-side pre-activations of step t for every sequence, [B, 3H] at row pitch ld_gi (b_ih included);
// gh: hidden-side pre-activations W_hh h_{t-1} + b_hh, [B, 3H].  Sequences shorter than t + 1 keep their state.
__global__ void gru_cell_kernel(const float* __restrict__ gi, long long ld_gi, const float* __restrict__ gh, long long ld_gh,
                                const float* __restrict__ h_prev, const int32_t* __restrict__ lengths, int t, int B, int H,
                                float* __restrict__ h_out, float* __restrict__ sum, float* __restrict__ last) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * H) return;
  const int b = static_cast<int>(i / H), j = static_cast<int>(i - static_cast<long long>(b) * H);
  const float hp = h_prev[i];
  const int len = lengths[b];
  if (t >= len) {
    h_out[i] = hp;
    return;
  }
  const float* gib = gi + b * ld_gi;
  const float* ghb = gh + b * ld_gh;
  const float r = sigmoid_acc(gib[j] + ghb[j]);
  const float z = sigmoid_acc(gib[H + j] + ghb[H + j]);
  const float n = tanhf(gib[2 * H + j] + r * ghb[2 * H + j]);
  const float h = (1.0f - z) * n + z * hp;
  h_out[i] = h;
  if (sum) sum[i] += h;
  if (last && t == len - 1) last[i] = h;
}

// Backward of one GRU step (BPTT).  dh_carry [B, H] holds d loss / d h_t flowing in from step t + 1 through the
// z * h path, dh_gemm (may be NULL) the part that came through W_hh (dGh_{t+1} @ W_hh, computed by the GEMM engine).
// The pooling's own contribution is added here: mean -> dout / len for t < len, last -> dlast at t == len - 1.
// Outputs: dgi_t, dgh_t [B, 3H] (gate order r, z, n) and the new dh_carry (= dh * z, or dh unchanged on frozen steps).
__global__ void gru_cell_bwd_kernel(const float* __restrict__ gi, long long ld_gi, const float* __restrict__ gh, long long ld_gh,
                                    const float* __restrict__ h_prev, const float* __restrict__ dmean, const float* __restrict__ dlast,
                                    const float* __restrict__ dh_gemm, const int32_t* __restrict__ lengths, int t, int B, int H,
                                    float* __restrict__ dh_carry, float* __restrict__ dgi, long long ld_dgi, float* __restrict__ dgh,
                                    long long ld_dgh) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * H) return;
  const int b = static_cast<int>(i / H), j = static_cast<int>(i - static_cast<long long>(b) * H);
  const int len = lengths[b];
  float dh = dh_carry[i] + (dh_gemm ? dh_gemm[i] : 0.f);
  float* dgib = dgi + b * ld_dgi;
  float* dghb = dgh + b * ld_dgh;
  if (t >= len) {  // the state was carried through unchanged
    dgib[j] = dgib[H + j] = dgib[2 * H + j] = 0.f;
    dghb[j] = dghb[H + j] = dghb[2 * H + j] = 0.f;
    dh_carry[i] = dh;
    return;
  }
  if (dmean) dh += dmean[i] / static_cast<float>(len);
  if (dlast && t == len - 1) dh += dlast[i];
  const float* gib = gi + b * ld_gi;
  const float* ghb = gh + b * ld_gh;
  const float r = sigmoid_acc(gib[j] + ghb[j]);
  const float z = sigmoid_acc(gib[H + j] + ghb[H + j]);
  const float ghn = ghb[2 * H + j];
  const float n = tanhf(gib[2 * H + j] + r * ghn);
  const float hp = h_prev[i];
  const float dn = dh * (1.0f - z);
  const float dz = dh * (hp - n);
  const float dpre_n = dn * (1.0f - n * n);
  const float dpre_z = dz * z * (1.0f - z);
  const float dpre_r = dpre_n * ghn * r * (1.0f - r);
  dgib[j] = dpre_r;
  dgib[H + j] = dpre_z;
  dgib[2 * H + j] = dpre_n;
  dghb[j] = dpre_r;
  dghb[H + j] = dpre_z;
  dghb[2 * H + j] = dpre_n * r;
  dh_carry[i] = dh * z;
}

// nn.Embedding backward: table_grad[ids[i], :] += dx[i, :]  (table_grad zeroed by the caller; fp32 atomics)
__global__ void scatter_add_rows_kernel(const float* __restrict__ dx, long long ld, const int32_t* __restrict__ ids, long long n,
                                        int dim, long long n_table, float* __restrict__ table_grad, long long ld_t) {
  const long long row = blockIdx.x;
  if (row >= n) return;
  const long long id = ids[row];
  if (id < 0 || id >= n_table) return;
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    const float v = dx[row * ld + c];
    if (v != 0.f) atomicAdd(table_grad + id * ld_t + c, v);
  }
}

// out[c] = sum_r x[r, c] in a fixed order (bias gradients)
__global__ void column_sum_kernel(const float* __restrict__ x, long long ld, long long rows, int cols, float* __restrict__ out) {
  __shared__ double sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0;
  if (c < cols)
    for (long long r = threadIdx.y; r < rows; r += 8) s += x[r * ld + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += sh[k][threadIdx.x];
    out[c] = static_cast<float>(t);
  }
}

__global__ void scale_rows_by_length_kernel(float* __restrict__ x, const int32_t* __restrict__ lengths, int B, int H) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * H) return;
  const int len = lengths[i / H];
  x[i] = len > 0 ? x[i] / static_cast<float>(len) : 0.f;
}

}  // namespace laff

using namespace laff;

extern "C" int laff_bow_counts(const long long* tok_offsets, const int32_t* tok_ids, int rows, int ndims, float* out,
                               long long ld, void* stream) {
  if (rows == 0) return LAFF_OK;
  LAFF_REQUIRE(tok_offsets && out && rows > 0 && ndims > 0 && ld >= ndims, LAFF_EINVAL, "laff_bow_counts: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if (rows == 0) return LAFF_OK;
  bow_counts_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(tok_offsets, tok_ids, ndims, out, ld);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_bow_project(const long long* tok_offsets, const int32_t* tok_ids, long long id_base, int rows, int vocab,
                                const float* wt, long long ld_wt, int D, const float* bias, int activation,
                                const float* bn_scale, const float* bn_shift, const float* row_scale, float* y, long long ld_y,
                                void* stream) {
  if (rows == 0) return LAFF_OK;
  LAFF_REQUIRE(tok_offsets && wt && y && rows > 0 && vocab > 0 && D > 0 && D % 4 == 0 && ld_wt >= D && ld_wt % 4 == 0 &&
                   ld_y >= D && ld_y % 4 == 0 && (reinterpret_cast<uintptr_t>(wt) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
               LAFF_EINVAL, "laff_bow_project: bad arguments (D and the row pitches must be multiples of 4, 16-byte aligned)");
  LAFF_REQUIRE((bn_scale == nullptr) == (bn_shift == nullptr), LAFF_EINVAL, "laff_bow_project: bn_scale / bn_shift mismatch");
  LAFF_REQUIRE(activation >= LAFF_ACT_NONE && activation <= LAFF_ACT_SIGMOID, LAFF_EINVAL, "laff_bow_project: activation %d", activation);
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  bow_project_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(tok_offsets, tok_ids, id_base, vocab, wt, ld_wt, D, bias,
                                                                          activation, bn_scale, bn_shift, row_scale, y, ld_y);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_gather_mean(const float* table, long long ld_table, long long n_table, const long long* offsets,
                                const int32_t* ids, int rows, int dim, float* out, long long ld, void* stream) {
  if (rows == 0) return LAFF_OK;
  LAFF_REQUIRE(offsets && out && rows > 0 && dim > 0 && ld >= dim && ld_table >= dim && n_table >= 0 && (table || n_table == 0),
               LAFF_EINVAL, "laff_gather_mean: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if (rows == 0) return LAFF_OK;
  gather_mean_kernel<<<rows, 128, 0, static_cast<cudaStream_t>(stream)>>>(table, ld_table, n_table, offsets, ids, dim, out, ld);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_gather_rows(const float* table, long long ld_table, long long n_table, const int32_t* ids, long long n,
                                int dim, float* out, long long ld, void* stream) {
  if (n == 0) return LAFF_OK;
  LAFF_REQUIRE(table && ids && out && n > 0 && n < (1LL << 31) && dim > 0 && ld >= dim && ld_table >= dim, LAFF_EINVAL,
               "laff_gather_rows: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if (n == 0) return LAFF_OK;
  gather_rows_kernel<<<static_cast<unsigned>(n), 128, 0, static_cast<cudaStream_t>(stream)>>>(table, ld_table, n_table, ids, n,
                                                                                             dim, out, ld);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_gru_cell(const float* gi, long long ld_gi, const float* gh, long long ld_gh, const float* h_prev,
                             const int32_t* lengths, int t, int B, int H, float* h_out, float* sum, float* last,
                             void* stream) {
  if (B == 0) return LAFF_OK;
  LAFF_REQUIRE(gi && gh && h_prev && lengths && h_out && B > 0 && H > 0 && t >= 0 && ld_gi >= 3LL * H && ld_gh >= 3LL * H,
               LAFF_EINVAL, "laff_gru_cell: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if (B == 0) return LAFF_OK;
  const long long n = static_cast<long long>(B) * H;
  gru_cell_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      gi, ld_gi, gh, ld_gh, h_prev, lengths, t, B, H, h_out, sum, last);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_gru_cell_backward(const float* gi, long long ld_gi, const float* gh, long long ld_gh, const float* h_prev,
                                      const float* dmean, const float* dlast, const float* dh_gemm, const int32_t* lengths, int t,
                                      int B, int H, float* dh_carry, float* dgi, long long ld_dgi, float* dgh, long long ld_dgh,
                                      void* stream) {
  if (B == 0) return LAFF_OK;
  LAFF_REQUIRE(gi && gh && h_prev && lengths && dh_carry && dgi && dgh && B > 0 && H > 0 && t >= 0 && ld_gi >= 3LL * H &&
                   ld_gh >= 3LL * H && ld_dgi >= 3LL * H && ld_dgh >= 3LL * H,
               LAFF_EINVAL, "laff_gru_cell_backward: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const long long n = static_cast<long long>(B) * H;
  gru_cell_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      gi, ld_gi, gh, ld_gh, h_prev, dmean, dlast, dh_gemm, lengths, t, B, H, dh_carry, dgi, ld_dgi, dgh, ld_dgh);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_scatter_add_rows(const float* dx, long long ld, const int32_t* ids, long long n, int dim, long long n_table,
                                     float* table_grad, long long ld_table, void* stream) {
  if (n == 0) return LAFF_OK;
  LAFF_REQUIRE(dx && ids && table_grad && n > 0 && n < (1LL << 31) && dim > 0 && ld >= dim && ld_table >= dim && n_table > 0,
               LAFF_EINVAL, "laff_scatter_add_rows: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  scatter_add_rows_kernel<<<static_cast<unsigned>(n), 128, 0, static_cast<cudaStream_t>(stream)>>>(dx, ld, ids, n, dim, n_table,
                                                                                                 table_grad, ld_table);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_column_sum(const float* x, long long ld, long long rows, int cols, float* out, void* stream) {
  LAFF_REQUIRE(x && out && rows >= 0 && cols > 0 && ld >= cols, LAFF_EINVAL, "laff_column_sum: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  column_sum_kernel<<<(cols + 31) / 32, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(x, ld, rows, cols, out);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}

extern "C" int laff_mean_over_length(float* x, const int32_t* lengths, int B, int H, void* stream) {
  if (B == 0) return LAFF_OK;
  LAFF_REQUIRE(x && lengths && B > 0 && H > 0, LAFF_EINVAL, "laff_mean_over_length: bad arguments");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if (B == 0) return LAFF_OK;
  const long long n = static_cast<long long>(B) * H;
  scale_rows_by_length_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, lengths, B, H);
  count_launch();
  LAFF_CUDA(cudaGetLastError());
  return LAFF_OK;
}
