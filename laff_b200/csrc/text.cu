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
