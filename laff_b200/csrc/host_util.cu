#include "host_util.cuh"

#include <cstdarg>
#include <cstdio>
#include <atomic>
#include <mutex>

namespace laff {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", static_cast<int>(e), cudaGetErrorString(e), file, line, what);
  return static_cast<int>(e);
}

// SM budget of the persistent kernels launched from now on (0 = the whole device): the retrieval pipeline runs the
// similarity sweep on all but a couple of SMs and the next step's query fusion / collectives' neighbours on those, so
// that the two overlap instead of queueing behind each other (laff_b200/retrieval.py::Retriever.submit).
static std::atomic<int> g_sm_limit{0};

int get_device_info(DeviceInfo* info) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    return LAFF_ENODEV;
  }
  int major = 0, sms = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    set_error("cannot query device %d", dev);
    return LAFF_ENODEV;
  }
  if (major != 10) {
    set_error("laff_b200 needs an sm_100-class GPU (B200); device %d is sm_%d x", dev, major);
    return LAFF_ENODEV;
  }
  info->device = dev;
  info->sms_total = sms;
  const int limit = g_sm_limit.load();
  info->sms = (limit > 0 && limit < sms) ? limit : sms;
  info->cc_major = major;
  return LAFF_OK;
}

int sm_limit() { return g_sm_limit.load(); }

// Measured on B200 (profiles/): single-tile units in n-major order with ~10 query row-tiles per group keep the query
// block L2-resident and let the CTA pairs that share a gallery tile run within a fraction of a tile of each other.
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count(bool reset) {
  return reset ? g_launches.exchange(0, std::memory_order_relaxed) : g_launches.load(std::memory_order_relaxed);
}

static Tuning g_tuning = {2, 1, 10};
static std::mutex g_tuning_mu;

Tuning get_tuning() {
  std::lock_guard<std::mutex> lk(g_tuning_mu);
  return g_tuning;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* tm, const void* base, int dtype, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                 uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  LAFF_REQUIRE(fn != nullptr, LAFF_ENODEV, "cuTensorMapEncodeTiled not available from the driver");
  LAFF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, LAFF_EINVAL, "operand base %p not 16-byte aligned", base);
  LAFF_REQUIRE((pitch_elems * 2) % 16 == 0, LAFF_EINVAL, "operand row pitch %llu elements is not a multiple of 8",
               (unsigned long long)pitch_elems);
  LAFF_REQUIRE(box_rows >= 1 && box_rows <= 256, LAFF_EINVAL, "bad TMA box rows %u", box_rows);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = dtype == LAFF_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = fn(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LAFF_REQUIRE(r == CUDA_SUCCESS, LAFF_EINVAL, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu pitch=%llu",
               (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_elems);
  return LAFF_OK;
}

}  // namespace laff

namespace laff { long long launch_count(bool reset); }

extern "C" {

const char* laff_last_error(void) { return laff::g_err; }

int laff_abi_version(void) { return 1; }

long long laff_launch_count(int reset) { return laff::launch_count(reset != 0); }

int laff_set_tuning(int cta_group, int chunk_tiles, int m_group) {
  if ((cta_group != 0 && cta_group != 1 && cta_group != 2) || chunk_tiles < 0 || m_group < 0) {
    laff::set_error("laff_set_tuning: bad value");
    return LAFF_EINVAL;
  }
  std::lock_guard<std::mutex> lk(laff::g_tuning_mu);
  if (cta_group) laff::g_tuning.cta_group = cta_group;
  if (chunk_tiles) laff::g_tuning.chunk_tiles = chunk_tiles;
  if (m_group) laff::g_tuning.m_group = m_group;
  return LAFF_OK;
}

int laff_get_tuning(int* cta_group, int* chunk_tiles, int* m_group) {
  laff::Tuning t = laff::get_tuning();
  if (cta_group) *cta_group = t.cta_group;
  if (chunk_tiles) *chunk_tiles = t.chunk_tiles;
  if (m_group) *m_group = t.m_group;
  return LAFF_OK;
}

}  // extern "C"

extern "C" int laff_set_sm_limit(int sms) {
  if (sms < 0) sms = 0;
  return laff::g_sm_limit.exchange(sms);
}
