// Host-side helpers shared by the C-ABI translation units: error reporting, device checks, TMA descriptor creation.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/laff_b200.h"

namespace laff {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

struct DeviceInfo {
  int device;
  int sms;        // SMs the caller may fill: the device's count, or the budget set with laff_set_sm_limit
  int sms_total;  // the device's SM count
  int cc_major;
};
int sm_limit();
// Fails (LAFF_ENODEV) unless the current device is sm_100-class: there is no fallback path.
int get_device_info(DeviceInfo* info);

struct Tuning {
  int cta_group;
  int chunk_tiles;
  int m_group;
};
Tuning get_tuning();

// 2-D K-major tensor map: global [rows, cols] 16-bit elements with row pitch `pitch_elems`, box = 64 x box_rows,
// SWIZZLE_128B, out-of-bounds elements read as zero.
int make_tmap_2d(CUtensorMap* tm, const void* base, int dtype, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                 uint32_t box_rows);

// Number of kernels this library has launched (bench.py reports it as gpu_launches).
// (declared with its default argument in gemm_engine.cuh when that header is included first)
#ifndef LAFF_COUNT_LAUNCH_DECLARED
void count_launch(int n = 1);
#endif

inline bool is16(int dtype) { return dtype == LAFF_F16 || dtype == LAFF_BF16; }

}  // namespace laff

#define LAFF_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t laff_e_ = (expr);                                                    \
    if (laff_e_ != cudaSuccess) return laff::cuda_fail(laff_e_, #expr, __FILE__, __LINE__); \
  } while (0)

#define LAFF_REQUIRE(cond, code, ...)   \
  do {                                  \
    if (!(cond)) {                      \
      laff::set_error(__VA_ARGS__);     \
      return (code);                    \
    }                                   \
  } while (0)
