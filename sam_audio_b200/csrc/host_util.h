// Host-side helpers: error plumbing, device arena, TMA tensor-map construction, GEMM op records.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <vector>
#include <stdexcept>

#include "gemm_tc.cuh"

namespace sab {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline std::string fmt(const char* f, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof(buf), f, ap);
  va_end(ap);
  return std::string(buf);
}

#define SAB_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw sab::Error(sab::fmt("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__)); \
  } while (0)

#define SAB_CHECK(cond, ...)                                                     \
  do {                                                                           \
    if (!(cond)) throw sab::Error(sab::fmt(__VA_ARGS__) + sab::fmt(" [%s:%d]", __FILE__, __LINE__)); \
  } while (0)

// ---- owned device allocations (freed with the engine / plan) ----
struct DevicePool {
  std::vector<void*> ptrs;
  int64_t bytes = 0;
  template <typename T>
  T* alloc(int64_t n, bool zero = false) {
    void* p = nullptr;
    const int64_t nb = ((n * (int64_t)sizeof(T) + 255) / 256) * 256;
    SAB_CUDA(cudaMalloc(&p, nb > 0 ? nb : 256));
    if (zero) SAB_CUDA(cudaMemset(p, 0, nb > 0 ? nb : 256));
    ptrs.push_back(p);
    bytes += nb;
    return reinterpret_cast<T*>(p);
  }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
    bytes = 0;
  }
  ~DevicePool() { release(); }
};

// ---- TMA descriptors ----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    SAB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    SAB_CHECK(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// bf16 tensor viewed as (cols, rows, items): element (c, r, i) at base + i*item_pitch + r*row_pitch + c.
// box = (box_cols, box_rows, 1); swizzle = box_cols*2 bytes (128 or 64).
inline CUtensorMap make_tmap_3d(const void* base, int64_t cols, int64_t rows, int64_t items, int64_t row_pitch_elems,
                                int64_t item_pitch_elems, int box_cols, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)items};
  cuuint64_t strides[2] = {(cuuint64_t)row_pitch_elems * 2, (cuuint64_t)item_pitch_elems * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  SAB_CHECK(box_cols * 2 == 128 || box_cols * 2 == 64, "box_cols must be 64 or 32 bf16");
  SAB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  SAB_CHECK((strides[0] % 16) == 0 && (strides[1] % 16) == 0, "TMA strides must be multiples of 16B (%lld, %lld)",
            (long long)strides[0], (long long)strides[1]);
  if (items == 1) strides[1] = strides[0] * (cuuint64_t)rows;  // any valid value
  CUtensorMapSwizzle sw = (box_cols * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SAB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed: %d (cols=%lld rows=%lld items=%lld pitch=%lld/%lld box=%d,%d)",
            (int)r, (long long)cols, (long long)rows, (long long)items, (long long)row_pitch_elems,
            (long long)item_pitch_elems, box_cols, box_rows);
  return m;
}

inline CUtensorMap make_tmap_2d(const void* base, int64_t cols, int64_t rows, int64_t row_pitch_elems, int box_cols,
                                int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  SAB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  SAB_CHECK((strides[0] % 16) == 0, "TMA stride must be a multiple of 16B");
  CUtensorMapSwizzle sw = (box_cols * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SAB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: %d (cols=%lld rows=%lld pitch=%lld box=%d,%d)", (int)r,
            (long long)cols, (long long)rows, (long long)row_pitch_elems, box_cols, box_rows);
  return m;
}

// ---- a planned GEMM launch ----
struct GemmOp {
  CUtensorMap tmA, tmB, tmW;   // tmW: resident second weight of a back-to-back pair (b2b)
  GemmParams P;
  int BN = 0, BK = 64, mode = EPI_AFFINE, cg = 1;
  bool b2b = false;
  int grid = 0;
  const char* tag = "";
  double flops = 0;  // algorithmic 2*M*N*K
  double in_bytes = 0;   // algorithmic operand bytes: every A element once + the packed weight once
  double rows = 0;       // output rows (for the epilogue's output / residual bytes)
};

}  // namespace sab
