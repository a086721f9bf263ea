// common.cuh -- shared helpers of libadmm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace admmb200 {

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const char* fmt, ...);

struct CudaFail {
  cudaError_t err;
};

#define ADMM_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::admmb200::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,  \
                            __LINE__, cudaGetErrorString(_e));                             \
      throw ::admmb200::CudaFail{_e};                                                      \
    }                                                                                      \
  } while (0)

struct ArgFail {
  int code;
};
#define ADMM_REQUIRE(cond, code, ...)           \
  do {                                          \
    if (!(cond)) {                              \
      ::admmb200::set_error(__VA_ARGS__);       \
      throw ::admmb200::ArgFail{code};          \
    }                                           \
  } while (0)

// ---- device helpers -----------------------------------------------------------------------
constexpr int kNumSM = 148;  // B200

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 16-byte load that does not allocate in L1 (matrix data read exactly once)
__device__ __forceinline__ double2 ldg_stream2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double ldg_stream1(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// read-only 16-byte / 8-byte loads through L1 (vectors that are re-read by many warps).  `asm volatile`
// on purpose: a non-volatile asm load (as __ldcs / __ldg are in the CUDA headers) may be hoisted out
// of its bounds guard and executed speculatively past the end of an allocation.
__device__ __forceinline__ double2 ldg_nc2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double ldg_nc1(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// FP64 tensor-core tile: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// lane = 4*g + t:  a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

}  // namespace admmb200
