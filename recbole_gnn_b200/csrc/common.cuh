// Shared helpers for the b200gcn kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "b200gcn.h"

namespace b200gcn {

// thread-local error text behind b200gcn_last_error()
void set_error(const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)              \
  do {                                         \
    if (!(cond)) {                             \
      ::b200gcn::set_error(__VA_ARGS__);       \
      return B200GCN_ERR_INVALID;              \
    }                                          \
  } while (0)

#define B200_CHECK_CUDA(expr)                                                            \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::b200gcn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                           __FILE__, __LINE__);                                          \
      return B200GCN_ERR_CUDA;                                                           \
    }                                                                                    \
  } while (0)

#define B200_CHECK_LAUNCH() B200_CHECK_CUDA(cudaGetLastError())

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Simple bump allocator over a caller-provided workspace.
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    T* r = reinterpret_cast<T*>(base + off);
    off += align_up(count * sizeof(T));
    return r;
  }
};

inline int bits_for(int64_t n) {  // number of bits needed to represent ids in [0, n)
  int b = 1;
  while ((int64_t(1) << b) < n) ++b;
  return b;
}

// streaming (read-once) loads: keep them out of L1 so the gathered embedding rows own the cache
__device__ __forceinline__ int ld_stream_i32(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_gather_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_f4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// stores that leave the GPU over NVLink: plain weak stores to a peer-mapped address ...
__device__ __forceinline__ void st_peer_f4(float* p, const float4& v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// ... or one multimem store to an NVSwitch multicast address (replicated in the switch)
__device__ __forceinline__ void st_multimem_f4(float* p, const float4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace b200gcn
