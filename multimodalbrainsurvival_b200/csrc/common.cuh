// Shared host/device helpers for libmmbs (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mmbs.h"

namespace mmbs {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
bool debug_sync();   // MMBS_DEBUG_SYNC=1: synchronise after every launch so that a device fault names its kernel

#define MMBS_CUDA_TRY(expr)                                                        \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      ::mmbs::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
                        cudaGetErrorString(_e));                                   \
      return MMBS_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)

#define MMBS_LAUNCH_CHECK()                                                        \
  do {                                                                             \
    ::mmbs::count_launch();                                                        \
    cudaError_t _e = ::mmbs::debug_sync() ? cudaDeviceSynchronize() : cudaPeekAtLastError(); \
    if (_e != cudaSuccess) {                                                       \
      ::mmbs::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,          \
                        cudaGetErrorString(_e));                                   \
      return MMBS_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)

#define MMBS_REQUIRE(cond, ...)                                                    \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      ::mmbs::set_error(__VA_ARGS__);                                              \
      return MMBS_ERR_ARG;                                                         \
    }                                                                              \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();

// cudaFuncSetAttribute is per DEVICE: a function-local "configured once" flag would leave the kernel unconfigured on
// the second GPU of a process.  One flag per (call site, device); returns true the first time it is asked for the
// current device.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;   // unknown device: always configure
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};

// Bump allocator over the caller's workspace.
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
};

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Order-preserving u32 image of (0 - t): ascending key == descending time; +0/-0 tie.
__device__ __forceinline__ uint32_t time_key(float t) {
  uint32_t b = __float_as_uint(__fsub_rn(0.0f, t));
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
// ---- dropout masks: Philox-4x32-10 keyed by (seed, tag), counter = (column / 8, row).  Every kernel that applies or
// re-applies a mask (the GEMM epilogue, the cast kernels, the MLP backward) derives it from dropout_keep8.
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
  uint32_t c2 = 0x5851f42du, c3 = 0x14057b7eu;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep-mask of the 8 elements [8*q8, 8*q8+8) of row `row`: one Philox call yields eight 16-bit uniforms, element j is
// kept when its uniform is >= thresh16 = round(p * 65536).  Every kernel (forward / backward, scalar / vector)
// derives its mask from this function, so the backward pass regenerates exactly the forward mask.
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint32_t tag, uint32_t row, uint32_t q8,
                                                  uint32_t thresh16) {
  const uint4 r = philox4x32(q8, row, uint32_t(seed) ^ tag, uint32_t(seed >> 32));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t keep = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    keep |= ((w[i] & 0xffffu) >= thresh16 ? 1u : 0u) << (2 * i);
    keep |= ((w[i] >> 16) >= thresh16 ? 1u : 0u) << (2 * i + 1);
  }
  return keep;
}
// keep-mask of the 4 elements [4*q, 4*q+4) of row `row` (q = column / 4): one nibble of the 8-element group
__device__ __forceinline__ uint32_t dropout_keep4(uint64_t seed, uint32_t tag, uint32_t row, uint32_t q,
                                                  uint32_t thresh16) {
  return (dropout_keep8(seed, tag, row, q >> 1, thresh16) >> (4 * (q & 1u))) & 0xfu;
}
__device__ __forceinline__ uint32_t drop_threshold(float p) {
  const float t = p * 65536.0f + 0.5f;
  return t >= 65535.0f ? 65535u : uint32_t(t);
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_f64(double* p, double v) {
  asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
#endif

}  // namespace mmbs
