// common.cuh -- shared helpers for the sm_100a kernels of libbdm_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <utility>

#include "../../include/bdm_b200.h"

#define BDM_CHECK_PTR(p) \
  do {                   \
    if ((p) == nullptr) return BDM_ERR_NULL_POINTER; \
  } while (0)

#define BDM_CHECK_SIZE(cond) \
  do {                       \
    if (!(cond)) return BDM_ERR_BAD_SIZE; \
  } while (0)

// Every launcher ends with this: a launch error becomes the (positive) return code.
#define BDM_RETURN_LAUNCH_STATUS()           \
  do {                                       \
    cudaError_t e__ = cudaGetLastError();    \
    return e__ == cudaSuccess ? BDM_OK : (int)e__; \
  } while (0)

namespace bdm {

constexpr int kWarp = 32;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Number of SMs of the current device (148 on B200); cached per process.
inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Opt a kernel in to `bytes` of dynamic shared memory (> 48 KB needs it) once per (kernel, device);
// cudaFuncSetAttribute is too slow to repeat on every launch and is a per-device setting.
inline cudaError_t ensure_dynamic_smem(const void *func, size_t bytes) {
  if (bytes <= 32 * 1024) return cudaSuccess;  // static + dynamic stays under the default 48 KB cap
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, size_t> configured;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  size_t &have = configured[std::make_pair(func, dev)];
  if (bytes <= have) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

// The reference's squared distance `dx*dx + dy*dy + dz*dz` as nvcc contracts it (SASS of the
// reference: FMUL dy*dy, FFMA dx*dx+t, FFMA dz*dz+t).  Spelled with intrinsics so that the result
// does not depend on this file's optimisation flags: bit-exact integer outputs hinge on it.
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
  float t = __fmul_rn(dy, dy);
  t = __fmaf_rn(dx, dx, t);
  return __fmaf_rn(dz, dz, t);
}

// Streaming 128-bit store / read-only 128-bit load (no L1 allocation for data touched once).
__device__ __forceinline__ void st_stream_f4(float *p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f1(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: one instruction moves a
// whole 4 KB slice row; the issuing thread needs no registers for the data and no address loop.
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // init visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
               "l"(gmem_src), "r"(bytes), "r"(a)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}

// Warp-private staging of `count` (<= TW) consecutive coordinates of three planes (x at src, y at
// src+stride, z at src+2*stride) into shared memory dst[3][TW]; ends with __syncwarp().  All loads of the
// warp are issued before the first use, so the (L2) latency is paid once per tile, not once per step of
// the scan that follows.
template <int TW>
__device__ __forceinline__ void warp_stage_xyz(const float *__restrict__ src, size_t stride, int count,
                                               float *__restrict__ dst, int lane, bool vec4) {
  if (vec4) {  // src + k*stride 16-byte aligned, count % 4 == 0
#pragma unroll
    for (int p = 0; p < 3; ++p)
      for (int q = lane * 4; q < count; q += 128)
        *reinterpret_cast<float4 *>(dst + p * TW + q) = __ldg(reinterpret_cast<const float4 *>(src + p * stride + q));
  } else {
#pragma unroll
    for (int p = 0; p < 3; ++p)
      for (int q = lane; q < count; q += 32) dst[p * TW + q] = __ldg(src + p * stride + q);
  }
  __syncwarp();
}

}  // namespace bdm
