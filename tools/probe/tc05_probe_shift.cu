// tc05_probe_shift.cu -- stand-alone check of the operand addressing csrc/conv3_tc05.cu relies on: the A operand of
// a tcgen05.mma (kind::f16, M=128, K-major, no swizzle) is a window of 128 consecutive rows of a taller
// "chunk-planar" slab in shared memory, [K/8 chunks][NR rows][8 halves]: a core matrix is 8 consecutive rows x 16
// bytes (128 contiguous bytes), SBO = 128, LBO = NR * 16 (not a multiple of 128 in general), and the window is
// selected by a start address that is only 16-byte aligned (base + s * 16 for a row shift s).
// D[m][n] = sum_k A[m + s][k] * B[n][k] is compared with a host product for several shifts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc05_probe_shift tc05_probe_shift.cu && ./tc05_probe_shift
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 64, K = 64, NR = 203;   // slab rows: window + up to 75 rows of shift

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __half *A, const __half *B, float *D, int shift) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem;                               // [K/8][NR][16 bytes]
  unsigned char *sB = smem + ((K / 8) * NR * 16 + 127) / 128 * 128;   // [K/8][N][16 bytes]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NR * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__half *>(sA + ((k / 8) * NR + r) * 16 + (k % 8) * 2) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    *reinterpret_cast<__half *>(sB + ((k / 8) * N + n) * 16 + (k % 8) * 2) = B[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int j = 0; j < K / 16; ++j) {
      const uint64_t da = make_desc(smem_u32(sA) + shift * 16 + 2 * j * NR * 16, NR * 16, 128);
      const uint64_t db = make_desc(smem_u32(sB) + 2 * j * N * 16, N * 16, 128);
      const uint32_t acc = j > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 8; ++c) D[row * N + c0 + c] = __uint_as_float(r[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int main() {
  std::vector<__half> hA(NR * K), hB(N * K);
  std::vector<float> fA(NR * K), fB(N * K), got(M * N);
  srand(11);
  for (int i = 0; i < NR * K; ++i) { hA[i] = __float2half((rand() % 2001 - 1000) / 500.0f); fA[i] = __half2float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { hB[i] = __float2half((rand() % 2001 - 1000) / 500.0f); fB[i] = __half2float(hB[i]); }
  __half *dA, *dB; float *dD;
  cudaMalloc(&dA, NR * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), NR * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  const int smem_bytes = ((K / 8) * NR * 16 + 127) / 128 * 128 + (K / 8) * N * 16;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  int all_ok = 1;
  const int shifts[] = {0, 1, 2, 3, 7, 8, 9, 33, 35, 75};
  for (int s : shifts) {
    cudaMemset(dD, 0xff, M * N * 4);
    probe_kernel<<<1, 128, smem_bytes>>>(dA, dB, dD, s);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift=%d: CUDA error %s\n", s, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(got.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)fA[(m + s) * K + k] * fB[n * K + k];
        double d = fabs((double)got[m * N + n] - ref);
        if (!(d <= 1e30)) d = 1e30;
        if (d > maxerr) maxerr = d;
        if (d > 1e-2) ++bad;
      }
    printf("row shift %2d (start address %% 128 = %3d): max_err=%.3e mismatches=%d/%d %s\n", s, (s * 16) % 128, maxerr, bad, M * N,
           bad == 0 ? "OK" : "WRONG");
    if (bad) all_ok = 0;
  }
  return all_ok ? 0 : 2;
}
