// tc05_probe_tf32.cu -- stand-alone check of tcgen05.mma kind::tf32 with MN-major operands in the no-swizzle
// canonical layout (core matrix = 8 K-rows x 16 bytes = 4 MN-contiguous tf32), as csrc/tap_gemm.cu uses them.
// D[128 x 64] = A[128 x 32] . B[64 x 32]^T, A and B given M-/N-contiguous (i.e. stored [K][M] and [K][N]).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc05_probe_tf32 tc05_probe_tf32.cu && ./tc05_probe_tf32
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 64, K = 32;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const float *A, const float *B, float *D, int swap, int a_mn, int b_mn) {
  __shared__ __align__(1024) unsigned char sA[M * K * 4];
  __shared__ __align__(1024) unsigned char sB[N * K * 4];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A given as [K][M] (M contiguous), B as [K][N].  MN-major core: 8 k-rows x 16 bytes (4 mn); K-major core: 8 mn-rows x 16 bytes (4 k)
  for (int i = tid; i < K * M; i += 128) {
    const int k = i / M, m = i % M;
    const int off = a_mn ? ((k / 8) * (M / 4) + m / 4) * 128 + (k % 8) * 16 + (m % 4) * 4
                         : ((k / 4) * (M / 8) + m / 8) * 128 + (m % 8) * 16 + (k % 4) * 4;
    *reinterpret_cast<float *>(sA + off) = A[i];
  }
  for (int i = tid; i < K * N; i += 128) {
    const int k = i / N, n = i % N;
    const int off = b_mn ? ((k / 8) * (N / 4) + n / 4) * 128 + (k % 8) * 16 + (n % 4) * 4
                         : ((k / 4) * (N / 8) + n / 8) * 128 + (n % 8) * 16 + (k % 4) * 4;
    *reinterpret_cast<float *>(sB + off) = B[i];
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
    // byte strides between core matrices: along K (LBO) and along M/N (SBO)
    uint32_t a_lbo = a_mn ? (M / 4) * 128 : (M / 8) * 128, a_sbo = 128, b_lbo = b_mn ? (N / 4) * 128 : (N / 8) * 128, b_sbo = 128;
    if (swap) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int j = 0; j < K / 8; ++j) {   // K = 8 per instruction: one 8-k group (MN-major) or two 4-k cores (K-major)
      const uint64_t da = make_desc(smem_u32(sA) + (a_mn ? j * (M / 4) * 128 : 2 * j * (M / 8) * 128), a_lbo, a_sbo);
      const uint64_t db = make_desc(smem_u32(sB) + (b_mn ? j * (N / 4) * 128 : 2 * j * (N / 8) * 128), b_lbo, b_sbo);
      const uint32_t acc = j > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 8; ++c) D[row * N + c0 + c] = __uint_as_float(r[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int main(int argc, char **argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;   // a_mn*4 + b_mn*2 + swap
  const int swap = variant & 1, a_mn = (variant >> 2) & 1, b_mn = (variant >> 1) & 1;
  std::vector<float> hA(K * M), hB(K * N), ref(M * N), got(M * N);
  srand(11);
  for (auto &v : hA) v = (rand() % 2001 - 1000) / 500.0f;
  for (auto &v : hB) v = (rand() % 2001 - 1000) / 500.0f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)hA[k * M + m] * hB[k * N + n];
      ref[m * N + n] = (float)s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, K * M * 4); cudaMalloc(&dB, K * N * 4); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), K * M * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), K * N * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, M * N * 4);
  probe_kernel<<<1, 128>>>(dA, dB, dD, swap, a_mn, b_mn);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("tf32 variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(got.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int i = 0; i < M * N; ++i) {
    double d = fabs((double)got[i] - ref[i]);
    if (!(d <= 1e30)) d = 1e30;
    if (d > maxerr) maxerr = d;
    if (d > 5e-2) ++bad;
  }
  printf("kind::tf32 a_mn=%d b_mn=%d swap=%d: max_err=%.3e (tf32 rounding ~1e-3 expected) mismatches=%d/%d %s\n", a_mn, b_mn, swap, maxerr, bad,
         M * N, bad == 0 ? "OK" : "WRONG");
  return bad == 0 ? 0 : 2;
}
