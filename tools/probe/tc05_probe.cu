// tc05_probe.cu -- stand-alone check of the tcgen05 operand descriptors used by csrc/attention_tc05.cu.
// D[128 x 64] (fp32, TMEM) = A[128 x 64] . B[64 x 64]^T with fp16 operands staged in shared memory in the
// no-swizzle canonical core-matrix layouts (8 x 16-byte core matrices), for both K-major and MN-major
// operands and for both assignments of the descriptor's two byte offsets.  Prints the max error of each
// variant against a host fp32 product of the same fp16-rounded inputs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc05_probe tc05_probe.cu && ./tc05_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// a_mn / b_mn: 1 = MN-major operand, 0 = K-major.  swap: exchange the two byte offsets in the descriptors.
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __half *A, const __half *B, float *D, int a_mn, int b_mn, int swap, int a_tmem) {
  __shared__ __align__(1024) unsigned char sA[M * K * 2];
  __shared__ __align__(1024) unsigned char sB[N * K * 2];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // core matrix (mn8 = mn/8, k8 = k/8) at (k8 * (MN/8) + mn8) * 128 bytes.
  // K-major core: row = mn % 8 (16 bytes per row), 8 k-elements per row.  MN-major core: row = k % 8, 8 mn-elements per row.
  for (int i = tid; i < M * K; i += 128) {
    const int m = i / K, k = i % K;
    const int core = (k / 8) * (M / 8) + m / 8;
    const int off = a_mn ? core * 128 + (k % 8) * 16 + (m % 8) * 2 : core * 128 + (m % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<__half *>(sA + off) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    const int core = (k / 8) * (N / 8) + n / 8;
    const int off = b_mn ? core * 128 + (k % 8) * 16 + (n % 8) * 2 : core * 128 + (n % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<__half *>(sB + off) = B[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  if (a_tmem) {
    // A operand in tensor memory: lane = row, 32-bit column c = (A[row][2c], A[row][2c+1]); columns 64..95
    uint32_t ar[32];
    const int row = warp * 32 + lane;
    for (int c2 = 0; c2 < 32; ++c2) {
      const __half2 h = __halves2half2(A[row * K + 2 * c2], A[row * K + 2 * c2 + 1]);
      ar[c2] = *reinterpret_cast<const uint32_t *>(&h);
    }
    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + 64;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(ta), "r"(ar[0]), "r"(ar[1]), "r"(ar[2]), "r"(ar[3]), "r"(ar[4]), "r"(ar[5]), "r"(ar[6]), "r"(ar[7]),
          "r"(ar[8]), "r"(ar[9]), "r"(ar[10]), "r"(ar[11]), "r"(ar[12]), "r"(ar[13]), "r"(ar[14]), "r"(ar[15]),
          "r"(ar[16]), "r"(ar[17]), "r"(ar[18]), "r"(ar[19]), "r"(ar[20]), "r"(ar[21]), "r"(ar[22]), "r"(ar[23]),
          "r"(ar[24]), "r"(ar[25]), "r"(ar[26]), "r"(ar[27]), "r"(ar[28]), "r"(ar[29]), "r"(ar[30]), "r"(ar[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    // byte offset between core matrices adjacent in K (k8 -> k8+1) and adjacent in MN (mn8 -> mn8+1)
    const uint32_t a_kstride = (M / 8) * 128, a_mnstride = 128;
    const uint32_t b_kstride = (N / 8) * 128, b_mnstride = 128;
    // K-major: LBO = k stride, SBO = mn stride.  MN-major (no swizzle): LBO = k stride (between 8-k groups), SBO = mn stride.
    uint32_t a_lbo = a_kstride, a_sbo = a_mnstride, b_lbo = b_kstride, b_sbo = b_mnstride;
    if (swap) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
    const uint32_t idesc = (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(M >> 4) << 24);
    for (int j = 0; j < K / 16; ++j) {
      const uint64_t da = make_desc(smem_u32(sA) + 2 * j * a_kstride, a_lbo, a_sbo);
      const uint64_t db = make_desc(smem_u32(sB) + 2 * j * b_kstride, b_lbo, b_sbo);
      const uint32_t acc = j > 0;
      if (a_tmem) {
        const uint32_t ta = tmem + 64 + j * 8;     // 16 fp16 of K = 8 columns
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
            ::"r"(tmem), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
      } else
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[64];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
        "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int row = warp * 32 + lane;
  for (int c = 0; c < N; ++c) D[row * N + c] = __uint_as_float(r[c]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main(int argc, char **argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;   // variant id = a_tmem*8 + a_mn*4 + b_mn*2 + swap (one per process: a fault is sticky)
  std::vector<__half> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K), ref(M * N), got(M * N);
  srand(7);
  for (int i = 0; i < M * K; ++i) { hA[i] = __float2half((rand() % 2001 - 1000) / 500.0f); fA[i] = __half2float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { hB[i] = __float2half((rand() % 2001 - 1000) / 500.0f); fB[i] = __half2float(hB[i]); }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k];
      ref[m * N + n] = (float)s;
    }
  __half *dA, *dB; float *dD;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  int ok_any = 0;
  for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
  for (int a_mn = 0; a_mn < 2; ++a_mn)
    for (int b_mn = 0; b_mn < 2; ++b_mn)
      for (int swap = 0; swap < 2; ++swap) {
        if (only >= 0 && only != a_tmem * 8 + a_mn * 4 + b_mn * 2 + swap) continue;
        if (a_tmem && (a_mn || swap)) continue;
        cudaMemset(dD, 0xff, M * N * 4);
        probe_kernel<<<1, 128>>>(dA, dB, dD, a_mn, b_mn, swap, a_tmem);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("a_mn=%d b_mn=%d swap=%d: CUDA error %s\n", a_mn, b_mn, swap, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(got.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int i = 0; i < M * N; ++i) {
          double d = fabs((double)got[i] - ref[i]);
          if (!(d <= 1e30)) d = 1e30;
          if (d > maxerr) maxerr = d;
          if (d > 1e-2) ++bad;
        }
        printf("a_tmem=%d a_mn=%d b_mn=%d swap=%d: max_err=%.3e mismatches=%d/%d %s\n", a_tmem, a_mn, b_mn, swap, maxerr, bad, M * N,
               bad == 0 ? "OK" : "WRONG");
        if (bad == 0) ok_any = 1;
      }
  return ok_any ? 0 : 2;
}
