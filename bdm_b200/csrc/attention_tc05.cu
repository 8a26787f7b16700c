// attention_tc05.cu -- the PVConv attention block (reference: experiments/model/pvcnn/modules/pvconv.py:36-63,
// class Attention: softmax(q^T k) applied to v, un-scaled logits, fp32) on Blackwell's 5th-generation tensor
// cores: tcgen05.mma with accumulators AND the A operands in tensor memory (TMEM), B operands fetched by TMA.
//
// Arithmetic is the one csrc/attention.cu established (and tests/test_dense_fused_gpu.py checks against
// float64): every fp32 operand is split into two fp16 numbers after a per-tensor power-of-two scaling
// (x*s = hi + lo), each product is three tensor-core products  lo*hi + hi*lo + hi*hi  accumulated in fp32,
// the softmax is online (running max / sum per query, fp32), and the P.V product of every key tile starts
// from zero and is merged into the running output with one rounded FMA.
//
// Kernels:
//   attention_prep_kernel   k, v f32[B,64,T] -> fp16 hi/lo planes laid out tile by tile exactly as the MMA reads
//                           its B operand from shared memory (8x8 "core matrices", no swizzle), so that a tile
//                           arrives in shared memory with ONE 1-D bulk copy (cp.async.bulk + mbarrier).
//   attention_tc05_kernel   CTA = NQ tiles of 128 queries x all keys in tiles of 64.
//       warps 0..4NQ-1      softmax: warp-group w owns query tile w, thread = one query row = one TMEM lane.
//                           Stores its Q row (fp16 hi/lo) into TMEM once; per key tile reads its S row with
//                           tcgen05.ld, exponentiates, stores P (fp16 hi/lo) back into TMEM with tcgen05.st as
//                           the A operand of the second GEMM; keeps the output row O[64] and the running
//                           max / sum in registers.
//       warp 4NQ            one thread: TMA producer (K ring and V ring of kStages tiles)
//       warps 4NQ+1+w       one thread each: issues the tcgen05.mma of query tile w and commits them to mbarriers
//   S = Q K^T   : M=128 queries, N=64 keys, K=64 channels; A = Q in TMEM, B = K tile in smem (MN-major, as stored)
//   O_tile = P V: M=128 queries, N=64 channels, K=64 keys; A = P in TMEM, B = V tile in smem (K-major, as stored)
//   TMEM per query tile (256 columns): S 64 | O_tile 64 | Q hi 32 | Q lo 32 | P hi 32 | P lo 32.
// Why the A operands live in TMEM: with both operands in shared memory an M=128, N=64, K=16 MMA reads 6 KB per 32
// tensor-pipe cycles -- 192 B/clk against a 128 B/clk shared-memory port; that version (still correct) ran at
// 341 us for B=16, T=4096.  With A in TMEM only the 2 KB B slice comes from shared memory.
// Operand layouts / descriptor conventions were verified on the device with tools/probe/tc05_probe.cu.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace bdm {
namespace tc05 {

constexpr int kD = 64;            // channels
constexpr int kQT = 128;          // queries per tile (UMMA M)
constexpr int kKT = 64;           // keys per tile (UMMA N of S, K of P.V)
constexpr int kStages = 4;        // K ring and V ring depth
constexpr int kKTileBytes = 2 * kKT * kD * 2;   // hi + lo planes of one key tile (16 KB), same for V

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrive on `bar` once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory operand descriptor, no swizzle: 8-row x 16-byte core matrices; lbo = byte distance between
// core matrices adjacent along K, sbo = along M/N
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16, fp16 inputs, fp32 accumulate, M=128, N=64; operand majors: 1 = MN-major, 0 = K-major
__host__ __device__ constexpr uint32_t instr_desc(uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((uint32_t)(kKT >> 3) << 17) | ((uint32_t)(kQT >> 4) << 24);
}

#define BDM_TMEM_LD32(r, taddr)                                                                                    \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                    \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                    \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),            \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),      \
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),    \
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])     \
      : "r"(taddr))

#define BDM_TMEM_ST32(taddr, r)                                                                                    \
  asm volatile(                                                                                                    \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                              \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                                   \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"                           \
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),        \
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),              \
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),            \
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])             \
      : "memory")

__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// (hi, lo) fp16 pairs of two fp32 values: hi = rn(x), lo = rn(x - hi)
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo) {
  const __half2 hh = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(hh);
  const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t *>(&hh);
  lo = *reinterpret_cast<const uint32_t *>(&ll);
}

__device__ __forceinline__ float fast_exp2(float x) {   // one MUFU.EX2; ex2(-inf) = +0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// power of two s with max|x| * s in [512, 1024); 1/s in *inv
__device__ __forceinline__ float pow2_scale(float amax, float *inv) {
  int e = 9;   // amax == 0 / non-finite: scale 1
  if (amax > 0.0f && amax < INFINITY) e = (int)((__float_as_uint(amax) >> 23) & 255u) - 127;
  const int se = min(max(9 - e, -60), 60);
  *inv = __uint_as_float((uint32_t)(127 - se) << 23);
  return __uint_as_float((uint32_t)(127 + se) << 23);
}

__device__ __forceinline__ void split8(const float (&x)[8], float s, uint4 &hi, uint4 &lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = x[2 * i] * s, b = x[2 * i + 1] * s;
    const __half2 hh = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
    h[i] = *reinterpret_cast<const uint32_t *>(&hh);
    l[i] = *reinterpret_cast<const uint32_t *>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// max |x| of three equally sized tensors -> amax[0..2] (bit pattern of a non-negative float; zeroed by the host)
__global__ void attention_tc05_amax_kernel(size_t n4, const float4 *__restrict__ q, const float4 *__restrict__ k,
                                           const float4 *__restrict__ v, unsigned *__restrict__ amax) {
  float m[3] = {0.0f, 0.0f, 0.0f};
  const float4 *src[3] = {q, k, v};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      const float4 x = __ldg(src[w] + i);
      m[w] = fmaxf(m[w], fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
    }
  }
#pragma unroll
  for (int w = 0; w < 3; ++w) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m[w] = fmaxf(m[w], __shfl_xor_sync(0xffffffffu, m[w], d));
    if ((threadIdx.x & 31) == 0) atomicMax(amax + w, __float_as_uint(m[w]));
  }
}

// the same for the fused projection layout qkv f32[rows][ld] (q | k | v in columns [0,c) [c,2c) [2c,3c), + bias)
__global__ void attention_tc05_amax_qkv_kernel(size_t rows, int ld, const float *__restrict__ qkv,
                                               const float *__restrict__ bias, unsigned *__restrict__ amax) {
  float m[3] = {0.0f, 0.0f, 0.0f};
  const int c4 = 3 * kD / 4;   // float4 units per row that belong to q, k, v
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * c4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / c4;
    const int col = (int)(i - r * c4) * 4;
    float4 x = __ldg(reinterpret_cast<const float4 *>(qkv + r * ld + col));
    if (bias != nullptr) {
      const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + col));
      x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
    }
    const float a = fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w)));
    const int w = col / kD;
    if (w == 0) m[0] = fmaxf(m[0], a); else if (w == 1) m[1] = fmaxf(m[1], a); else m[2] = fmaxf(m[2], a);
  }
#pragma unroll
  for (int w = 0; w < 3; ++w) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m[w] = fmaxf(m[w], __shfl_xor_sync(0xffffffffu, m[w], d));
    if ((threadIdx.x & 31) == 0) atomicMax(amax + w, __float_as_uint(m[w]));
  }
}

// Workspace (bytes): [0,16) amax | [256, ...) K planes | V planes, each b * t * 64 * 2 (hi, lo) * 2 bytes.
//   K tile (64 keys): plane hi then lo; element (key j, channel d) of a plane at half index
//       ((d/8) * 8 + j/8) * 64 + (d%8) * 8 + j%8          MN-major core matrices (8 channels x 8 keys)
//   V tile (64 keys):     ((j/8) * 8 + d/8) * 64 + (d%8) * 8 + j%8          K-major  (8 channels x 8 keys)
// Block = 256 threads = 64 channels x 4 runs of 8 consecutive tokens; 8 lanes with consecutive channels write one
// contiguous 128-byte core matrix, and read 8 x 32-byte sectors.
// TM (token-major): k, v point at the first k / v column of the fused projection qkv f32[b][T][ld] and `kbias` /
// `vbias` (or NULL) are added on the way; otherwise k, v are f32[b][64][T] (ld, biases unused).
template <bool TM>
__global__ void __launch_bounds__(256)
attention_prep_kernel(int T, int ld, const float *__restrict__ k, const float *__restrict__ v,
                      const float *__restrict__ kbias, const float *__restrict__ vbias,
                      const unsigned *__restrict__ amax, __half *__restrict__ kp, __half *__restrict__ vp) {
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int d = (tid >> 5) * 8 + (tid & 7);
  const int t0 = blockIdx.x * 32 + ((tid >> 3) & 3) * 8;
  float inv;
  const float sk = pow2_scale(__uint_as_float(__ldg(amax + 1)), &inv);
  const float sv = pow2_scale(__uint_as_float(__ldg(amax + 2)), &inv);
  const size_t src = TM ? ((size_t)b * T + t0) * ld + d : ((size_t)b * kD + d) * T + t0;
  float x[8];
  uint4 hi, lo;
  auto load8 = [&](const float *p, const float *bias) {
    if (TM) {   // 8 tokens of one channel: 8 lanes with consecutive channels share a 32-byte sector per token
      const float bb = bias != nullptr ? __ldg(bias + d) : 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = __ldg(p + src + (size_t)i * ld) + bb;
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4 *>(p + src)), c = __ldg(reinterpret_cast<const float4 *>(p + src + 4));
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = c.x; x[5] = c.y; x[6] = c.z; x[7] = c.w;
    }
  };
  const int dg = d >> 3, dr = d & 7;
  const int tile = t0 / kKT, j = t0 % kKT;
  {   // K
    load8(k, kbias);
    split8(x, sk, hi, lo);
    __half *base = kp + ((size_t)b * (T / kKT) + tile) * (2 * kKT * kD) + ((dg * 8 + j / 8) * 64 + dr * 8);
    *reinterpret_cast<uint4 *>(base) = hi;
    *reinterpret_cast<uint4 *>(base + kKT * kD) = lo;
  }
  {   // V
    load8(v, vbias);
    split8(x, sv, hi, lo);
    __half *base = vp + ((size_t)b * (T / kKT) + tile) * (2 * kKT * kD) + (((j / 8) * 8 + dg) * 64 + dr * 8);
    *reinterpret_cast<uint4 *>(base) = hi;
    *reinterpret_cast<uint4 *>(base + kKT * kD) = lo;
  }
}

// TMEM columns of one query tile
constexpr uint32_t kColS = 0, kColO = 64, kColQhi = 128, kColQlo = 160, kColPhi = 192, kColPlo = 224, kColsPerTile = 256;

template <int NQ>
struct Smem {
  static constexpr int kK = 0;
  static constexpr int kV = kK + kStages * kKTileBytes;
  static constexpr int kBars = kV + kStages * kKTileBytes;
  // barriers: k_full[S] k_empty[S] v_full[S] v_empty[S] | per w: q_ready s_full s_empty p_full o_full
  static constexpr int kPerTile = 5;
  static constexpr int kNumBars = 4 * kStages + NQ * kPerTile;
  static constexpr int kTmemSlot = kBars + kNumBars * 8;
  static constexpr int kBytes = kTmemSlot + 16;
};

// TM (token-major): q = fused projection f32[b][T][ld] (q in columns [0,64), + qbias), out f32[b][T][64];
// otherwise q, out f32[b][64][T].
template <int NQ, bool TM>
__global__ void __launch_bounds__((4 * NQ + 4) * 32, 1)
attention_tc05_kernel(int T, int ld, const float *__restrict__ q, const float *__restrict__ qbias,
                      const __half *__restrict__ kp, const __half *__restrict__ vp,
                      const unsigned *__restrict__ amax, float *__restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  using L = Smem<NQ>;
  constexpr uint32_t kTmemCols = NQ * kColsPerTile;
  const int b = blockIdx.y;
  const int qtile0 = blockIdx.x * NQ;          // first 128-query tile of this CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int J = T / kKT;

  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::kBars);
  uint64_t *k_full = bars, *k_empty = k_full + kStages, *v_full = k_empty + kStages, *v_empty = v_full + kStages;
  uint64_t *wbars = v_empty + kStages;           // + w * kPerTile: q_ready, s_full, s_empty, p_full, o_full
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::kTmemSlot);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(k_full + s, 1); mbar_init(v_full + s, 1);          // TMA producer (expect_tx)
      mbar_init(k_empty + s, NQ); mbar_init(v_empty + s, NQ);      // one tcgen05.commit per MMA thread
    }
    for (int w = 0; w < NQ; ++w) {
      uint64_t *wb = wbars + w * L::kPerTile;
      mbar_init(wb + 0, kQT);     // q_ready: every softmax thread has stored its Q row into TMEM
      mbar_init(wb + 1, 1);       // s_full   (tcgen05.commit)
      mbar_init(wb + 2, kQT);     // s_empty: every softmax thread has its S row in registers
      mbar_init(wb + 3, kQT);     // p_full:  every softmax thread has stored its P row into TMEM
      mbar_init(wb + 4, 1);       // o_full   (tcgen05.commit)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4 * NQ + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // Register budget: the kernel is compiled for (4NQ+4) warps sharing the register file evenly; the control
  // warp-group (TMA producer, MMA issuers, one idle warp) hands most of its share to the softmax warp-groups,
  // whose threads hold a 64-value output row and a 64-value S row each.
  // (NQ == 2: 12 warps x 168 registers at launch -> control warp-group 4 x 80, softmax warp-groups 8 x 208;
  // each setmaxnreg sits at the top of the branch whose code it governs)
  if (warp == 4 * NQ) {
    // ===================== TMA producer =====================
    if (NQ == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (lane == 0) {
      const __half *ksrc = kp + (size_t)b * J * (2 * kKT * kD);
      const __half *vsrc = vp + (size_t)b * J * (2 * kKT * kD);
      for (int j = 0; j < J; ++j) {
        const int s = j % kStages;
        const uint32_t ph = ((j / kStages) & 1) ^ 1;
        mbar_wait(k_empty + s, ph);
        mbar_expect_tx(k_full + s, kKTileBytes);
        tma_load(smem + L::kK + s * kKTileBytes, ksrc + (size_t)j * (2 * kKT * kD), kKTileBytes, k_full + s);
        mbar_wait(v_empty + s, ph);
        mbar_expect_tx(v_full + s, kKTileBytes);
        tma_load(smem + L::kV + s * kKTileBytes, vsrc + (size_t)j * (2 * kKT * kD), kKTileBytes, v_full + s);
      }
    }
  } else if (warp > 4 * NQ) {
    // ===================== MMA issuer of query tile w =====================
    const int w = warp - (4 * NQ + 1);
    if (NQ == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (lane == 0 && w < NQ) {
      constexpr uint32_t idesc_s = instr_desc(0, 1);     // A = Q (TMEM), B = K (MN-major)
      constexpr uint32_t idesc_o = instr_desc(0, 0);     // A = P (TMEM), B = V (K-major)
      const uint32_t k_base = smem_u32(smem + L::kK), v_base = smem_u32(smem + L::kV);
      uint64_t *wb = wbars + w * L::kPerTile;
      const uint32_t tw = tmem + w * kColsPerTile;
      mbar_wait(wb + 0, 0);                        // Q is in TMEM
      tc_fence_after();
      for (int j = 0; j <= J; ++j) {
        if (j < J) {
          const int s = j % kStages;
          mbar_wait(k_full + s, (j / kStages) & 1);
          mbar_wait(wb + 2, (j & 1) ^ 1);          // the softmax threads have taken S_{j-1} out of TMEM
          tc_fence_after();
          const uint32_t kb = k_base + s * kKTileBytes;
          // terms in ascending magnitude: lo*hi, hi*lo, hi*hi; 4 k-steps of 16 channels (8 TMEM columns of A)
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint32_t qa = tw + (term == 0 ? kColQlo : kColQhi);
            const uint32_t ka = kb + (term == 1 ? kKT * kD * 2 : 0);            // K lo plane for term 1
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16_ts(tw + kColS, qa + kk * 8, smem_desc(ka + kk * 2048, 1024, 128), idesc_s, (term | kk) != 0);
          }
          umma_commit(wb + 1);                     // s_full
          umma_commit(k_empty + s);
        }
        if (j >= 1) {
          const int jj = j - 1, s = jj % kStages;
          mbar_wait(v_full + s, (jj / kStages) & 1);
          mbar_wait(wb + 3, jj & 1);               // P_jj is in TMEM (and O_tile_{jj-1} has been taken out)
          tc_fence_after();
          const uint32_t vb = v_base + s * kKTileBytes;
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint32_t pa = tw + (term == 0 ? kColPlo : kColPhi);
            const uint32_t va = vb + (term == 1 ? kKT * kD * 2 : 0);            // V lo plane for term 1
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16_ts(tw + kColO, pa + kk * 8, smem_desc(va + kk * 2048, 1024, 128), idesc_o, (term | kk) != 0);
          }
          umma_commit(wb + 4);                     // o_full
          umma_commit(v_empty + s);
        }
      }
    }
  } else {
    // ===================== softmax warp-groups =====================
    if (NQ == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int w = warp >> 2;                        // query tile of this warp-group
    const int row = (warp & 3) * 32 + lane;         // query row inside the tile == TMEM lane
    uint64_t *wb = wbars + w * L::kPerTile;
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16) + w * kColsPerTile;
    float inv_sq, inv_sk, inv_sv;
    const float sq = pow2_scale(__uint_as_float(__ldg(amax + 0)), &inv_sq);
    pow2_scale(__uint_as_float(__ldg(amax + 1)), &inv_sk);
    pow2_scale(__uint_as_float(__ldg(amax + 2)), &inv_sv);
    const float c = inv_sq * inv_sk * 1.4426950408889634f;     // raw logit -> log2 units

    {   // this thread's query row -> fp16 hi / lo pairs -> TMEM (A operand of S = Q K^T): column = channel pair
      uint32_t qh[32], ql[32];
      if (TM) {
        const float *qrow = q + ((size_t)b * T + (size_t)(qtile0 + w) * kQT + row) * ld;
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          float4 v = __ldg(reinterpret_cast<const float4 *>(qrow) + c4);
          if (qbias != nullptr) {
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(qbias) + c4);
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
          }
          split2(v.x * sq, v.y * sq, qh[2 * c4], ql[2 * c4]);
          split2(v.z * sq, v.w * sq, qh[2 * c4 + 1], ql[2 * c4 + 1]);
        }
      } else {
        const float *qrow = q + (size_t)b * kD * T + (size_t)(qtile0 + w) * kQT + row;
#pragma unroll
        for (int c2 = 0; c2 < 32; ++c2)
          split2(__ldg(qrow + (size_t)(2 * c2) * T) * sq, __ldg(qrow + (size_t)(2 * c2 + 1) * T) * sq, qh[c2], ql[c2]);
      }
      BDM_TMEM_ST32(t_lane + kColQhi, qh);
      BDM_TMEM_ST32(t_lane + kColQlo, ql);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(wb + 0);                         // q_ready
    }

    float o[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) o[i] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f, alpha_prev = 0.0f;

    for (int j = 0; j < J; ++j) {
      uint32_t sr[kKT];
      mbar_wait(wb + 1, j & 1);                    // S_j is in TMEM
      tc_fence_after();
      BDM_TMEM_LD32(sr, t_lane + kColS);
      {
        uint32_t(&hi32)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[32]);
        BDM_TMEM_LD32(hi32, t_lane + kColS + 32);
      }
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(wb + 2);                         // s_empty: S_{j+1} may be written
      float mx4[4] = {m_run, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < kKT; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sr[i]));
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // exp(s - max) = 2^((s - max) * c): the difference first -- it is exact or nearly so for the entries that
      // matter -- then the scale.  (Folding it into one FFMA with a pre-rounded max*c costs 5x in accuracy.)
      const float alpha = fast_exp2((m_run - mx) * c);
      m_run = mx;
      float sum4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int i = 0; i < kKT; ++i) {
        const float p = fast_exp2((__uint_as_float(sr[i]) - mx) * c);
        sr[i] = __float_as_uint(p);
        sum4[i & 3] += p;
      }
      l_run = fmaf(l_run, alpha, (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
      if (j > 0) {
        // O_{j-1} = O_{j-2} * alpha_{j-1} + P_{j-1} V_{j-1}: one rounded FMA per element
        mbar_wait(wb + 4, (j - 1) & 1);            // o_full: P_{j-1} V_{j-1} is in TMEM (and P_{j-1} has been read)
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t ot[32];
          BDM_TMEM_LD32(ot, t_lane + kColO + h * 32);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[h * 32 + i] = fmaf(o[h * 32 + i], alpha_prev, __uint_as_float(ot[i]));
        }
      }
      alpha_prev = alpha;
      // P_j -> TMEM as fp16 hi / lo pairs (A operand of O_tile = P V): column = key pair
      {
        uint32_t ph[32], pl[32];
#pragma unroll
        for (int c2 = 0; c2 < 32; ++c2) split2(__uint_as_float(sr[2 * c2]), __uint_as_float(sr[2 * c2 + 1]), ph[c2], pl[c2]);
        BDM_TMEM_ST32(t_lane + kColPhi, ph);
        BDM_TMEM_ST32(t_lane + kColPlo, pl);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(wb + 3);                         // p_full
    }
    {
      mbar_wait(wb + 4, (J - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t ot[32];
        BDM_TMEM_LD32(ot, t_lane + kColO + h * 32);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[h * 32 + i] = fmaf(o[h * 32 + i], alpha_prev, __uint_as_float(ot[i]));
      }
      tc_fence_before();
    }
    const float scale = inv_sv / l_run;
    if (TM) {   // the thread's 64 channels are one 256-byte row of out[b][T][64]
      float4 *ob = reinterpret_cast<float4 *>(out + ((size_t)b * T + (size_t)(qtile0 + w) * kQT + row) * kD);
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4)
        ob[c4] = make_float4(o[4 * c4] * scale, o[4 * c4 + 1] * scale, o[4 * c4 + 2] * scale, o[4 * c4 + 3] * scale);
    } else {
      float *ob = out + (size_t)b * kD * T + (size_t)(qtile0 + w) * kQT + row;
#pragma unroll
      for (int ch = 0; ch < kD; ++ch) ob[(size_t)ch * T] = o[ch] * scale;   // 32 lanes = 128 contiguous bytes per channel
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4 * NQ + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
  }
}

}  // namespace tc05
}  // namespace bdm

extern "C" size_t bdm_attention_workspace_bytes(int b, int c, int t) {
  if (b <= 0 || c <= 0 || t <= 0) return 16;
  return 256 + (size_t)2 * b * t * c * 2 * 2;
}

// defined in attention.cu: the mma.sync (legacy tensor path) version, kept for A/B comparison (BDM_ATTENTION=mma)
extern "C" int bdm_attention_mma(int b, int c, int t, const float *q, const float *k, const float *v, float *out,
                                 void *workspace, size_t workspace_bytes, bdm_stream_t stream);

namespace bdm {
namespace tc05 {

// common tail of the two entry points: statistics are in amax, planes in the workspace
template <bool TM>
static int launch_attention(int b, int t, int ld, const float *q, const float *qbias, const float *k, const float *v,
                            const float *kbias, const float *vbias, float *out, void *workspace, int variant, cudaStream_t st) {
  unsigned *amax = static_cast<unsigned *>(workspace);
  const size_t plane = (size_t)b * t * kD * 2;   // halves per tensor (hi + lo)
  __half *kp = reinterpret_cast<__half *>(static_cast<unsigned char *>(workspace) + 256);
  __half *vp = kp + plane;
  attention_prep_kernel<TM><<<dim3(t / 32, b), 256, 0, st>>>(t, ld, k, v, kbias, vbias, amax, kp, vp);
  const int nq = (variant == 2 && t % (2 * kQT) == 0) ? 2 : 1;
  cudaError_t e;
  if (nq == 2) {
    e = ensure_dynamic_smem(reinterpret_cast<const void *>(attention_tc05_kernel<2, TM>), Smem<2>::kBytes);
    if (e != cudaSuccess) return (int)e;
    attention_tc05_kernel<2, TM><<<dim3(t / (2 * kQT), b), (4 * 2 + 4) * 32, Smem<2>::kBytes, st>>>(t, ld, q, qbias, kp, vp, amax, out);
  } else {
    e = ensure_dynamic_smem(reinterpret_cast<const void *>(attention_tc05_kernel<1, TM>), Smem<1>::kBytes);
    if (e != cudaSuccess) return (int)e;
    attention_tc05_kernel<1, TM><<<dim3(t / kQT, b), (4 * 1 + 4) * 32, Smem<1>::kBytes, st>>>(t, ld, q, qbias, kp, vp, amax, out);
  }
  BDM_RETURN_LAUNCH_STATUS();
}

static int attention_variant() {   // A/B hook: BDM_ATTENTION=mma | tc05x1 | tc05x2 (default: tc05x2 when t % 256 == 0)
  static const int variant = [] {
    const char *e = std::getenv("BDM_ATTENTION");
    if (e == nullptr) return 2;
    if (e[0] == 'm') return 0;
    return (e[0] == 't' && e[1] == 'c' && e[2] == '0' && e[3] == '5' && e[4] == 'x' && e[5] == '1') ? 1 : 2;
  }();
  return variant;
}

}  // namespace tc05
}  // namespace bdm

// q, k, v, out: f32[b][64][t] (channel-first, as the 1x1 convolutions of the block produce them);
// out[b][c][i] = sum_j softmax_j(q[b][:,i] . k[b][:,j]) * v[b][c][j].  c must be 64, t a multiple of 128 (256
// for the two-tile CTA).  workspace: bdm_attention_workspace_bytes(b, c, t) bytes, 256-byte aligned.
extern "C" int bdm_attention(int b, int c, int t, const float *q, const float *k, const float *v, float *out,
                             void *workspace, size_t workspace_bytes, bdm_stream_t stream) {
  using namespace bdm;
  using namespace bdm::tc05;
  BDM_CHECK_SIZE(b >= 0 && c == kD && t >= kQT && t % kQT == 0);
  if (b == 0) return BDM_OK;
  BDM_CHECK_PTR(q); BDM_CHECK_PTR(k); BDM_CHECK_PTR(v); BDM_CHECK_PTR(out); BDM_CHECK_PTR(workspace);
  if (((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
        reinterpret_cast<uintptr_t>(out)) & 15) != 0 || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return BDM_ERR_MISALIGNED;
  const int variant = attention_variant();
  if (variant == 0) return bdm_attention_mma(b, c, t, q, k, v, out, workspace, workspace_bytes, stream);
  if (workspace_bytes < bdm_attention_workspace_bytes(b, c, t)) return BDM_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned *amax = static_cast<unsigned *>(workspace);
  cudaMemsetAsync(amax, 0, 16, st);
  const size_t n4 = (size_t)b * kD * t / 4;
  attention_tc05_amax_kernel<<<2 * sm_count(), 512, 0, st>>>(n4, reinterpret_cast<const float4 *>(q),
                                                           reinterpret_cast<const float4 *>(k),
                                                           reinterpret_cast<const float4 *>(v), amax);
  return launch_attention<false>(b, t, 0, q, nullptr, k, v, nullptr, nullptr, out, workspace, variant, st);
}

// The same attention fed by ONE fused projection: qkv f32[b][t][ld] holds q | k | v of a token in columns [0,c),
// [c,2c), [2c,3c) (what x[b][t][:] @ [Wq;Wk;Wv]^T yields for channels-last activations), bias f32[3c] (or NULL) is
// added on the way in (the three 1x1 convolutions' biases), and the result is written token-major, out f32[b][t][c]:
// no bias-add kernels, no layout copies around the kernel (modules/pvconv.py:40-57).  ld >= 3c, ld % 4 == 0.
extern "C" int bdm_attention_qkv(int b, int c, int t, const float *qkv, int ld, const float *bias, float *out,
                                 void *workspace, size_t workspace_bytes, bdm_stream_t stream) {
  using namespace bdm;
  using namespace bdm::tc05;
  BDM_CHECK_SIZE(b >= 0 && c == kD && t >= kQT && t % kQT == 0 && ld >= 3 * c && ld % 4 == 0);
  if (b == 0) return BDM_OK;
  BDM_CHECK_PTR(qkv); BDM_CHECK_PTR(out); BDM_CHECK_PTR(workspace);
  if (((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias)) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return BDM_ERR_MISALIGNED;
  if (workspace_bytes < bdm_attention_workspace_bytes(b, c, t)) return BDM_ERR_WORKSPACE_TOO_SMALL;
  int variant = attention_variant();
  if (variant == 0) variant = 2;   // the mma.sync kernel has no token-major form
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned *amax = static_cast<unsigned *>(workspace);
  cudaMemsetAsync(amax, 0, 16, st);
  attention_tc05_amax_qkv_kernel<<<2 * sm_count(), 512, 0, st>>>((size_t)b * t, ld, qkv, bias, amax);
  return launch_attention<true>(b, t, ld, qkv, bias, qkv + c, qkv + 2 * c, bias != nullptr ? bias + c : nullptr,
                                bias != nullptr ? bias + 2 * c : nullptr, out, workspace, variant, st);
}
