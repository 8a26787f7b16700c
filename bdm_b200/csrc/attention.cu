// attention.cu -- fused single-head self-attention over voxels for the PVConv attention block.
//
// The reference (experiments/model/pvcnn/modules/pvconv.py:36-63, class Attention) computes, per shape,
//     attn = softmax_j( q[:, i] . k[:, j] )          (un-scaled logits, [T, T])
//     h[c, i] = sum_j attn[i, j] * v[c, j]
// with two fp32 matmuls and a softmax: at the PC^2 stage that has it (C = 64 channels, T = 16^3 = 4096
// voxels, 16 shapes) the [16, 4096, 4096] logits are written, read, written and read again -- 4.3 GB of
// HBM traffic and 2.3 ms of an 8.9 ms step.  This kernel never materialises them (online softmax over key
// tiles, running max / sum per query in registers).
//
// Precision: torch runs these matmuls in IEEE fp32 (allow_tf32 is off for matmul by default) and the
// logits are un-scaled, so a single reduced-precision pass is not acceptable.  Every fp32 operand is split
// into two fp16 numbers (x*s = hi + lo, 11 + 11 mantissa bits; s is a per-tensor power of two that puts
// max|x| in [512, 1024) so that neither half leaves fp16's range) and each product is three tensor-core
// products, hi*hi + hi*lo + lo*hi, accumulated in fp32, small terms first: ~22 mantissa bits per product,
// the same order as fp32 accumulation error over the 64- and 4096-long sums.  Tensor-core accumulation
// truncates, so the P.V product of each key tile is accumulated from zero and merged into the running
// output with one rounded FMA (512 chained MMAs into one accumulator showed as a 2e-5 bias).
// tests/test_dense_fused_gpu.py checks the result against float64 next to torch's fp32 route.  (A 3xTF32
// version of the same kernel measured 1.32 ms at B=16, T=4096 against 0.80 ms for this one: m16n8k16.f16
// covers twice the reduction depth per tensor-pipe cycle and needs half the shared-memory operand reads.)
//
// Tiling: CTA = 128 queries (8 warps x 16 rows) x all keys in tiles of 64; raw K/V tiles arrive by
// cp.async (double-buffered) and are converted once per CTA into fp16 hi / lo planes in their natural
// [channel][key] layout; fragments come through ldmatrix (.trans for K, whose reduction index is the row).
// The 16-row query fragment stays in registers for the whole kernel, and the S accumulator fragment is
// re-packed in registers as the A operand of P.V (no shuffles between the two GEMMs).
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace bdm {

constexpr int kHD = 64;             // channels
constexpr int kBM = 128;            // queries per CTA
constexpr int kBN = 64;             // keys per tile
constexpr int kLd = kBN + 8;        // raw K/V tile row stride (floats)
constexpr int kLdQ = kBM + 8;       // Q staging row stride
constexpr int kLdO = kBM + 4;       // output staging row stride
constexpr int kAttnThreads = 256;
constexpr int kTile = kHD * kLd;               // one [64][72] tile
constexpr int kStageFloats = 2 * kTile;        // raw K tile + raw V tile (cp.async destination)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kLdH = kBN + 8;          // halves per plane row (144 bytes: conflict-free ldmatrix rows)
constexpr int kPlaneH = kHD * kLdH;    // halves per plane
constexpr int kSmemBytesF16 = 2 * kStageFloats * 4 + 2 * 4 * kPlaneH * 2;   // 2 raw stages + 2 x {K,V} x {hi,lo} planes

__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void *p) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void *p) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// power of two s with max|x| * s in [512, 1024); exponent of 1/s in *inv_exp
__device__ __forceinline__ float pow2_scale(float amax, float *inv) {
  int e = 9;   // amax == 0 / non-finite: scale 1
  if (amax > 0.0f && amax < INFINITY) e = (int)((__float_as_uint(amax) >> 23) & 255u) - 127;
  const int se = min(max(9 - e, -60), 60);
  *inv = __uint_as_float((uint32_t)(127 - se) << 23);
  return __uint_as_float((uint32_t)(127 + se) << 23);
}

// max |x| of three equally sized tensors -> amax[0..2] (bit pattern of a non-negative float; zeroed by the host)
__global__ void attention_amax_kernel(size_t n4, const float4 *__restrict__ q, const float4 *__restrict__ k,
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

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_hd64_f16_kernel(int T, const float *__restrict__ q, const float *__restrict__ k,
                          const float *__restrict__ v, float *__restrict__ out,
                          const unsigned *__restrict__ amax) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.y, i0 = blockIdx.x * kBM;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const float *qb = q + (size_t)b * kHD * T, *kb = k + (size_t)b * kHD * T, *vb = v + (size_t)b * kHD * T;

  float inv_sq, inv_sk, inv_sv;
  const float sq = pow2_scale(__uint_as_float(__ldg(amax + 0)), &inv_sq);
  const float sk = pow2_scale(__uint_as_float(__ldg(amax + 1)), &inv_sk);
  const float sv = pow2_scale(__uint_as_float(__ldg(amax + 2)), &inv_sv);
  const float inv_sqk = inv_sq * inv_sk;

  // ---- Q tile -> shared -> per-warp A fragments (hi/lo halves), kept for the whole kernel ----
  for (int idx = tid; idx < kHD * (kBM / 4); idx += kAttnThreads) {
    const int c = idx / (kBM / 4), f = (idx % (kBM / 4)) * 4;
    *reinterpret_cast<float4 *>(smem + c * kLdQ + f) = __ldg(reinterpret_cast<const float4 *>(qb + (size_t)c * T + i0 + f));
  }
  __syncthreads();
  uint32_t qh[4][4], ql[4][4];
  {
    const int r0 = warp * 16;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float *c0 = smem + (kk * 16 + 2 * t) * kLdQ + r0 + g;   // channel kk*16+2t, query row g
      split_h2(c0[0] * sq, c0[kLdQ] * sq, qh[kk][0], ql[kk][0]);
      split_h2(c0[8] * sq, c0[kLdQ + 8] * sq, qh[kk][1], ql[kk][1]);
      split_h2(c0[8 * kLdQ] * sq, c0[9 * kLdQ] * sq, qh[kk][2], ql[kk][2]);
      split_h2(c0[8 * kLdQ + 8] * sq, c0[9 * kLdQ + 8] * sq, qh[kk][3], ql[kk][3]);
    }
  }
  __syncthreads();

  auto load_tile = [&](int stage, int j0) {
    float *ks_ = smem + stage * kStageFloats;
    float *vs_ = ks_ + kTile;
    for (int idx = tid; idx < kHD * (kBN / 4); idx += kAttnThreads) {
      const int c = idx / (kBN / 4), f = (idx % (kBN / 4)) * 4;
      cp_async16(ks_ + c * kLd + f, kb + (size_t)c * T + j0 + f);
      cp_async16(vs_ + c * kLd + f, vb + (size_t)c * T + j0 + f);
    }
    cp_async_commit();
  };

  __half *planes = reinterpret_cast<__half *>(smem + 2 * kStageFloats);   // [2][K hi, K lo, V hi, V lo][kPlaneH]
  // ldmatrix row of this lane: matrix (lane >> 3), row (lane & 7)
  const int lm = lane >> 3, lr = lane & 7;
  const int k_off = (lr + (lm & 1) * 8) * kLdH + (lm >> 1) * 8;   // K (.trans): rows = channels, + n-tile of the pair
  const int v_off = ((lm >> 1) * 8 + lr) * kLdH + (lm & 1) * 8;   // V: rows = channels (n), + key half of the slice

  // raw fp32 stage -> scaled fp16 hi / lo planes (same [channel][key] layout)
  auto convert_tile = [&](int stage) {
    const float *raw = smem + stage * kStageFloats;
    __half *pl = planes + stage * 4 * kPlaneH;
    for (int idx = tid; idx < 2 * kHD * (kBN / 4); idx += kAttnThreads) {
      const int which = idx / (kHD * (kBN / 4));          // 0 = K, 1 = V
      const int rem = idx - which * (kHD * (kBN / 4));
      const int c = rem / (kBN / 4), f = (rem % (kBN / 4)) * 4;
      const float sc = which ? sv : sk;
      const float4 x = *reinterpret_cast<const float4 *>(raw + which * kTile + c * kLd + f);
      uint2 hi, lo;
      split_h2(x.x * sc, x.y * sc, hi.x, lo.x);
      split_h2(x.z * sc, x.w * sc, hi.y, lo.y);
      *reinterpret_cast<uint2 *>(pl + (2 * which) * kPlaneH + c * kLdH + f) = hi;
      *reinterpret_cast<uint2 *>(pl + (2 * which + 1) * kPlaneH + c * kLdH + f) = lo;
    }
  };

  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.0f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;

  // Pipeline, one barrier per key tile: while tile jt is consumed from planes[jt & 1], the same warps
  // convert raw tile jt+1 into the other plane set and cp.async fetches raw tile jt+2; conversion (ALU) and
  // the MMAs of different warps overlap because nothing separates them inside an iteration.
  const int ntiles = T / kBN;
  load_tile(0, 0);
  cp_async_wait<0>();
  __syncthreads();
  if (ntiles > 1) load_tile(1, kBN);
  convert_tile(0);
  for (int jt = 0; jt < ntiles; ++jt) {
    cp_async_wait<0>();
    __syncthreads();   // planes[jt & 1] complete, raw tile jt+1 landed, everyone is done with tile jt-1
    if (jt + 2 < ntiles) load_tile(jt & 1, (jt + 2) * kBN);
    if (jt + 1 < ntiles) convert_tile((jt + 1) & 1);
    const __half *kh_ = planes + (jt & 1) * 4 * kPlaneH;
    const __half *kl_ = kh_ + kPlaneH, *vh_ = kl_ + kPlaneH, *vl_ = vh_ + kPlaneH;

    // ---- S = Q^T K (16 queries x 64 keys per warp) ----
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bh[4], bl[4];
        const int off = kk * 16 * kLdH + np * 16 + k_off;
        ldsm_x4_trans(bh, kh_ + off);
        ldsm_x4_trans(bl, kl_ + off);
        mma_f16(s[2 * np], ql[kk], bh[0], bh[1]);
        mma_f16(s[2 * np + 1], ql[kk], bh[2], bh[3]);
        mma_f16(s[2 * np], qh[kk], bl[0], bl[1]);
        mma_f16(s[2 * np + 1], qh[kk], bl[2], bl[3]);
        mma_f16(s[2 * np], qh[kk], bh[0], bh[1]);
        mma_f16(s[2 * np + 1], qh[kk], bh[2], bh[3]);
      }
    }

    // ---- online softmax (logits un-scaled back by the exact power of two) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] *= inv_sqk; s[nt][1] *= inv_sqk; s[nt][2] *= inv_sqk; s[nt][3] *= inv_sqk;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float sc0 = __expf(m0 - mn0), sc1 = __expf(m1 - mn1);
    m0 = mn0;
    m1 = mn1;
    float sum0 = 0.0f, sum1 = 0.0f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = __expf(s[nt][0] - mn0);
      s[nt][1] = __expf(s[nt][1] - mn0);
      s[nt][2] = __expf(s[nt][2] - mn1);
      s[nt][3] = __expf(s[nt][3] - mn1);
      sum0 += s[nt][0] + s[nt][1];
      sum1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * sc0 + sum0;
    l1 = l1 * sc1 + sum1;

    // ---- O = O * scale + P V^T; the tile's product starts from zero (truncating accumulation, see top) ----
    float part[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) part[nt][0] = part[nt][1] = part[nt][2] = part[nt][3] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ph[4], pl[4];   // the S accumulator fragments of key tiles 2kk, 2kk+1 are the A fragment of slice kk
      split_h2(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
      split_h2(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
      split_h2(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
      split_h2(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t vh[4], vl[4];
        const int off = np * 16 * kLdH + kk * 16 + v_off;
        ldsm_x4(vh, vh_ + off);
        ldsm_x4(vl, vl_ + off);
        mma_f16(part[2 * np], pl, vh[0], vh[1]);
        mma_f16(part[2 * np + 1], pl, vh[2], vh[3]);
        mma_f16(part[2 * np], ph, vl[0], vl[1]);
        mma_f16(part[2 * np + 1], ph, vl[2], vl[3]);
        mma_f16(part[2 * np], ph, vh[0], vh[1]);
        mma_f16(part[2 * np + 1], ph, vh[2], vh[3]);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o[nt][0] = fmaf(o[nt][0], sc0, part[nt][0]);
      o[nt][1] = fmaf(o[nt][1], sc0, part[nt][1]);
      o[nt][2] = fmaf(o[nt][2], sc1, part[nt][2]);
      o[nt][3] = fmaf(o[nt][3], sc1, part[nt][3]);
    }
  }
  __syncthreads();

  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = inv_sv / l0, inv1 = inv_sv / l1;
  {
    const int r0 = warp * 16;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + 2 * t;
      smem[c * kLdO + r0 + g] = o[nt][0] * inv0;
      smem[(c + 1) * kLdO + r0 + g] = o[nt][1] * inv0;
      smem[c * kLdO + r0 + g + 8] = o[nt][2] * inv1;
      smem[(c + 1) * kLdO + r0 + g + 8] = o[nt][3] * inv1;
    }
  }
  __syncthreads();
  float *ob = out + (size_t)b * kHD * T + i0;
  for (int idx = tid; idx < kHD * (kBM / 4); idx += kAttnThreads) {
    const int c = idx / (kBM / 4), f = (idx % (kBM / 4)) * 4;
    *reinterpret_cast<float4 *>(ob + (size_t)c * T + f) = *reinterpret_cast<const float4 *>(smem + c * kLdO + f);
  }
}

}  // namespace bdm

// q, k, v, out: f32[b][64][t] (channel-first, as the 1x1 convolutions of the block produce them);
// out[b][c][i] = sum_j softmax_j(q[b][:,i] . k[b][:,j]) * v[b][c][j].  c must be 64, t a multiple of 128.
// workspace: 16 bytes (the three max|.|).
// Not part of the public header: the entry point is bdm_attention (attention_tc05.cu), which routes here when
// BDM_ATTENTION=mma is set (A/B timing of the legacy tensor path against the tcgen05 kernel).
extern "C" int bdm_attention_mma(int b, int c, int t, const float *q, const float *k, const float *v, float *out,
                                 void *workspace, size_t workspace_bytes, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c == kHD && t >= kBM && t % kBM == 0);
  if (b == 0) return BDM_OK;
  BDM_CHECK_PTR(q); BDM_CHECK_PTR(k); BDM_CHECK_PTR(v); BDM_CHECK_PTR(out);
  if (((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
        reinterpret_cast<uintptr_t>(out)) & 15) != 0)
    return BDM_ERR_MISALIGNED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  BDM_CHECK_PTR(workspace);
  if (workspace_bytes < 16) return BDM_ERR_WORKSPACE_TOO_SMALL;
  unsigned *amax = static_cast<unsigned *>(workspace);
  cudaMemsetAsync(amax, 0, 16, st);
  const size_t n4 = (size_t)b * kHD * t / 4;
  attention_amax_kernel<<<2 * sm_count(), 512, 0, st>>>(n4, reinterpret_cast<const float4 *>(q),
                                                      reinterpret_cast<const float4 *>(k),
                                                      reinterpret_cast<const float4 *>(v), amax);
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(attention_hd64_f16_kernel), kSmemBytesF16);
  if (e != cudaSuccess) return (int)e;
  attention_hd64_f16_kernel<<<dim3(t / kBM, b), kAttnThreads, kSmemBytesF16, st>>>(t, q, k, v, out, amax);
  BDM_RETURN_LAUNCH_STATUS();
}
