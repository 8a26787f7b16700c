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
// logits are un-scaled, so single-pass TF32 is not acceptable.  Both products use the 3xTF32 split
// (x = hi + lo, hi = tf32(x), lo = tf32(x - hi); hi*hi + hi*lo + lo*hi accumulated in fp32, small terms
// first), which keeps ~21 mantissa bits per product -- the same order as fp32 accumulation error over the
// 64- and 4096-long sums.  tests/test_dense_fused_gpu.py checks it against float64 next to torch's fp32.
//
// Tiling: CTA = 128 queries (8 warps x 16 rows) x all keys in tiles of 64; raw K/V tiles arrive by
// cp.async (double-buffered) and are split into tf32 hi / lo planes once per CTA.  Tensor cores through mma.sync.m16n8k8.tf32 (the 16-row query fragment
// stays in registers for the whole kernel).  The S accumulator fragment is reused directly as the A
// operand of P.V by numbering the key slots of each 8-key slice as (2t, 2t+1) <-> (t, t+4): the sum over
// keys does not care about their order, and V is read from shared memory under the same numbering
// (one 64-bit load per fragment), so no shuffles are needed between the two GEMMs.
#include "common.cuh"

namespace bdm {

constexpr int kHD = 64;             // channels
constexpr int kBM = 128;            // queries per CTA
constexpr int kBN = 64;             // keys per tile
constexpr int kLd = kBN + 8;        // K/V tile row stride (floats): = 8 mod 32 -> conflict-free fragment loads
constexpr int kLdQ = kBM + 8;       // Q staging row stride
constexpr int kLdO = kBM + 4;       // output staging row stride
constexpr int kAttnThreads = 256;
constexpr int kTile = kHD * kLd;               // one [64][72] tile
constexpr int kStageFloats = 2 * kTile;        // raw K tile + raw V tile (cp.async destination)
constexpr int kSmemFloats = 2 * kStageFloats + 6 * kTile;   // 2 raw stages + K {hi,lo} + 2 x V {hi,lo} planes

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_hd64_kernel(int T, const float *__restrict__ q, const float *__restrict__ k,
                      const float *__restrict__ v, float *__restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.y, i0 = blockIdx.x * kBM;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const float *qb = q + (size_t)b * kHD * T, *kb = k + (size_t)b * kHD * T, *vb = v + (size_t)b * kHD * T;

  // ---- Q tile -> shared -> per-warp A fragments (hi/lo), kept for the whole kernel ----
  for (int idx = tid; idx < kHD * (kBM / 4); idx += kAttnThreads) {
    const int c = idx / (kBM / 4), f = (idx % (kBM / 4)) * 4;
    *reinterpret_cast<float4 *>(smem + c * kLdQ + f) = __ldg(reinterpret_cast<const float4 *>(qb + (size_t)c * T + i0 + f));
  }
  __syncthreads();
  uint32_t qh[8][4], ql[8][4];
  {
    const int r0 = warp * 16;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      split_tf32(smem[(ks * 8 + t) * kLdQ + r0 + g], qh[ks][0], ql[ks][0]);
      split_tf32(smem[(ks * 8 + t) * kLdQ + r0 + g + 8], qh[ks][1], ql[ks][1]);
      split_tf32(smem[(ks * 8 + t + 4) * kLdQ + r0 + g], qh[ks][2], ql[ks][2]);
      split_tf32(smem[(ks * 8 + t + 4) * kLdQ + r0 + g + 8], qh[ks][3], ql[ks][3]);
    }
  }
  __syncthreads();

  auto load_tile = [&](int stage, int j0) {
    float *ks_ = smem + stage * kStageFloats;
    float *vs_ = ks_ + kHD * kLd;
    for (int idx = tid; idx < kHD * (kBN / 4); idx += kAttnThreads) {
      const int c = idx / (kBN / 4), f = (idx % (kBN / 4)) * 4;
      cp_async16(ks_ + c * kLd + f, kb + (size_t)c * T + j0 + f);
      cp_async16(vs_ + c * kLd + f, vb + (size_t)c * T + j0 + f);
    }
    cp_async_commit();
  };

  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.0f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;   // rows g and g+8 of this warp's 16

  // tf32 hi / lo planes of the current tile, split once per CTA instead of once per warp.  The V planes
  // are double-buffered because half of the warps consume them one tile late (see below).
  uint32_t *kh_ = reinterpret_cast<uint32_t *>(smem + 2 * kStageFloats);
  uint32_t *kl_ = kh_ + kTile;
  uint32_t *vplanes = kl_ + kTile;   // [2 stages][hi, lo][kTile]

  // O = O * scale + P V^T for one key tile.  Key slots of slice ks are numbered (t, t+4) <-> tile columns
  // (2t, 2t+1).  The tile's product is accumulated from zero and merged with one rounded FMA per element:
  // tensor-core accumulation truncates, and 512 chained MMAs into one running accumulator showed as a
  // 2e-5 bias.
  auto pv_merge = [&](const float (&p)[8][4], float sc0, float sc1, const uint32_t *vh_, const uint32_t *vl_) {
    float part[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) part[nt][0] = part[nt][1] = part[nt][2] = part[nt][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t ph[4], pl[4];
      split_tf32(p[ks][0], ph[0], pl[0]);
      split_tf32(p[ks][2], ph[1], pl[1]);
      split_tf32(p[ks][1], ph[2], pl[2]);
      split_tf32(p[ks][3], ph[3], pl[3]);
      // four channel tiles at a time, the three passes interleaved across them so that consecutive MMAs
      // never wait on each other's accumulator
#pragma unroll
      for (int n0 = 0; n0 < 8; n0 += 4) {
        uint2 vh[4], vl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int iv = ((n0 + u) * 8 + g) * kLd + ks * 8 + 2 * t;
          vh[u] = *reinterpret_cast<const uint2 *>(vh_ + iv);
          vl[u] = *reinterpret_cast<const uint2 *>(vl_ + iv);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) mma_tf32(part[n0 + u], pl, vh[u].x, vh[u].y);
#pragma unroll
        for (int u = 0; u < 4; ++u) mma_tf32(part[n0 + u], ph, vl[u].x, vl[u].y);
#pragma unroll
        for (int u = 0; u < 4; ++u) mma_tf32(part[n0 + u], ph, vh[u].x, vh[u].y);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o[nt][0] = fmaf(o[nt][0], sc0, part[nt][0]);
      o[nt][1] = fmaf(o[nt][1], sc0, part[nt][1]);
      o[nt][2] = fmaf(o[nt][2], sc1, part[nt][2]);
      o[nt][3] = fmaf(o[nt][3], sc1, part[nt][3]);
    }
  };

  // The barriers keep the 8 warps in step, so if all of them ran S -> softmax -> PV the tensor pipe would
  // idle through every softmax.  Warps 4-7 therefore run one PV behind (PV of tile jt-1, then S and softmax
  // of tile jt): on each scheduler one warp is in its softmax while the other issues MMAs.
  const bool late_pv = warp >= kAttnThreads / 64;
  float s[8][4];
  float sc0 = 0.0f, sc1 = 0.0f;

  const int ntiles = T / kBN;
  load_tile(0, 0);
  for (int jt = 0; jt < ntiles; ++jt) {
    cp_async_wait<0>();
    __syncthreads();   // raw tile jt landed; every warp is done with the planes this tile will overwrite
    if (jt + 1 < ntiles) load_tile((jt + 1) & 1, (jt + 1) * kBN);
    uint32_t *vh_cur = vplanes + (jt & 1) * 2 * kTile, *vl_cur = vh_cur + kTile;
    {
      const float *raw = smem + (jt & 1) * kStageFloats;
      for (int idx = tid; idx < 2 * kHD * (kBN / 4); idx += kAttnThreads) {
        const int which = idx / (kHD * (kBN / 4));          // 0 = K, 1 = V
        const int rem = idx - which * (kHD * (kBN / 4));
        const int off = (rem / (kBN / 4)) * kLd + (rem % (kBN / 4)) * 4;
        const float4 x = *reinterpret_cast<const float4 *>(raw + which * kTile + off);
        uint4 hi, lo;
        split_tf32(x.x, hi.x, lo.x);
        split_tf32(x.y, hi.y, lo.y);
        split_tf32(x.z, hi.z, lo.z);
        split_tf32(x.w, hi.w, lo.w);
        *reinterpret_cast<uint4 *>((which ? vh_cur : kh_) + off) = hi;
        *reinterpret_cast<uint4 *>((which ? vl_cur : kl_) + off) = lo;
      }
    }
    __syncthreads();

    if (late_pv && jt > 0) {
      const uint32_t *vh_prev = vplanes + ((jt - 1) & 1) * 2 * kTile;
      pv_merge(s, sc0, sc1, vh_prev, vh_prev + kTile);
    }

    // ---- S = Q^T K (16 queries x 64 keys per warp) ----
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
      for (int n0 = 0; n0 < 8; n0 += 4) {
        uint32_t bh0[4], bh1[4], bl0[4], bl1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i0_ = (ks * 8 + t) * kLd + (n0 + u) * 8 + g, i1_ = i0_ + 4 * kLd;
          bh0[u] = kh_[i0_];
          bh1[u] = kh_[i1_];
          bl0[u] = kl_[i0_];
          bl1[u] = kl_[i1_];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) mma_tf32(s[n0 + u], ql[ks], bh0[u], bh1[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) mma_tf32(s[n0 + u], qh[ks], bl0[u], bl1[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) mma_tf32(s[n0 + u], qh[ks], bh0[u], bh1[u]);
      }
    }

    // ---- online softmax ----
    float mx0 = s[0][0], mx1 = s[0][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    sc0 = __expf(m0 - mn0);
    sc1 = __expf(m1 - mn1);
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
    l0 = l0 * sc0 + sum0;   // per-thread partial; the 4 lanes of a row are combined at the end
    l1 = l1 * sc1 + sum1;

    if (!late_pv) pv_merge(s, sc0, sc1, vh_cur, vl_cur);
  }
  if (late_pv) {
    const uint32_t *vh_prev = vplanes + ((ntiles - 1) & 1) * 2 * kTile;
    pv_merge(s, sc0, sc1, vh_prev, vh_prev + kTile);
  }
  __syncthreads();   // planes no longer read: the front of shared memory becomes the output staging area

  // ---- normalise, transpose through shared memory, coalesced store of h[c][i] ----
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
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
extern "C" int bdm_attention(int b, int c, int t, const float *q, const float *k, const float *v, float *out,
                             bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c == kHD && t >= kBM && t % kBM == 0);
  if (b == 0) return BDM_OK;
  BDM_CHECK_PTR(q); BDM_CHECK_PTR(k); BDM_CHECK_PTR(v); BDM_CHECK_PTR(out);
  if (((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
        reinterpret_cast<uintptr_t>(out)) & 15) != 0)
    return BDM_ERR_MISALIGNED;
  const size_t smem_bytes = sizeof(float) * kSmemFloats;
  static_assert(2 * kStageFloats >= kHD * kLdQ && 2 * kStageFloats >= kHD * kLdO, "Q / output staging fits");
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(attention_hd64_kernel), smem_bytes);
  if (e != cudaSuccess) return (int)e;
  attention_hd64_kernel<<<dim3(t / kBM, b), kAttnThreads, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(
      t, q, k, v, out);
  BDM_RETURN_LAUNCH_STATUS();
}
