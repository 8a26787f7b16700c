// groupnorm_cluster.cuh -- one-pass GroupNorm (+ conv bias, + Swish, + max over neighbours) for groups a little too
// large for one CTA, on thread-block clusters with distributed shared memory.  Included by groupnorm.cu (same
// arithmetic as its two-kernel path; same reference lines: modules/shared_mlp.py:25-31, modules/pointnet.py:86).
//
// The two-kernel path reads a tensor twice (statistics, then normalise).  Here the CTAs of a cluster split one
// (sample, group) block of cg*s CONTIGUOUS floats between them, fetch their slices into shared memory by TMA bulk
// copy (cp.async.bulk + mbarrier, no register staging), exchange their partial moments through DSMEM (one cluster
// barrier) and normalise from shared memory: 1 read + 1 write, one launch.
//
// Used for clusters of 1 - 4 CTAs (groups up to 256 KB).  Larger clusters were built and measured: 16 CTAs for the
// 1 MB groups of [32,64,1024,32] ran at 192 us against 146 us for statistics + apply, and a channels-last variant
// (16-CTA cluster per 8 MB sample, first rows stashed in shared memory, the rest re-read through L2) at 188 us
// against 140 us: the load-all / barrier / store-all phases of a cluster leave HBM idle more than two streaming
// kernels do.  Those variants were removed.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace bdm {
namespace gnc {

namespace cgx = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float swish_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// ---------------------------------------------------------------------------------------------------
// channel-first
// ---------------------------------------------------------------------------------------------------
constexpr int kCfThreads = 512;
constexpr int kCfSliceFloats = 16384;      // 64 KB of shared memory per CTA -> 3 CTAs per SM
constexpr int kCfMaxCg = 128;              // channels per group the per-channel (A, B) table holds

// MODE 0: y[b][c][s] elementwise;  MODE 1: y[b][c][s/u] = max over each run of u values (u/4 lanes per run)
template <bool SWISH, int MODE>
__global__ void __launch_bounds__(kCfThreads)
gn_cluster_kernel(int c, int s, int groups, float eps, int u, int per4, const float *__restrict__ x,
                  const float *__restrict__ conv_bias, const float *__restrict__ gamma,
                  const float *__restrict__ beta, float *__restrict__ y) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  float4 *tile = reinterpret_cast<float4 *>(dyn_smem);
  __shared__ double2 part;                  // this CTA's (sum, sum of squares) of (x + conv_bias - k)
  __shared__ double s_red[2][kCfThreads / 32];
  __shared__ float2 ab[kCfMaxCg];
  __shared__ __align__(8) uint64_t bar;
  cgx::cluster_group cluster = cgx::this_cluster();
  const unsigned rank = cluster.block_rank(), nranks = cluster.num_blocks();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cg = c / groups;
  const int sample = blockIdx.y / groups, g = blockIdx.y - sample * groups;
  const int ch0 = g * cg;
  const size_t base = ((size_t)sample * c + ch0) * s;
  const int n4 = (cg * s) >> 2;
  const int lo4 = min((int)rank * per4, n4), hi4 = min(lo4 + per4, n4), m4 = hi4 - lo4;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && m4 > 0) {
    const uint32_t bytes = (uint32_t)m4 * 16u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    const float *src = x + base + 4 * (size_t)lo4;
    for (uint32_t off = 0; off < bytes; off += 32768u) {
      const uint32_t n = min(32768u, bytes - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(dyn_smem + off)), "l"(reinterpret_cast<const unsigned char *>(src) + off), "r"(n),
                     "r"(smem_u32(&bar)) : "memory");
    }
  }
  // shift of the moments: the unit's first value (+ its conv bias), the same number in every CTA of the cluster
  const float k = __ldg(x + base) + (conv_bias != nullptr ? __ldg(conv_bias + ch0) : 0.0f);
  if (m4 > 0) {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  }

  float s1 = 0.0f, s2 = 0.0f;
  double d1 = 0.0, d2 = 0.0;
  int since = 0;
  for (int i = tid; i < m4; i += kCfThreads) {
    const float4 v = tile[i];
    const int ch = ch0 + (4 * (lo4 + i)) / s;     // s % 4 == 0: a float4 never straddles two channels
    const float kk = conv_bias != nullptr ? k - __ldg(conv_bias + ch) : k;
    const float a = v.x - kk, b2 = v.y - kk, c2 = v.z - kk, e2 = v.w - kk;
    s1 += (a + b2) + (c2 + e2);
    s2 += (a * a + b2 * b2) + (c2 * c2 + e2 * e2);
    if (++since == 8) { d1 += s1; d2 += s2; s1 = s2 = 0.0f; since = 0; }   // bound the fp32 run length (32 values)
  }
  d1 += s1; d2 += s2;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    d1 += __shfl_xor_sync(0xffffffffu, d1, d);
    d2 += __shfl_xor_sync(0xffffffffu, d2, d);
  }
  if (lane == 0) { s_red[0][warp] = d1; s_red[1][warp] = d2; }
  __syncthreads();
  if (tid == 0) {
    double a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int w = 0; w < kCfThreads / 32; ++w) { a1 += s_red[0][w]; a2 += s_red[1][w]; }
    part = make_double2(a1, a2);
  }
  cluster.sync();                           // every CTA's `part` is written and visible cluster-wide
  __shared__ double s_mean;
  __shared__ float s_rstd;
  if (tid == 0) {
    double a1 = 0.0, a2 = 0.0;
    for (unsigned r = 0; r < nranks; ++r) {  // fixed order: deterministic, identical in every CTA
      const double2 v = *cluster.map_shared_rank(&part, r);
      a1 += v.x; a2 += v.y;
    }
    const double n = (double)cg * (double)s;
    const double m = a1 / n;                 // mean of (x + conv_bias - k)
    const double var = fmax(a2 / n - m * m, 0.0);
    s_mean = (double)k + m;
    s_rstd = (float)(1.0 / sqrt(var + (double)eps));
  }
  cluster.sync();                           // nobody leaves (or reuses `part`) while a neighbour may still read it
  if (tid < cg) {
    const int ch = ch0 + tid;
    const float ga = gamma != nullptr ? __ldg(gamma + ch) : 1.0f;
    const float be = beta != nullptr ? __ldg(beta + ch) : 0.0f;
    const float cb = conv_bias != nullptr ? __ldg(conv_bias + ch) : 0.0f;
    const float A = s_rstd * ga;                                                     // same folding as gn_apply_kernel
    ab[tid] = make_float2(A, (float)((double)be + ((double)cb - s_mean) * (double)A));
  }
  __syncthreads();
  auto act = [](float t) { return SWISH ? swish_fast(t) : t; };
  if (MODE == 0) {
    float *py = y + base + 4 * (size_t)lo4;
    for (int i = tid; i < m4; i += kCfThreads) {
      const float4 v = tile[i];
      const float2 p = ab[(4 * (lo4 + i)) / s];
      st_stream_f4(py + 4 * (size_t)i, make_float4(act(fmaf(v.x, p.x, p.y)), act(fmaf(v.y, p.x, p.y)),
                                                   act(fmaf(v.z, p.x, p.y)), act(fmaf(v.w, p.x, p.y))));
    }
  } else {
    const int lpr = u >> 2;                  // lanes per run of u values; per4 and kCfThreads are multiples of it
    float *py = y + ((size_t)sample * c + ch0) * (size_t)(s / u);
    for (int i0 = 0; i0 < m4; i0 += kCfThreads) {
      const int i = i0 + tid;
      // the largest activated value of a run belongs to its largest or its smallest input (fma monotone, Swish
      // unimodal): two activations per run instead of u
      const float inf = __int_as_float(0x7f800000);
      float hi_v = -inf, lo_v = inf;
      float2 p = make_float2(0.0f, 0.0f);
      if (i < m4) {
        const float4 v = tile[i];
        p = ab[(4 * (lo4 + i)) / s];
        hi_v = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        lo_v = fminf(fminf(v.x, v.y), fminf(v.z, v.w));
      }
      for (int d = 1; d < lpr; d <<= 1) {
        hi_v = fmaxf(hi_v, __shfl_xor_sync(0xffffffffu, hi_v, d));
        lo_v = fminf(lo_v, __shfl_xor_sync(0xffffffffu, lo_v, d));
      }
      if (i < m4 && (tid & (lpr - 1)) == 0) py[(lo4 + i) / lpr] = fmaxf(act(fmaf(hi_v, p.x, p.y)), act(fmaf(lo_v, p.x, p.y)));
    }
  }
}

// cluster size for a (sample, group) block of `gelems` floats: the smallest power of two whose slices fit
static inline int cf_cluster_size(long long gelems) {
  int cl = 1;
  while (cl < 16 && gelems > (long long)cl * kCfSliceFloats) cl <<= 1;
  return gelems <= (long long)cl * kCfSliceFloats ? cl : 0;
}

}  // namespace gnc
}  // namespace bdm
