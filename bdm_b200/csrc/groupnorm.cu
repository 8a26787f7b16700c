// groupnorm.cu -- fused [conv bias +] GroupNorm [+ Swish] [+ max over neighbours | + channel sums], sm_100a.
//
// Every conv of the denoisers is followed by GroupNorm(8) and Swish (x * sigmoid(x)):
// modules/shared_mlp.py:25-31 (1x1 conv -> GroupNorm -> Swish), modules/pvconv.py:75-88 (Conv3d -> GroupNorm
// -> Swish), pvconv.py:59-61 (attention: GroupNorm -> Swish).  Through torch that is a bias-add kernel
// after the conv, then four kernels and ~5 reads + 3 writes of the tensor for norm + activation
// (RowwiseMoments with ONE CTA per (sample, group) row of up to 512 KB, the normalise pass, sigmoid, mul),
// often followed by yet another full pass: the max over the 32 neighbours in PointNetSAModule
// (modules/pointnet.py:86) or the squeeze-excite mean (modules/se.py:19).  All of it is HBM streaming.
// Fused here into 2 reads + (at most) 1 write:
//   gn_stats_kernel   per (sample, channel) row of S contiguous floats, split over several CTAs when long:
//                     per-thread fp32 partial sums over <= 64 elements, then double precision through the
//                     warp / block reduction; one (sum, sumsq) partial per CTA in a fixed slot
//                     (deterministic).
//   gn_apply_kernel   one CTA per (sample, channel, tile): warp 0 folds the group's partials -- and, when
//                     the conv was run without its bias, the per-channel bias terms, analytically -- into
//                     mean / rstd in double, then y = act((x + b_conv) * rstd*gamma + beta - mean*rstd*gamma)
//                     with 128-bit loads; optionally reduces max over the innermost U (<= 128) values
//                     instead of writing them all, or emits per-tile sums of y for the SE squeeze.
// Tolerance against torch (conv bias add, F.group_norm, x*sigmoid(x), max / mean): 1e-5 relative to the
// output's peak (tests/test_dense_fused_gpu.py); biased variance, eps inside the sqrt, like torch.
#include <cstdlib>

#include "common.cuh"
#include "groupnorm_cluster.cuh"

namespace bdm {

constexpr int kGnThreads = 256;
constexpr int kGnMaxChunks = 32;

// kernels launched by the last bdm_groupnorm_act / _cl call of this thread (1 = one-pass or apply-only,
// 2 = statistics + apply): lets the host-side launch accounting (bench.py's gpu_launches) stay exact.
static thread_local int g_last_launches = 0;

__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(long long row_len, int nchunks, const float *__restrict__ x, double2 *__restrict__ partials) {
  // CTAs are dispatched in ascending block order; the tensor was written front to back by its producer, so its END is what
  // the 126 MB L2 still holds: walk it back to front
  const long long row = gridDim.y - 1 - blockIdx.y;
  const int chunk = gridDim.x - 1 - blockIdx.x;
  long long per = (row_len + nchunks - 1) / nchunks;
  per = (per + 3) & ~3LL;  // chunks start on 16-byte boundaries when the row does
  const long long lo = min(chunk * per, row_len), hi = min(lo + per, row_len);
  const float *p = x + row * row_len;
  // Sums are taken of (x - k), k = the row's first element: with a shift near the mean the fp32 partial
  // sums of squares keep the variance even when |mean| >> std; the apply kernel undoes the shift in double.
  const float k = __ldg(p);
  float s = 0.0f, q = 0.0f;
  double ds = 0.0, dq = 0.0;
  int since = 0;
  if ((reinterpret_cast<uintptr_t>(p + lo) & 15) == 0) {
    const long long n4 = (hi - lo) >> 2;
    long long i = threadIdx.x;
    // 4 independent 128-bit loads in flight per thread, 4 independent fp32 accumulator pairs
    for (; i + 3 * kGnThreads < n4; i += 4 * kGnThreads) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = ld_stream_f4(p + lo + 4 * (i + (long long)j * kGnThreads));
      float s4 = 0.0f, q4 = 0.0f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = v[j].x - k, b2 = v[j].y - k, c2 = v[j].z - k, d2 = v[j].w - k;
        s4 += (a + b2) + (c2 + d2);
        q4 += (a * a + b2 * b2) + (c2 * c2 + d2 * d2);
      }
      s += s4; q += q4;
      if (++since == 4) { ds += s; dq += q; s = q = 0.0f; since = 0; }  // bound the fp32 run length (64 values)
    }
    for (; i < n4; i += kGnThreads) {
      float4 v = ld_stream_f4(p + lo + 4 * i);
      v.x -= k; v.y -= k; v.z -= k; v.w -= k;
      s += (v.x + v.y) + (v.z + v.w);
      q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    for (long long t = lo + (n4 << 2) + threadIdx.x; t < hi; t += kGnThreads) { const float v = p[t] - k; s += v; q += v * v; }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += kGnThreads) {
      const float v = p[i] - k;
      s += v; q += v * v;
      if (++since == 64) { ds += s; dq += q; s = q = 0.0f; since = 0; }
    }
  }
  ds += s; dq += q;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, d);
    dq += __shfl_xor_sync(0xffffffffu, dq, d);
  }
  __shared__ double sh[2][kGnThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = ds; sh[1][warp] = dq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b2 = 0.0;
#pragma unroll
    for (int w = 0; w < kGnThreads / 32; ++w) { a += sh[0][w]; b2 += sh[1][w]; }
    partials[row * nchunks + chunk] = make_double2(a, b2);
  }
}

// MODE 0: y[b,c,s] elementwise (+ optional per-tile sums of y);  MODE 1: y[b,c,s/u] = max over innermost u
template <bool SWISH, int MODE>
__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(int c, long long s, int groups, int nchunks, float eps, int tile4, int u,
                const float *__restrict__ x, const float *__restrict__ conv_bias,
                const float *__restrict__ gamma, const float *__restrict__ beta,
                const double2 *__restrict__ partials, float *__restrict__ y, float *__restrict__ tile_sums) {
  const long long bc = blockIdx.x;       // sample * c + channel
  const int ch = (int)(bc % c);
  const long long sample = bc / c;
  const int cg = c / groups;
  const int ch0 = (ch / cg) * cg;        // first channel of this channel's group
  __shared__ float s_ab[2];
  __shared__ float s_red[kGnThreads / 32];
  if (threadIdx.x < 32) {
    // fold the group's per-channel partial sums of (x - k_c).  With t = k_c + conv_bias_c the channel's
    // true sums follow analytically:  sum(x+b) = S1 + S t,   sum((x+b)^2) = S2 + 2 t S1 + S t^2
    double S1 = 0.0, S2 = 0.0;
    const double ds = (double)s;
    for (int e = threadIdx.x; e < cg * nchunks; e += 32) {
      const int cc = e / nchunks;
      const double2 v = partials[(sample * c + ch0 + cc) * nchunks + (e - cc * nchunks)];
      double t = (double)__ldg(x + (sample * c + ch0 + cc) * s);
      if (conv_bias != nullptr) t += (double)conv_bias[ch0 + cc];
      S1 += v.x;
      S2 += v.y + 2.0 * t * v.x;
      if (e - cc * nchunks == 0) { S1 += ds * t; S2 += ds * t * t; }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      S1 += __shfl_xor_sync(0xffffffffu, S1, d);
      S2 += __shfl_xor_sync(0xffffffffu, S2, d);
    }
    if (threadIdx.x == 0) {
      const double n = (double)cg * ds;
      const double mean = S1 / n;
      const double var = fmax(S2 / n - mean * mean, 0.0);
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      const float ga = gamma != nullptr ? gamma[ch] : 1.0f;
      const float be = beta != nullptr ? beta[ch] : 0.0f;
      const float cb = conv_bias != nullptr ? conv_bias[ch] : 0.0f;
      const float scale = rstd * ga;
      s_ab[0] = scale;
      s_ab[1] = (float)((double)be + ((double)cb - mean) * (double)scale);
    }
  }
  __syncthreads();
  const float A = s_ab[0], Bc = s_ab[1];
  const float *px = x + bc * s;
  auto act = [](float v) { return SWISH ? gnc::swish_fast(v) : v; };
  const bool vec = (s & 3) == 0 && (reinterpret_cast<uintptr_t>(px) & 15) == 0;

  if (MODE == 0) {
    float *py = y + bc * s;
    float local = 0.0f;
    if (vec && (reinterpret_cast<uintptr_t>(py) & 15) == 0) {
      const long long n4 = s >> 2;
      const long long lo = (long long)blockIdx.y * tile4, hi = min(lo + tile4, n4);
      long long i = lo + threadIdx.x;
      {   // software-pipelined: the next 4 loads are in flight while the current 4 values are activated and stored
        float4 v[4], w[4];
        bool have = i + 3 * kGnThreads < hi;
        if (have) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = ld_stream_f4(px + 4 * (i + (long long)j * kGnThreads));
        }
        while (have) {
          const long long ni = i + 4 * kGnThreads;
          const bool more = ni + 3 * kGnThreads < hi;
          if (more) {
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = ld_stream_f4(px + 4 * (ni + (long long)j * kGnThreads));
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[j].x = act(fmaf(v[j].x, A, Bc)); v[j].y = act(fmaf(v[j].y, A, Bc));
            v[j].z = act(fmaf(v[j].z, A, Bc)); v[j].w = act(fmaf(v[j].w, A, Bc));
            *reinterpret_cast<float4 *>(py + 4 * (i + (long long)j * kGnThreads)) = v[j];
            local += (v[j].x + v[j].y) + (v[j].z + v[j].w);
          }
          if (more) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = w[j];
          }
          i = ni;
          have = more;
        }
      }
      for (; i < hi; i += kGnThreads) {
        float4 v = ld_stream_f4(px + 4 * i);
        v.x = act(fmaf(v.x, A, Bc)); v.y = act(fmaf(v.y, A, Bc)); v.z = act(fmaf(v.z, A, Bc)); v.w = act(fmaf(v.w, A, Bc));
        *reinterpret_cast<float4 *>(py + 4 * i) = v;
        local += (v.x + v.y) + (v.z + v.w);
      }
    } else {
      const long long lo = (long long)blockIdx.y * tile4 * 4, hi = min(lo + (long long)tile4 * 4, s);
      for (long long i = lo + threadIdx.x; i < hi; i += kGnThreads) {
        const float v = act(fmaf(px[i], A, Bc));
        py[i] = v;
        local += v;
      }
    }
    if (tile_sums != nullptr) {  // per-tile sum of the outputs (SE squeeze), fixed slot: deterministic
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kGnThreads / 32; ++w) t += s_red[w];
        tile_sums[bc * gridDim.y + blockIdx.y] = t;
      }
    }
  } else {
    // u innermost values -> 1: lanes_per_row = u/4 consecutive lanes own one row (u is a power of two, 4..128)
    const long long m = s / u;
    float *py = y + bc * m;
    const int lpr = u >> 2;
    const long long n4 = s >> 2;
    const long long lo = (long long)blockIdx.y * tile4, hi = min(lo + tile4, n4);   // tile4 is a multiple of lpr
    const float ninf = -__int_as_float(0x7f800000);
    for (long long i0 = lo; i0 < hi; i0 += 4LL * kGnThreads) {   // 4 loads in flight per thread
      float4 v[4];
      long long idx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        idx[j] = i0 + (long long)j * kGnThreads + threadIdx.x;
        if (idx[j] < hi) v[j] = ld_stream_f4(px + 4 * idx[j]);
      }
      // v -> fma(v, A, B) is monotone and Swish falls to its minimum (at -1.278) and rises from it: the largest
      // activated value of a run belongs to its largest or to its smallest input -- two activations per run instead of
      // u (the MUFU pipe, not memory, bounded this loop)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float hi_v = ninf, lo_v = -ninf;
        if (idx[j] < hi) {
          hi_v = fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w));
          lo_v = fminf(fminf(v[j].x, v[j].y), fminf(v[j].z, v[j].w));
        }
        for (int d = 1; d < lpr; d <<= 1) {
          hi_v = fmaxf(hi_v, __shfl_xor_sync(0xffffffffu, hi_v, d));
          lo_v = fminf(lo_v, __shfl_xor_sync(0xffffffffu, lo_v, d));
        }
        if (idx[j] < hi && (threadIdx.x & (lpr - 1)) == 0) py[idx[j] / lpr] = fmaxf(act(fmaf(hi_v, A, Bc)), act(fmaf(lo_v, A, Bc)));
      }
    }
  }
}

static int gn_nchunks(long long rows, long long row_len) {
  long long n = (4LL * sm_count() + rows - 1) / rows;
  const long long by_len = (row_len + 8191) / 8192;
  if (n > by_len) n = by_len;
  if (n > kGnMaxChunks) n = kGnMaxChunks;
  return n < 1 ? 1 : (int)n;
}

static int gn_tiles(long long bc, long long n4, int align4) {
  int tiles = 1;
  while (bc * tiles < 8LL * sm_count() && n4 / (tiles * 2) >= 2 * kGnThreads && tiles < 16384) tiles *= 2;
  long long tile4 = (n4 + tiles - 1) / tiles;
  tile4 = (tile4 + align4 - 1) / align4 * align4;
  return (int)tile4;
}

// ------------------------------------------------------------------------------------------------
// One-pass variant for small groups.  Two thirds of the step's norm layers sit on the coarse stages
// ([16,256,8,8,8], [16,256,256], [16,128,1024] ...): 8-17 MB tensors whose (sample, group) block --
// contiguous in memory -- is at most 32K floats.  One CTA per (sample, group) reads its block ONCE into
// registers (up to 8 x 128-bit loads per thread, all in flight together), reduces the shifted moments
// through shared memory in double, and normalises / activates from the registers: 1 read + 1 write and one
// launch instead of 2 reads + 1 write and two.  Same arithmetic as the two-kernel path.
// ------------------------------------------------------------------------------------------------
constexpr int kOneMaxV = 8;

template <int V, bool SWISH>
__global__ void __launch_bounds__(1024)
gn_onepass_kernel(int c, int s, int groups, float eps, int tiles, const float *__restrict__ x,
                  const float *__restrict__ conv_bias, const float *__restrict__ gamma,
                  const float *__restrict__ beta, float *__restrict__ y, float *__restrict__ tile_sums) {
  const int cg = c / groups;
  const int sample = blockIdx.x / groups, g = blockIdx.x - sample * groups;
  const int ch0 = g * cg;
  const size_t base = ((size_t)sample * c + ch0) * s;
  const int n4 = (cg * s) >> 2;
  const int nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ double s_red[2][32];
  __shared__ float s_stat[2];
  __shared__ double s_dmean;
  __shared__ float s_part[32 * kOneMaxV];

  const float k = __ldg(x + base) + (conv_bias != nullptr ? __ldg(conv_bias + ch0) : 0.0f);
  float4 v[V];
  int chn[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int idx = tid + j * nt;
    chn[j] = ch0 + min((idx * 4) / s, cg - 1);
    v[j] = idx < n4 ? ld_stream_f4(x + base + 4 * (size_t)idx) : make_float4(k, k, k, k);
  }
  float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int idx = tid + j * nt;
    // moments of (x + conv_bias - k); padding lanes hold k with no bias and add 0
    const float kk = (conv_bias != nullptr && idx < n4) ? k - __ldg(conv_bias + chn[j]) : k;
    const float a = v[j].x - kk, b2 = v[j].y - kk, c2 = v[j].z - kk, d2 = v[j].w - kk;
    s1 += (a + b2) + (c2 + d2);
    s2 += (a * a + b2 * b2) + (c2 * c2 + d2 * d2);
  }
  double d1 = (double)s1, d2_ = (double)s2;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    d1 += __shfl_xor_sync(0xffffffffu, d1, d);
    d2_ += __shfl_xor_sync(0xffffffffu, d2_, d);
  }
  if (lane == 0) { s_red[0][warp] = d1; s_red[1][warp] = d2_; }
  __syncthreads();
  if (warp == 0) {
    const int nw = nt >> 5;
    double a1 = lane < nw ? s_red[0][lane] : 0.0, a2 = lane < nw ? s_red[1][lane] : 0.0;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, d);
      a2 += __shfl_xor_sync(0xffffffffu, a2, d);
    }
    if (lane == 0) {
      const double n = (double)cg * (double)s;
      const double m = a1 / n;                       // mean of (x - k)
      const double var = fmax(a2 / n - m * m, 0.0);
      s_dmean = (double)k + m;
      s_stat[1] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
  __syncthreads();
  const double mean = s_dmean;
  const float rstd = s_stat[1];
  auto act = [](float t) { return SWISH ? gnc::swish_fast(t) : t; };
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int idx = tid + j * nt;
    const float ga = gamma != nullptr ? __ldg(gamma + chn[j]) : 1.0f;
    const float be = beta != nullptr ? __ldg(beta + chn[j]) : 0.0f;
    const float cb = conv_bias != nullptr ? __ldg(conv_bias + chn[j]) : 0.0f;
    const float A = rstd * ga;                                                       // same folding as gn_apply_kernel
    const float Bc = (float)((double)be + ((double)cb - mean) * (double)A);
    float4 o;
    o.x = act(fmaf(v[j].x, A, Bc));
    o.y = act(fmaf(v[j].y, A, Bc));
    o.z = act(fmaf(v[j].z, A, Bc));
    o.w = act(fmaf(v[j].w, A, Bc));
    if (idx < n4) *reinterpret_cast<float4 *>(y + base + 4 * (size_t)idx) = o;
    if (tile_sums != nullptr) {   // s % 128 == 0: the 32 lanes of a warp row share one channel
      float t = idx < n4 ? (o.x + o.y) + (o.z + o.w) : 0.0f;
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
      if (lane == 0) s_part[j * 32 + warp] = t;
    }
  }
  if (tile_sums != nullptr) {
    __syncthreads();
    if (tid < cg) {   // fixed order: deterministic
      float acc = 0.0f;
      const int nw = nt >> 5;
      for (int j = 0; j < V; ++j)
        for (int w = 0; w < nw; ++w) {
          const int idx = w * 32 + j * nt;
          if (idx < n4 && (idx * 4) / s == tid) acc += s_part[j * 32 + w];
        }
      float *ts = tile_sums + ((size_t)sample * c + ch0 + tid) * tiles;
      ts[0] = acc;
      for (int q = 1; q < tiles; ++q) ts[q] = 0.0f;
    }
  }
}

template <int V>
static void launch_onepass(bool swish, int ctas, int nt, cudaStream_t st, int c, int s, int groups, float eps, int tiles,
                           const float *x, const float *cb, const float *ga, const float *be, float *y, float *ts) {
  if (swish) gn_onepass_kernel<V, true><<<ctas, nt, 0, st>>>(c, s, groups, eps, tiles, x, cb, ga, be, y, ts);
  else gn_onepass_kernel<V, false><<<ctas, nt, 0, st>>>(c, s, groups, eps, tiles, x, cb, ga, be, y, ts);
}

// ------------------------------------------------------------------------------------------------
// Channels-last variant, x f32[b][s][c] (the layout cuDNN's tensor-core convolutions work in; keeping the
// voxel branch in it removes the NCDHW<->NDHWC transposes cuDNN otherwise wraps around every Conv3d).
// A thread owns 4 consecutive channels (one 128-bit column, the same for every row it visits), so the
// per-channel constants live in registers and a pass of 256 threads covers 256*4/c whole rows -- fully
// contiguous 4 KB reads and writes.  Statistics are per channel (shifted by the channel's first value),
// folded per group, with the conv bias analytically, in the apply kernel's prologue, as above.
// Needs c in {16, 32, 64, 128, 256} (c/4 divides 256, c <= 256).
// ------------------------------------------------------------------------------------------------
constexpr int kClThreads = 256;
constexpr int kClMaxChunks = 32;

static inline bool gn_cl_supported(int c, int groups) {
  return c >= 16 && c <= 256 && (c & (c - 1)) == 0 && c % groups == 0 && (c / groups) >= 1;
}
static inline int gn_cl_chunks(int b, long long s, int c) {
  const int rpp = kClThreads / (c / 4);
  long long want = (2LL * sm_count() + b - 1) / b;
  long long maxc = (s + rpp - 1) / rpp;
  return (int)max(1LL, min(min(want, (long long)kClMaxChunks), maxc));
}
static inline bool gn_cl_onepass(int b, int c, long long s, int groups) {
  const int cg = c / groups;
  const long long gelems = (long long)cg * s;
  return cg % 4 == 0 && cg <= 128 && (32 % (cg / 4) == 0 || (cg / 4) % 32 == 0) && gelems <= 4LL * 1024 * 8 &&
         (long long)b * groups <= 0x7fffffffLL && s <= 0x7fffffffLL / c;
}
static inline int gn_tune(const char *name, int dflt) {   // tuning hooks (tools/gn_bench.py)
  const char *e = std::getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}
static inline int gn_cl_tiles(int b, long long s, int c) {
  static const int per_sm = gn_tune("BDM_GN_CL_CTAS_PER_SM", 8);
  const int rpp = kClThreads / (c / 4);
  long long want = ((long long)per_sm * sm_count() + b - 1) / b;
  long long maxc = (s + 4LL * rpp - 1) / (4LL * rpp);
  return (int)max(1LL, min(min(want, 64LL), maxc));
}

__global__ void __launch_bounds__(kClThreads)
gn_cl_stats_kernel(int c, long long s, int nchunks, const float *__restrict__ x, double2 *__restrict__ partials) {
  __shared__ double2 red[kClThreads * 4];   // [rows per pass][c]
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int c4 = c >> 2, rpp = kClThreads / c4;
  const int q = threadIdx.x % c4, r0 = threadIdx.x / c4;
  const float *px = x + (size_t)b * s * c;
  long long per = (s + nchunks - 1) / nchunks;
  const long long lo = min((long long)chunk * per, s), hi = min(lo + per, s);
  const float4 k = __ldg(reinterpret_cast<const float4 *>(px) + q);   // the channels' values at the first voxel
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  double d1[4] = {0., 0., 0., 0.}, d2[4] = {0., 0., 0., 0.};
  int since = 0;
  auto add = [&](const float4 &v) {
    const float e0 = v.x - k.x, e1 = v.y - k.y, e2 = v.z - k.z, e3 = v.w - k.w;
    s1[0] += e0; s1[1] += e1; s1[2] += e2; s1[3] += e3;
    s2[0] += e0 * e0; s2[1] += e1 * e1; s2[2] += e2 * e2; s2[3] += e3 * e3;
  };
  long long row = lo + r0;
  for (; row + 3LL * rpp < hi; row += 4LL * rpp) {   // 4 loads in flight per thread
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ld_stream_f4(px + (size_t)(row + (long long)j * rpp) * c + 4 * q);
#pragma unroll
    for (int j = 0; j < 4; ++j) add(v[j]);
    if (++since == 16) {   // bound the fp32 run length (64 values)
#pragma unroll
      for (int j = 0; j < 4; ++j) { d1[j] += s1[j]; d2[j] += s2[j]; s1[j] = s2[j] = 0.f; }
      since = 0;
    }
  }
  for (; row < hi; row += rpp) add(ld_stream_f4(px + (size_t)row * c + 4 * q));
#pragma unroll
  for (int j = 0; j < 4; ++j) red[r0 * c + 4 * q + j] = make_double2(d1[j] + s1[j], d2[j] + s2[j]);
  __syncthreads();
  if (threadIdx.x < c) {
    double a1 = 0.0, a2 = 0.0;
    for (int r = 0; r < rpp; ++r) { a1 += red[r * c + threadIdx.x].x; a2 += red[r * c + threadIdx.x].y; }
    partials[((size_t)b * nchunks + chunk) * c + threadIdx.x] = make_double2(a1, a2);
  }
}

// Producer statistics at group level, f64[b][blocks][groups][2] = (sum, sum of squares) of the tensor the norm is applied to
// (bias included) over disjoint blocks of voxels -- what bdm_conv3_tc05 leaves per unit.  Folded by 256 threads into the
// per-channel table the kernels below work from: the group's sums in its first channel's slot, zeros in the others.
__device__ __forceinline__ void fold_group_partials(const double2 *__restrict__ gp, int b, int blocks, int groups, int c, int cg,
                                                    double2 *slice /* [256] */, double2 *chan /* [c] */) {
  const int t = threadIdx.x;
  const int nsl = 256 / groups, g = t % groups, sl = t / groups;
  double S1 = 0.0, S2 = 0.0;
  if (sl < nsl) {
    for (int blk = sl; blk < blocks; blk += nsl) {
      const double2 v = gp[((size_t)b * blocks + blk) * groups + g];
      S1 += v.x; S2 += v.y;
    }
  }
  slice[t] = make_double2(S1, S2);
  __syncthreads();
  if (t < c) {
    double A1 = 0.0, A2 = 0.0;
    if (t % cg == 0) {
      for (int k = 0; k < nsl; ++k) { A1 += slice[k * groups + t / cg].x; A2 += slice[k * groups + t / cg].y; }
    }
    chan[t] = make_double2(A1, A2);
  }
  __syncthreads();
}

// STORE == false: nothing is written back except the per-tile sums and (tile 0) the per-channel coefficients
// coef[b][c] = (A, B) of y = act(x * A + B): the consumer normalises on the fly (bdm_trilinear_devoxelize_cl_norm).
template <bool SWISH, int UNR, bool STORE = true>
__global__ void __launch_bounds__(kClThreads)
gn_cl_apply_kernel(int c, long long s, int groups, int nchunks, int pstride, int ntiles, float eps, int zero_shift,
                   const float *__restrict__ x, const float *__restrict__ conv_bias,
                   const float *__restrict__ gamma, const float *__restrict__ beta,
                   const double2 *__restrict__ partials, float *__restrict__ y, float *__restrict__ tile_sums,
                   float2 *__restrict__ coef = nullptr) {
  __shared__ double2 chan[256];          // per-channel folded (sum, sumsq) of x + conv_bias
  __shared__ double2 grp[256];           // per-group (mean, rstd)
  __shared__ float2 ab[256];             // per-channel (A, B): y = act(x * A + B)
  __shared__ float sums[kClThreads * 4]; // [rows per pass][c] for the SE squeeze
  const int b = gridDim.y - 1 - blockIdx.y, tile = gridDim.x - 1 - blockIdx.x;     // back to front: see gn_stats_kernel
  const int c4 = c >> 2, rpp = kClThreads / c4, cg = c / groups;
  const float *px = x + (size_t)b * s * c;
  const int t = threadIdx.x;
  if (nchunks < 0) {
    fold_group_partials(partials, b, -nchunks, groups, c, cg, grp, chan);      // group-level producer statistics (bias included)
  } else {
    {
      // the sample's partial moments, folded by all 256 threads: thread (channel t % c, slice t / c) sums every
      // (256/c)-th block, the slices are combined in order below (producer-made statistics come in up to 128 blocks;
      // with c threads alone this prologue cost every CTA several microseconds)
      const int nsl = kClThreads / c, tc = t % c, sl = t / c;
      double S1 = 0.0, S2 = 0.0;
      for (int ch = sl; ch < nchunks; ch += nsl) {
        const double2 v = partials[((size_t)b * pstride + ch) * c + tc];   // pstride = blocks per sample in memory
        S1 += v.x; S2 += v.y;
      }
      grp[t] = make_double2(S1, S2);          // grp doubles as the slice buffer until the group fold below
    }
    __syncthreads();
    if (t < c) {
      // true sums of (x + bias) from the shifted partials: with tt = k_c + bias_c,
      //   sum = S1 + s*tt,  sumsq = S2 + 2*tt*S1 + s*tt^2
      double S1 = 0.0, S2 = 0.0;
      for (int sl = 0; sl < kClThreads / c; ++sl) { S1 += grp[sl * c + t].x; S2 += grp[sl * c + t].y; }
      double tt = zero_shift ? 0.0 : (double)__ldg(px + t);   // producer-made partials are of x itself
      if (conv_bias != nullptr) tt += (double)conv_bias[t];
      const double ds = (double)s;
      chan[t] = make_double2(S1 + ds * tt, S2 + 2.0 * tt * S1 + ds * tt * tt);
    }
    __syncthreads();
  }
  if (t < groups) {
    double S1 = 0.0, S2 = 0.0;
    for (int j = 0; j < cg; ++j) { S1 += chan[t * cg + j].x; S2 += chan[t * cg + j].y; }
    const double n = (double)cg * (double)s;
    const double mean = S1 / n;
    const double var = fmax(S2 / n - mean * mean, 0.0);
    grp[t] = make_double2(mean, 1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if (t < c) {
    const double2 g = grp[t / cg];
    const float ga = gamma != nullptr ? gamma[t] : 1.0f;
    const float be = beta != nullptr ? beta[t] : 0.0f;
    const float cb = conv_bias != nullptr ? conv_bias[t] : 0.0f;
    const float A = (float)g.y * ga;
    ab[t] = make_float2(A, (float)((double)be + ((double)cb - g.x) * (double)A));
    if (!STORE && coef != nullptr && tile == 0) coef[(size_t)b * c + t] = ab[t];
  }
  __syncthreads();
  const int q = t % c4, r0 = t / c4;
  const float2 p0 = ab[4 * q], p1 = ab[4 * q + 1], p2 = ab[4 * q + 2], p3 = ab[4 * q + 3];
  auto act = [](float v) { return SWISH ? gnc::swish_fast(v) : v; };
  long long per = (s + ntiles - 1) / ntiles;
  per = (per + rpp - 1) / rpp * rpp;
  const long long lo = min((long long)tile * per, s), hi = min(lo + per, s);
  float *py = y + (size_t)b * s * c;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  long long row = lo + r0;
  // software-pipelined: the next UNR loads are issued before the current UNR values are activated and stored, so a
  // thread always has loads in flight (a plain load -> compute -> store loop left HBM idle during the compute phase:
  // 4.1 TB/s against 6.5 TB/s for a copy of the same tensor)
  {
    const long long step = (long long)UNR * rpp;
    float4 v[UNR], w[UNR];
    bool have = row + (long long)(UNR - 1) * rpp < hi;
    if (have) {
#pragma unroll
      for (int j = 0; j < UNR; ++j) v[j] = ld_stream_f4(px + (size_t)(row + (long long)j * rpp) * c + 4 * q);
    }
    while (have) {
      const long long nrow = row + step;
      const bool more = nrow + (long long)(UNR - 1) * rpp < hi;
      if (more) {
#pragma unroll
        for (int j = 0; j < UNR; ++j) w[j] = ld_stream_f4(px + (size_t)(nrow + (long long)j * rpp) * c + 4 * q);
      }
#pragma unroll
      for (int j = 0; j < UNR; ++j) {
        v[j].x = act(fmaf(v[j].x, p0.x, p0.y)); v[j].y = act(fmaf(v[j].y, p1.x, p1.y));
        v[j].z = act(fmaf(v[j].z, p2.x, p2.y)); v[j].w = act(fmaf(v[j].w, p3.x, p3.y));
        if (STORE) *reinterpret_cast<float4 *>(py + (size_t)(row + (long long)j * rpp) * c + 4 * q) = v[j];
        // (__fadd_rn: never contracted with the multiply inside the activation, so the sums are those of the rounded
        // values whether or not they are also stored)
        acc[0] = __fadd_rn(acc[0], v[j].x); acc[1] = __fadd_rn(acc[1], v[j].y);
        acc[2] = __fadd_rn(acc[2], v[j].z); acc[3] = __fadd_rn(acc[3], v[j].w);
      }
      if (more) {
#pragma unroll
        for (int j = 0; j < UNR; ++j) v[j] = w[j];
      }
      row = nrow;
      have = more;
    }
  }
  for (; row < hi; row += rpp) {
    float4 v = ld_stream_f4(px + (size_t)row * c + 4 * q);
    v.x = act(fmaf(v.x, p0.x, p0.y)); v.y = act(fmaf(v.y, p1.x, p1.y));
    v.z = act(fmaf(v.z, p2.x, p2.y)); v.w = act(fmaf(v.w, p3.x, p3.y));
    if (STORE) *reinterpret_cast<float4 *>(py + (size_t)row * c + 4 * q) = v;
    acc[0] = __fadd_rn(acc[0], v.x); acc[1] = __fadd_rn(acc[1], v.y); acc[2] = __fadd_rn(acc[2], v.z); acc[3] = __fadd_rn(acc[3], v.w);
  }
  if (tile_sums != nullptr) {   // per-tile, per-channel sums of y in fixed slots: deterministic
#pragma unroll
    for (int j = 0; j < 4; ++j) sums[r0 * c + 4 * q + j] = acc[j];
    __syncthreads();
    if (t < c) {
      float a = 0.0f;
      for (int r = 0; r < rpp; ++r) a += sums[r * c + t];
      tile_sums[((size_t)b * ntiles + tile) * c + t] = a;
    }
  }
}

// One-pass channels-last variant: one CTA per (sample, group).  The group's data are S segments of cg
// contiguous channels (cg*4 bytes each, S = voxels); float4 unit idx -> voxel idx / (cg/4), quad idx % (cg/4).
// The block size is a multiple of 32 and cg/4 divides 32, so a thread's quad -- its 4 channels -- is the same
// for all of its units: gamma / beta / conv bias are loaded once as float4.
template <int V, bool SWISH>
__global__ void __launch_bounds__(1024)
gn_onepass_cl_kernel(int c, int s, int groups, float eps, int ntiles, const float *__restrict__ x,
                     const float *__restrict__ conv_bias, const float *__restrict__ gamma,
                     const float *__restrict__ beta, float *__restrict__ y, float *__restrict__ tile_sums) {
  const int cg = c / groups, q4 = cg >> 2;
  const int sample = blockIdx.x / groups, g = blockIdx.x - sample * groups;
  const int n4 = q4 * s;
  const int nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int quad = tid % q4;
  const int ch = g * cg + 4 * quad;                       // this thread's 4 channels
  const size_t base = (size_t)sample * s * c + ch;        // + voxel * c
  __shared__ double s_red[2][32];
  __shared__ double s_dmean;
  __shared__ float s_rstd;
  __shared__ float s_sum[1024 * 4];

  const float4 cb = conv_bias != nullptr ? __ldg(reinterpret_cast<const float4 *>(conv_bias + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float k = __ldg(x + (size_t)sample * s * c + g * cg) + (conv_bias != nullptr ? __ldg(conv_bias + g * cg) : 0.0f);
  float4 v[V];
  float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int idx = tid + j * nt;
    v[j] = idx < n4 ? ld_stream_f4(x + base + (size_t)(idx / q4) * c) : make_float4(k - cb.x, k - cb.y, k - cb.z, k - cb.w);
  }
#pragma unroll
  for (int j = 0; j < V; ++j) {   // moments of (x + conv_bias - k); padding units add 0
    const float a = v[j].x + cb.x - k, b2 = v[j].y + cb.y - k, c2 = v[j].z + cb.z - k, d2 = v[j].w + cb.w - k;
    s1 += (a + b2) + (c2 + d2);
    s2 += (a * a + b2 * b2) + (c2 * c2 + d2 * d2);
  }
  double d1 = (double)s1, d2_ = (double)s2;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    d1 += __shfl_xor_sync(0xffffffffu, d1, d);
    d2_ += __shfl_xor_sync(0xffffffffu, d2_, d);
  }
  if (lane == 0) { s_red[0][warp] = d1; s_red[1][warp] = d2_; }
  __syncthreads();
  if (warp == 0) {
    const int nw = nt >> 5;
    double a1 = lane < nw ? s_red[0][lane] : 0.0, a2 = lane < nw ? s_red[1][lane] : 0.0;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, d);
      a2 += __shfl_xor_sync(0xffffffffu, a2, d);
    }
    if (lane == 0) {
      const double n = (double)cg * (double)s;
      const double m = a1 / n;
      const double var = fmax(a2 / n - m * m, 0.0);
      s_dmean = (double)k + m;
      s_rstd = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
  __syncthreads();
  const double mean = s_dmean;
  const float rstd = s_rstd;
  const float4 ga = gamma != nullptr ? __ldg(reinterpret_cast<const float4 *>(gamma + ch)) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 be = beta != nullptr ? __ldg(reinterpret_cast<const float4 *>(beta + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float A0 = rstd * ga.x, A1 = rstd * ga.y, A2 = rstd * ga.z, A3 = rstd * ga.w;    // same folding as gn_apply_kernel
  const float B0 = (float)((double)be.x + ((double)cb.x - mean) * (double)A0);
  const float B1 = (float)((double)be.y + ((double)cb.y - mean) * (double)A1);
  const float B2 = (float)((double)be.z + ((double)cb.z - mean) * (double)A2);
  const float B3 = (float)((double)be.w + ((double)cb.w - mean) * (double)A3);
  auto act = [](float t) { return SWISH ? gnc::swish_fast(t) : t; };
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int idx = tid + j * nt;
    if (idx < n4) {
      float4 o;
      o.x = act(fmaf(v[j].x, A0, B0));
      o.y = act(fmaf(v[j].y, A1, B1));
      o.z = act(fmaf(v[j].z, A2, B2));
      o.w = act(fmaf(v[j].w, A3, B3));
      *reinterpret_cast<float4 *>(y + base + (size_t)(idx / q4) * c) = o;
      acc[0] += o.x; acc[1] += o.y; acc[2] += o.z; acc[3] += o.w;
    }
  }
  if (tile_sums != nullptr) {   // [b][ntiles][c]: the whole sum goes to tile 0, fixed order
#pragma unroll
    for (int j = 0; j < 4; ++j) s_sum[tid * 4 + j] = acc[j];
    __syncthreads();
    if (tid < cg) {
      const int qd = tid >> 2, comp = tid & 3;
      float a = 0.0f;
      for (int t2 = qd; t2 < nt; t2 += q4) a += s_sum[t2 * 4 + comp];
      float *ts = tile_sums + (size_t)sample * ntiles * c + g * cg + tid;
      ts[0] = a;
      for (int q = 1; q < ntiles; ++q) ts[(size_t)q * c] = 0.0f;
    }
  }
}

template <int V>
static void launch_onepass_cl(bool swish, int ctas, int nt, cudaStream_t st, int c, int s, int groups, float eps, int ntiles,
                              const float *x, const float *cb, const float *ga, const float *be, float *y, float *ts) {
  if (swish) gn_onepass_cl_kernel<V, true><<<ctas, nt, 0, st>>>(c, s, groups, eps, ntiles, x, cb, ga, be, y, ts);
  else gn_onepass_cl_kernel<V, false><<<ctas, nt, 0, st>>>(c, s, groups, eps, ntiles, x, cb, ga, be, y, ts);
}

// ------------------------------------------------------------------------------------------------
// Squeeze-excite gate (modules/se.py:8-19) from the per-channel sums the norm kernels emit:
//   pooled = sums / count;  gate = sigmoid(W2 . act(W1 . pooled)),  act = ReLU or Swish, no biases.
// Through torch that is 6 tiny launches per block (sum over tiles, divide, two Linears, two activations)
// on the critical path of every PVConv; here one CTA per sample does it.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
se_gate_kernel(int c, int hidden, int tiles, float count, long long sb, long long st_, long long sc,
               const float *__restrict__ sums, const float *__restrict__ w1, const float *__restrict__ w2,
               int use_relu, float *__restrict__ gate) {
  extern __shared__ float sh[];   // pooled[c], hid[hidden], part[256]
  float *pooled = sh, *hid = sh + c, *part = hid + hidden;
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // everything here is latency-bound: independent loads are issued together, and the sum over tiles is
  // split over 256 / c thread groups (fixed assignment and order: deterministic)
  if (c <= 256) {
    const int groups = 256 / c, ch = t % c, grp = t / c;
    float a = 0.0f;
    if (grp < groups) {
      int q = grp;
      for (; q + 3 * groups < tiles; q += 4 * groups) {
        const float v0 = __ldg(sums + b * sb + (long long)q * st_ + ch * sc);
        const float v1 = __ldg(sums + b * sb + (long long)(q + groups) * st_ + ch * sc);
        const float v2 = __ldg(sums + b * sb + (long long)(q + 2 * groups) * st_ + ch * sc);
        const float v3 = __ldg(sums + b * sb + (long long)(q + 3 * groups) * st_ + ch * sc);
        a += (v0 + v1) + (v2 + v3);
      }
      for (; q < tiles; q += groups) a += __ldg(sums + b * sb + (long long)q * st_ + ch * sc);
    }
    part[t] = a;
    __syncthreads();
    if (t < c) {
      float tot = 0.0f;
      for (int g2 = 0; g2 < groups; ++g2) tot += part[g2 * c + t];
      pooled[t] = tot / count;
    }
  } else {
    for (int ch = t; ch < c; ch += 256) {
      float a = 0.0f;
      for (int q = 0; q < tiles; ++q) a += __ldg(sums + b * sb + (long long)q * st_ + ch * sc);
      pooled[ch] = a / count;
    }
  }
  __syncthreads();
  for (int h = warp; h < hidden; h += 8) {
    float a = 0.0f;
    for (int ch = lane; ch < c; ch += 32) a = fmaf(__ldg(w1 + (size_t)h * c + ch), pooled[ch], a);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if (lane == 0) hid[h] = use_relu ? fmaxf(a, 0.0f) : a / (1.0f + expf(-a));
  }
  __syncthreads();
  const bool vec = (hidden & 3) == 0 && (reinterpret_cast<uintptr_t>(w2) & 15) == 0;
  for (int ch = t; ch < c; ch += 256) {
    float a = 0.0f;
    if (vec) {   // the row's loads are independent: one latency, not `hidden` of them
      const float4 *wr = reinterpret_cast<const float4 *>(w2 + (size_t)ch * hidden);
#pragma unroll 8
      for (int h4 = 0; h4 < hidden / 4; ++h4) {
        const float4 w = __ldg(wr + h4);
        a = fmaf(w.x, hid[4 * h4], a);
        a = fmaf(w.y, hid[4 * h4 + 1], a);
        a = fmaf(w.z, hid[4 * h4 + 2], a);
        a = fmaf(w.w, hid[4 * h4 + 3], a);
      }
    } else {
      for (int h = 0; h < hidden; ++h) a = fmaf(__ldg(w2 + (size_t)ch * hidden + h), hid[h], a);
    }
    gate[(size_t)b * c + ch] = 1.0f / (1.0f + expf(-a));
  }
}


// ---- thread-block-cluster launches (groupnorm_cluster.cuh) ----
// Whether a cluster of `cl` CTAs of `func` with `smem` dynamic bytes can be co-scheduled on this device; the answer
// (and the one-time function attributes it needs) is cached per (function, cluster size).
static bool cluster_launchable(const void *func, int cl, int threads, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, int> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  int &state = cache[std::make_pair(func, cl * 64 + dev)];
  if (state != 0) return state > 0;
  state = -1;
  if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
  if (cl > 8 && cudaFuncSetAttribute(func, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return false; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cl, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, func, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
  if (nclusters >= 1) state = 1;
  return state > 0;
}

template <typename... Args>
static cudaError_t launch_cluster(void (*kernel)(Args...), dim3 grid, int threads, size_t smem, int cl, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

}  // namespace bdm

extern "C" size_t bdm_groupnorm_workspace_bytes(int b, int c, long long s) {
  if (b <= 0 || c <= 0 || s <= 0) return 16;
  return sizeof(double2) * (size_t)b * c * bdm::gn_nchunks((long long)b * c, s);
}

// number of per-(sample,channel) tile sums bdm_groupnorm_act writes when tile_sums != NULL
extern "C" int bdm_groupnorm_tiles(int b, int c, long long s) {
  if (b <= 0 || c <= 0 || s <= 0) return 1;
  const long long n4 = (s + 3) >> 2;
  const int tile4 = bdm::gn_tiles((long long)b * c, n4, 1);
  return (int)((n4 + tile4 - 1) / tile4);
}

extern "C" int bdm_groupnorm_act(int b, int c, long long s, int groups, float eps, int swish, int max_over_u,
                                 const float *x, const float *conv_bias, const float *gamma, const float *beta,
                                 float *y, float *tile_sums, void *workspace, size_t workspace_bytes,
                                 bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 1 && s >= 0 && groups >= 1 && c % groups == 0);
  if (b == 0 || s == 0) return BDM_OK;
  BDM_CHECK_PTR(x); BDM_CHECK_PTR(y); BDM_CHECK_PTR(workspace);
  const long long rows = (long long)b * c;
  BDM_CHECK_SIZE(rows <= 65535LL * 64);
  if (workspace_bytes < bdm_groupnorm_workspace_bytes(b, c, s)) return BDM_ERR_WORKSPACE_TOO_SMALL;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return BDM_ERR_MISALIGNED;
  if (max_over_u) {
    // power of two in [4,128], rows of u values contiguous and 16-byte aligned
    BDM_CHECK_SIZE(max_over_u >= 4 && max_over_u <= 128 && (max_over_u & (max_over_u - 1)) == 0 && s % max_over_u == 0);
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return BDM_ERR_MISALIGNED;
    BDM_CHECK_SIZE(tile_sums == nullptr);
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // small groups: one CTA per (sample, group), one pass
  {
    const long long gelems = (long long)(c / groups) * s;
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (!max_over_u && aligned && s % 4 == 0 && gelems <= 4LL * 1024 * kOneMaxV && (long long)b * groups <= 0x7fffffffLL &&
        (tile_sums == nullptr || (s % 128 == 0 && c / groups <= 1024 && c / groups <= 32 * 32))) {
      const int n4 = (int)(gelems >> 2);
      int v = 1;
      while (v < kOneMaxV && (long long)v * 1024 < n4) v <<= 1;
      int nt = (n4 + v - 1) / v;
      nt = min(1024, max(32, (nt + 31) & ~31));
      if (tile_sums != nullptr) nt = max(nt, min(1024, ((c / groups) + 31) & ~31));
      const int tiles = tile_sums != nullptr ? bdm_groupnorm_tiles(b, c, s) : 1;
      const int ctas = b * groups;
      if (v == 1) launch_onepass<1>(swish != 0, ctas, nt, st, c, (int)s, groups, eps, tiles, x, conv_bias, gamma, beta, y, tile_sums);
      else if (v == 2) launch_onepass<2>(swish != 0, ctas, nt, st, c, (int)s, groups, eps, tiles, x, conv_bias, gamma, beta, y, tile_sums);
      else if (v == 4) launch_onepass<4>(swish != 0, ctas, nt, st, c, (int)s, groups, eps, tiles, x, conv_bias, gamma, beta, y, tile_sums);
      else launch_onepass<8>(swish != 0, ctas, nt, st, c, (int)s, groups, eps, tiles, x, conv_bias, gamma, beta, y, tile_sums);
      g_last_launches = 1;
      BDM_RETURN_LAUNCH_STATUS();
    }
  }
  // mid-sized groups (64 KB .. 256 KB; any size up to that when the max over neighbours is wanted): a cluster of up to 4
  // CTAs per (sample, group), one pass.  (Measured on B200, 32 shapes: [32,512,16,32]+max 103 -> 33 us, [32,256,64,32]+max
  // 71 -> 53 us; clusters of 8 / 16 for 0.5 - 1 MB groups were SLOWER than the two-kernel path -- 192 vs 146 us for
  // [32,64,1024,32] -- and are not used.)
  {
    const int cg = c / groups;
    const long long gelems = (long long)cg * s;
    const int cl = gnc::cf_cluster_size(gelems);
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    static const bool enabled = [] { const char *e = std::getenv("BDM_GN_CLUSTER"); return e == nullptr || e[0] != '0'; }();
    if (enabled && cl > 0 && cl <= 4 && tile_sums == nullptr && aligned && s % 4 == 0 && cg <= gnc::kCfMaxCg && s <= 0x7fffffffLL / 4 &&
        (long long)b * groups <= 65535) {
      const int n4 = (int)(gelems >> 2);
      const int lpr = max_over_u ? max_over_u / 4 : 1;
      int per4 = (n4 + cl - 1) / cl;
      per4 = (per4 + lpr - 1) / lpr * lpr;
      const size_t smem = (size_t)per4 * 16;
      const dim3 grid(cl, (unsigned)(b * groups));
#define BDM_GNC_LAUNCH(SW, MODE)                                                                                   \
      do {                                                                                                         \
        auto kfn = gnc::gn_cluster_kernel<SW, MODE>;                                                               \
        if (cluster_launchable(reinterpret_cast<const void *>(kfn), cl, gnc::kCfThreads, (size_t)gnc::kCfSliceFloats * 4)) { \
          cudaError_t e = launch_cluster(kfn, grid, gnc::kCfThreads, smem, cl, st, c, (int)s, groups, eps, max_over_u, per4, x, \
                                         conv_bias, gamma, beta, y);                                             \
          if (e != cudaSuccess) return (int)e;                                                                     \
          g_last_launches = 1;                                                                                     \
          BDM_RETURN_LAUNCH_STATUS();                                                                              \
        }                                                                                                          \
      } while (0)
      if (max_over_u) { if (swish) BDM_GNC_LAUNCH(true, 1); else BDM_GNC_LAUNCH(false, 1); }
      else            { if (swish) BDM_GNC_LAUNCH(true, 0); else BDM_GNC_LAUNCH(false, 0); }
#undef BDM_GNC_LAUNCH
    }
  }
  const int nchunks = gn_nchunks(rows, s);
  double2 *partials = static_cast<double2 *>(workspace);
  // rows can exceed the 65535 limit of gridDim.y: fold them in slabs
  for (long long r0 = 0; r0 < rows; r0 += 65535) {
    const unsigned nr = (unsigned)min(65535LL, rows - r0);
    gn_stats_kernel<<<dim3(nchunks, nr), kGnThreads, 0, st>>>(s, nchunks, x + r0 * s, partials + r0 * nchunks);
  }
  const long long n4 = (s + 3) >> 2;
  const int tile4 = gn_tiles(rows, n4, max_over_u ? max_over_u / 4 : 1);
  const unsigned tiles = (unsigned)((n4 + tile4 - 1) / tile4);
  BDM_CHECK_SIZE(rows <= 0x7fffffffLL && tiles <= 65535);
  const dim3 grid((unsigned)rows, tiles);
#define BDM_GN_LAUNCH(SW, MODE)                                                                               \
  gn_apply_kernel<SW, MODE><<<grid, kGnThreads, 0, st>>>(c, s, groups, nchunks, eps, tile4, max_over_u, x,    \
                                                          conv_bias, gamma, beta, partials, y, tile_sums)
  if (max_over_u) { if (swish) BDM_GN_LAUNCH(true, 1); else BDM_GN_LAUNCH(false, 1); }
  else            { if (swish) BDM_GN_LAUNCH(true, 0); else BDM_GN_LAUNCH(false, 0); }
#undef BDM_GN_LAUNCH
  g_last_launches = 2;
  BDM_RETURN_LAUNCH_STATUS();
}

// ---- channels-last entry points: x, y f32[b][s][c]; tile_sums f32[b][tiles][c] or NULL ----
extern "C" int bdm_groupnorm_cl_supported(int c, int groups) { return bdm::gn_cl_supported(c, groups) ? 1 : 0; }

extern "C" size_t bdm_groupnorm_cl_workspace_bytes(int b, int c, long long s) {
  if (b <= 0 || c < 16 || s <= 0) return 16;
  return sizeof(double2) * (size_t)b * bdm::gn_cl_chunks(b, s, c) * c;
}

extern "C" int bdm_groupnorm_cl_tiles(int b, int c, long long s, int groups) {
  if (b <= 0 || c < 16 || s <= 0 || groups < 1 || c % groups != 0) return 1;
  if (bdm::gn_cl_onepass(b, c, s, groups)) return 1;   // the one-pass kernel emits whole sums
  return bdm::gn_cl_tiles(b, s, c);
}

// precomputed_chunks > 0: `workspace` already holds f64[b][precomputed_chunks][c][2] = per-channel (sum, sum of
// squares) of x over disjoint blocks of voxels, written by the producer of x (bdm_sparse_conv3_gather's
// `stats`); the statistics pass over x is skipped.
extern "C" int bdm_groupnorm_act_cl(int b, int c, long long s, int groups, float eps, int swish, const float *x,
                                    const float *conv_bias, const float *gamma, const float *beta, float *y,
                                    float *tile_sums, void *workspace, size_t workspace_bytes,
                                    int precomputed_chunks, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && s >= 0 && groups >= 1 && gn_cl_supported(c, groups) && b <= 65535);
  BDM_CHECK_SIZE(precomputed_chunks >= 0 || (conv_bias == nullptr && groups <= 32));   // group partials include the bias
  if (b == 0 || s == 0) return BDM_OK;
  BDM_CHECK_PTR(x); BDM_CHECK_PTR(y); BDM_CHECK_PTR(workspace);
  // (a group small enough for the one-pass kernel ignores producer statistics: one read either way, and the
  // tile count reported by bdm_groupnorm_cl_tiles stays consistent)
  if (precomputed_chunks != 0 && !gn_cl_onepass(b, c, s, groups)) {
    const size_t need = precomputed_chunks > 0 ? sizeof(double2) * (size_t)b * precomputed_chunks * c
                                               : sizeof(double2) * (size_t)b * (size_t)(-precomputed_chunks) * groups;
    if (workspace_bytes < need) return BDM_ERR_WORKSPACE_TOO_SMALL;
    if (((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) != 0)
      return BDM_ERR_MISALIGNED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int ntiles = gn_cl_tiles(b, s, c);   // precomputed statistics: always the two-kernel apply
    const double2 *partials = static_cast<const double2 *>(workspace);
    const int use_chunks = precomputed_chunks;
    g_last_launches = 1;
    if (swish)
      gn_cl_apply_kernel<true, 4><<<dim3(ntiles, b), kClThreads, 0, st>>>(c, s, groups, use_chunks, precomputed_chunks, ntiles, eps, 1, x,
                                                                     conv_bias, gamma, beta, partials, y, tile_sums);
    else
      gn_cl_apply_kernel<false, 4><<<dim3(ntiles, b), kClThreads, 0, st>>>(c, s, groups, use_chunks, precomputed_chunks, ntiles, eps, 1, x,
                                                                      conv_bias, gamma, beta, partials, y, tile_sums);
    BDM_RETURN_LAUNCH_STATUS();
  }
  if (!gn_cl_onepass(b, c, s, groups) && workspace_bytes < bdm_groupnorm_cl_workspace_bytes(b, c, s))
    return BDM_ERR_WORKSPACE_TOO_SMALL;
  if (((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) != 0)
    return BDM_ERR_MISALIGNED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nchunks = gn_cl_chunks(b, s, c), ntiles = bdm_groupnorm_cl_tiles(b, c, s, groups);
  {   // small groups: one CTA per (sample, group), one pass
    const int cg = c / groups;
    const long long gelems = (long long)cg * s;
    if (gn_cl_onepass(b, c, s, groups)) {
      const int n4 = (int)(gelems >> 2);
      int v = 1;
      while (v < kOneMaxV && (long long)v * 1024 < n4) v <<= 1;
      int nt = (n4 + v - 1) / v;
      nt = min(1024, max(32, (nt + 31) & ~31));
      nt = max(nt, min(1024, (cg + 31) & ~31));
      const int ctas = b * groups;
      if (v == 1) launch_onepass_cl<1>(swish != 0, ctas, nt, st, c, (int)s, groups, eps, ntiles, x, conv_bias, gamma, beta, y, tile_sums);
      else if (v == 2) launch_onepass_cl<2>(swish != 0, ctas, nt, st, c, (int)s, groups, eps, ntiles, x, conv_bias, gamma, beta, y, tile_sums);
      else if (v == 4) launch_onepass_cl<4>(swish != 0, ctas, nt, st, c, (int)s, groups, eps, ntiles, x, conv_bias, gamma, beta, y, tile_sums);
      else launch_onepass_cl<8>(swish != 0, ctas, nt, st, c, (int)s, groups, eps, ntiles, x, conv_bias, gamma, beta, y, tile_sums);
      g_last_launches = 1;
      BDM_RETURN_LAUNCH_STATUS();
    }
  }
  double2 *partials = static_cast<double2 *>(workspace);
  gn_cl_stats_kernel<<<dim3(nchunks, b), kClThreads, 0, st>>>(c, s, nchunks, x, partials);
  if (swish)
    gn_cl_apply_kernel<true, 4><<<dim3(ntiles, b), kClThreads, 0, st>>>(c, s, groups, nchunks, nchunks, ntiles, eps, 0, x, conv_bias,
                                                                   gamma, beta, partials, y, tile_sums);
  else
    gn_cl_apply_kernel<false, 4><<<dim3(ntiles, b), kClThreads, 0, st>>>(c, s, groups, nchunks, nchunks, ntiles, eps, 0, x, conv_bias,
                                                                    gamma, beta, partials, y, tile_sums);
  g_last_launches = 2;
  BDM_RETURN_LAUNCH_STATUS();
}

// The statistics-only half of bdm_groupnorm_act_cl for a consumer that normalises on the fly: x f32[b][s][c] with producer
// statistics partials f64[b][chunks][c][2] -> tile_sums f32[b][tiles][c] (tiles = bdm_groupnorm_cl_sums_tiles: the sums of
// y = act(group_norm(x + conv_bias)) the SE gate needs) and coef f32[b][c][2] = (A, B) with y = act(x * A + B).  y itself is
// never written: the pass reads x once.
extern "C" int bdm_groupnorm_cl_sums_tiles(int b, int c, long long s) { return bdm::gn_cl_tiles(b, s, c); }
extern "C" int bdm_groupnorm_cl_sums(int b, int c, long long s, int groups, float eps, int swish, const float *x,
                                     const float *conv_bias, const float *gamma, const float *beta, const double *partials,
                                     int chunks, float *tile_sums, float *coef, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && s >= 0 && groups >= 1 && gn_cl_supported(c, groups) && b <= 65535 && chunks != 0);
  BDM_CHECK_SIZE(chunks > 0 || (conv_bias == nullptr && groups <= 32));      // chunks < 0: -chunks blocks of group partials
  if (b == 0 || s == 0) return BDM_OK;
  BDM_CHECK_PTR(x); BDM_CHECK_PTR(partials); BDM_CHECK_PTR(tile_sums); BDM_CHECK_PTR(coef);
  if (((reinterpret_cast<uintptr_t>(partials) | reinterpret_cast<uintptr_t>(x)) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(coef) & 7) != 0)
    return BDM_ERR_MISALIGNED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ntiles = gn_cl_tiles(b, s, c);
  const double2 *pp = reinterpret_cast<const double2 *>(partials);
  float2 *cf = reinterpret_cast<float2 *>(coef);
  if (swish)
    gn_cl_apply_kernel<true, 4, false><<<dim3(ntiles, b), kClThreads, 0, st>>>(c, s, groups, chunks, chunks, ntiles, eps, 1, x, conv_bias,
                                                                           gamma, beta, pp, nullptr, tile_sums, cf);
  else
    gn_cl_apply_kernel<false, 4, false><<<dim3(ntiles, b), kClThreads, 0, st>>>(c, s, groups, chunks, chunks, ntiles, eps, 1, x, conv_bias,
                                                                            gamma, beta, pp, nullptr, tile_sums, cf);
  BDM_RETURN_LAUNCH_STATUS();
}

// gate f32[b][c] = sigmoid(w2 . act(w1 . (sum_t sums[b,t,c] / count))); sums addressed with element strides
// (stride_b, stride_t, stride_c) so that both the [b*c][tiles] layout of bdm_groupnorm_act and the
// [b][tiles][c] layout of bdm_groupnorm_act_cl can be passed as they are.  w1 f32[hidden][c], w2 f32[c][hidden].
extern "C" int bdm_se_gate(int b, int c, int hidden, int tiles, float count, long long stride_b, long long stride_t,
                           long long stride_c, const float *sums, const float *w1, const float *w2, int use_relu,
                           float *gate, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 1 && hidden >= 1 && tiles >= 1 && count > 0.0f);
  const size_t smem = sizeof(float) * ((size_t)c + hidden + 256);
  BDM_CHECK_SIZE(smem <= 200 * 1024);
  if (b == 0) return BDM_OK;
  BDM_CHECK_PTR(sums); BDM_CHECK_PTR(w1); BDM_CHECK_PTR(w2); BDM_CHECK_PTR(gate);
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(se_gate_kernel), smem);
  if (e != cudaSuccess) return (int)e;
  se_gate_kernel<<<b, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      c, hidden, tiles, count, stride_b, stride_t, stride_c, sums, w1, w2, use_relu, gate);
  BDM_RETURN_LAUNCH_STATUS();
}

extern "C" int bdm_groupnorm_last_launches(void) { return bdm::g_last_launches; }
