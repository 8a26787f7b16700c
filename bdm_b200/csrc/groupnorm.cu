// groupnorm.cu -- fused GroupNorm (+ Swish) for the dense side of the point-voxel blocks, sm_100a.
//
// Every conv of the denoisers is followed by GroupNorm(8) and Swish (x * sigmoid(x)):
// modules/shared_mlp.py:25-31 (1x1 conv -> GroupNorm -> Swish), modules/pvconv.py:75-88 (Conv3d -> GroupNorm
// -> Swish), pvconv.py:59-61 (attention: GroupNorm -> Swish).  Through torch that is four kernels and
// ~5 reads + 3 writes of the tensor (RowwiseMoments with ONE CTA per (sample, group) row of up to 512 KB,
// the normalise pass, sigmoid, mul) -- 7 of the 17 ms of a PC^2 step at B=16.  Both ops are pure HBM
// streaming; fused they are 2 reads + 1 write:
//   gn_stats_kernel   each (sample, group) row -- cg*S contiguous floats -- is split over several CTAs;
//                     per-thread fp32 partial sums over <= 64 elements, then double precision through the
//                     warp / block reduction; one (sum, sumsq) partial per CTA (fixed slots: deterministic)
//   gn_apply_kernel   one CTA per (sample, channel, tile): warp 0 folds the row's partials in a fixed
//                     order into mean / rstd, then y = swish(x * (rstd*gamma) + (beta - mean*rstd*gamma))
//                     with 128-bit loads and streaming stores.
// Tolerance against torch (F.group_norm followed by x*sigmoid(x)): 1e-5 relative to the output's max
// (tests/test_dense_fused_gpu.py); biased variance, eps inside the sqrt, like torch.
#include "common.cuh"

namespace bdm {

constexpr int kGnThreads = 256;
constexpr int kGnMaxChunks = 64;

__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(long long row_len, int nchunks, const float *__restrict__ x, double2 *__restrict__ partials) {
  const long long row = blockIdx.y;
  const int chunk = blockIdx.x;
  const long long per = (row_len + nchunks - 1) / nchunks;
  const long long lo = chunk * per, hi = min(lo + per, row_len);
  const float *p = x + row * row_len;
  float s = 0.0f, q = 0.0f;
  double ds = 0.0, dq = 0.0;
  int since = 0;
  if (((row * row_len + lo) & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const long long n4 = (hi - lo) >> 2;
    const float4 *p4 = reinterpret_cast<const float4 *>(p + lo);
    for (long long i = threadIdx.x; i < n4; i += kGnThreads) {
      const float4 v = ld_stream_f4(reinterpret_cast<const float *>(p4 + i));
      s += (v.x + v.y) + (v.z + v.w);
      q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      if (++since == 16) { ds += s; dq += q; s = q = 0.0f; since = 0; }  // bound the fp32 run length
    }
    for (long long i = lo + (n4 << 2) + threadIdx.x; i < hi; i += kGnThreads) { const float v = p[i]; s += v; q += v * v; }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += kGnThreads) {
      const float v = p[i];
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

template <bool SWISH>
__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(int c, long long s, int groups, int nchunks, float eps, int tile4,
                const float *__restrict__ x, const float *__restrict__ gamma,
                const float *__restrict__ beta, const double2 *__restrict__ partials,
                float *__restrict__ y) {
  const long long bc = blockIdx.x;       // sample * c + channel
  const int ch = (int)(bc % c);
  const long long sample = bc / c;
  const int cg = c / groups;
  const long long row = sample * groups + ch / cg;
  __shared__ float s_ab[2];
  if (threadIdx.x < 32) {
    double a = 0.0, b2 = 0.0;
    for (int k = threadIdx.x; k < nchunks; k += 32) { const double2 v = partials[row * nchunks + k]; a += v.x; b2 += v.y; }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, d);
      b2 += __shfl_xor_sync(0xffffffffu, b2, d);
    }
    if (threadIdx.x == 0) {
      const double n = (double)cg * (double)s;
      const double mean = a / n;
      const double var = fmax(b2 / n - mean * mean, 0.0);
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      const float ga = gamma != nullptr ? gamma[ch] : 1.0f;
      const float be = beta != nullptr ? beta[ch] : 0.0f;
      const float scale = rstd * ga;
      s_ab[0] = scale;
      s_ab[1] = be - (float)mean * scale;
    }
  }
  __syncthreads();
  const float A = s_ab[0], Bc = s_ab[1];
  const float *px = x + bc * s;
  float *py = y + bc * s;
  auto act = [](float v) { return SWISH ? v / (1.0f + expf(-v)) : v; };
  if ((s & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    const long long n4 = s >> 2;
    const long long lo = (long long)blockIdx.y * tile4, hi = min(lo + tile4, n4);
    for (long long i = lo + threadIdx.x; i < hi; i += kGnThreads) {
      float4 v = ld_stream_f4(px + 4 * i);
      v.x = act(fmaf(v.x, A, Bc)); v.y = act(fmaf(v.y, A, Bc)); v.z = act(fmaf(v.z, A, Bc)); v.w = act(fmaf(v.w, A, Bc));
      *reinterpret_cast<float4 *>(py + 4 * i) = v;
    }
  } else {
    const long long lo = (long long)blockIdx.y * tile4 * 4, hi = min(lo + (long long)tile4 * 4, s);
    for (long long i = lo + threadIdx.x; i < hi; i += kGnThreads) py[i] = act(fmaf(px[i], A, Bc));
  }
}

}  // namespace bdm

extern "C" size_t bdm_groupnorm_workspace_bytes(long long rows) {
  return sizeof(double2) * (size_t)(rows > 0 ? rows : 1) * bdm::kGnMaxChunks;
}

extern "C" int bdm_groupnorm_act(int b, int c, long long s, int groups, float eps, int swish, const float *x,
                                 const float *gamma, const float *beta, float *y, void *workspace,
                                 size_t workspace_bytes, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 1 && s >= 0 && groups >= 1 && c % groups == 0);
  if (b == 0 || s == 0) return BDM_OK;
  BDM_CHECK_PTR(x); BDM_CHECK_PTR(y); BDM_CHECK_PTR(workspace);
  const long long rows = (long long)b * groups;
  const long long row_len = (long long)(c / groups) * s;
  if (workspace_bytes < bdm_groupnorm_workspace_bytes(rows)) return BDM_ERR_WORKSPACE_TOO_SMALL;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return BDM_ERR_MISALIGNED;
  BDM_CHECK_SIZE(rows <= 65535);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // enough CTAs to fill the machine ~4x over, at least 8 K elements per CTA
  int nchunks = (int)((4LL * sm_count() + rows - 1) / rows);
  const long long max_by_len = (row_len + 8191) / 8192;
  if (nchunks > max_by_len) nchunks = (int)max_by_len;
  if (nchunks > kGnMaxChunks) nchunks = kGnMaxChunks;
  if (nchunks < 1) nchunks = 1;
  double2 *partials = static_cast<double2 *>(workspace);
  gn_stats_kernel<<<dim3(nchunks, (unsigned)rows), kGnThreads, 0, st>>>(row_len, nchunks, x, partials);
  const long long bc = (long long)b * c;
  const long long n4 = (s + 3) >> 2;
  int tiles = 1;
  while (bc * tiles < 8LL * sm_count() && n4 / (tiles * 2) >= 2 * kGnThreads && tiles < 65535 / 2) tiles *= 2;
  const int tile4 = (int)((n4 + tiles - 1) / tiles);
  BDM_CHECK_SIZE(bc <= 0x7fffffffLL);
  const dim3 grid((unsigned)bc, (unsigned)tiles);
  if (swish)
    gn_apply_kernel<true><<<grid, kGnThreads, 0, st>>>(c, s, groups, nchunks, eps, tile4, x, gamma, beta, partials, y);
  else
    gn_apply_kernel<false><<<grid, kGnThreads, 0, st>>>(c, s, groups, nchunks, eps, tile4, x, gamma, beta, partials, y);
  BDM_RETURN_LAUNCH_STATUS();
}
