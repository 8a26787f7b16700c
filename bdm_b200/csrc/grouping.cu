// grouping.cu -- neighbour grouping (forward/backward) for sm_100a.
//
// Replaces grouping_kernel / grouping_grad_kernel
// (experiments/model/pvcnn/modules/functional/src/grouping/grouping.cu:18-36, :58-77; one CTA per
// batch element, scalar stores, output pre-zeroed by the wrapper).
//
//   out[b,c,m,u] = features[b,c,indices[b,m,u]]
//
// The op is bound by the [B,C,M,U] output write (134 MB for C=64 at the first SA stage, B=16).  Each
// thread owns 4 consecutive (m,u) outputs: one 128-bit index load reused for CT channels, 4 read-only
// gathers per channel from a feature row that is L1/L2 resident (N*4 bytes), one 128-bit streaming
// store per channel.
#include "common.cuh"

namespace bdm {

constexpr int kGrpThreads = 256;
constexpr int kGrpCT = 8;

// Where the output goes and what is subtracted on the way: channel `coff + c` of a tensor with `oc`
// channels (so that several groupings fill one concatenated tensor, ball_query.py:26-33), minus
// centers[b,c,m] when given (the neighbour coordinates are made relative to their centre, ball_query.py:25;
// one fp32 subtraction per value, exactly the reference's `grouping(...) - centers.unsqueeze(-1)`).
struct GroupDst {
  float *out;
  const float *centers;   // f32[b][c][m] or nullptr
  int oc, coff, u;
};

template <bool VEC4>
__global__ void __launch_bounds__(kGrpThreads)
grouping_kernel(int c, int n, int mu, const float *__restrict__ features,
                const int *__restrict__ indices, GroupDst dst) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * kGrpCT;
  const int c1 = min(c0 + kGrpCT, c);
  const float *f = features + ((size_t)b * c + c0) * n;
  float *o = dst.out + ((size_t)b * dst.oc + dst.coff + c0) * mu;
  const int *ix = indices + (size_t)b * mu;
  const int m = mu / dst.u;
  const float *cen = dst.centers ? dst.centers + ((size_t)b * c + c0) * m : nullptr;
  if (VEC4) {
    const int q = (blockIdx.x * kGrpThreads + threadIdx.x) * 4;
    if (q >= mu) return;
    const int4 id = __ldg(reinterpret_cast<const int4 *>(ix + q));
    for (int cc = c0; cc < c1; ++cc) {
      float4 v;
      v.x = __ldg(f + id.x); v.y = __ldg(f + id.y); v.z = __ldg(f + id.z); v.w = __ldg(f + id.w);
      if (cen) {
        v.x = __fsub_rn(v.x, __ldg(cen + q / dst.u)); v.y = __fsub_rn(v.y, __ldg(cen + (q + 1) / dst.u));
        v.z = __fsub_rn(v.z, __ldg(cen + (q + 2) / dst.u)); v.w = __fsub_rn(v.w, __ldg(cen + (q + 3) / dst.u));
        cen += m;
      }
      st_stream_f4(o + q, v);
      f += n;
      o += mu;
    }
  } else {
    const int q = blockIdx.x * kGrpThreads + threadIdx.x;
    if (q >= mu) return;
    const int id = __ldg(ix + q);
    for (int cc = c0; cc < c1; ++cc) {
      float v = __ldg(f + id);
      if (cen) { v = __fsub_rn(v, __ldg(cen + q / dst.u)); cen += m; }
      o[q] = v;
      f += n;
      o += mu;
    }
  }
}

// Shared-memory variant: the gathers of one output row hit random addresses of a feature row of n
// floats.  From global memory each warp-level gather costs up to 32 L1 wavefronts; from shared memory
// about 3.  A CTA stages CT rows [CT][n] and serves a chunk of the (m,u) index space from them:
// one 128-bit index load, CT x 4 shared-memory reads, CT 128-bit streaming stores per thread step.
template <int CT>
__global__ void __launch_bounds__(kGrpThreads)
grouping_rows_kernel(int c, int n, int mu, int chunk4, const float *__restrict__ features,
                     const int *__restrict__ indices, GroupDst dst) {
  extern __shared__ __align__(128) float rows[];  // [CT][n]
  __shared__ __align__(8) uint64_t mbar;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * CT;
  const int nrows = min(CT, c - c0);
  const float *f = features + ((size_t)b * c + c0) * n;
  // the CT rows are one contiguous block of nrows*n floats: a single TMA bulk copy when 16-byte aligned
  const bool bulk = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(f) & 15) == 0);
  if (bulk) {
    if (threadIdx.x == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned bytes = (unsigned)(sizeof(float) * (size_t)nrows * n);
      mbar_expect_tx(&mbar, bytes);
      bulk_g2s(rows, f, bytes, &mbar);
    }
  } else {
    for (int q = threadIdx.x; q < nrows * n; q += kGrpThreads) rows[q] = ld_stream_f1(f + q);
  }
  const int4 *ix = reinterpret_cast<const int4 *>(indices + (size_t)b * mu);
  float *o = dst.out + ((size_t)b * dst.oc + dst.coff + c0) * mu;
  const int m = mu / dst.u;
  const float *cen = dst.centers ? dst.centers + ((size_t)b * c + c0) * m : nullptr;   // only with u % 4 == 0
  const int g_end = min((blockIdx.x + 1) * chunk4, mu / 4);
  int g = blockIdx.x * chunk4 + threadIdx.x;
  int4 id = g < g_end ? __ldg(ix + g) : make_int4(0, 0, 0, 0);   // first indices arrive while the rows do
  if (bulk) mbar_wait(&mbar, 0); else __syncthreads();
  for (; g < g_end; g += kGrpThreads) {
    const int gn = g + kGrpThreads;
    const int4 idn = gn < g_end ? __ldg(ix + gn) : make_int4(0, 0, 0, 0);
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) {
      if (cc < nrows) {
        const float *r = rows + cc * n;
        float4 v = make_float4(r[id.x], r[id.y], r[id.z], r[id.w]);
        if (cen) {   // the four values share one centre
          const float ce = __ldg(cen + (size_t)cc * m + (4 * g) / dst.u);
          v.x = __fsub_rn(v.x, ce); v.y = __fsub_rn(v.y, ce); v.z = __fsub_rn(v.z, ce); v.w = __fsub_rn(v.w, ce);
        }
        st_stream_f4(o + (size_t)cc * mu + 4 * (size_t)g, v);
      }
    }
    id = idn;
  }
}

// backward: grad_x[b,c,indices[b,m,u]] += grad_y[b,c,m,u]   (grouping.cu:71-76)
__global__ void __launch_bounds__(kGrpThreads)
grouping_grad_kernel(int c, int n, int mu, const float *__restrict__ grad_y,
                     const int *__restrict__ indices, float *__restrict__ grad_x) {
  const int b = blockIdx.z;
  const int q = blockIdx.x * kGrpThreads + threadIdx.x;
  if (q >= mu) return;
  const int id = __ldg(indices + (size_t)b * mu + q);
  const int c0 = blockIdx.y * kGrpCT;
  const int c1 = min(c0 + kGrpCT, c);
  for (int cc = c0; cc < c1; ++cc)
    atomicAdd(grad_x + ((size_t)b * c + cc) * n + id, grad_y[((size_t)b * c + cc) * mu + q]);
}

}  // namespace bdm

// out[b, coff + cc, m, u] = features[b, cc, indices[b,m,u]] - (centers ? centers[b, cc, m] : 0) for a tensor
// `out` of oc channels: bdm_grouping with a destination slice and the centre subtraction folded in.
extern "C" int bdm_grouping_into(int b, int c, int n, int m, int u, const float *features, const int *indices,
                                 const float *centers, float *out, int out_channels, int channel_offset,
                                 bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && m >= 0 && u >= 0 && b <= 65535);
  BDM_CHECK_SIZE(channel_offset >= 0 && channel_offset + c <= out_channels);
  BDM_CHECK_SIZE((long long)m * u <= 0x7fffffffLL);
  const int mu = m * u;
  if (b == 0 || c == 0 || mu == 0) return BDM_OK;
  BDM_CHECK_PTR(features); BDM_CHECK_PTR(indices); BDM_CHECK_PTR(out);
  BDM_CHECK_SIZE(ceil_div(c, kGrpCT) <= 65535);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const GroupDst dst{out, centers, out_channels, channel_offset, u};
  const bool vec4 = (mu % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(indices) & 15) == 0);
  if (vec4 && n >= 1 && sizeof(float) * (size_t)n <= 64 * 1024 && mu / 4 >= kGrpThreads &&
      (centers == nullptr || u % 4 == 0)) {
    int ct = 4;
    while (ct > 1 && (sizeof(float) * (size_t)ct * n > 64 * 1024 || b * ceil_div(c, ct) < sm_count())) ct >>= 1;
    const int tiles = ceil_div(c, ct) * b;
    int chunks = 1;
    while (tiles * chunks < 3 * sm_count() && ceil_div(mu / 4, chunks * 2) >= 4 * kGrpThreads) chunks *= 2;
    const int chunk4 = ceil_div(mu / 4, chunks);
    const size_t smem = sizeof(float) * (size_t)ct * n;
    const dim3 grid(chunks, ceil_div(c, ct), b);
    cudaError_t e;
    if (ct == 4) {
      e = ensure_dynamic_smem(reinterpret_cast<const void *>(grouping_rows_kernel<4>), smem);
      if (e == cudaSuccess) grouping_rows_kernel<4><<<grid, kGrpThreads, smem, st>>>(c, n, mu, chunk4, features, indices, dst);
    } else if (ct == 2) {
      e = ensure_dynamic_smem(reinterpret_cast<const void *>(grouping_rows_kernel<2>), smem);
      if (e == cudaSuccess) grouping_rows_kernel<2><<<grid, kGrpThreads, smem, st>>>(c, n, mu, chunk4, features, indices, dst);
    } else {
      e = ensure_dynamic_smem(reinterpret_cast<const void *>(grouping_rows_kernel<1>), smem);
      if (e == cudaSuccess) grouping_rows_kernel<1><<<grid, kGrpThreads, smem, st>>>(c, n, mu, chunk4, features, indices, dst);
    }
    if (e != cudaSuccess) return (int)e;
  } else if (vec4)
    grouping_kernel<true><<<dim3(ceil_div(mu / 4, kGrpThreads), ceil_div(c, kGrpCT), b), kGrpThreads, 0, st>>>(
        c, n, mu, features, indices, dst);
  else
    grouping_kernel<false><<<dim3(ceil_div(mu, kGrpThreads), ceil_div(c, kGrpCT), b), kGrpThreads, 0, st>>>(
        c, n, mu, features, indices, dst);
  BDM_RETURN_LAUNCH_STATUS();
}

extern "C" int bdm_grouping(int b, int c, int n, int m, int u, const float *features,
                            const int *indices, float *out, bdm_stream_t stream) {
  return bdm_grouping_into(b, c, n, m, u, features, indices, nullptr, out, c, 0, stream);
}

extern "C" int bdm_grouping_grad(int b, int c, int n, int m, int u, const float *grad_y,
                                 const int *indices, float *grad_x, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && m >= 0 && u >= 0 && b <= 65535);
  BDM_CHECK_SIZE((long long)m * u <= 0x7fffffffLL);
  const int mu = m * u;
  if (b == 0 || c == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(grad_x);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)b * c * n, st);
  if (mu > 0) {
    BDM_CHECK_PTR(grad_y); BDM_CHECK_PTR(indices);
    grouping_grad_kernel<<<dim3(ceil_div(mu, kGrpThreads), ceil_div(c, kGrpCT), b), kGrpThreads, 0, st>>>(
        c, n, mu, grad_y, indices, grad_x);
  }
  BDM_RETURN_LAUNCH_STATUS();
}
