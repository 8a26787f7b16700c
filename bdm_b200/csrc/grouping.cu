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

template <bool VEC4>
__global__ void __launch_bounds__(kGrpThreads)
grouping_kernel(int c, int n, int mu, const float *__restrict__ features,
                const int *__restrict__ indices, float *__restrict__ out) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * kGrpCT;
  const int c1 = min(c0 + kGrpCT, c);
  const float *f = features + ((size_t)b * c + c0) * n;
  float *o = out + ((size_t)b * c + c0) * mu;
  const int *ix = indices + (size_t)b * mu;
  if (VEC4) {
    const int q = (blockIdx.x * kGrpThreads + threadIdx.x) * 4;
    if (q >= mu) return;
    const int4 id = __ldg(reinterpret_cast<const int4 *>(ix + q));
    for (int cc = c0; cc < c1; ++cc) {
      float4 v;
      v.x = __ldg(f + id.x); v.y = __ldg(f + id.y); v.z = __ldg(f + id.z); v.w = __ldg(f + id.w);
      st_stream_f4(o + q, v);
      f += n;
      o += mu;
    }
  } else {
    const int q = blockIdx.x * kGrpThreads + threadIdx.x;
    if (q >= mu) return;
    const int id = __ldg(ix + q);
    for (int cc = c0; cc < c1; ++cc) {
      o[q] = __ldg(f + id);
      f += n;
      o += mu;
    }
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

extern "C" int bdm_grouping(int b, int c, int n, int m, int u, const float *features,
                            const int *indices, float *out, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && m >= 0 && u >= 0 && b <= 65535);
  BDM_CHECK_SIZE((long long)m * u <= 0x7fffffffLL);
  const int mu = m * u;
  if (b == 0 || c == 0 || mu == 0) return BDM_OK;
  BDM_CHECK_PTR(features); BDM_CHECK_PTR(indices); BDM_CHECK_PTR(out);
  BDM_CHECK_SIZE(ceil_div(c, kGrpCT) <= 65535);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec4 = (mu % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(indices) & 15) == 0);
  if (vec4)
    grouping_kernel<true><<<dim3(ceil_div(mu / 4, kGrpThreads), ceil_div(c, kGrpCT), b), kGrpThreads, 0, st>>>(
        c, n, mu, features, indices, out);
  else
    grouping_kernel<false><<<dim3(ceil_div(mu, kGrpThreads), ceil_div(c, kGrpCT), b), kGrpThreads, 0, st>>>(
        c, n, mu, features, indices, out);
  BDM_RETURN_LAUNCH_STATUS();
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
