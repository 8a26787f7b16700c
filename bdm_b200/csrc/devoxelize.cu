// devoxelize.cu -- trilinear_devoxelize forward/backward for sm_100a.
//
// Replaces trilinear_devoxelize_kernel / _grad_kernel
// (experiments/model/pvcnn/modules/functional/src/interpolate/trilinear_devox.cu:21-105, :119-162;
// launched with one CTA per batch element, each thread looping serially over all C channels with
// 8 uncoalesced 4-byte loads per channel).
//
// Arithmetic contract (bit-exact with the reference's SASS, see DESIGN.md "fp contract"):
//   weights  (xd*yd)*zd left to right (:52-59); corner indices via the `(d>0 ? -1 : 0) & stride`
//   trick (:61-75), so no +1 neighbour is touched when a coordinate is an exact integer;
//   out = w000*f000 + ... + w111*f111 contracted as
//         acc = w001*f001;  acc = fma(w000,f000,acc);  acc = fma(w010,f010,acc); ... ; fma(w111,f111,acc)
//
// Kernels:
//   devox_gather_kernel   generic path: thread per point, channel chunk per blockIdx.y, 8 read-only
//                         gathers per channel, 4 channels in flight.
#include "common.cuh"

namespace bdm {

struct Corner8 {
  int id[8];
  float w[8];
};

// Per-point weights and flat corner indices, exactly as trilinear_devox.cu:37-75.
__device__ __forceinline__ void devox_corners(float x, float y, float z, int r, int r2, Corner8 &k) {
  const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
  const float xd1 = __fsub_rn(x, xl), yd1 = __fsub_rn(y, yl), zd1 = __fsub_rn(z, zl);
  const float xd0 = __fsub_rn(1.0f, xd1), yd0 = __fsub_rn(1.0f, yd1), zd0 = __fsub_rn(1.0f, zd1);
  const float w00 = __fmul_rn(xd0, yd0), w01 = __fmul_rn(xd0, yd1);
  const float w10 = __fmul_rn(xd1, yd0), w11 = __fmul_rn(xd1, yd1);
  k.w[0] = __fmul_rn(w00, zd0); k.w[1] = __fmul_rn(w00, zd1);
  k.w[2] = __fmul_rn(w01, zd0); k.w[3] = __fmul_rn(w01, zd1);
  k.w[4] = __fmul_rn(w10, zd0); k.w[5] = __fmul_rn(w10, zd1);
  k.w[6] = __fmul_rn(w11, zd0); k.w[7] = __fmul_rn(w11, zd1);
  const int xlo = (int)xl, ylo = (int)yl, zlo = (int)zl;
  const int xo = (xd1 > 0.0f) ? r2 : 0, yo = (yd1 > 0.0f) ? r : 0, zo = (zd1 > 0.0f) ? 1 : 0;
  k.id[0] = xlo * r2 + ylo * r + zlo;
  k.id[1] = k.id[0] + zo;
  k.id[2] = k.id[0] + yo;
  k.id[3] = k.id[2] + zo;
  k.id[4] = k.id[0] + xo;
  k.id[5] = k.id[4] + zo;
  k.id[6] = k.id[4] + yo;
  k.id[7] = k.id[6] + zo;
}

__device__ __forceinline__ float devox_blend(const float *__restrict__ f, const Corner8 &k) {
  const float f0 = __ldg(f + k.id[0]), f1 = __ldg(f + k.id[1]), f2 = __ldg(f + k.id[2]),
              f3 = __ldg(f + k.id[3]), f4 = __ldg(f + k.id[4]), f5 = __ldg(f + k.id[5]),
              f6 = __ldg(f + k.id[6]), f7 = __ldg(f + k.id[7]);
  float acc = __fmul_rn(k.w[1], f1);
  acc = __fmaf_rn(k.w[0], f0, acc);
  acc = __fmaf_rn(k.w[2], f2, acc);
  acc = __fmaf_rn(k.w[3], f3, acc);
  acc = __fmaf_rn(k.w[4], f4, acc);
  acc = __fmaf_rn(k.w[5], f5, acc);
  acc = __fmaf_rn(k.w[6], f6, acc);
  acc = __fmaf_rn(k.w[7], f7, acc);
  return acc;
}

constexpr int kDevoxThreads = 128;
constexpr int kDevoxChunk = 8;  // channels per CTA in the generic path

__global__ void __launch_bounds__(kDevoxThreads)
devox_gather_kernel(int c, int n, int r, int is_training, const float *__restrict__ coords,
                    const float *__restrict__ feat, int *__restrict__ inds,
                    float *__restrict__ wgts, float *__restrict__ outs) {
  const int b = blockIdx.z;
  const int i = blockIdx.x * kDevoxThreads + threadIdx.x;
  if (i >= n) return;
  const int r2 = r * r;
  const size_t r3 = (size_t)r2 * r;
  const float *co = coords + (size_t)b * 3 * n;
  Corner8 k;
  devox_corners(co[i], co[i + n], co[i + n + n], r, r2, k);
  if (is_training && blockIdx.y == 0) {
    int *in = inds + (size_t)b * 8 * n;
    float *wg = wgts + (size_t)b * 8 * n;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      wg[i + (size_t)n * q] = k.w[q];
      in[i + (size_t)n * q] = k.id[q];
    }
  }
  const int c0 = blockIdx.y * kDevoxChunk;
  const int c1 = min(c0 + kDevoxChunk, c);
  const float *f = feat + ((size_t)b * c + c0) * r3;
  float *o = outs + ((size_t)b * c + c0) * n + i;
  int cc = c0;
  for (; cc + 4 <= c1; cc += 4) {
    const float v0 = devox_blend(f, k), v1 = devox_blend(f + r3, k), v2 = devox_blend(f + 2 * r3, k),
                v3 = devox_blend(f + 3 * r3, k);
    o[0] = v0; o[n] = v1; o[2 * (size_t)n] = v2; o[3 * (size_t)n] = v3;
    f += 4 * r3;
    o += 4 * (size_t)n;
  }
  for (; cc < c1; ++cc) {
    o[0] = devox_blend(f, k);
    f += r3;
    o += n;
  }
}

// backward (trilinear_devox.cu:119-162): 8 atomic scatter-adds of fl(w*g) per (point, channel).
__global__ void __launch_bounds__(kDevoxThreads)
devox_grad_kernel(int c, int n, int r3, const int *__restrict__ inds, const float *__restrict__ wgts,
                  const float *__restrict__ grad_y, float *__restrict__ grad_x) {
  const int b = blockIdx.z;
  const int i = blockIdx.x * kDevoxThreads + threadIdx.x;
  if (i >= n) return;
  const int *in = inds + (size_t)b * 8 * n;
  const float *wg = wgts + (size_t)b * 8 * n;
  int id[8];
  float w[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    id[q] = in[i + (size_t)n * q];
    w[q] = wg[i + (size_t)n * q];
  }
  const int c0 = blockIdx.y * kDevoxChunk;
  const int c1 = min(c0 + kDevoxChunk, c);
  for (int cc = c0; cc < c1; ++cc) {
    const float g = grad_y[((size_t)b * c + cc) * n + i];
    float *gx = grad_x + ((size_t)b * c + cc) * r3;
#pragma unroll
    for (int q = 0; q < 8; ++q) atomicAdd(gx + id[q], __fmul_rn(w[q], g));
  }
}

}  // namespace bdm

extern "C" int bdm_trilinear_devoxelize(int b, int c, int n, int r, int is_training,
                                        const float *coords, const float *feat, int *inds,
                                        float *wgts, float *outs, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && r >= 1 && (long long)r * r * r <= 0x7fffffffLL);
  if (b == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(coords);
  if (c > 0) { BDM_CHECK_PTR(feat); BDM_CHECK_PTR(outs); }
  if (is_training) { BDM_CHECK_PTR(inds); BDM_CHECK_PTR(wgts); }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int cy = c > 0 ? ceil_div(c, kDevoxChunk) : 1;
  BDM_CHECK_SIZE(cy <= 65535 && b <= 65535);
  devox_gather_kernel<<<dim3(ceil_div(n, kDevoxThreads), cy, b), kDevoxThreads, 0, st>>>(
      c, n, r, is_training, coords, feat, inds, wgts, outs);
  BDM_RETURN_LAUNCH_STATUS();
}

extern "C" int bdm_trilinear_devoxelize_grad(int b, int c, int n, int r3, const int *inds,
                                             const float *wgts, const float *grad_y, float *grad_x,
                                             bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && r3 >= 0);
  if (b == 0 || c == 0) return BDM_OK;
  BDM_CHECK_PTR(grad_x);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)b * c * r3, st);
  if (n > 0) {
    BDM_CHECK_PTR(inds); BDM_CHECK_PTR(wgts); BDM_CHECK_PTR(grad_y);
    const int cy = ceil_div(c, kDevoxChunk);
    BDM_CHECK_SIZE(cy <= 65535 && b <= 65535);
    devox_grad_kernel<<<dim3(ceil_div(n, kDevoxThreads), cy, b), kDevoxThreads, 0, st>>>(c, n, r3, inds, wgts,
                                                                                         grad_y, grad_x);
  }
  BDM_RETURN_LAUNCH_STATUS();
}
