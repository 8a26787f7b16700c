// voxel_coords.cu -- the coordinate half of Voxelization.forward as ONE kernel.
//
// Reference: experiments/model/pvcnn/modules/voxelization.py:16-25 -- eight torch launches per call, fourteen
// calls per denoiser forward:
//     norm = coords - coords.mean(2, keepdim=True)
//     norm = norm / (norm.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + eps) + 0.5     (normalize)
//     norm = (norm + 1) / 2.0                                                                              (otherwise)
//     norm = clamp(norm * r, 0, r - 1);   vox = round(norm).int32          (round half to even)
// One CTA per shape: three passes over the shape's 3*N floats (mean, largest norm, write), the first from global
// memory, the others from L1/L2.  Every elementwise step is the torch op's own fp32 arithmetic, spelled with
// round-to-nearest intrinsics so that nothing is contracted: centred = x - mean, |c|^2 = (cx*cx + cy*cy) + cz*cz
// (the three squares rounded one by one, summed left to right, as torch's reduction over a 3-long dimension
// does), sqrt, * 2, + eps, IEEE division, + 0.5, * r, clamp, rint.  The one deliberate difference: the mean is
// accumulated in double and rounded once (torch sums in fp32 in a build-specific tree order), i.e. it is the
// correctly rounded mean, at most one ulp from any fp32 summation order.
#include "common.cuh"

namespace bdm {

constexpr int kCoordThreads = 1024;

__global__ void __launch_bounds__(kCoordThreads)
voxel_coords_kernel(int n, int r, int normalize, float eps, const float *__restrict__ coords,
                    float *__restrict__ norm_coords, int *__restrict__ vox_coords) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *px = coords + (size_t)b * 3 * n, *py = px + n, *pz = py + n;
  __shared__ double s_sum[3][kCoordThreads / 32];
  __shared__ float s_max[kCoordThreads / 32];
  __shared__ float s_mean[3];
  __shared__ float s_extent;

  double sx = 0.0, sy = 0.0, sz = 0.0;
  for (int i = tid; i < n; i += kCoordThreads) { sx += (double)__ldg(px + i); sy += (double)__ldg(py + i); sz += (double)__ldg(pz + i); }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, d);
    sy += __shfl_xor_sync(0xffffffffu, sy, d);
    sz += __shfl_xor_sync(0xffffffffu, sz, d);
  }
  if (lane == 0) { s_sum[0][warp] = sx; s_sum[1][warp] = sy; s_sum[2][warp] = sz; }
  __syncthreads();
  if (tid < 3) {
    double a = 0.0;
    for (int w = 0; w < kCoordThreads / 32; ++w) a += s_sum[tid][w];
    s_mean[tid] = (float)(a / (double)n);
  }
  __syncthreads();
  const float mx = s_mean[0], my = s_mean[1], mz = s_mean[2];

  if (normalize) {
    float best = 0.0f;
    for (int i = tid; i < n; i += kCoordThreads) {
      const float cx = __fsub_rn(__ldg(px + i), mx), cy = __fsub_rn(__ldg(py + i), my), cz = __fsub_rn(__ldg(pz + i), mz);
      const float sq = __fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz));
      best = fmaxf(best, __fsqrt_rn(sq));
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, d));
    if (lane == 0) s_max[warp] = best;
    __syncthreads();
    if (tid == 0) {
      float m = 0.0f;
      for (int w = 0; w < kCoordThreads / 32; ++w) m = fmaxf(m, s_max[w]);
      s_extent = __fadd_rn(__fmul_rn(m, 2.0f), eps);
    }
    __syncthreads();
  }
  const float extent = normalize ? s_extent : 1.0f;
  const float fr = (float)r, hi = (float)(r - 1);
  float *ox = norm_coords + (size_t)b * 3 * n;
  int *vx = vox_coords + (size_t)b * 3 * n;
  for (int i = tid; i < n; i += kCoordThreads) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float c = __fsub_rn(__ldg(px + (size_t)a * n + i), a == 0 ? mx : (a == 1 ? my : mz));
      float u;
      if (normalize) u = __fadd_rn(__fdiv_rn(c, extent), 0.5f);
      else u = __fdiv_rn(__fadd_rn(c, 1.0f), 2.0f);
      float v = __fmul_rn(u, fr);
      v = fminf(fmaxf(v, 0.0f), hi);        // torch.clamp(x, 0, r - 1)
      ox[(size_t)a * n + i] = v;
      vx[(size_t)a * n + i] = __float2int_rn(v);   // round half to even, like torch.round
    }
  }
}

}  // namespace bdm

// coords f32[b][3][n] -> norm_coords f32[b][3][n] (float voxel coordinates in [0, r-1]) and vox_coords i32[b][3][n]
// (their round-half-even), as Voxelization.forward computes them (modules/voxelization.py:17-24).
extern "C" int bdm_voxelize_coords(int b, int n, int r, int normalize, float eps, const float *coords,
                                   float *norm_coords, int *vox_coords, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && n >= 0 && r >= 1);
  if (b == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(coords); BDM_CHECK_PTR(norm_coords); BDM_CHECK_PTR(vox_coords);
  voxel_coords_kernel<<<b, kCoordThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(n, r, normalize ? 1 : 0, eps, coords,
                                                                                      norm_coords, vox_coords);
  BDM_RETURN_LAUNCH_STATUS();
}
