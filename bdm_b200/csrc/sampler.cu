// sampler.cu -- the reverse-diffusion update of one sampling step as ONE elementwise kernel with a
// device-side timestep, so that a whole denoising step (conditioning -> denoiser -> update) is a single
// CUDA-graph replay with no host-side coefficient math and no host->device traffic.
//
// Replaces, per step, the ~8 eager torch launches of
//   * the PC^2 side: diffusers DDPMScheduler.step(...).prev_sample as driven by
//     experiments/model/model.py:182-194 / :273-289 (epsilon prediction, `fixed_small` variance, no clipping;
//     diffusers is not vendored: published algorithm, PARITY UNPINNED), and
//   * the PVD side: GaussianDiffusion.p_sample, experiments/pvd/__init__.py:136-224
//     (_predict_xstart_from_eps :184-193, q_posterior_mean_variance :112-134, noise masked at t == 0 :212-219).
// Every product and sum below is rounded separately (__fmul_rn / __fadd_rn / __fsub_rn, never contracted):
// the torch sequence it replaces is one kernel per arithmetic op, so the results are bit-identical to it.
#include "common.cuh"

namespace bdm {

constexpr int kCoefStride = 8;   // floats per timestep row of the coefficient table

template <int MODE>
__device__ __forceinline__ float update_one(float x, float e, float z, float c0, float c1, float c2, float c3,
                                            float c4, bool add_noise) {
  float x0;
  if (MODE == 0) {
    // x0 = (x - sqrt(1-abar) * eps) / sqrt(abar): torch divides by a host scalar as a multiplication with
    // its reciprocal (computed in double, rounded to float) -- c1 is that reciprocal.
    x0 = __fmul_rn(__fsub_rn(x, __fmul_rn(c0, e)), c1);
  } else {
    // x0 = sqrt(1/abar) * x - sqrt(1/abar - 1) * eps        (pvd/__init__.py:184-193)
    x0 = __fsub_rn(__fmul_rn(c0, x), __fmul_rn(c1, e));
  }
  float prev = __fadd_rn(__fmul_rn(c2, x0), __fmul_rn(c3, x));       // posterior mean
  if (add_noise) prev = __fadd_rn(prev, __fmul_rn(c4, z));            // + sigma_t * z, t > 0 only
  return prev;
}

template <int MODE>
__global__ void __launch_bounds__(256)
sampler_update_kernel(long long n4, long long n, const float *__restrict__ x, const float *__restrict__ eps,
                      const float *__restrict__ noise, const float *__restrict__ table, int rows,
                      const int *__restrict__ t_dev, float *__restrict__ out) {
  int t = __ldg(t_dev);
  t = min(max(t, 0), rows - 1);
  const float *c = table + (size_t)t * kCoefStride;
  const float c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3), c4 = __ldg(c + 4);
  const bool add_noise = t > 0;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    const float4 xv = *reinterpret_cast<const float4 *>(x + 4 * i);
    const float4 ev = *reinterpret_cast<const float4 *>(eps + 4 * i);
    const float4 zv = *reinterpret_cast<const float4 *>(noise + 4 * i);
    float4 o;
    o.x = update_one<MODE>(xv.x, ev.x, zv.x, c0, c1, c2, c3, c4, add_noise);
    o.y = update_one<MODE>(xv.y, ev.y, zv.y, c0, c1, c2, c3, c4, add_noise);
    o.z = update_one<MODE>(xv.z, ev.z, zv.z, c0, c1, c2, c3, c4, add_noise);
    o.w = update_one<MODE>(xv.w, ev.w, zv.w, c0, c1, c2, c3, c4, add_noise);
    *reinterpret_cast<float4 *>(out + 4 * i) = o;
  } else if (i == n4) {   // the <= 3 trailing elements
    for (long long j = 4 * n4; j < n; ++j)
      out[j] = update_one<MODE>(x[j], eps[j], noise[j], c0, c1, c2, c3, c4, add_noise);
  }
}

}  // namespace bdm

// out[i] = update(x[i], eps[i], noise[i]; table[*t_dev]) for i < n.  `out` may alias `x` (each element is read
// and written by the same thread).  table f32[rows][8] = per-timestep coefficients (c0..c4, 3 pad); mode 0 =
// DDPM / diffusers form, mode 1 = PVD form (see the top of sampler.cu).  *t_dev is clamped to [0, rows-1].
extern "C" int bdm_sampler_update(long long n, int mode, const float *x, const float *eps, const float *noise,
                                  const float *table, int rows, const int *t_dev, float *out,
                                  bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(n >= 0 && rows >= 1 && (mode == 0 || mode == 1));
  if (n == 0) return BDM_OK;
  BDM_CHECK_PTR(x); BDM_CHECK_PTR(eps); BDM_CHECK_PTR(noise); BDM_CHECK_PTR(table); BDM_CHECK_PTR(t_dev); BDM_CHECK_PTR(out);
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(eps) | reinterpret_cast<uintptr_t>(noise) |
        reinterpret_cast<uintptr_t>(out)) & 15) != 0)
    return BDM_ERR_MISALIGNED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n4 = n / 4;
  const long long threads = n4 + 1;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  if (mode == 0)
    sampler_update_kernel<0><<<blocks, 256, 0, st>>>(n4, n, x, eps, noise, table, rows, t_dev, out);
  else
    sampler_update_kernel<1><<<blocks, 256, 0, st>>>(n4, n, x, eps, noise, table, rows, t_dev, out);
  BDM_RETURN_LAUNCH_STATUS();
}
