// projection.cu -- PC^2 projection conditioning for sm_100a.
//
// Replaces, for a whole batch in three launches, what
// PointCloudProjectionModel.surface_projection (experiments/model/projection_model.py:127-157) does
// per sample through pytorch3d's naive point rasteriser (K=1, radius 0.0075 NDC, bin_size=0:
// O(H*W*N) point-pixel tests per sample) plus boolean-mask indexing, inside a Python loop over the
// camera list (:205-212).
//
// Semantics restated from pytorch3d (un-vendored; PARITY UNPINNED, see oracle/bdm_oracle.c):
//   view = X R + T (row vectors); ndc.xy = focal * view.xy / view.z + principal; depth = view.z;
//   pixel (row,col) centre: x = -1 + (2(W-1-col)+1)/W, y = -1 + (2(H-1-row)+1)/H  (+X left, +Y up);
//   a point covers a pixel iff depth >= 0 and dx^2+dy^2 < radius^2 (strict); the smallest depth wins,
//   the earlier point index wins depth ties; a winning point receives its pixel's feature vector
//   (lowest pixel index if it wins several -- the reference's order there is unspecified), every other
//   point receives zeros.
//
// Design: point-parallel splat instead of pixel-parallel search.  Each point tests only the few
// pixels whose centre can be inside its radius and does a 64-bit atomicMin of
// (depth_bits << 32 | point_index) into a [B,H,W] z-buffer (depth >= 0, so the float bit pattern
// orders like an unsigned int and the index breaks ties exactly as required).  A second pass gives
// every winner its lowest pixel; a third gathers the C-vector per point with channel-contiguous
// writes.  With a channel-last (HWC) copy of the step-invariant feature map the gather reads are
// contiguous too (feat_is_hwc = 1).
#include "common.cuh"

namespace bdm {

__global__ void proj_init_kernel(unsigned long long *zbuf, size_t npix, int *pix, size_t npts) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npix) zbuf[i] = 0xffffffffffffffffull;
  if (i < npts) pix[i] = 0x7fffffff;
}

__device__ __forceinline__ float pix_to_ndc(int i, int S) {
  // pytorch3d PixToNonSquareNdc for a square image: -1 + (2*i + 1) / S with i already flipped
  return -1.0f + __fdiv_rn(__fmaf_rn(2.0f, (float)i, 1.0f), (float)S);
}

__global__ void proj_splat_kernel(int n, int H, int W, float radius, const float *__restrict__ points,
                                  const float *__restrict__ R, const float *__restrict__ T,
                                  const float *__restrict__ focal, const float *__restrict__ principal,
                                  unsigned long long *__restrict__ zbuf) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *p = points + ((size_t)b * n + i) * 3;
  const float *r = R + (size_t)b * 9;
  const float *t = T + (size_t)b * 3;
  const float p0 = p[0], p1 = p[1], p2 = p[2];
  float v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float acc = __fmul_rn(p0, __ldg(r + k));
    acc = __fmaf_rn(p1, __ldg(r + 3 + k), acc);
    acc = __fmaf_rn(p2, __ldg(r + 6 + k), acc);
    v[k] = __fadd_rn(acc, __ldg(t + k));
  }
  const float px = __fmaf_rn(__ldg(focal + b * 2), __fdiv_rn(v[0], v[2]), __ldg(principal + b * 2));
  const float py = __fmaf_rn(__ldg(focal + b * 2 + 1), __fdiv_rn(v[1], v[2]), __ldg(principal + b * 2 + 1));
  const float pz = v[2];
  if (!(pz >= 0.0f)) return;
  const float r2 = __fmul_rn(radius, radius);
  // candidate window: generous cull, the exact strict test decides
  const int ky = (int)ceilf(radius * (float)H * 0.5f) + 2, kx = (int)ceilf(radius * (float)W * 0.5f) + 2;
  const float ycf = (float)(H - 1) - ((py + 1.0f) * (float)H - 1.0f) * 0.5f;
  const float xcf = (float)(W - 1) - ((px + 1.0f) * (float)W - 1.0f) * 0.5f;
  if (!(ycf > -1e6f && ycf < 1e6f && xcf > -1e6f && xcf < 1e6f)) return;
  const int y0 = max((int)floorf(ycf) - ky, 0), y1 = min((int)floorf(ycf) + ky + 1, H - 1);
  const int x0 = max((int)floorf(xcf) - kx, 0), x1 = min((int)floorf(xcf) + kx + 1, W - 1);
  const unsigned long long key = ((unsigned long long)__float_as_uint(pz) << 32) | (unsigned)i;
  unsigned long long *zb = zbuf + (size_t)b * H * W;
  for (int yi = y0; yi <= y1; ++yi) {
    const float dy = __fsub_rn(pix_to_ndc(H - 1 - yi, H), py);
    const float dy2 = __fmul_rn(dy, dy);
    for (int xi = x0; xi <= x1; ++xi) {
      const float dx = __fsub_rn(pix_to_ndc(W - 1 - xi, W), px);
      const float d2 = __fmaf_rn(dx, dx, dy2);
      if (d2 < r2) atomicMin(zb + (size_t)yi * W + xi, key);
    }
  }
}

__global__ void proj_resolve_kernel(int n, int HW, const unsigned long long *__restrict__ zbuf,
                                    int *__restrict__ pix) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= HW) return;
  const unsigned long long key = zbuf[(size_t)b * HW + q];
  if (key == 0xffffffffffffffffull) return;
  atomicMin(pix + (size_t)b * n + (unsigned)(key & 0xffffffffull), q);
}

// one warp per point: lanes stride over channels -> out[b,i,:] written contiguously
__global__ void proj_gather_kernel(int n, int C, int HW, int feat_is_hwc, const float *__restrict__ feat,
                                   int *__restrict__ pix, float *__restrict__ out) {
  const int b = blockIdx.y;
  const int warps_per_block = blockDim.x >> 5;
  const int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  int q = pix[(size_t)b * n + i];
  const bool won = q != 0x7fffffff;
  float *o = out + ((size_t)b * n + i) * C;
  if (won) {
    if (feat_is_hwc) {
      const float *f = feat + ((size_t)b * HW + q) * C;
      for (int ch = lane; ch < C; ch += 32) o[ch] = __ldg(f + ch);
    } else {
      const float *f = feat + (size_t)b * C * HW + q;
      for (int ch = lane; ch < C; ch += 32) o[ch] = __ldg(f + (size_t)ch * HW);
    }
  } else {
    for (int ch = lane; ch < C; ch += 32) o[ch] = 0.0f;
  }
  __syncwarp();
  if (lane == 0 && !won) pix[(size_t)b * n + i] = -1;
}

// Channel-first variant for the fused conditioning input: out_cf[b, ch_off + ch, i].  A CTA transposes a
// 32-point x 32-channel tile through shared memory: reads are channel-contiguous (HWC map), writes are
// point-contiguous.
__global__ void __launch_bounds__(256)
proj_gather_cf_kernel(int n, int C, int HW, int c_total, int ch_off, const float *__restrict__ feat_hwc,
                      int *__restrict__ pix, float *__restrict__ out_cf) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int pl = warp * 4 + r, i = p0 + pl, ch = c0 + lane;
    float v = 0.0f;
    if (i < n && ch < C) {
      const int q = pix[(size_t)b * n + i];
      if (q != 0x7fffffff && q >= 0) v = __ldg(feat_hwc + ((size_t)b * HW + q) * C + ch);
    }
    tile[pl][lane] = v;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int cl = warp * 4 + r, ch = c0 + cl, i = p0 + lane;
    if (i < n && ch < C) out_cf[((size_t)b * c_total + ch_off + ch) * n + i] = tile[lane][cl];
  }
}

__global__ void proj_finalize_pix_kernel(size_t npts, int *__restrict__ pix) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npts && pix[i] == 0x7fffffff) pix[i] = -1;
}

}  // namespace bdm

static int bdm_surface_projection_impl(int b, int n, int C, int H, int W, float radius,
                                       const float *points, const float *R, const float *T,
                                       const float *focal, const float *principal, const float *feat,
                                       int feat_is_hwc, unsigned long long *zbuf, int *pix, float *out,
                                       bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && n >= 0 && C >= 0 && H >= 1 && W >= 1 && b <= 65535);
  BDM_CHECK_SIZE((long long)H * W <= 0x7fffffffLL);
  if (b == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(points); BDM_CHECK_PTR(R); BDM_CHECK_PTR(T); BDM_CHECK_PTR(focal); BDM_CHECK_PTR(principal);
  BDM_CHECK_PTR(zbuf); BDM_CHECK_PTR(pix);
  if (C > 0) { BDM_CHECK_PTR(feat); BDM_CHECK_PTR(out); }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int HW = H * W;
  const size_t npix = (size_t)b * HW, npts = (size_t)b * n;
  const size_t ninit = npix > npts ? npix : npts;
  proj_init_kernel<<<(unsigned)((ninit + 255) / 256), 256, 0, st>>>(zbuf, npix, pix, npts);
  proj_splat_kernel<<<dim3(ceil_div(n, 128), b), 128, 0, st>>>(n, H, W, radius, points, R, T, focal, principal,
                                                              zbuf);
  proj_resolve_kernel<<<dim3(ceil_div(HW, 256), b), 256, 0, st>>>(n, HW, zbuf, pix);
  proj_gather_kernel<<<dim3(ceil_div(n, 8), b), 256, 0, st>>>(n, C, HW, feat_is_hwc, feat, pix, out);
  BDM_RETURN_LAUNCH_STATUS();
}

// Fused conditioning input: writes the projected features straight into channels [ch_off, ch_off+C) of a
// channel-first tensor out_cf f32[b, c_total, n] (what the denoiser consumes), skipping the
// [b,n,C] intermediate, the concat and the transpose of get_input_with_conditioning
// (projection_model.py:179-231 + point_cloud_model.py:65).  feat_hwc f32[b,H,W,C].
extern "C" int bdm_surface_projection_cf(int b, int n, int C, int H, int W, float radius, const float *points,
                                         const float *R, const float *T, const float *focal,
                                         const float *principal, const float *feat_hwc,
                                         unsigned long long *zbuf, int *pix, float *out_cf, int c_total,
                                         int ch_off, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && n >= 0 && C >= 0 && H >= 1 && W >= 1 && b <= 65535 && ch_off >= 0 && ch_off + C <= c_total);
  BDM_CHECK_SIZE((long long)H * W <= 0x7fffffffLL);
  if (b == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(points); BDM_CHECK_PTR(R); BDM_CHECK_PTR(T); BDM_CHECK_PTR(focal); BDM_CHECK_PTR(principal);
  BDM_CHECK_PTR(zbuf); BDM_CHECK_PTR(pix);
  if (C > 0) { BDM_CHECK_PTR(feat_hwc); BDM_CHECK_PTR(out_cf); }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int HW = H * W;
  const size_t npix = (size_t)b * HW, npts = (size_t)b * n;
  const size_t ninit = npix > npts ? npix : npts;
  proj_init_kernel<<<(unsigned)((ninit + 255) / 256), 256, 0, st>>>(zbuf, npix, pix, npts);
  proj_splat_kernel<<<dim3(ceil_div(n, 128), b), 128, 0, st>>>(n, H, W, radius, points, R, T, focal, principal, zbuf);
  proj_resolve_kernel<<<dim3(ceil_div(HW, 256), b), 256, 0, st>>>(n, HW, zbuf, pix);
  if (C > 0)
    proj_gather_cf_kernel<<<dim3(ceil_div(n, 32), ceil_div(C, 32), b), 256, 0, st>>>(n, C, HW, c_total, ch_off, feat_hwc, pix, out_cf);
  proj_finalize_pix_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, st>>>(npts, pix);
  BDM_RETURN_LAUNCH_STATUS();
}

extern "C" int bdm_surface_projection(int b, int n, int C, int H, int W, float radius,
                                      const float *points, const float *R, const float *T,
                                      const float *focal, const float *principal, const float *feat,
                                      unsigned long long *zbuf, int *pix, float *out,
                                      bdm_stream_t stream) {
  return bdm_surface_projection_impl(b, n, C, H, W, radius, points, R, T, focal, principal, feat, 0, zbuf, pix,
                                     out, stream);
}

extern "C" int bdm_surface_projection_hwc(int b, int n, int C, int H, int W, float radius,
                                          const float *points, const float *R, const float *T,
                                          const float *focal, const float *principal,
                                          const float *feat_hwc, unsigned long long *zbuf, int *pix,
                                          float *out, bdm_stream_t stream) {
  return bdm_surface_projection_impl(b, n, C, H, W, radius, points, R, T, focal, principal, feat_hwc, 1, zbuf,
                                     pix, out, stream);
}
