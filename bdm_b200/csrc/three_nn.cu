// three_nn.cu -- three nearest neighbours + inverse-distance interpolation (fwd/bwd) for sm_100a.
//
// Replaces three_nearest_neighbors_kernel, three_nearest_neighbors_interpolate_kernel and
// three_nearest_neighbors_interpolate_grad_kernel
// (experiments/model/pvcnn/modules/functional/src/interpolate/neighbor_interpolate.cu:20-75, :90-116,
// :145-170; one CTA per batch element, serial scan of the centres from global memory).
//
// Bit-exact contract for the search:
//   d = fma(dz,dz, fma(dx,dx, dy*dy)) with d = point - centre (:43 as contracted by nvcc);
//   the reference compares the float d against double bests initialised to 1e40 with strict '<'
//   (:44-57).  float -> double widening is exact, and 1e40 only ever compares against finite floats
//   or +inf/NaN, for which `d < 1e40` and `d < +inf` agree -- so float bests initialised to +inf give
//   the same insertion decisions.  The clamp and the three pair products are evaluated in double and
//   rounded to float (:61-66), the rest in fp32 (:67-72) -- reproduced literally.
// Interpolation: out = f[i1]*w1 + f[i2]*w2 + f[i3]*w3, contracted as fma(f3,w3, fma(f1,w1, f2*w2)).
//
// Search: one lane per point, uniform 128-bit loads of 4 consecutive centre x/y/z values.
// Interpolation: one thread per point, (idx,w) in registers, CT channels per CTA, coalesced stores.
#include "common.cuh"

namespace bdm {

constexpr int kNnThreads = 128;
constexpr int kNnCT = 8;

__device__ __forceinline__ void nn3_insert(float d, int k, float &b0, float &b1, float &b2, int &i0,
                                           int &i1, int &i2) {
  if (d < b2) {
    b2 = d; i2 = k;
    if (d < b1) {
      b2 = b1; i2 = i1; b1 = d; i1 = k;
      if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = k; }
    }
  }
}

// One warp = 32 points (one per lane) x one contiguous SPLIT of the centre range; the SPLITS warps of
// a CTA work on the same 32 points and merge their sorted top-3 lists through shared memory in split
// order, which preserves the reference's "earlier index first on equal distance" rule because split
// s only holds indices smaller than split s+1.
constexpr int kNnMaxSplits = 8;
constexpr int kNnTile = 256;  // centres staged per warp at a time (3 KB)

template <bool VEC4>
__global__ void __launch_bounds__(32 * kNnMaxSplits)
three_nn_kernel(int n, int m, int splits, const float *__restrict__ points,
                const float *__restrict__ centers, float *__restrict__ weights,
                int *__restrict__ indices) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, split = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  points += (size_t)b * 3 * n;
  centers += (size_t)b * 3 * m;
  weights += (size_t)b * 3 * n;
  indices += (size_t)b * 3 * n;
  const bool valid = j < n;
  const float ux = valid ? points[j] : 0.0f;
  const float uy = valid ? points[j + n] : 0.0f;
  const float uz = valid ? points[j + n + n] : 0.0f;

  int len = ceil_div(m, splits);
  len = (len + 3) & ~3;
  const int k0 = min(split * len, m), k1 = min(k0 + len, m);

  const float inf = __int_as_float(0x7f800000);
  float b0 = inf, b1 = inf, b2 = inf;
  int i0 = 0, i1 = 0, i2 = 0;
  // centres of this split are staged through a warp-private shared-memory tile: the loads of a tile are
  // all in flight together, the scan then reads broadcast 128-bit values at shared-memory latency
  __shared__ __align__(16) float s_tile[kNnMaxSplits][3 * kNnTile];
  float *tile = s_tile[split];
  for (int t0 = k0; t0 < k1; t0 += kNnTile) {
    const int tn = min(kNnTile, k1 - t0);
    __syncwarp();
    warp_stage_xyz<kNnTile>(centers + t0, (size_t)m, tn, tile, lane, VEC4 && (tn & 3) == 0);
    int q = 0;
    for (; q + 4 <= tn; q += 4) {
      const float4 X = *reinterpret_cast<const float4 *>(tile + q);
      const float4 Y = *reinterpret_cast<const float4 *>(tile + kNnTile + q);
      const float4 Z = *reinterpret_cast<const float4 *>(tile + 2 * kNnTile + q);
      const float d0 = sqdist_ref(__fsub_rn(ux, X.x), __fsub_rn(uy, Y.x), __fsub_rn(uz, Z.x));
      const float d1 = sqdist_ref(__fsub_rn(ux, X.y), __fsub_rn(uy, Y.y), __fsub_rn(uz, Z.y));
      const float d2 = sqdist_ref(__fsub_rn(ux, X.z), __fsub_rn(uy, Y.z), __fsub_rn(uz, Z.z));
      const float d3 = sqdist_ref(__fsub_rn(ux, X.w), __fsub_rn(uy, Y.w), __fsub_rn(uz, Z.w));
      // none of the four can enter the top-3 unless the smallest beats the current third best
      if (fminf(fminf(d0, d1), fminf(d2, d3)) < b2) {
        const int k = t0 + q;
        nn3_insert(d0, k, b0, b1, b2, i0, i1, i2);
        nn3_insert(d1, k + 1, b0, b1, b2, i0, i1, i2);
        nn3_insert(d2, k + 2, b0, b1, b2, i0, i1, i2);
        nn3_insert(d3, k + 3, b0, b1, b2, i0, i1, i2);
      }
    }
    for (; q < tn; ++q) {
      const float d = sqdist_ref(__fsub_rn(ux, tile[q]), __fsub_rn(uy, tile[kNnTile + q]), __fsub_rn(uz, tile[2 * kNnTile + q]));
      nn3_insert(d, t0 + q, b0, b1, b2, i0, i1, i2);
    }
  }

  // merge the per-split top-3 lists (ascending split = ascending index, strict '<' keeps ties stable)
  __shared__ float s_d[kNnMaxSplits][3][32];
  __shared__ int s_i[kNnMaxSplits][3][32];
  if (splits > 1) {
    s_d[split][0][lane] = b0; s_d[split][1][lane] = b1; s_d[split][2][lane] = b2;
    s_i[split][0][lane] = i0; s_i[split][1][lane] = i1; s_i[split][2][lane] = i2;
    __syncthreads();
    if (split != 0) return;
    for (int s2 = 1; s2 < splits; ++s2) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float d = s_d[s2][q][lane];
        // unset slots are +inf with index 0: never inserted (strict '<' against a finite or inf best)
        nn3_insert(d, s_i[s2][q][lane], b0, b1, b2, i0, i1, i2);
      }
    }
  }
  if (!valid) return;
  // neighbor_interpolate.cu:61-72.  An unset best is 1e40 in the reference and +inf here; both clamp
  // to 1e10.  NaN bests follow CUDA's max/min (non-NaN operand wins) in both.
  const double lo = (double)1e-10f, hi = (double)1e10f;
  const double e0 = fmax(fmin(hi, (double)b0), lo);
  const double e1 = fmax(fmin(hi, (double)b1), lo);
  const double e2 = fmax(fmin(hi, (double)b2), lo);
  const float d0d1 = __double2float_rn(__dmul_rn(e0, e1));
  const float d0d2 = __double2float_rn(__dmul_rn(e0, e2));
  const float d1d2 = __double2float_rn(__dmul_rn(e1, e2));
  const float inv = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(d0d1, d0d2), d1d2));
  weights[j] = __fmul_rn(d1d2, inv);
  indices[j] = i0;
  weights[j + n] = __fmul_rn(d0d2, inv);
  indices[j + n] = i1;
  weights[j + n + n] = __fmul_rn(d0d1, inv);
  indices[j + n + n] = i2;
}

__global__ void __launch_bounds__(kNnThreads)
three_interp_kernel(int c, int m, int n, const float *__restrict__ feat,
                    const int *__restrict__ indices, const float *__restrict__ weights,
                    float *__restrict__ out) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * kNnThreads + threadIdx.x;
  if (j >= n) return;
  const int *ix = indices + (size_t)b * 3 * n;
  const float *w = weights + (size_t)b * 3 * n;
  const int i1 = __ldg(ix + j), i2 = __ldg(ix + j + n), i3 = __ldg(ix + j + n + n);
  const float w1 = __ldg(w + j), w2 = __ldg(w + j + n), w3 = __ldg(w + j + n + n);
  const int c0 = blockIdx.y * kNnCT;
  const int c1 = min(c0 + kNnCT, c);
  const float *f = feat + ((size_t)b * c + c0) * m;
  float *o = out + ((size_t)b * c + c0) * n + j;
  for (int cc = c0; cc < c1; ++cc) {
    float acc = __fmul_rn(__ldg(f + i2), w2);
    acc = __fmaf_rn(__ldg(f + i1), w1, acc);
    acc = __fmaf_rn(__ldg(f + i3), w3, acc);
    *o = acc;
    f += m;
    o += n;
  }
}

// Shared-memory variant: the 3 gathers per output element hit random addresses of a feature row; from
// global memory every warp-level gather costs up to 32 L1 wavefronts, from shared memory ~3 (bank
// conflicts).  A CTA stages CT rows [CT][m] once and serves a chunk of points from them.
constexpr int kNnRowThreads = 256;

template <int CT>
__global__ void __launch_bounds__(kNnRowThreads)
three_interp_rows_kernel(int c, int m, int n, int chunk, const float *__restrict__ feat,
                         const int *__restrict__ indices, const float *__restrict__ weights,
                         float *__restrict__ out) {
  // Channel-interleaved tile [m][CT]: one neighbour's CT channel values are contiguous, so a point
  // needs 3 * CT/4 128-bit shared loads instead of 3 * CT scalar ones.
  extern __shared__ __align__(16) float rows[];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * CT;
  const int nrows = min(CT, c - c0);
  const float *f = feat + ((size_t)b * c + c0) * m;
  for (int k = threadIdx.x; k < m; k += kNnRowThreads) {
    float t[CT];
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) t[cc] = cc < nrows ? ld_stream_f1(f + (size_t)cc * m + k) : 0.0f;
    if constexpr (CT >= 4) {
#pragma unroll
      for (int h = 0; h < CT / 4; ++h)
        reinterpret_cast<float4 *>(rows + (size_t)k * CT)[h] = make_float4(t[4 * h], t[4 * h + 1], t[4 * h + 2], t[4 * h + 3]);
    } else {
#pragma unroll
      for (int cc = 0; cc < CT; ++cc) rows[(size_t)k * CT + cc] = t[cc];
    }
  }
  __syncthreads();
  const int *ix = indices + (size_t)b * 3 * n;
  const float *w = weights + (size_t)b * 3 * n;
  const int j_end = min((blockIdx.x + 1) * chunk, n);
  for (int j = blockIdx.x * chunk + threadIdx.x; j < j_end; j += kNnRowThreads) {
    const int i1 = __ldg(ix + j), i2 = __ldg(ix + j + n), i3 = __ldg(ix + j + n + n);
    const float w1 = __ldg(w + j), w2 = __ldg(w + j + n), w3 = __ldg(w + j + n + n);
    float f1[CT], f2[CT], f3[CT];
    if constexpr (CT >= 4) {
#pragma unroll
      for (int h = 0; h < CT / 4; ++h) {
        const float4 a = reinterpret_cast<const float4 *>(rows + (size_t)i1 * CT)[h];
        const float4 bq = reinterpret_cast<const float4 *>(rows + (size_t)i2 * CT)[h];
        const float4 cq = reinterpret_cast<const float4 *>(rows + (size_t)i3 * CT)[h];
        f1[4 * h] = a.x; f1[4 * h + 1] = a.y; f1[4 * h + 2] = a.z; f1[4 * h + 3] = a.w;
        f2[4 * h] = bq.x; f2[4 * h + 1] = bq.y; f2[4 * h + 2] = bq.z; f2[4 * h + 3] = bq.w;
        f3[4 * h] = cq.x; f3[4 * h + 1] = cq.y; f3[4 * h + 2] = cq.z; f3[4 * h + 3] = cq.w;
      }
    } else {
#pragma unroll
      for (int cc = 0; cc < CT; ++cc) {
        f1[cc] = rows[(size_t)i1 * CT + cc];
        f2[cc] = rows[(size_t)i2 * CT + cc];
        f3[cc] = rows[(size_t)i3 * CT + cc];
      }
    }
    float *o = out + ((size_t)b * c + c0) * n + j;
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) {
      if (cc < nrows) {
        float acc = __fmul_rn(f2[cc], w2);
        acc = __fmaf_rn(f1[cc], w1, acc);
        acc = __fmaf_rn(f3[cc], w3, acc);
        o[(size_t)cc * n] = acc;
      }
    }
  }
}

// backward (neighbor_interpolate.cu:155-169): 3 atomic scatter-adds of fl(g*w) per (point, channel)
__global__ void __launch_bounds__(kNnThreads)
three_interp_grad_kernel(int c, int n, int m, const float *__restrict__ grad_y,
                         const int *__restrict__ indices, const float *__restrict__ weights,
                         float *__restrict__ grad_x) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * kNnThreads + threadIdx.x;
  if (j >= n) return;
  const int *ix = indices + (size_t)b * 3 * n;
  const float *w = weights + (size_t)b * 3 * n;
  const int i1 = ix[j], i2 = ix[j + n], i3 = ix[j + n + n];
  const float w1 = w[j], w2 = w[j + n], w3 = w[j + n + n];
  const int c0 = blockIdx.y * kNnCT;
  const int c1 = min(c0 + kNnCT, c);
  for (int cc = c0; cc < c1; ++cc) {
    const float g = grad_y[((size_t)b * c + cc) * n + j];
    float *gx = grad_x + ((size_t)b * c + cc) * m;
    atomicAdd(gx + i1, __fmul_rn(g, w1));
    atomicAdd(gx + i2, __fmul_rn(g, w2));
    atomicAdd(gx + i3, __fmul_rn(g, w3));
  }
}

}  // namespace bdm

extern "C" int bdm_three_nn_search(int b, int n, int m, const float *points_coords,
                                   const float *centers_coords, float *weights, int *indices,
                                   bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && n >= 0 && m >= 0 && b <= 65535);
  if (b == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(points_coords); BDM_CHECK_PTR(weights); BDM_CHECK_PTR(indices);
  if (m > 0) BDM_CHECK_PTR(centers_coords);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec4 = (m % 4 == 0) && ((reinterpret_cast<uintptr_t>(centers_coords) & 15) == 0);
  // enough warps to fill the machine; at least 64 centres per split
  const int ctas = ceil_div(n, 32) * b;
  int splits = 1;
  while (splits < kNnMaxSplits && ctas * splits < 48 * sm_count() && m / (splits * 2) >= 64) splits *= 2;  // warps
  if (vec4)
    three_nn_kernel<true><<<dim3(ceil_div(n, 32), b), 32 * splits, 0, st>>>(n, m, splits, points_coords,
                                                                           centers_coords, weights, indices);
  else
    three_nn_kernel<false><<<dim3(ceil_div(n, 32), b), 32 * splits, 0, st>>>(n, m, splits, points_coords,
                                                                            centers_coords, weights, indices);
  BDM_RETURN_LAUNCH_STATUS();
}

extern "C" int bdm_three_nn_interpolate(int b, int c, int m, int n, const float *centers_features,
                                        const int *indices, const float *weights, float *out,
                                        bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && m >= 0 && b <= 65535);
  if (b == 0 || c == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(centers_features); BDM_CHECK_PTR(indices); BDM_CHECK_PTR(weights); BDM_CHECK_PTR(out);
  BDM_CHECK_SIZE(ceil_div(c, kNnCT) <= 65535);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // rows in shared memory when they fit (CT*m*4 <= 64 KB), else gathers straight from global
  int ct = 8;
  while (ct > 1 && sizeof(float) * (size_t)ct * m > 64 * 1024) ct >>= 1;
  if (m >= 1 && sizeof(float) * (size_t)ct * m <= 64 * 1024 && ceil_div(c, ct) <= 65535) {
    const int tiles = ceil_div(c, ct) * b;
    int chunks = 1;
    while (tiles * chunks < 2 * sm_count() && ceil_div(n, chunks * 2) >= 2 * kNnRowThreads) chunks *= 2;
    const int chunk = ceil_div(n, chunks);
    const size_t smem = sizeof(float) * (size_t)ct * m;
    const dim3 grid(chunks, ceil_div(c, ct), b);
    cudaError_t e = cudaSuccess;
    if (ct == 8) {
      e = ensure_dynamic_smem(reinterpret_cast<const void *>(three_interp_rows_kernel<8>), smem);
      if (e == cudaSuccess) three_interp_rows_kernel<8><<<grid, kNnRowThreads, smem, st>>>(c, m, n, chunk, centers_features, indices, weights, out);
    } else if (ct == 4) {
      e = ensure_dynamic_smem(reinterpret_cast<const void *>(three_interp_rows_kernel<4>), smem);
      if (e == cudaSuccess) three_interp_rows_kernel<4><<<grid, kNnRowThreads, smem, st>>>(c, m, n, chunk, centers_features, indices, weights, out);
    } else if (ct == 2) {
      e = ensure_dynamic_smem(reinterpret_cast<const void *>(three_interp_rows_kernel<2>), smem);
      if (e == cudaSuccess) three_interp_rows_kernel<2><<<grid, kNnRowThreads, smem, st>>>(c, m, n, chunk, centers_features, indices, weights, out);
    } else {
      e = ensure_dynamic_smem(reinterpret_cast<const void *>(three_interp_rows_kernel<1>), smem);
      if (e == cudaSuccess) three_interp_rows_kernel<1><<<grid, kNnRowThreads, smem, st>>>(c, m, n, chunk, centers_features, indices, weights, out);
    }
    if (e != cudaSuccess) return (int)e;
    BDM_RETURN_LAUNCH_STATUS();
  }
  three_interp_kernel<<<dim3(ceil_div(n, kNnThreads), ceil_div(c, kNnCT), b), kNnThreads, 0, st>>>(
      c, m, n, centers_features, indices, weights, out);
  BDM_RETURN_LAUNCH_STATUS();
}

extern "C" int bdm_three_nearest_neighbors_interpolate(int b, int c, int m, int n,
                                                       const float *points_coords,
                                                       const float *centers_coords,
                                                       const float *centers_features, int *indices,
                                                       float *weights, float *out,
                                                       bdm_stream_t stream) {
  int rc = bdm_three_nn_search(b, n, m, points_coords, centers_coords, weights, indices, stream);
  if (rc != BDM_OK) return rc;
  return bdm_three_nn_interpolate(b, c, m, n, centers_features, indices, weights, out, stream);
}

extern "C" int bdm_three_nearest_neighbors_interpolate_grad(int b, int c, int n, int m,
                                                            const float *grad_y, const int *indices,
                                                            const float *weights, float *grad_x,
                                                            bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && m >= 0 && b <= 65535);
  if (b == 0 || c == 0 || m == 0) return BDM_OK;
  BDM_CHECK_PTR(grad_x);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)b * c * m, st);
  if (n > 0) {
    BDM_CHECK_PTR(grad_y); BDM_CHECK_PTR(indices); BDM_CHECK_PTR(weights);
    three_interp_grad_kernel<<<dim3(ceil_div(n, kNnThreads), ceil_div(c, kNnCT), b), kNnThreads, 0, st>>>(
        c, n, m, grad_y, indices, weights, grad_x);
  }
  BDM_RETURN_LAUNCH_STATUS();
}
