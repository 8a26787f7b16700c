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
//   devox_grid_kernel     inference, R <= 16: the whole channel-interleaved grid tile [R^3][CT] in shared
//                         memory (one LDS.128 per corner for 4 channels), points in original order.
//   devox_bin_kernel +    inference, R = 17..32.  The 8 corner reads of a point are random addresses in a
//   devox_slab_kernel     4*R^3-byte channel row: served from global memory every warp-level gather costs up
//                         to 32 L1 wavefronts and the kernel is L1-bound at ~30 % of the HBM roofline
//                         (measured, profiles/).  Instead: points are binned by x-slice once per call (one
//                         small CTA per shape); then one CTA per (shape, 2 channels, slab of 3 x-slices)
//                         bulk-loads its slab + 1 halo slice into shared memory with TMA (cp.async.bulk,
//                         one mbarrier per slice) and serves the slab's points from there; 7 CTAs per SM
//                         hide the load latency.  HBM traffic = the grid once + the output once (the halo
//                         slices are L2 hits: neighbouring slabs run concurrently).
//   devox_gather_kernel   generic path (training: also writes inds/wgts; large R or N): thread per
//                         point, channel chunk per blockIdx.y, 8 read-only gathers per channel.
#include <cstdlib>

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

// Corner indices of a coordinate outside [0, r-1] would address outside the grid (the reference has no
// check either: trilinear_devox.cu:64-75); loads go through this clamp, results on valid input are unchanged.
__device__ __forceinline__ void devox_clamp_ids(Corner8 &k, int r3) {
#pragma unroll
  for (int q = 0; q < 8; ++q) k.id[q] = min(max(k.id[q], 0), r3 - 1);
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

// ------------------------------------------------------------------------------------------------
// fast path: x-slice binning + slice-ring gather
// ------------------------------------------------------------------------------------------------
constexpr int kSliceThreads = 256;
constexpr size_t kSlabBudget = 32 * 1024;  // slab + halo per CTA: 3+1 slices x 2 channels at R=32, 7 CTAs per SM
                                             // (swept on B200: tools/devox_sweep.sh, profiles/)
constexpr int kSliceMaxR = 32;
constexpr int kBinThreads = 1024;

struct DevoxPlanLayout {
  size_t xstart;  // i32[r+1]
  size_t spid;    // i32[n]   sorted position -> point
  size_t sxyz;    // f32[3][n] coordinates in sorted order
  size_t stride;
};

__host__ __device__ inline DevoxPlanLayout devox_plan_layout(int n, int r) {
  DevoxPlanLayout L;
  size_t off = 0;
  L.xstart = off; off = align_up(off + sizeof(int) * (r + 1), 16);
  L.spid = off;   off = align_up(off + sizeof(int) * (size_t)n, 16);
  L.sxyz = off;   off = align_up(off + sizeof(float) * 3 * (size_t)n, 16);
  L.stride = off;
  return L;
}

__device__ __forceinline__ int devox_xbin(float x, int r) { return min(max((int)floorf(x), 0), r - 1); }

__global__ void __launch_bounds__(kBinThreads)
devox_bin_kernel(int n, int r, const float *__restrict__ coords, unsigned char *__restrict__ ws,
                 DevoxPlanLayout L) {
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  coords += (size_t)b * 3 * n;
  ws += (size_t)b * L.stride;
  int *g_xstart = reinterpret_cast<int *>(ws + L.xstart);
  int *g_spid = reinterpret_cast<int *>(ws + L.spid);
  float *g_sxyz = reinterpret_cast<float *>(ws + L.sxyz);
  __shared__ int hist[kSliceMaxR];
  __shared__ int cursor[kSliceMaxR];
  if (tid < kSliceMaxR) hist[tid] = 0;
  __syncthreads();
  // up to kBinKeep points per thread stay in registers between the histogram and the scatter pass
  constexpr int kBinKeep = 4;
  float kx[kBinKeep], ky[kBinKeep], kz[kBinKeep];
#pragma unroll
  for (int j = 0; j < kBinKeep; ++j) {
    const int i = tid + j * kBinThreads;
    if (i < n) {
      kx[j] = coords[i]; ky[j] = coords[i + n]; kz[j] = coords[i + n + n];
      atomicAdd(&hist[devox_xbin(kx[j], r)], 1);
    }
  }
  for (int i = tid + kBinKeep * kBinThreads; i < n; i += kBinThreads) atomicAdd(&hist[devox_xbin(coords[i], r)], 1);
  __syncthreads();
  if (tid < 32) {  // kSliceMaxR == 32: one warp scans the bins
    const int c = tid < r ? hist[tid] : 0;
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (tid >= d) incl += t;
    }
    cursor[tid] = incl - c;
    if (tid < r) g_xstart[tid] = incl - c;
    if (tid == 31) g_xstart[r] = n;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kBinKeep; ++j) {
    const int i = tid + j * kBinThreads;
    if (i < n) {
      const int pos = atomicAdd(&cursor[devox_xbin(kx[j], r)], 1);
      g_spid[pos] = i;
      g_sxyz[pos] = kx[j];
      g_sxyz[pos + n] = ky[j];
      g_sxyz[pos + 2 * (size_t)n] = kz[j];
    }
  }
  for (int i = tid + kBinKeep * kBinThreads; i < n; i += kBinThreads) {
    const float x = coords[i], y = coords[i + n], z = coords[i + n + n];
    const int pos = atomicAdd(&cursor[devox_xbin(x, r)], 1);
    g_spid[pos] = i;
    g_sxyz[pos] = x;
    g_sxyz[pos + n] = y;
    g_sxyz[pos + 2 * (size_t)n] = z;
  }
}

template <int CT> struct VecOf;
template <> struct VecOf<4> { using type = float4; };
template <> struct VecOf<2> { using type = float2; };
template <> struct VecOf<1> { using type = float; };

// CT consecutive floats from shared memory in as few (<= 128-bit) loads as possible
template <int CT>
__device__ __forceinline__ void lds_vec(const float *p, float (&v)[CT]) {
  if constexpr (CT > 4) {
#pragma unroll
    for (int h = 0; h < CT / 4; ++h) {
      const float4 t = reinterpret_cast<const float4 *>(p)[h];
      v[4 * h] = t.x; v[4 * h + 1] = t.y; v[4 * h + 2] = t.z; v[4 * h + 3] = t.w;
    }
  } else {
    using V = typename VecOf<CT>::type;
    const V t = *reinterpret_cast<const V *>(p);
    const float *tf = reinterpret_cast<const float *>(&t);
#pragma unroll
    for (int q = 0; q < CT; ++q) v[q] = tf[q];
  }
}

template <int CT>
__device__ __forceinline__ void sts_vec(float *p, const float (&v)[CT]) {
  if constexpr (CT > 4) {
#pragma unroll
    for (int h = 0; h < CT / 4; ++h)
      reinterpret_cast<float4 *>(p)[h] = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
  } else {
    using V = typename VecOf<CT>::type;
    V t;
    float *tf = reinterpret_cast<float *>(&t);
#pragma unroll
    for (int q = 0; q < CT; ++q) tf[q] = v[q];
    *reinterpret_cast<V *>(p) = t;
  }
}

// Small grids (R^3 * CT * 4 <= 64 KB: R <= 16): the whole channel-interleaved grid tile [R^3][CT] sits
// in shared memory, points are served in their original order (no binning pass, coalesced stores).
constexpr int kGridThreads = 256;

template <int CT>
__global__ void __launch_bounds__(kGridThreads)
devox_grid_kernel(int c, int n, int r, int chunk, const float *__restrict__ coords,
                  const float *__restrict__ feat, float *__restrict__ outs) {
  extern __shared__ __align__(16) float tile[];  // [r3][CT]
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * CT;
  const int nch = min(CT, c - c0);
  const int r2 = r * r, r3 = r2 * r;
  const float *fbase = feat + ((size_t)b * c + c0) * r3;
  for (int v = threadIdx.x; v < r3; v += kGridThreads) {
    float t[CT];
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) t[cc] = cc < nch ? ld_stream_f1(fbase + (size_t)cc * r3 + v) : 0.0f;
    sts_vec<CT>(tile + (size_t)v * CT, t);
  }
  __syncthreads();
  const float *co = coords + (size_t)b * 3 * n;
  float *o = outs + ((size_t)b * c + c0) * n;
  const int i_end = min((blockIdx.y + 1) * chunk, n);
  for (int i = blockIdx.y * chunk + threadIdx.x; i < i_end; i += kGridThreads) {
    Corner8 k;
    devox_corners(co[i], co[i + n], co[i + n + n], r, r2, k);
    float f[8][CT];
#pragma unroll
    for (int q = 0; q < 8; ++q) lds_vec<CT>(tile + (size_t)min(max(k.id[q], 0), r3 - 1) * CT, f[q]);
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) {
      float acc = __fmul_rn(k.w[1], f[1][cc]);
      acc = __fmaf_rn(k.w[0], f[0][cc], acc);
#pragma unroll
      for (int q = 2; q < 8; ++q) acc = __fmaf_rn(k.w[q], f[q][cc], acc);
      if (cc < nch) o[(size_t)cc * n + i] = acc;
    }
  }
}

// R = 17..32 (even): one CTA per (shape, CT channels, slab of XS x-slices).  The slab plus one halo
// slice is bulk-loaded into shared memory by TMA (cp.async.bulk: one 4 KB copy per slice row issued by
// one thread, completion on an mbarrier, no register staging), then the slab's points -- a contiguous run of the x-sorted
// order -- gather their 8 corners from shared memory.  No per-slice pipeline: latency is hidden by the
// several CTAs resident per SM (7 at R=32).  Results are stored straight to global memory by original point index.
template <int CT>
__global__ void __launch_bounds__(1024)
devox_slab_kernel(int c, int n, int r, int xs_per_slab, const float *__restrict__ feat,
                  const unsigned char *__restrict__ ws, DevoxPlanLayout L, float *__restrict__ outs) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * CT;
  const int nch = min(CT, c - c0);
  const int x0 = blockIdx.y * xs_per_slab;
  const int x1 = min(x0 + xs_per_slab, r);           // slab = slices [x0, x1), halo = slice x1 (if < r)
  const int nslices = min(x1 + 1, r) - x0;
  const int tid = threadIdx.x;
  const int r2 = r * r;
  const size_t r3 = (size_t)r2 * r;

  // One mbarrier per slice: thread `sl` of warp 0 arms barrier sl and issues that slice's CT row copies.
  // A point of bin xb only waits for slices xb and xb+1, so the gather on the first bins overlaps the
  // arrival of the later slices (points are visited in x-sorted order).
  extern __shared__ __align__(128) float slab[];       // [nslices][CT][r2]
  __shared__ __align__(8) uint64_t mbar[kSliceMaxR + 1];
  const float *fbase = feat + ((size_t)b * c + c0) * r3 + (size_t)x0 * r2;
  const unsigned row_bytes = (unsigned)(sizeof(float) * r2);
  if (tid < nslices) mbar_init(&mbar[tid], 1);
  __syncthreads();
  if (tid < nslices) {
    mbar_expect_tx(&mbar[tid], row_bytes * (unsigned)nch);
    for (int cc = 0; cc < nch; ++cc)
      bulk_g2s(slab + ((size_t)tid * CT + cc) * r2, fbase + (size_t)cc * r3 + (size_t)tid * r2, row_bytes, &mbar[tid]);
  }

  ws += (size_t)b * L.stride;
  const int *g_xstart = reinterpret_cast<const int *>(ws + L.xstart);
  const int p_begin = __ldg(g_xstart + x0), p_end = __ldg(g_xstart + x1);
  const int *g_spid = reinterpret_cast<const int *>(ws + L.spid);
  const float *g_sx = reinterpret_cast<const float *>(ws + L.sxyz);
  const float *g_sy = g_sx + n;
  const float *g_sz = g_sy + n;
  float *obase = outs + ((size_t)b * c + c0) * n;

  int p = p_begin + tid;
  float px = 0.f, py = 0.f, pz = 0.f;
  int pid = 0;
  if (p < p_end) { px = g_sx[p]; py = g_sy[p]; pz = g_sz[p]; pid = g_spid[p]; }
  int ready = -1;  // slices [0, ready] are known to have landed

  while (p < p_end) {
    const int pn = p + (int)blockDim.x;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    int npid = 0;
    if (pn < p_end) { nx = g_sx[pn]; ny = g_sy[pn]; nz = g_sz[pn]; npid = g_spid[pn]; }
    // trilinear_devox.cu:37-75 with slab-relative offsets
    const float xl = floorf(px), yl = floorf(py), zl = floorf(pz);
    const float xd1 = __fsub_rn(px, xl), yd1 = __fsub_rn(py, yl), zd1 = __fsub_rn(pz, zl);
    const float xd0 = __fsub_rn(1.0f, xd1), yd0 = __fsub_rn(1.0f, yd1), zd0 = __fsub_rn(1.0f, zd1);
    const float w00 = __fmul_rn(xd0, yd0), w01 = __fmul_rn(xd0, yd1);
    const float w10 = __fmul_rn(xd1, yd0), w11 = __fmul_rn(xd1, yd1);
    const float w0 = __fmul_rn(w00, zd0), w1 = __fmul_rn(w00, zd1), w2 = __fmul_rn(w01, zd0),
                w3 = __fmul_rn(w01, zd1), w4 = __fmul_rn(w10, zd0), w5 = __fmul_rn(w10, zd1),
                w6 = __fmul_rn(w11, zd0), w7 = __fmul_rn(w11, zd1);
    const int xb = devox_xbin(px, r);  // the bin this point was sorted into: x0 <= xb < x1
    const int need = min(xb - x0 + 1, nslices - 1);
    while (ready < need) mbar_wait(&mbar[++ready], 0);
    const int ylo = min(max((int)yl, 0), r - 1), zlo = min(max((int)zl, 0), r - 1);
    const int yo = (yd1 > 0.0f && ylo < r - 1) ? r : 0, zo = (zd1 > 0.0f && zlo < r - 1) ? 1 : 0;
    const float *A = slab + (size_t)(xb - x0) * CT * r2;
    const float *Hi = (xd1 > 0.0f && xb + 1 < r) ? A + (size_t)CT * r2 : A;
    const int o00 = ylo * r + zlo, o01 = o00 + zo, o10 = o00 + yo, o11 = o10 + zo;
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) {
      const float *a = A + cc * r2, *h = Hi + cc * r2;
      float acc = __fmul_rn(w1, a[o01]);
      acc = __fmaf_rn(w0, a[o00], acc);
      acc = __fmaf_rn(w2, a[o10], acc);
      acc = __fmaf_rn(w3, a[o11], acc);
      acc = __fmaf_rn(w4, h[o00], acc);
      acc = __fmaf_rn(w5, h[o01], acc);
      acc = __fmaf_rn(w6, h[o10], acc);
      acc = __fmaf_rn(w7, h[o11], acc);
      if (cc < nch) obase[(size_t)cc * n + pid] = acc;
    }
    p = pn; px = nx; py = ny; pz = nz; pid = npid;
  }
  // a CTA must not retire while bulk copies into its shared memory are still in flight
  while (ready < nslices - 1) mbar_wait(&mbar[++ready], 0);
}

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
  devox_clamp_ids(k, (int)r3);
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

// Channels-last grid (inference): feat f32[b][r^3][c] as the channels-last voxel branch leaves it
// (modules/point_voxel.py).  A warp takes one point: its 8 corner rows are c contiguous floats each
// (lane = channel, fully coalesced), blended with the same weights in the same order as everywhere else
// in this file; a [32 points][c] tile is transposed through shared memory so that outs[b][c][n] is
// written as 128-byte row segments.
constexpr int kDevoxClWarps = 8;

// PTS = points per CTA: 32 (128-byte output rows) when there are enough points to fill the GPU that way,
// 8 (one point per warp) for the coarse stages, whose 64-1024 points per shape would otherwise leave most
// SMs idle while each warp walks 4 points x C/32 dependent-latency steps.
// NORM: feat is the un-normalised output of the block's last convolution and every corner value goes through
// y = swish(x * A + B) (coef[b][c] = (A, B), the GroupNorm + Swish of modules/pvconv.py:82-83 with the statistics
// folded in) before it is weighted -- the same expression the stand-alone norm kernel evaluates, so the result is
// bit-identical to devoxelizing the normalised grid, which is then never written.
__device__ __forceinline__ float devox_swish(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
template <int PTS, bool NORM = false>
__global__ void __launch_bounds__(kDevoxClWarps * 32)
devox_cl_kernel(int c, int n, int r, const float *__restrict__ coords, const float *__restrict__ feat,
                const float *__restrict__ gate, const float *residual, float *outs,
                const float2 *__restrict__ coef = nullptr, int swish = 0) {
  extern __shared__ float tile[];   // [PTS][c + 1]
  const int b = blockIdx.y, i0 = blockIdx.x * PTS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r2 = r * r;
  const size_t r3 = (size_t)r2 * r;
  const float *co = coords + (size_t)b * 3 * n;
  const float *f = feat + (size_t)b * r3 * c;
  const int ld = c + 1;
  for (int p = warp; p < PTS; p += kDevoxClWarps) {
    const int i = i0 + p;
    if (i >= n) break;
    Corner8 k;
    devox_corners(__ldg(co + i), __ldg(co + i + n), __ldg(co + i + n + n), r, r2, k);
    devox_clamp_ids(k, (int)r3);
    for (int cc = lane; cc < c; cc += 32) {
      float f0 = __ldg(f + (size_t)k.id[0] * c + cc), f1 = __ldg(f + (size_t)k.id[1] * c + cc),
            f2 = __ldg(f + (size_t)k.id[2] * c + cc), f3 = __ldg(f + (size_t)k.id[3] * c + cc),
            f4 = __ldg(f + (size_t)k.id[4] * c + cc), f5 = __ldg(f + (size_t)k.id[5] * c + cc),
            f6 = __ldg(f + (size_t)k.id[6] * c + cc), f7 = __ldg(f + (size_t)k.id[7] * c + cc);
      if (NORM) {
        const float2 ab = __ldg(coef + (size_t)b * c + cc);
        f0 = fmaf(f0, ab.x, ab.y); f1 = fmaf(f1, ab.x, ab.y); f2 = fmaf(f2, ab.x, ab.y); f3 = fmaf(f3, ab.x, ab.y);
        f4 = fmaf(f4, ab.x, ab.y); f5 = fmaf(f5, ab.x, ab.y); f6 = fmaf(f6, ab.x, ab.y); f7 = fmaf(f7, ab.x, ab.y);
        if (swish) {
          f0 = devox_swish(f0); f1 = devox_swish(f1); f2 = devox_swish(f2); f3 = devox_swish(f3);
          f4 = devox_swish(f4); f5 = devox_swish(f5); f6 = devox_swish(f6); f7 = devox_swish(f7);
        }
      }
      float acc = __fmul_rn(k.w[1], f1);
      acc = __fmaf_rn(k.w[0], f0, acc);
      acc = __fmaf_rn(k.w[2], f2, acc);
      acc = __fmaf_rn(k.w[3], f3, acc);
      acc = __fmaf_rn(k.w[4], f4, acc);
      acc = __fmaf_rn(k.w[5], f5, acc);
      acc = __fmaf_rn(k.w[6], f6, acc);
      acc = __fmaf_rn(k.w[7], f7, acc);
      tile[p * ld + cc] = acc;
    }
  }
  __syncthreads();
  const int np = min(PTS, n - i0);
  float *o = outs + (size_t)b * c * n + i0;
  // optional epilogue of the PVConv block (pvconv.py:97 after se.py:19): out = devox * gate[b,c] + residual
  constexpr int kRowsPerPass = kDevoxClWarps * 32 / PTS;   // channels written per pass of the CTA
  const int pl = threadIdx.x % PTS, c_first = threadIdx.x / PTS;
  for (int cc = c_first; cc < c; cc += kRowsPerPass)
    if (pl < np) {
      float v = tile[pl * ld + cc];
      if (gate != nullptr) v *= __ldg(gate + (size_t)b * c + cc);
      if (residual != nullptr) v += residual[(size_t)b * c * n + i0 + (size_t)cc * n + pl];
      o[(size_t)cc * n + pl] = v;
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

namespace bdm {

// slab width (x-slices per CTA) such that slab + halo fits `budget` bytes of shared memory
static int devox_slab_width(int ct, int r, size_t budget) {
  const size_t per_slice = sizeof(float) * (size_t)ct * r * r;
  const int fit = (int)(budget / per_slice) - 1;
  return fit < 1 ? 0 : (fit > r ? r : fit);
}

// channel tile for the slab kernel, or 0 when the fast path does not apply
static int devox_slice_ct(int b, int c, int n, int r, int is_training) {
  if (is_training || r > kSliceMaxR || r < 1 || n < 1 || c < 1) return 0;
  if (((r * r) & 3) != 0) return 0;  // 16-byte cp.async chunks need r^2 % 4 == 0
  if (b > 65535) return 0;
  int ct = 2;
  if (c == 1) ct = 1;
  if (devox_slab_width(ct, r, kSlabBudget) < 1) ct = 1;
  if (devox_slab_width(ct, r, kSlabBudget) < 1) return 0;
  if (ceil_div(c, ct) > 0x7fffffff) return 0;
  return ct;
}

// channel tile for the whole-grid kernel, or 0 when the grid tile does not fit
static int devox_grid_ct(int b, int c, int n, int r, int is_training) {
  if (is_training || n < 1 || c < 1) return 0;
  const size_t r3 = (size_t)r * r * r;
  if (r3 * sizeof(float) > 64 * 1024) return 0;
  int ct = 8;
  while (ct > 1 && (r3 * ct * sizeof(float) > 64 * 1024 || b * ceil_div(c, ct) < 2 * sm_count())) ct >>= 1;
  if (ceil_div(c, ct) > 65535 || b > 65535) return 0;
  return ct;
}

template <int CT>
static cudaError_t launch_grid(int b, int c, int n, int r, const float *coords, const float *feat, float *outs,
                               cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t)r * r * r * CT;
  auto kern = devox_grid_kernel<CT>;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
  if (e != cudaSuccess) return e;
  const int tiles = ceil_div(c, CT) * b;
  int chunks = 1;
  while (tiles * chunks < 2 * sm_count() && ceil_div(n, chunks * 2) >= kGridThreads && chunks < 64) chunks *= 2;
  kern<<<dim3(ceil_div(c, CT), chunks, b), kGridThreads, smem, st>>>(c, n, r, ceil_div(n, chunks), coords, feat, outs);
  return cudaGetLastError();
}

template <int CT>
static cudaError_t launch_slab(int b, int c, int n, int r, const float *feat, const unsigned char *ws,
                               const DevoxPlanLayout &L, float *outs, cudaStream_t st) {
  static const char *env_t = getenv("BDM_DEVOX_THREADS");   // tuning hooks
  static const char *env_b = getenv("BDM_DEVOX_BUDGET_KB");
  const int threads = env_t ? atoi(env_t) : kSliceThreads;
  const size_t budget = env_b ? (size_t)atoi(env_b) * 1024 : kSlabBudget;
  int width = devox_slab_width(CT, r, budget);
  if (width < 1) width = 1;
  const int nslabs = ceil_div(r, width);
  const size_t smem = sizeof(float) * (size_t)(width + 1) * CT * r * r;
  auto kern = devox_slab_kernel<CT>;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
  if (e != cudaSuccess) return e;
  kern<<<dim3(ceil_div(c, CT), nslabs, b), threads, smem, st>>>(c, n, r, width, feat, ws, L, outs);
  return cudaGetLastError();
}

}  // namespace bdm

extern "C" size_t bdm_trilinear_devoxelize_workspace_bytes(int b, int n, int r) {
  if (b <= 0 || n <= 0 || r <= 0 || r > bdm::kSliceMaxR) return 16;
  return bdm::devox_plan_layout(n, r).stride * (size_t)b;
}

// Coordinate-only half of the inference fast path: x-slice binning of the points into `workspace`.
// A no-op for sizes the slab kernel does not serve.  Callers that devoxelize several grids over the
// same coordinates (consecutive PVConv blocks of a stage) run it once and pass planned=1 afterwards.
extern "C" int bdm_trilinear_devoxelize_plan(int b, int n, int r, const float *coords, void *workspace,
                                             size_t workspace_bytes, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && n >= 0 && r >= 1);
  if (b == 0 || n == 0 || r > kSliceMaxR || ((r * r) & 3) != 0) return BDM_OK;
  if (sizeof(float) * (size_t)r * r * r <= 64 * 1024) return BDM_OK;  // served by devox_grid_kernel: no plan
  BDM_CHECK_PTR(coords);
  const DevoxPlanLayout L = devox_plan_layout(n, r);
  if (workspace == nullptr) return BDM_ERR_NULL_POINTER;
  if (workspace_bytes < L.stride * (size_t)b) return BDM_ERR_WORKSPACE_TOO_SMALL;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return BDM_ERR_MISALIGNED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  devox_bin_kernel<<<b, kBinThreads, 0, st>>>(n, r, coords, static_cast<unsigned char *>(workspace), L);
  BDM_RETURN_LAUNCH_STATUS();
}

extern "C" int bdm_trilinear_devoxelize(int b, int c, int n, int r, int is_training,
                                        const float *coords, const float *feat, int *inds,
                                        float *wgts, float *outs, void *workspace,
                                        size_t workspace_bytes, int planned, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && r >= 1 && (long long)r * r * r <= 0x7fffffffLL);
  if (b == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(coords);
  if (c > 0) { BDM_CHECK_PTR(feat); BDM_CHECK_PTR(outs); }
  if (is_training) { BDM_CHECK_PTR(inds); BDM_CHECK_PTR(wgts); }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  const int gct = devox_grid_ct(b, c, n, r, is_training);
  if (gct > 0) {
    cudaError_t e;
    if (gct == 8) e = launch_grid<8>(b, c, n, r, coords, feat, outs, st);
    else if (gct == 4) e = launch_grid<4>(b, c, n, r, coords, feat, outs, st);
    else if (gct == 2) e = launch_grid<2>(b, c, n, r, coords, feat, outs, st);
    else e = launch_grid<1>(b, c, n, r, coords, feat, outs, st);
    return e == cudaSuccess ? BDM_OK : (int)e;
  }

  const int ct = devox_slice_ct(b, c, n, r, is_training);
  const DevoxPlanLayout L = devox_plan_layout(n, r);
  if (ct > 0 && workspace != nullptr && workspace_bytes >= L.stride * (size_t)b &&
      (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0) {
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    cudaError_t e = cudaSuccess;
    if (!planned) {
      devox_bin_kernel<<<b, kBinThreads, 0, st>>>(n, r, coords, ws, L);
      e = cudaGetLastError();
      if (e != cudaSuccess) return (int)e;
    }
    if (ct == 2) e = launch_slab<2>(b, c, n, r, feat, ws, L, outs, st);
    else e = launch_slab<1>(b, c, n, r, feat, ws, L, outs, st);
    return e == cudaSuccess ? BDM_OK : (int)e;
  }

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

// Inference devoxelization from a channels-last grid feat f32[b][r^3][c] -> outs f32[b][c][n]; same
// arithmetic (weights, corner order, fma chain) as bdm_trilinear_devoxelize.  Optional epilogue:
// outs = devox * gate[b][c] + residual[b][c][n] (either may be NULL; residual may alias outs).
// coef (or NULL): f32[b][c][2] = (A, B); when given, feat is the un-normalised grid and every corner value is taken through
// act(x * A + B) first (act = Swish when swish != 0): devoxelization of GroupNorm+Swish(feat) without that tensor.
extern "C" int bdm_trilinear_devoxelize_cl_norm(int b, int c, int n, int r, const float *coords, const float *feat,
                                                const float *coef, int swish, const float *gate, const float *residual,
                                                float *outs, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && r >= 1 && b <= 65535);
  BDM_CHECK_SIZE((long long)r * r * r <= 0x7fffffffLL && c <= 8192);
  if (b == 0 || c == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(coords); BDM_CHECK_PTR(feat); BDM_CHECK_PTR(outs);
  if ((reinterpret_cast<uintptr_t>(coef) & 7) != 0) return BDM_ERR_MISALIGNED;
  const float2 *cf = reinterpret_cast<const float2 *>(coef);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((long long)ceil_div(n, 32) * b >= 4LL * sm_count()) {
    const size_t smem = sizeof(float) * 32 * (size_t)(c + 1);
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(devox_cl_kernel<32, false>), smem);
    if (e == cudaSuccess) e = ensure_dynamic_smem(reinterpret_cast<const void *>(devox_cl_kernel<32, true>), smem);
    if (e != cudaSuccess) return (int)e;
    if (cf != nullptr)
      devox_cl_kernel<32, true><<<dim3(ceil_div(n, 32), b), kDevoxClWarps * 32, smem, st>>>(c, n, r, coords, feat, gate, residual, outs, cf, swish);
    else
      devox_cl_kernel<32, false><<<dim3(ceil_div(n, 32), b), kDevoxClWarps * 32, smem, st>>>(c, n, r, coords, feat, gate, residual, outs);
  } else {
    const size_t smem = sizeof(float) * 8 * (size_t)(c + 1);
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(devox_cl_kernel<8, false>), smem);
    if (e == cudaSuccess) e = ensure_dynamic_smem(reinterpret_cast<const void *>(devox_cl_kernel<8, true>), smem);
    if (e != cudaSuccess) return (int)e;
    if (cf != nullptr)
      devox_cl_kernel<8, true><<<dim3(ceil_div(n, 8), b), kDevoxClWarps * 32, smem, st>>>(c, n, r, coords, feat, gate, residual, outs, cf, swish);
    else
      devox_cl_kernel<8, false><<<dim3(ceil_div(n, 8), b), kDevoxClWarps * 32, smem, st>>>(c, n, r, coords, feat, gate, residual, outs);
  }
  BDM_RETURN_LAUNCH_STATUS();
}

extern "C" int bdm_trilinear_devoxelize_cl(int b, int c, int n, int r, const float *coords, const float *feat,
                                           const float *gate, const float *residual, float *outs,
                                           bdm_stream_t stream) {
  return bdm_trilinear_devoxelize_cl_norm(b, c, n, r, coords, feat, nullptr, 0, gate, residual, outs, stream);
}
