// voxelize.cu -- avg_voxelize forward/backward for sm_100a.
//
// Replaces the reference's grid_stats_kernel + avg_voxelize_kernel
// (experiments/model/pvcnn/modules/functional/src/voxelization/vox.cu:18-72, launched at :112-119 with
// one CTA per batch element and C sequential global fp32 atomics per point onto a torch::zeros'd
// 818 MB output at the first PC^2 layer).
//
// Design (B200-first): the op is bound by the dense [B,C,R^3] output write (94 % zeros at R=32), so
// the output is produced exactly once, in full 16-byte streaming stores, with no atomics and no
// pre-zeroing pass:
//   1. vox_sort_kernel   one CTA per shape: voxel index of every point, a shared-memory histogram
//                        (-> cnt), a counting sort of the <=16384 points by voxel id with a stable
//                        (ascending point index) order inside each voxel, and three small lookup
//                        tables per shape in the workspace: an occupancy bitmask (1 bit/voxel), the
//                        occupied-rank base of every 32-voxel word, and the start offset of every
//                        occupied voxel in the sorted order.
//   2. vox_fill_kernel   one CTA per (shape, tile of CT channels): (a) stages its channels' point
//                        features into shared memory *in sorted order* (coalesced global reads,
//                        permuting shared-memory writes); (b) one thread per OCCUPIED voxel sums the
//                        voxel's contiguous run of sorted points and compacts the averages in place
//                        (value j of a channel = average of the j-th occupied voxel); (c) streams over
//                        the voxel grid four voxels per thread: bitmask nibble == 0 -> zeros, else a
//                        popcount-ranked look-up of the precomputed averages -> one st.global.cs.v4
//                        per channel.  HBM traffic = read feat once + write out once; the streaming
//                        loop is ~17 instructions per 4 x 512-byte warp stores.
// Summation order inside a voxel is ascending point index (deterministic run to run); each addend is
// fl(feat * fl(1/cnt)) exactly like the reference (vox.cu:66-68), so voxels holding one or two
// points are bit-identical to the reference and the rest differ only by fp32 summation order
// (the reference's own order is unspecified: RED.ADD.F32).
//
// Sizes outside the shared-memory fast path (R^3 > 32768 or N > 16384) take a generic path with
// global atomics (memset + stats + scatter): still CUDA, still on the caller's stream.
#include "common.cuh"
#include "voxel_plan.cuh"

namespace bdm {

constexpr int kSortThreads = 1024;
constexpr int kFillThreads = 512;
// ------------------------------------------------------------------------------------------------
// 1. per-shape index / histogram / counting sort
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads, 1)
vox_sort_kernel(int n, int r, const int *__restrict__ coords, int *__restrict__ ind,
                int *__restrict__ cnt, unsigned char *__restrict__ ws, VoxAuxLayout L) {
  const int b = blockIdx.x;
  const int r2 = r * r, r3 = r2 * r;
  const int nw = L.nw;
  const int nblk = (nw + 3) >> 2;        // 128-voxel blocks (4 words), the unit of the vectorised passes
  const int nwp = nblk << 2;             // padded word count
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  extern __shared__ __align__(16) uint32_t smem_u32[];
  uint32_t *hist = smem_u32;             // [nwp*32]  count -> (cursor<<16 | in-word exclusive prefix)
  uint32_t *wbase = hist + nwp * 32;     // [nwp]     word totals -> exclusive point-offset base
  uint32_t *obase = wbase + nwp;         // [nwp]     occupied-voxel rank base
  uint32_t *wmask = obase + nwp;         // [nwp]     occupancy bits
  uint16_t *perm = reinterpret_cast<uint16_t *>(wmask + nwp);  // [n]  sorted position -> point
  uint16_t *vbuf = perm + ((n + 1) & ~1);                      // [n]  point -> voxel
  __shared__ uint32_t warp_tot[32];

  coords += (size_t)b * 3 * n;
  ind += (size_t)b * n;
  cnt += (size_t)b * r3;
  ws += (size_t)b * L.stride;
  uint32_t *g_bitmask = reinterpret_cast<uint32_t *>(ws + L.bitmask);
  uint16_t *g_obase = reinterpret_cast<uint16_t *>(ws + L.obase);
  uint16_t *g_ostart = reinterpret_cast<uint16_t *>(ws + L.ostart);
  uint16_t *g_rank = reinterpret_cast<uint16_t *>(ws + L.rank);

  {
    uint4 *h4 = reinterpret_cast<uint4 *>(hist);
    for (int q = tid; q < nwp * 8; q += kSortThreads) h4[q] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();

  // pass 1: voxel index (vox.cu:31) + histogram
  for (int i = tid; i < n; i += kSortThreads) {
    const int v = coords[i] * r2 + coords[i + n] * r + coords[i + n + n];
    ind[i] = v;
    // Out-of-range coordinates are undefined behaviour in the reference (it writes out of bounds);
    // here they are clamped into the grid for the histogram so that memory stays safe.
    const int vc = min(max(v, 0), r3 - 1);
    vbuf[i] = (uint16_t)vc;
    atomicAdd(&hist[vc], 1u);
  }
  __syncthreads();

  // pass 2 (one warp per 128-voxel block, 4 voxels per lane): cnt out, occupancy masks, in-word
  // exclusive prefix of the counts
  const bool cnt_vec = (r3 & 3) == 0 && (reinterpret_cast<uintptr_t>(cnt) & 15) == 0;
  for (int blk = warp; blk < nblk; blk += kSortThreads / 32) {
    const int v0 = blk * 128 + lane * 4;
    const uint4 c4 = reinterpret_cast<const uint4 *>(hist)[blk * 32 + lane];
    if (cnt_vec) {
      if (v0 < r3) *reinterpret_cast<int4 *>(cnt + v0) = make_int4((int)c4.x, (int)c4.y, (int)c4.z, (int)c4.w);
    } else {
      if (v0 < r3) cnt[v0] = (int)c4.x;
      if (v0 + 1 < r3) cnt[v0 + 1] = (int)c4.y;
      if (v0 + 2 < r3) cnt[v0 + 2] = (int)c4.z;
      if (v0 + 3 < r3) cnt[v0 + 3] = (int)c4.w;
    }
    const uint32_t nib = (c4.x > 0 ? 1u : 0u) | (c4.y > 0 ? 2u : 0u) | (c4.z > 0 ? 4u : 0u) | (c4.w > 0 ? 8u : 0u);
    uint32_t mask = nib << ((lane & 7) * 4);          // OR over the 8 lanes that share one 32-voxel word
    mask |= __shfl_xor_sync(0xffffffffu, mask, 1);
    mask |= __shfl_xor_sync(0xffffffffu, mask, 2);
    mask |= __shfl_xor_sync(0xffffffffu, mask, 4);
    const uint32_t tot = c4.x + c4.y + c4.z + c4.w;
    uint32_t incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    const uint32_t excl = incl - tot;                                            // within the 128-voxel block
    const uint32_t wstart = __shfl_sync(0xffffffffu, excl, lane & ~7);           // ... at the word's first lane
    const uint32_t wend = __shfl_sync(0xffffffffu, incl, lane | 7);              // ... after its last lane
    const uint32_t e0 = excl - wstart;
    reinterpret_cast<uint4 *>(hist)[blk * 32 + lane] = make_uint4(e0, e0 + c4.x, e0 + c4.x + c4.y, e0 + c4.x + c4.y + c4.z);
    if ((lane & 7) == 0) {
      const int w = blk * 4 + (lane >> 3);
      wbase[w] = (wend - wstart) | ((uint32_t)__popc(mask) << 16);  // packed: points | occupied voxels
      wmask[w] = mask;
    }
  }
  __syncthreads();

  // pass 3: block-wide exclusive scan over the (<=1024) packed word totals
  {
    const uint32_t val = (tid < nwp) ? wbase[tid] : 0u;
    uint32_t incl = val;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const uint32_t t0 = warp_tot[lane];
      uint32_t s = t0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += t;
      }
      warp_tot[lane] = s - t0;  // exclusive warp base
    }
    __syncthreads();
    const uint32_t excl = warp_tot[warp] + incl - val;
    if (tid < nwp) {
      wbase[tid] = excl & 0xffffu;
      obase[tid] = excl >> 16;
      if (tid < nw) {
        g_bitmask[tid] = wmask[tid];
        g_obase[tid] = (uint16_t)(excl >> 16);
      }
      if (tid == nwp - 1) {
        g_ostart[(excl + val) >> 16] = (uint16_t)n;  // sentinel after the last run
        reinterpret_cast<uint32_t *>(ws + L.header)[0] = (excl + val) >> 16;
      }
    }
  }
  __syncthreads();

  // pass 4: place every point inside its voxel's run (arrival order)
  for (int i = tid; i < n; i += kSortThreads) {
    const int v = vbuf[i];
    const uint32_t old = atomicAdd(&hist[v], 1u << 16);
    perm[wbase[v >> 5] + (old & 0xffffu) + (old >> 16)] = (uint16_t)i;
  }
  __syncthreads();

  // pass 5: stable rank = run start + number of run members with a smaller point index; the run's
  // first point (stable rank 0) also records the run start under the voxel's occupied rank
  for (int i = tid; i < n; i += kSortThreads) {
    const int v = vbuf[i];
    const uint32_t h = hist[v];
    const int w = v >> 5;
    const int start = wbase[w] + (h & 0xffffu);
    const int c = h >> 16;
    int rk = 0;
    for (int q = 0; q < c; ++q) rk += (perm[start + q] < i) ? 1 : 0;
    g_rank[i] = (uint16_t)(start + rk);
    if (rk == 0) g_ostart[obase[w] + __popc(wmask[w] & ((1u << (v & 31)) - 1u))] = (uint16_t)start;
  }
}

// ------------------------------------------------------------------------------------------------
// 2. dense fill: one CTA per (shape, CT channels)
// ------------------------------------------------------------------------------------------------
template <int CT, int VEC, bool COMPACT = false>
__global__ void __launch_bounds__(kFillThreads)
vox_fill_kernel(int c, int n, int r3, const float *__restrict__ feat, float *__restrict__ out,
                const unsigned char *__restrict__ ws, VoxAuxLayout L, unsigned *__restrict__ amax_bits = nullptr) {
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * CT;
  const int tid = threadIdx.x;
  const int nw = L.nw;

  extern __shared__ uint32_t smem_u32[];
  float *buf = reinterpret_cast<float *>(smem_u32);                    // [CT][n] sorted features, then
                                                                       //         per-voxel averages (in place)
  uint32_t *bitmask = smem_u32 + (size_t)CT * n;                       // [nw]
  uint16_t *obase = reinterpret_cast<uint16_t *>(bitmask + nw);        // [nw]

  ws += (size_t)b * L.stride;
  const uint32_t *g_bitmask = reinterpret_cast<const uint32_t *>(ws + L.bitmask);
  const uint16_t *g_obase = reinterpret_cast<const uint16_t *>(ws + L.obase);
  const uint16_t *g_ostart = reinterpret_cast<const uint16_t *>(ws + L.ostart);
  const uint16_t *g_rank = reinterpret_cast<const uint16_t *>(ws + L.rank);
  const int nocc = (int)reinterpret_cast<const uint32_t *>(ws + L.header)[0];

  for (int w = tid; w < nw; w += kFillThreads) {
    bitmask[w] = g_bitmask[w];
    obase[w] = g_obase[w];
  }

  // (a) stage this tile's features in sorted order: 4 points per step (one 64-bit load of 4 ranks,
  // one 128-bit load per channel), all loads of a step in flight together
  const float *f = feat + ((size_t)b * c + c0) * n;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0) {
    const int ng = n >> 2;
#pragma unroll 2
    for (int g = tid; g < ng; g += kFillThreads) {
      const uint2 rk = __ldg(reinterpret_cast<const uint2 *>(g_rank) + g);
      float4 v[CT];
#pragma unroll
      for (int cc = 0; cc < CT; ++cc)
        v[cc] = (c0 + cc < c) ? ld_stream_f4(f + (size_t)cc * n + 4 * (size_t)g) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int r0 = rk.x & 0xffffu, r1 = rk.x >> 16, r2 = rk.y & 0xffffu, r3_ = rk.y >> 16;
#pragma unroll
      for (int cc = 0; cc < CT; ++cc) {
        buf[cc * n + r0] = v[cc].x;
        buf[cc * n + r1] = v[cc].y;
        buf[cc * n + r2] = v[cc].z;
        buf[cc * n + r3_] = v[cc].w;
      }
    }
  } else {
    for (int i = tid; i < n; i += kFillThreads) {
      const int rk = g_rank[i];
#pragma unroll
      for (int cc = 0; cc < CT; ++cc)
        if (c0 + cc < c) buf[cc * n + rk] = ld_stream_f1(f + (size_t)cc * n + i);
    }
  }
  __syncthreads();

  // (b) one thread per occupied voxel: average of its run, compacted in place.  Voxel j's run starts
  // at ostart[j] >= j, so chunk k (voxels [k*T,(k+1)*T)) only reads positions >= k*T and only writes
  // positions < (k+1)*T: a barrier between the chunk's reads and its writes is all that is needed.
  for (int base = 0; base < nocc; base += kFillThreads) {
    const int j = base + tid;
    float acc[CT];
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) acc[cc] = 0.0f;
    if (j < nocc) {
      const int s = __ldg(g_ostart + j), e = __ldg(g_ostart + j + 1);  // L2-resident, read once per tile
      const float inv = __frcp_rn((float)(e - s));  // == 1.0 / float(cnt), vox.cu:65
      for (int p = s; p < e; ++p) {
#pragma unroll
        for (int cc = 0; cc < CT; ++cc) acc[cc] = __fadd_rn(acc[cc], __fmul_rn(buf[cc * n + p], inv));
      }
    }
    __syncthreads();
    if (j < nocc) {
#pragma unroll
      for (int cc = 0; cc < CT; ++cc) buf[cc * n + j] = acc[cc];
    }
  }
  __syncthreads();

  if constexpr (COMPACT) {
    // (c') sparse consumers (sparse_conv.cu) want only the occupied voxels: out[b][c][j] = average of the
    // j-th occupied voxel (ascending voxel id), zero-padded to n columns so the shape is static
    float *o = out + ((size_t)b * c + c0) * n;
    float amax = 0.0f;
    for (int j = tid; j < n; j += kFillThreads) {
#pragma unroll
      for (int cc = 0; cc < CT; ++cc)
        if (c0 + cc < c) {
          const float v = j < nocc ? buf[cc * n + j] : 0.0f;
          o[(size_t)cc * n + j] = v;
          amax = fmaxf(amax, fabsf(v));
        }
    }
    if (amax_bits != nullptr) {   // max |average| of the tensor for the fp16 consumer (conv3_tc05.cu): one atomic per warp
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, d));
      if ((tid & 31) == 0 && amax > 0.0f) atomicMax(amax_bits, __float_as_uint(amax));
    }
    return;
  }

  // (c) stream the dense grid
  float *o = out + ((size_t)b * c + c0) * r3;
  const int ngroups = r3 / VEC;
  for (int g = tid; g < ngroups; g += kFillThreads) {
    const int v0 = g * VEC;
    const uint32_t word = bitmask[v0 >> 5];
    const int sh = v0 & 31;
    const uint32_t nib = (word >> sh) & ((1u << VEC) - 1u);
    float val[CT][VEC];
#pragma unroll
    for (int cc = 0; cc < CT; ++cc)
#pragma unroll
      for (int k = 0; k < VEC; ++k) val[cc][k] = 0.0f;
    if (nib) {
      int j = obase[v0 >> 5] + __popc(word & ((1u << sh) - 1u));
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        if ((nib >> k) & 1u) {
#pragma unroll
          for (int cc = 0; cc < CT; ++cc) val[cc][k] = buf[cc * n + j];
          ++j;
        }
      }
    }
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) {
      if (c0 + cc < c) {
        if constexpr (VEC == 4) {
          st_stream_f4(o + (size_t)cc * r3 + v0,
                       make_float4(val[cc][0], val[cc][1], val[cc][2], val[cc][3]));
        } else {
          o[(size_t)cc * r3 + v0] = val[cc][0];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// generic path (large grids / clouds): global atomics, like the reference but parallel over points
// ------------------------------------------------------------------------------------------------
__global__ void vox_stats_generic_kernel(int n, int r, const int *__restrict__ coords,
                                         int *__restrict__ ind, int *__restrict__ cnt) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r2 = r * r, r3 = r2 * r;
  const int *co = coords + (size_t)b * 3 * n;
  const int v = co[i] * r2 + co[i + n] * r + co[i + n + n];
  ind[(size_t)b * n + i] = v;
  atomicAdd(cnt + (size_t)b * r3 + min(max(v, 0), r3 - 1), 1);
}

__global__ void vox_scatter_generic_kernel(int c, int n, int r3, const int *__restrict__ ind,
                                           const int *__restrict__ cnt,
                                           const float *__restrict__ feat, float *__restrict__ out) {
  const int b = blockIdx.z;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pos = min(max(ind[(size_t)b * n + i], 0), r3 - 1);
  const float inv = __frcp_rn((float)cnt[(size_t)b * r3 + pos]);
  const int c0 = blockIdx.y * 8;
  for (int cc = c0; cc < min(c0 + 8, c); ++cc)
    atomicAdd(out + ((size_t)b * c + cc) * r3 + pos, __fmul_rn(feat[((size_t)b * c + cc) * n + i], inv));
}

// backward (vox.cu:86-110): grad_x[b,c,i] = grad_y[b,c,ind[i]] * fl(1/cnt); no atomics needed --
// every (c,i) is written by exactly one thread (the reference atomically adds onto zeros).
__global__ void vox_grad_kernel(int c, int n, int s, const int *__restrict__ ind,
                                const int *__restrict__ cnt, const float *__restrict__ grad_y,
                                float *__restrict__ grad_x) {
  const int b = blockIdx.z;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pos = ind[(size_t)b * n + i];
  const int cur = cnt[(size_t)b * s + pos];
  const float inv = cur > 0 ? __frcp_rn((float)cur) : 0.0f;
  const int c0 = blockIdx.y * 8;
  for (int cc = c0; cc < min(c0 + 8, c); ++cc) {
    const float g = cur > 0 ? __fmul_rn(__ldg(grad_y + ((size_t)b * c + cc) * s + pos), inv) : 0.0f;
    grad_x[((size_t)b * c + cc) * n + i] = g + 0.0f;
  }
}



template <int CT, int VEC, bool COMPACT = false>
static cudaError_t launch_fill(int b, int c, int n, int r3, const float *feat, float *out,
                               const unsigned char *ws, const VoxAuxLayout &L, cudaStream_t st, unsigned *amax_bits = nullptr) {
  const size_t smem = sizeof(float) * (size_t)CT * n + sizeof(uint32_t) * L.nw +
                      sizeof(uint16_t) * ((L.nw + 1) & ~1);
  auto kern = vox_fill_kernel<CT, VEC, COMPACT>;
  cudaError_t e0 = ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
  if (e0 != cudaSuccess) return e0;
  dim3 grid(ceil_div(c, CT), b);
  kern<<<grid, kFillThreads, smem, st>>>(c, n, r3, feat, out, ws, L, amax_bits);
  return cudaGetLastError();
}

}  // namespace bdm

extern "C" size_t bdm_avg_voxelize_workspace_bytes(int b, int n, int r) {
  if (b <= 0 || n <= 0 || r <= 0) return 16;
  const long long r3 = (long long)r * r * r;
  if (r3 > bdm::kFastMaxR3 || !bdm::vox_fast_path(n, (int)r3)) return 16;
  return bdm::vox_aux_layout(n, (int)r3).stride * (size_t)b;
}

// Step 1 of avg_voxelize: everything that depends only on the coordinates (ind, cnt, and the sorted
// plan in the workspace).  Callers that voxelize several feature tensors over the same coordinates
// (consecutive PVConv blocks of one stage do) run this once.
extern "C" int bdm_voxel_plan(int b, int n, int r, const int *coords, int *ind, int *cnt,
                              void *workspace, size_t workspace_bytes, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && n >= 0 && r >= 1);
  const long long r3ll = (long long)r * r * r;
  BDM_CHECK_SIZE(r3ll <= 0x7fffffffLL);
  const int r3 = (int)r3ll;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (b == 0) return BDM_OK;
  BDM_CHECK_PTR(cnt);
  if (n == 0) {
    cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)b * r3, st);
    BDM_RETURN_LAUNCH_STATUS();
  }
  BDM_CHECK_PTR(coords); BDM_CHECK_PTR(ind);
  if (vox_fast_path(n, r3)) {
    const VoxAuxLayout L = vox_aux_layout(n, r3);
    const int rc = check_workspace(L, b, workspace, workspace_bytes);
    if (rc != BDM_OK) return rc;
    const size_t nwp = ((size_t)L.nw + 3) & ~(size_t)3;
    const size_t smem_sort = sizeof(uint32_t) * (nwp * 32 + 3 * nwp) + sizeof(uint16_t) * (2 * (size_t)((n + 1) & ~1));
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(vox_sort_kernel), smem_sort);
    if (e != cudaSuccess) return (int)e;
    vox_sort_kernel<<<b, kSortThreads, smem_sort, st>>>(n, r, coords, ind, cnt, static_cast<unsigned char *>(workspace), L);
    BDM_RETURN_LAUNCH_STATUS();
  }
  cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)b * r3, st);
  vox_stats_generic_kernel<<<dim3(ceil_div(n, 256), b), 256, 0, st>>>(n, r, coords, ind, cnt);
  BDM_RETURN_LAUNCH_STATUS();
}

// Step 2 of avg_voxelize: the dense [b,c,r^3] grid from features + the plan of bdm_voxel_plan
// (same b, n, r, workspace; ind/cnt are only read on the generic path).
extern "C" int bdm_avg_voxelize_fill(int b, int c, int n, int r, const int *ind, const int *cnt,
                                     const float *feat, float *out, const void *workspace,
                                     size_t workspace_bytes, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && r >= 1);
  const long long r3ll = (long long)r * r * r;
  BDM_CHECK_SIZE(r3ll <= 0x7fffffffLL);
  const int r3 = (int)r3ll;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (b == 0 || c == 0) return BDM_OK;
  BDM_CHECK_PTR(out);
  if (n == 0) {
    cudaMemsetAsync(out, 0, sizeof(float) * (size_t)b * c * r3, st);
    BDM_RETURN_LAUNCH_STATUS();
  }
  BDM_CHECK_PTR(feat);
  if (vox_fast_path(n, r3)) {
    const VoxAuxLayout L = vox_aux_layout(n, r3);
    const int rc = check_workspace(L, b, workspace, workspace_bytes);
    if (rc != BDM_OK) return rc;
    const unsigned char *ws = static_cast<const unsigned char *>(workspace);
    const bool vec4 = (r3 % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    // Channel tile: as large as possible (amortises the per-CTA lookup tables and the staging
    // barrier) while still giving every SM at least two CTAs.
    const int want = 2 * sm_count();
    int ct = 4;
    if (b * ceil_div(c, 4) < want) ct = 2;
    if (b * ceil_div(c, 2) < want) ct = 1;
    if (sizeof(float) * (size_t)ct * n > 160 * 1024) ct = (n > 8192) ? 1 : 2;
    cudaError_t e;
    if (vec4) {
      if (ct == 4) e = launch_fill<4, 4>(b, c, n, r3, feat, out, ws, L, st);
      else if (ct == 2) e = launch_fill<2, 4>(b, c, n, r3, feat, out, ws, L, st);
      else e = launch_fill<1, 4>(b, c, n, r3, feat, out, ws, L, st);
    } else {
      if (ct == 4) e = launch_fill<4, 1>(b, c, n, r3, feat, out, ws, L, st);
      else if (ct == 2) e = launch_fill<2, 1>(b, c, n, r3, feat, out, ws, L, st);
      else e = launch_fill<1, 1>(b, c, n, r3, feat, out, ws, L, st);
    }
    return e == cudaSuccess ? BDM_OK : (int)e;
  }
  BDM_CHECK_PTR(ind); BDM_CHECK_PTR(cnt);
  cudaMemsetAsync(out, 0, sizeof(float) * (size_t)b * c * r3, st);
  vox_scatter_generic_kernel<<<dim3(ceil_div(n, 256), ceil_div(c, 8), b), 256, 0, st>>>(c, n, r3, ind, cnt, feat, out);
  BDM_RETURN_LAUNCH_STATUS();
}

// Step 2', sparse flavour: only the occupied voxels' averages, out[b][c][j] for the j-th occupied voxel of
// shape b in ascending voxel id, columns nocc(b)..n-1 zero.  Same arithmetic as the dense fill (the values
// are the dense grid's non-empty entries, bit for bit).  Needs the sorted plan (fast path sizes only).
// amax_bits (or NULL): a device word that receives the bit pattern of max |out| (zeroed here first) -- the dynamic fp16
// scale of bdm_conv3_tc05_fill_planes without a separate pass over the tensor.
extern "C" int bdm_avg_voxelize_compact_amax(int b, int c, int n, int r, const float *feat, float *out,
                                             const void *workspace, size_t workspace_bytes, unsigned *amax_bits,
                                             bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 1 && r >= 1);
  const long long r3ll = (long long)r * r * r;
  BDM_CHECK_SIZE(r3ll <= kFastMaxR3 && vox_fast_path(n, (int)r3ll));
  const int r3 = (int)r3ll;
  if (b == 0 || c == 0) return BDM_OK;
  BDM_CHECK_PTR(feat); BDM_CHECK_PTR(out);
  const VoxAuxLayout L = vox_aux_layout(n, r3);
  const int rc = check_workspace(L, b, workspace, workspace_bytes);
  if (rc != BDM_OK) return rc;
  const unsigned char *ws = static_cast<const unsigned char *>(workspace);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (amax_bits != nullptr) cudaMemsetAsync(amax_bits, 0, sizeof(unsigned), st);
  const int want = 2 * sm_count();
  int ct = 4;
  if (b * ceil_div(c, 4) < want) ct = 2;
  if (b * ceil_div(c, 2) < want) ct = 1;
  if (sizeof(float) * (size_t)ct * n > 160 * 1024) ct = (n > 8192) ? 1 : 2;
  cudaError_t e;
  if (ct == 4) e = launch_fill<4, 4, true>(b, c, n, r3, feat, out, ws, L, st, amax_bits);
  else if (ct == 2) e = launch_fill<2, 4, true>(b, c, n, r3, feat, out, ws, L, st, amax_bits);
  else e = launch_fill<1, 4, true>(b, c, n, r3, feat, out, ws, L, st, amax_bits);
  return e == cudaSuccess ? BDM_OK : (int)e;
}

extern "C" int bdm_avg_voxelize_compact(int b, int c, int n, int r, const float *feat, float *out,
                                        const void *workspace, size_t workspace_bytes, bdm_stream_t stream) {
  return bdm_avg_voxelize_compact_amax(b, c, n, r, feat, out, workspace, workspace_bytes, nullptr, stream);
}

extern "C" int bdm_avg_voxelize(int b, int c, int n, int r, const int *coords, const float *feat,
                                int *ind, int *cnt, float *out, void *workspace,
                                size_t workspace_bytes, bdm_stream_t stream) {
  const int rc = bdm_voxel_plan(b, n, r, coords, ind, cnt, workspace, workspace_bytes, stream);
  if (rc != BDM_OK) return rc;
  return bdm_avg_voxelize_fill(b, c, n, r, ind, cnt, feat, out, workspace, workspace_bytes, stream);
}

extern "C" int bdm_avg_voxelize_grad(int b, int c, int n, int s, const int *ind, const int *cnt,
                                     const float *grad_y, float *grad_x, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && s >= 0);
  if (b == 0 || c == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(ind); BDM_CHECK_PTR(cnt); BDM_CHECK_PTR(grad_y); BDM_CHECK_PTR(grad_x);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  vox_grad_kernel<<<dim3(ceil_div(n, 256), ceil_div(c, 8), b), 256, 0, st>>>(c, n, s, ind, cnt, grad_y, grad_x);
  BDM_RETURN_LAUNCH_STATUS();
}
