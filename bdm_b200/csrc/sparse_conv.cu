// sparse_conv.cu -- the first 3x3x3 convolution of a PVConv block, exploiting that its input is a
// freshly voxelized point cloud.
//
// The reference materialises avg_voxelize's dense grid and runs nn.Conv3d over it
// (experiments/model/pvcnn/modules/pvconv.py:75-76,91-97: `voxel_features, voxel_coords =
// self.voxelization(features, coords); voxel_features = self.voxel_layers(voxel_features)`).  At R=32 a
// 4096-point cloud occupies ~5 % of the 32768 voxels, so 95 % of that convolution's multiply-adds have a
// zero operand and the 818 MB grid of the first PC^2 layer (C=390) is written, transposed and read only
// to carry ~1800 non-zero columns per shape.  Convolution is linear, so the same result is
//
//     taps[b][j][k][co] = sum_ci W[co][ci][k] * avg[b][ci][j]        j-th occupied voxel, k = (kd*3+kh)*3+kw
//     out[b][co][v]     = sum_k taps[b][slot(v + k - 1)][k][co]      over the occupied neighbours of v
//
// i.e. one GEMM over the occupied voxels only (cuBLAS, issued by the host side on the output of
// bdm_avg_voxelize_compact) followed by the gather below, which is the only pass that touches a dense
// grid -- the convolution's OUTPUT, written once.  Taps are accumulated in ascending k for every output
// voxel (deterministic).
//
// Kernel: one warp per z-row (x,y) of the output and 32 output channels.  It lists the occupied voxels of
// the 9 neighbouring rows (occupancy bits + popcount ranks, no search); an occupied neighbour (x',y',z')
// contributes its three kw taps -- 3 x 32 contiguous floats, lane-coalesced, four neighbours in flight at a
// time -- to outputs z'+1, z', z'-1 of a [32 z][32 co] tile in shared memory, which is then written
// transposed as 128-byte row segments.
#include <cstdlib>

#include "common.cuh"
#include "voxel_plan.cuh"

namespace bdm {

constexpr int kGatherWarps = 8;
constexpr int kGatherCo = 32;
constexpr int kGatherLd = kGatherCo + 1;   // tile row stride (conflict-free column and row access)

// BATCH = occupied neighbours whose taps are in flight together (3 loads each)
template <int BATCH>
__global__ void __launch_bounds__(kGatherWarps * 32)
sparse_conv3_gather_kernel(int n, int cout, int r, int channels_last, const float *__restrict__ taps,
                           const float *__restrict__ bias, float *__restrict__ out, double2 *__restrict__ stats,
                           const unsigned char *__restrict__ ws, VoxAuxLayout L) {
  __shared__ __align__(16) float tile[kGatherWarps][32][kGatherLd];
  __shared__ uint32_t entries[kGatherWarps][9 * 32];   // slot | z' << 16 | neighbour row << 24
  const int b = blockIdx.z, co0 = blockIdx.y * kGatherCo;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kGatherWarps + warp;  // x * r + y
  // per-channel (sum, sum of squares) of this CTA's rows of the (bias-less) output, for the GroupNorm that
  // follows: stats[b][blockIdx.x][channel].  Rows beyond the grid contribute zero.  The per-warp partials
  // reuse each warp's entry list (dead by then).
  float2(*wsum)[9 * 16] = reinterpret_cast<float2(*)[9 * 16]>(entries);
  // rows beyond the grid exist only when r*r is not a multiple of the warps per CTA (r = 1, 2).  Without
  // `stats` such a warp may leave at once; with it, it stays (as a row without neighbours: zero sums) until
  // the block-wide barrier of the statistics reduction, so that every warp of the CTA reaches that barrier.
  const bool valid = row < r * r;
  if (!valid && stats == nullptr) return;
  const int x = row / r, y = row - x * r;
  const int r3 = r * r * r;

  ws += (size_t)b * L.stride;
  const uint32_t *bitmask = reinterpret_cast<const uint32_t *>(ws + L.bitmask);
  const uint16_t *obase = reinterpret_cast<const uint16_t *>(ws + L.obase);
  const uint32_t rowmask = r >= 32 ? 0xffffffffu : ((1u << r) - 1u);

  // lane j < 9 fetches the occupancy bits of neighbouring row j = (dx+1)*3 + (dy+1) and the occupied
  // rank of that row's first voxel: one round of loads for the whole neighbourhood
  uint32_t my_bits = 0u;
  int my_slot = 0;
  if (lane < 9 && valid) {
    const int xx = x + lane / 3 - 1, yy = y + lane % 3 - 1;
    if (xx >= 0 && xx < r && yy >= 0 && yy < r) {
      const int bit0 = (xx * r + yy) * r;
      const uint32_t word = __ldg(bitmask + (bit0 >> 5));
      const int sh = bit0 & 31;
      my_bits = (word >> sh) & rowmask;
      my_slot = (int)__ldg(obase + (bit0 >> 5)) + __popc(word & ((1u << sh) - 1u));
    }
  }

  float(*t)[kGatherLd] = tile[warp];
  {
    float4 *t4 = reinterpret_cast<float4 *>(&t[0][0]);
#pragma unroll
    for (int q = lane; q < 32 * kGatherLd / 4; q += 32) t4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncwarp();

  // the occupied neighbours as a list, row-major then z ascending (= tap index k ascending for every
  // output voxel): lane z appends its own voxel of each row
  uint32_t *ent = entries[warp];
  int total = 0;
  const uint32_t below = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const uint32_t bits = __shfl_sync(0xffffffffu, my_bits, j);
    const int slot0 = __shfl_sync(0xffffffffu, my_slot, j);
    if ((bits >> lane) & 1u) {
      const int before = __popc(bits & below);
      ent[total + before] = (uint32_t)(slot0 + before) | ((uint32_t)lane << 16) | ((uint32_t)j << 24);
    }
    total += __popc(bits);
  }
  __syncwarp();

  const bool co_ok = co0 + lane < cout;
  const float *tp = taps + (size_t)b * n * 27 * cout + co0 + lane;
  for (int e0 = 0; e0 < total; e0 += BATCH) {
    uint32_t en[BATCH];
    float a[BATCH][3];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      en[u] = e0 + u < total ? ent[e0 + u] : 0xffffffffu;
      a[u][0] = a[u][1] = a[u][2] = 0.0f;
      if (en[u] != 0xffffffffu && co_ok) {
        const float *q = tp + ((size_t)(en[u] & 0xffffu) * 27 + (en[u] >> 24) * 3) * cout;
        a[u][0] = __ldg(q);
        a[u][1] = __ldg(q + cout);
        a[u][2] = __ldg(q + 2 * cout);
      }
    }
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      if (en[u] != 0xffffffffu) {
        const int zp = (en[u] >> 16) & 31;
        // input z' = z + kw - 1  =>  tap kw lands on output z = z' + 1 - kw
        if (zp + 1 < r) t[zp + 1][lane] += a[u][0];
        t[zp][lane] += a[u][1];
        if (zp >= 1) t[zp - 1][lane] += a[u][2];
      }
    }
  }
  __syncwarp();

  const int nco = min(kGatherCo, cout - co0);
  if (stats != nullptr) {
    float a1 = 0.0f, a2 = 0.0f;
    if (total != 0)
      for (int z = 0; z < r; ++z) { const float v = t[z][lane]; a1 += v; a2 = fmaf(v, v, a2); }
    __syncwarp();   // every lane is done reading this warp's entries
    wsum[warp][lane] = make_float2(a1, a2);
    __syncthreads();
    if (warp == 0 && lane < nco) {
      double a1 = 0.0, a2 = 0.0;
      for (int w = 0; w < kGatherWarps; ++w) { a1 += wsum[w][lane].x; a2 += wsum[w][lane].y; }
      stats[((size_t)b * gridDim.x + blockIdx.x) * cout + co0 + lane] = make_double2(a1, a2);
    }
  }
  if (!valid) return;
  if (channels_last) {
    // out[b][voxel][co]: the tile rows are already channel-contiguous
    float *obase_ = out + ((size_t)b * r3 + (size_t)row * r) * cout + co0;
    if (nco == kGatherCo && (cout & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      // 8 lanes x 16 bytes cover a voxel's 32 channels; the 4 lane groups take 4 voxels per store instruction
      const int c4 = (lane & 7) * 4, zs = lane >> 3;
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias) bb = __ldg(reinterpret_cast<const float4 *>(bias + co0 + c4));
      for (int z = zs; z < r; z += 4) {
        float4 v = total != 0 ? make_float4(t[z][c4], t[z][c4 + 1], t[z][c4 + 2], t[z][c4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
        *reinterpret_cast<float4 *>(obase_ + (size_t)z * cout + c4) = v;
      }
    } else if (lane < nco) {
      const float bb = bias ? __ldg(bias + co0 + lane) : 0.0f;
      for (int z = 0; z < r; ++z) obase_[(size_t)z * cout + lane] = (total != 0 ? t[z][lane] : 0.0f) + bb;
    }
    return;
  }
  // transposed write-out, one 4r-byte row segment per output channel
  float *orow = out + ((size_t)b * cout + co0) * r3 + (size_t)row * r;
  if (r == 32 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    // 8 lanes x 16 bytes cover a row; the 4 lane groups take 4 channels per store instruction
    const int z4 = (lane & 7) * 4, cs = lane >> 3;
    for (int co = cs; co < nco; co += 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (total != 0) v = make_float4(t[z4][co], t[z4 + 1][co], t[z4 + 2][co], t[z4 + 3][co]);
      if (bias) {
        const float bb = __ldg(bias + co0 + co);
        v.x += bb; v.y += bb; v.z += bb; v.w += bb;
      }
      *reinterpret_cast<float4 *>(orow + (size_t)co * r3 + z4) = v;
    }
  } else if (lane < r) {
    float *o = orow + lane;
    for (int co = 0; co < nco; ++co) {
      const float v = total != 0 ? t[lane][co] : 0.0f;
      o[(size_t)co * r3] = bias ? v + __ldg(bias + co0 + co) : v;
    }
  }
}

}  // namespace bdm

// taps f32[b][n][27][cout] (row j = the j-th occupied voxel of shape b in ascending voxel id, as produced
// from bdm_avg_voxelize_compact; rows >= the shape's occupied count are ignored), bias f32[cout] or NULL,
// out f32[b][cout][r^3], or f32[b][r^3][cout] when channels_last != 0.  stats (optional, f64[b][blocks][cout][2],
// blocks = bdm_sparse_conv3_stats_blocks(r)): per-channel (sum, sum of squares) of the bias-less output per block of
// rows, which bdm_groupnorm_act_cl accepts in place of its own statistics pass.  workspace = the plan bdm_voxel_plan left for these (b, n, r).  r in {1,2,4,8,16,32}.
// Number of per-channel statistics blocks bdm_sparse_conv3_gather writes per shape when `stats` is given.
extern "C" int bdm_sparse_conv3_stats_blocks(int r) {
  return r >= 1 ? bdm::ceil_div(r * r, bdm::kGatherWarps) : 0;
}

extern "C" int bdm_sparse_conv3_gather(int b, int cout, int n, int r, const float *taps, const float *bias,
                                       float *out, int channels_last, double *stats, const void *workspace,
                                       size_t workspace_bytes, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && cout >= 0 && n >= 1 && r >= 1 && r <= 32 && (r & (r - 1)) == 0);  // rows within a word
  const int r3 = r * r * r;
  BDM_CHECK_SIZE(vox_fast_path(n, r3));
  if (b == 0 || cout == 0) return BDM_OK;
  BDM_CHECK_PTR(taps); BDM_CHECK_PTR(out);
  const VoxAuxLayout L = vox_aux_layout(n, r3);
  const int rc = check_workspace(L, b, workspace, workspace_bytes);
  if (rc != BDM_OK) return rc;
  dim3 grid(ceil_div(r * r, kGatherWarps), ceil_div(cout, kGatherCo), b);
  static const int batch = [] {   // tuning hook (tools/sparse_conv_bench.py): BDM_GATHER_BATCH=4|8
    const char *e = std::getenv("BDM_GATHER_BATCH");
    return (e != nullptr && e[0] == '4') ? 4 : 8;
  }();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (batch == 4)
    sparse_conv3_gather_kernel<4><<<grid, kGatherWarps * 32, 0, st>>>(
        n, cout, r, channels_last, taps, bias, out, reinterpret_cast<double2 *>(stats),
        static_cast<const unsigned char *>(workspace), L);
  else
    sparse_conv3_gather_kernel<8><<<grid, kGatherWarps * 32, 0, st>>>(
        n, cout, r, channels_last, taps, bias, out, reinterpret_cast<double2 *>(stats),
        static_cast<const unsigned char *>(workspace), L);
  BDM_RETURN_LAUNCH_STATUS();
}
