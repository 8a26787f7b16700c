// ball_query.cu -- ball query for sm_100a.
//
// Replaces ball_query_kernel
// (experiments/model/pvcnn/modules/functional/src/ball_query/ball_query.cu:19-50; one CTA per batch
// element, one thread per centre walking all N points from global memory).
//
// Semantics kept bit-exact: d2 = fma(dz,dz, fma(dx,dx, dy*dy)) with d = centre - point (:35-38 as
// contracted by nvcc); hit iff d2 < r2 (strict); the row holds the first <=U hits in ascending point
// index, padded with the first hit; all zeros when there is no hit (wrapper zero-fills, .cpp:20-22).
//
// Design: the op is compute-bound on B*M*N distance tests (algorithmic bytes are ~4 MB per call).
// One warp = 32 centres (one per lane) x one contiguous SPLIT of the point range; all lanes read the
// same point -> uniform 128-bit loads of 4 consecutive x / y / z values (L1 broadcast), 7 FP32
// instructions per test, no ballot / compaction in the inner loop.  A CTA is SPLITS warps working on
// the same 32 centres; hits go to per-(centre,split) lists in shared memory and are concatenated in
// split order (= ascending point index) when the row is written, as one coalesced 128-byte store
// per centre for U = 32.
#include "common.cuh"

namespace bdm {

constexpr int kBqMaxSplits = 8;

constexpr int kBqTile = 512;  // points staged per warp at a time (6 KB)

template <bool VEC4>
__global__ void __launch_bounds__(32 * kBqMaxSplits)
ball_query_kernel(int n, int m, float r2, int u, int splits, const float *__restrict__ centers,
                  const float *__restrict__ points, int *__restrict__ neighbors) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, split = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;  // centre handled by this lane
  centers += (size_t)b * 3 * m;
  points += (size_t)b * 3 * n;
  neighbors += (size_t)b * m * u;

  extern __shared__ __align__(16) int s_dyn[];    // [splits][32 centres][us] hit lists, us = u|1 (odd stride:
                                                  // simultaneous appends of different lanes hit different
                                                  // banks), then [splits][3][kBqTile] staged points
  __shared__ int s_cnt[kBqMaxSplits][32];
  const int us = u | 1;
  int *s_hits = s_dyn;
  int *my = s_hits + ((size_t)split * 32 + lane) * us;
  float *tile = reinterpret_cast<float *>(s_dyn + (((size_t)splits * 32 * us + 3) & ~(size_t)3)) + (size_t)split * 3 * kBqTile;

  const bool valid = j < m;
  const float cx = valid ? centers[j] : 0.0f;
  const float cy = valid ? centers[j + m] : 0.0f;
  const float cz = valid ? centers[j + m + m] : 0.0f;

  // contiguous point range of this split, multiple of 4 so vector loads stay aligned
  int len = ceil_div(n, splits);
  len = (len + 3) & ~3;
  const int k0 = min(split * len, n), k1 = min(k0 + len, n);

  // A lane that has no centre, or whose list is full, gets a negative threshold: no distance is
  // below it, so the hot loop needs no `cnt < u` test.
  int cnt = 0;
  float thr = valid ? r2 : -1.0f;
  for (int t0 = k0; t0 < k1; t0 += kBqTile) {
    if (__all_sync(0xffffffffu, thr < 0.0f)) break;
    const int tn = min(kBqTile, k1 - t0);
    __syncwarp();
    warp_stage_xyz<kBqTile>(points + t0, (size_t)n, tn, tile, lane, VEC4 && (tn & 3) == 0);
    int q = 0;
    for (; q + 4 <= tn; q += 4) {
      const float4 X = *reinterpret_cast<const float4 *>(tile + q);
      const float4 Y = *reinterpret_cast<const float4 *>(tile + kBqTile + q);
      const float4 Z = *reinterpret_cast<const float4 *>(tile + 2 * kBqTile + q);
      const float d0 = sqdist_ref(__fsub_rn(cx, X.x), __fsub_rn(cy, Y.x), __fsub_rn(cz, Z.x));
      const float d1 = sqdist_ref(__fsub_rn(cx, X.y), __fsub_rn(cy, Y.y), __fsub_rn(cz, Z.y));
      const float d2 = sqdist_ref(__fsub_rn(cx, X.z), __fsub_rn(cy, Y.z), __fsub_rn(cz, Z.z));
      const float d3 = sqdist_ref(__fsub_rn(cx, X.w), __fsub_rn(cy, Y.w), __fsub_rn(cz, Z.w));
      if (fminf(fminf(d0, d1), fminf(d2, d3)) < thr) {  // rare when balls hold a handful of points
        const int k = t0 + q;
        if (d0 < thr) { my[cnt++] = k;     if (cnt == u) thr = -1.0f; }
        if (d1 < thr) { my[cnt++] = k + 1; if (cnt == u) thr = -1.0f; }
        if (d2 < thr) { my[cnt++] = k + 2; if (cnt == u) thr = -1.0f; }
        if (d3 < thr) { my[cnt++] = k + 3; if (cnt == u) thr = -1.0f; }
      }
    }
    for (; q < tn; ++q) {
      const float d = sqdist_ref(__fsub_rn(cx, tile[q]), __fsub_rn(cy, tile[kBqTile + q]), __fsub_rn(cz, tile[2 * kBqTile + q]));
      if (d < thr) { my[cnt++] = t0 + q; if (cnt == u) thr = -1.0f; }
    }
  }
  s_cnt[split][lane] = valid ? cnt : 0;
  __syncthreads();

  // merge: warp `split` writes rows split, split+splits, ... ; lane = output slot (strided over u)
  for (int row = split; row < 32; row += splits) {
    const int jj = blockIdx.x * 32 + row;
    if (jj >= m) break;
    int first = 0;
    bool have_first = false;
    for (int s = 0; s < splits && !have_first; ++s)
      if (s_cnt[s][row] > 0) { first = s_hits[((size_t)s * 32 + row) * us]; have_first = true; }
    for (int slot = lane; slot < u; slot += 32) {
      int val = first, rem = slot;
      for (int s = 0; s < splits; ++s) {
        const int cs = s_cnt[s][row];
        if (rem < cs) { val = s_hits[((size_t)s * 32 + row) * us + rem]; break; }
        rem -= cs;
      }
      neighbors[(size_t)jj * u + slot] = val;
    }
  }
}

}  // namespace bdm

extern "C" int bdm_ball_query(int b, int n, int m, float r2, int u, const float *centers_coords,
                              const float *points_coords, int *neighbors_indices,
                              bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && n >= 0 && m >= 0 && u >= 0 && b <= 65535);
  if (b == 0 || m == 0 || u == 0) return BDM_OK;
  BDM_CHECK_PTR(centers_coords); BDM_CHECK_PTR(neighbors_indices);
  if (n > 0) BDM_CHECK_PTR(points_coords);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // enough warps to fill the machine (>= 8 warps per SM), at most kBqMaxSplits per CTA, and at
  // least 256 points per split so the per-row merge stays negligible
  const int ctas = ceil_div(m, 32) * b;
  int splits = 1;
  while (splits < kBqMaxSplits && ctas * splits < 16 * sm_count() && n / (splits * 2) >= 64) splits *= 2;
  auto smem_for = [&](int sp) {
    return sizeof(int) * ((((size_t)sp * 32 * (u | 1)) + 3) & ~(size_t)3) + sizeof(float) * (size_t)sp * 3 * kBqTile;
  };
  size_t smem = smem_for(splits);
  while (smem > 200 * 1024 && splits > 1) { splits /= 2; smem = smem_for(splits); }
  BDM_CHECK_SIZE(smem <= 200 * 1024);
  const bool vec4 = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(points_coords) & 15) == 0);
  auto kern = vec4 ? ball_query_kernel<true> : ball_query_kernel<false>;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<dim3(ceil_div(m, 32), b), 32 * splits, smem, st>>>(n, m, r2, u, splits, centers_coords, points_coords,
                                                            neighbors_indices);
  BDM_RETURN_LAUNCH_STATUS();
}
