// knn_eval.cu -- fp64 nearest-neighbour distances for Chamfer distance / F-score, sm_100a.
//
// Replaces the K=1 nearest-neighbour search inside pytorch3d.loss.chamfer_distance
// (called at experiments/evaluation/evaluation_cd.py:125; pytorch3d is un-vendored -> PARITY UNPINNED,
// restated from its published algorithm: squared L2, lowest index on ties) and the dense
// N x M fp64 distance matrix + row-min of compute_pc_to_pc_dist
// (experiments/evaluation/evaluation_f1.py:90-98; expansion form -2ab + |a|^2 + |b|^2, clamp 1e-12).
// The reference handles one pair at a time and materialises the 4096^2 fp64 matrix (134 MB); here a
// whole batch of pairs is one launch and nothing but the row minima leaves the SM.
//
// One thread per source point; target points are staged through shared memory in tiles and read as
// broadcasts.  Arithmetic is spelled with __dmul_rn/__dadd_rn (no fma) so that it matches the
// CPU oracle's plain C expressions bit for bit.
#include "common.cuh"

namespace bdm {

constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 512;

template <bool EXPANDED>
__global__ void __launch_bounds__(kKnnThreads)
nn_f64_kernel(int n, int m, const double *__restrict__ src, const double *__restrict__ tgt,
              double *__restrict__ dist, int *__restrict__ idx) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * kKnnThreads + threadIdx.x;
  src += (size_t)b * n * 3;
  tgt += (size_t)b * m * 3;
  __shared__ double st[kKnnTile * 3];
  __shared__ double stt[kKnnTile];
  const bool valid = i < n;
  const double s0 = valid ? src[(size_t)i * 3] : 0.0, s1 = valid ? src[(size_t)i * 3 + 1] : 0.0,
               s2 = valid ? src[(size_t)i * 3 + 2] : 0.0;
  const double ss = __dadd_rn(__dadd_rn(__dmul_rn(s0, s0), __dmul_rn(s1, s1)), __dmul_rn(s2, s2));
  double best = __longlong_as_double(0x7ff0000000000000ll);
  int besti = 0;
  for (int t0 = 0; t0 < m; t0 += kKnnTile) {
    const int tn = min(kKnnTile, m - t0);
    __syncthreads();
    for (int q = threadIdx.x; q < tn * 3; q += kKnnThreads) st[q] = tgt[(size_t)t0 * 3 + q];
    __syncthreads();
    if (EXPANDED) {
      for (int q = threadIdx.x; q < tn; q += kKnnThreads) {
        const double a = st[q * 3], c = st[q * 3 + 1], e = st[q * 3 + 2];
        stt[q] = __dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(c, c)), __dmul_rn(e, e));
      }
      __syncthreads();
    }
#pragma unroll 4
    for (int j = 0; j < tn; ++j) {
      const double q0 = st[j * 3], q1 = st[j * 3 + 1], q2 = st[j * 3 + 2];
      double d;
      if (EXPANDED) {
        const double ab = __dadd_rn(__dadd_rn(__dmul_rn(s0, q0), __dmul_rn(s1, q1)), __dmul_rn(s2, q2));
        d = __dmul_rn(-2.0, ab);
        d = __dadd_rn(d, ss);
        d = __dadd_rn(d, stt[j]);
        d = d < 1e-12 ? 1e-12 : d;
      } else {
        const double dx = __dsub_rn(s0, q0), dy = __dsub_rn(s1, q1), dz = __dsub_rn(s2, q2);
        d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      }
      if (d < best) { best = d; besti = t0 + j; }
    }
  }
  if (valid) {
    dist[(size_t)b * n + i] = best;
    if (idx != nullptr) idx[(size_t)b * n + i] = besti;
  }
}


// The same search with the metric's reduction fused in: per CTA of 256 threads x 2 source points the sum of the
// nearest squared distances (Chamfer: evaluation_cd.py:125 takes their mean) and the number of them below `thr`
// (F-score: evaluation_f1.py:104-106), written to fixed slots [b][blockIdx.x] -- the caller adds up ceil(n/512)
// numbers per pair instead of reading n distances back.  Reduction order is fixed (shuffle tree, then warps in order).
// (256 threads x 2 source points per CTA when that still fills the GPU, else 128 x 1)
template <bool EXPANDED, int kRedThreads, int kRedPerThread>
__global__ void __launch_bounds__(kRedThreads)
nn_f64_reduce_kernel(int n, int m, double thr, const double *__restrict__ src, const double *__restrict__ tgt,
                     double *__restrict__ part_sum, int *__restrict__ part_cnt) {
  const int b = blockIdx.y;
  src += (size_t)b * n * 3;
  tgt += (size_t)b * m * 3;
  __shared__ double st[kKnnTile * 3];
  __shared__ double stt[kKnnTile];
  __shared__ double w_sum[kRedThreads / 32];
  __shared__ int w_cnt[kRedThreads / 32];
  double s[kRedPerThread][3], ss[kRedPerThread], best[kRedPerThread];
  bool valid[kRedPerThread];
#pragma unroll
  for (int u = 0; u < kRedPerThread; ++u) {
    const int i = (blockIdx.x * kRedPerThread + u) * kRedThreads + threadIdx.x;
    valid[u] = i < n;
#pragma unroll
    for (int a = 0; a < 3; ++a) s[u][a] = valid[u] ? src[(size_t)i * 3 + a] : 0.0;
    ss[u] = __dadd_rn(__dadd_rn(__dmul_rn(s[u][0], s[u][0]), __dmul_rn(s[u][1], s[u][1])), __dmul_rn(s[u][2], s[u][2]));
    best[u] = __longlong_as_double(0x7ff0000000000000ll);
  }
  for (int t0 = 0; t0 < m; t0 += kKnnTile) {
    const int tn = min(kKnnTile, m - t0);
    __syncthreads();
    for (int q = threadIdx.x; q < tn * 3; q += kRedThreads) st[q] = tgt[(size_t)t0 * 3 + q];
    __syncthreads();
    if (EXPANDED) {
      for (int q = threadIdx.x; q < tn; q += kRedThreads) {
        const double a = st[q * 3], c = st[q * 3 + 1], e = st[q * 3 + 2];
        stt[q] = __dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(c, c)), __dmul_rn(e, e));
      }
      __syncthreads();
    }
#pragma unroll 4
    for (int j = 0; j < tn; ++j) {
      const double q0 = st[j * 3], q1 = st[j * 3 + 1], q2 = st[j * 3 + 2];
#pragma unroll
      for (int u = 0; u < kRedPerThread; ++u) {
        double d;
        if (EXPANDED) {
          const double ab = __dadd_rn(__dadd_rn(__dmul_rn(s[u][0], q0), __dmul_rn(s[u][1], q1)), __dmul_rn(s[u][2], q2));
          d = __dmul_rn(-2.0, ab);
          d = __dadd_rn(d, ss[u]);
          d = __dadd_rn(d, stt[j]);
          d = d < 1e-12 ? 1e-12 : d;
        } else {
          const double dx = __dsub_rn(s[u][0], q0), dy = __dsub_rn(s[u][1], q1), dz = __dsub_rn(s[u][2], q2);
          d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        }
        best[u] = d < best[u] ? d : best[u];
      }
    }
  }
  double sum = 0.0;
  int cnt = 0;
#pragma unroll
  for (int u = 0; u < kRedPerThread; ++u)
    if (valid[u]) { sum += best[u]; cnt += best[u] < thr ? 1 : 0; }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, d);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  }
  if ((threadIdx.x & 31) == 0) { w_sum[threadIdx.x >> 5] = sum; w_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    int c = 0;
#pragma unroll
    for (int w = 0; w < kRedThreads / 32; ++w) { a += w_sum[w]; c += w_cnt[w]; }
    part_sum[(size_t)b * gridDim.x + blockIdx.x] = a;
    part_cnt[(size_t)b * gridDim.x + blockIdx.x] = c;
  }
}

}  // namespace bdm

extern "C" int bdm_nn_f64(int b, int n, int m, int expanded, const double *src, const double *tgt,
                          double *dist, int *idx, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && n >= 0 && m >= 0 && b <= 65535);
  if (b == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(src); BDM_CHECK_PTR(dist);
  if (m > 0) BDM_CHECK_PTR(tgt);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (expanded)
    nn_f64_kernel<true><<<dim3(ceil_div(n, kKnnThreads), b), kKnnThreads, 0, st>>>(n, m, src, tgt, dist, idx);
  else
    nn_f64_kernel<false><<<dim3(ceil_div(n, kKnnThreads), b), kKnnThreads, 0, st>>>(n, m, src, tgt, dist, idx);
  BDM_RETURN_LAUNCH_STATUS();
}

namespace bdm {
// source points per CTA: 512 (256 threads x 2) when b pairs still give every SM a CTA that way, else 128
static inline int nn_reduce_points_per_cta(int b, int n) {
  return (long long)b * ceil_div(n, 512) >= sm_count() ? 512 : 128;
}
}  // namespace bdm

// Number of partial (sum, count) slots per pair bdm_nn_f64_reduce writes for b pairs of n source points.
extern "C" int bdm_nn_f64_reduce_blocks(int b, int n) {
  return n > 0 && b > 0 ? bdm::ceil_div(n, bdm::nn_reduce_points_per_cta(b, n)) : 0;
}

// Nearest-neighbour search + the metric's reduction: for each pair, partial sums of min_j |s_i - t_j|^2 over blocks of
// source points (part_sum f64[b][blocks]) and the number of those minima below thr (part_cnt i32[b][blocks]);
// blocks = bdm_nn_f64_reduce_blocks(b, n).  Chamfer term = sum(part_sum) / n; F-score term = sum(part_cnt) / n.
extern "C" int bdm_nn_f64_reduce(int b, int n, int m, int expanded, double thr, const double *src, const double *tgt,
                                 double *part_sum, int *part_cnt, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && n >= 0 && m >= 0 && b <= 65535);
  if (b == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(src); BDM_CHECK_PTR(part_sum); BDM_CHECK_PTR(part_cnt);
  if (m > 0) BDM_CHECK_PTR(tgt);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid(bdm_nn_f64_reduce_blocks(b, n), b);
  if (nn_reduce_points_per_cta(b, n) == 512) {
    if (expanded) nn_f64_reduce_kernel<true, 256, 2><<<grid, 256, 0, st>>>(n, m, thr, src, tgt, part_sum, part_cnt);
    else nn_f64_reduce_kernel<false, 256, 2><<<grid, 256, 0, st>>>(n, m, thr, src, tgt, part_sum, part_cnt);
  } else {
    if (expanded) nn_f64_reduce_kernel<true, 128, 1><<<grid, 128, 0, st>>>(n, m, thr, src, tgt, part_sum, part_cnt);
    else nn_f64_reduce_kernel<false, 128, 1><<<grid, 128, 0, st>>>(n, m, thr, src, tgt, part_sum, part_cnt);
  }
  BDM_RETURN_LAUNCH_STATUS();
}
