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
