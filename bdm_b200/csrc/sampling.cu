// sampling.cu -- furthest point sampling + gather (forward/backward) for sm_100a.
//
// Replaces furthest_point_sampling_kernel, gather_features_kernel, gather_features_grad_kernel
// (experiments/model/pvcnn/modules/functional/src/sampling/sampling.cu:86-167, :17-31, :52-66).
//
// FPS is M-1 strictly dependent rounds per shape; it is latency-bound, not bandwidth-bound.  The
// reference spends each round on a global-memory read-modify-write of the running distances and a
// 9-level shared-memory tree with 10 __syncthreads.  Here one CTA per shape keeps coordinates AND
// running distances in registers, and a round is: distance update (3 FADD, FMUL, 2 FFMA, 2 FMNMX per
// point), two 32-bit warp REDUX ops (max over the distance bit pattern -- non-negative floats order
// like unsigned ints -- then min over a tie key), one shared-memory exchange with ONE __syncthreads
// (slots double-buffered by round parity), and two more REDUX ops.
//
// Bit-exact tie rule of the reference (sampling.cu:141-160, block size hard-wired to 512 at :171):
// thread t keeps the first strict maximum over k = t, t+512, ...; the tree keeps the lower slot on
// equal distances.  Net effect: among points with the maximal running distance the winner minimises
// (k mod 512, k div 512) lexicographically.  tie_key(k) = (k & 511) << 22 | (k >> 9) encodes that
// order in one unsigned int (k < 2^31 / ... fine for k < 2^22*512).
#include <cstdlib>

#include "common.cuh"

namespace bdm {

__device__ __forceinline__ unsigned fps_tie_key(int k) { return ((unsigned)(k & 511) << 22) | ((unsigned)k >> 9); }
__device__ __forceinline__ int fps_tie_key_decode(unsigned t) { return (int)(((t & 0x3fffffu) << 9) | (t >> 22)); }

template <int N>
__device__ __forceinline__ float tree_max(const float (&a)[N]) {
  float t[N];
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = a[i];
#pragma unroll
  for (int w = N / 2; w >= 1; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; ++i) t[i] = fmaxf(t[i], t[i + w]);
  return t[0];
}

template <int N>
__device__ __forceinline__ unsigned tree_min(const unsigned (&a)[N]) {
  unsigned t[N];
#pragma unroll
  for (int i = 0; i < N; ++i) t[i] = a[i];
#pragma unroll
  for (int w = N / 2; w >= 1; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; ++i) t[i] = min(t[i], t[i + w]);
  return t[0];
}

// Register-resident FPS: PPT points per thread, T <= MAXT threads, n <= PPT*T.  A round is
// issue-bound on one SM (every SMSP retires ~N*10/128 instructions) plus a fixed reduction/barrier
// latency per warp, so few fat threads (PPT=16, 8 warps for N=4096) beat many thin ones.
template <int PPT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
fps_register_kernel(int n, int m, const float *__restrict__ coords, int *__restrict__ indices) {
  const int b = blockIdx.x;
  const int T = blockDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  coords += (size_t)b * 3 * n;
  indices += (size_t)b * m;

  extern __shared__ float sco[];  // [3][n] coordinate planes, for the winner look-up
  __shared__ unsigned long long slot[2][32];

  float px[PPT], py[PPT], pz[PPT], dist[PPT];
  unsigned tk[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int k = tid + j * T;
    const bool ok = k < n;
    px[j] = ok ? coords[k] : 0.0f;
    py[j] = ok ? coords[k + n] : 0.0f;
    pz[j] = ok ? coords[k + n + n] : 0.0f;
    if (ok) { sco[k] = px[j]; sco[k + n] = py[j]; sco[k + n + n] = pz[j]; }
    // running distance starts at 1e38 (sampling.cpp:53-54); slots without a point hold -1, which
    // fminf keeps at -1 forever, so they never reach the (non-negative) maximum
    dist[j] = ok ? 1e38f : -1.0f;
    tk[j] = ok ? fps_tie_key(k) : 0xffffffffu;
  }
  if (tid == 0) indices[0] = 0;
  __syncthreads();

  int old = 0;
  for (int s = 1; s < m; ++s) {
    const float x1 = sco[old], y1 = sco[old + n], z1 = sco[old + n + n];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const float d = sqdist_ref(__fsub_rn(px[j], x1), __fsub_rn(py[j], y1), __fsub_rn(pz[j], z1));
      dist[j] = fminf(d, dist[j]);
    }
    // per-thread reductions as trees (log depth): the round is a dependent chain end to end
    const float best = fmaxf(tree_max<PPT>(dist), 0.0f);  // slots without a point hold -1
    const unsigned wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(best));
    unsigned cand[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) cand[j] = (__float_as_uint(dist[j]) == wmax) ? tk[j] : 0xffffffffu;
    const unsigned wkey = __reduce_min_sync(0xffffffffu, tree_min<PPT>(cand));
    if (lane == 0) slot[s & 1][warp] = ((unsigned long long)wmax << 32) | (unsigned long long)(~wkey);
    __syncthreads();
    const unsigned long long v = (lane < nwarps) ? slot[s & 1][lane] : 0ull;
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    const unsigned gmax = __reduce_max_sync(0xffffffffu, hi);
    const unsigned gkey = __reduce_max_sync(0xffffffffu, (hi == gmax && lane < nwarps) ? lo : 0u);
    old = fps_tie_key_decode(~gkey);
    if (tid == 0) indices[s] = old;
  }
}

// Large clouds (n > BDM_FPS_REGISTER_MAX_N): running distances in a global workspace, coordinates
// re-read from global (L2-resident); same reduction scheme.
__global__ void __launch_bounds__(1024, 1)
fps_global_kernel(int n, int m, const float *__restrict__ coords, float *__restrict__ distances,
                  int *__restrict__ indices) {
  const int b = blockIdx.x;
  const int T = blockDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  coords += (size_t)b * 3 * n;
  distances += (size_t)b * n;
  indices += (size_t)b * m;
  __shared__ unsigned long long slot[2][32];
  for (int k = tid; k < n; k += T) distances[k] = 1e38f;
  if (tid == 0) indices[0] = 0;
  int old = 0;
  for (int s = 1; s < m; ++s) {
    const float x1 = coords[old], y1 = coords[old + n], z1 = coords[old + n + n];
    float best = 0.0f;
    unsigned bestk = 0xffffffffu;
    for (int k = tid; k < n; k += T) {
      const float d = sqdist_ref(__fsub_rn(coords[k], x1), __fsub_rn(coords[k + n], y1), __fsub_rn(coords[k + n + n], z1));
      const float d2 = fminf(d, distances[k]);
      distances[k] = d2;
      const unsigned key = fps_tie_key(k);
      if (d2 > best || (d2 == best && key < bestk)) { best = d2; bestk = key; }
    }
    const unsigned wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(best));
    const unsigned wkey = __reduce_min_sync(0xffffffffu, (__float_as_uint(best) == wmax) ? bestk : 0xffffffffu);
    if (lane == 0) slot[s & 1][warp] = ((unsigned long long)wmax << 32) | (unsigned long long)(~wkey);
    __syncthreads();
    const unsigned long long v = (lane < nwarps) ? slot[s & 1][lane] : 0ull;
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    const unsigned gmax = __reduce_max_sync(0xffffffffu, hi);
    const unsigned gkey = __reduce_max_sync(0xffffffffu, (hi == gmax && lane < nwarps) ? lo : 0u);
    old = fps_tie_key_decode(~gkey);
    if (tid == 0) indices[s] = old;
  }
}

__global__ void fill_int_kernel(int *p, size_t count, int v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) p[i] = v;
}

// gather: out[b,c,j] = features[b,c,indices[b,j]]   (sampling.cu:28-30)
__global__ void gather_kernel(int c, int n, int m, const float *__restrict__ features,
                              const int *__restrict__ indices, float *__restrict__ out) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int src = indices[(size_t)b * m + j];
  for (int cc = blockIdx.y; cc < c; cc += gridDim.y)
    out[((size_t)b * c + cc) * m + j] = __ldg(features + ((size_t)b * c + cc) * n + src);
}

// gather backward: grad_x[b,c,indices[b,j]] += grad_y[b,c,j]   (sampling.cu:63-65)
__global__ void gather_grad_kernel(int c, int n, int m, const float *__restrict__ grad_y,
                                   const int *__restrict__ indices, float *__restrict__ grad_x) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int dst = indices[(size_t)b * m + j];
  for (int cc = blockIdx.y; cc < c; cc += gridDim.y)
    atomicAdd(grad_x + ((size_t)b * c + cc) * n + dst, grad_y[((size_t)b * c + cc) * m + j]);
}

template <int PPT, int MAXT>
static cudaError_t launch_fps_reg(int b, int n, int m, int threads, const float *coords, int *indices,
                                  cudaStream_t st) {
  const size_t smem = sizeof(float) * 3 * (size_t)n;
  auto kern = fps_register_kernel<PPT, MAXT>;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
  if (e != cudaSuccess) return e;
  kern<<<b, threads, smem, st>>>(n, m, coords, indices);
  return cudaGetLastError();
}

}  // namespace bdm

extern "C" size_t bdm_furthest_point_sampling_workspace_bytes(int b, int n) {
  if (b <= 0 || n <= BDM_FPS_REGISTER_MAX_N) return 16;
  return sizeof(float) * (size_t)b * n;
}

extern "C" int bdm_furthest_point_sampling(int b, int n, int m, const float *coords, int *indices,
                                           void *workspace, size_t workspace_bytes,
                                           bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && n >= 0);
  if (b == 0 || m <= 0) return BDM_OK;  // sampling.cu:90-91: m <= 0 leaves the (empty) output untouched
  BDM_CHECK_PTR(indices);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0) {  // no points: the reference's kernel would leave torch::zeros -> all indices 0
    const size_t count = (size_t)b * m;
    fill_int_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(indices, count, 0);
    BDM_RETURN_LAUNCH_STATUS();
  }
  BDM_CHECK_PTR(coords);
  cudaError_t e;
  if (n <= BDM_FPS_REGISTER_MAX_N) {
    if (n <= 1024) {         // small pyramid levels: 4 points per thread, <= 8 warps
      const int threads = ((ceil_div(n, 4) + 31) / 32) * 32;
      e = launch_fps_reg<4, 256>(b, n, m, threads, coords, indices, st);
    } else if (n <= 4096) {  // 16 points per thread, <= 8 warps (BDM_FPS_VARIANT: tuning hook)
      static const char *variant = getenv("BDM_FPS_VARIANT");
      if (variant != nullptr && variant[0] == '8') {
        e = launch_fps_reg<8, 512>(b, n, m, ((ceil_div(n, 8) + 31) / 32) * 32, coords, indices, st);
      } else if (variant != nullptr && variant[0] == '4') {
        e = launch_fps_reg<4, 1024>(b, n, m, ((ceil_div(n, 4) + 31) / 32) * 32, coords, indices, st);
      } else {
        e = launch_fps_reg<16, 256>(b, n, m, ((ceil_div(n, 16) + 31) / 32) * 32, coords, indices, st);
      }
    } else {                 // <= 8192 points: 8 per thread, 32 warps (64-register budget)
      e = launch_fps_reg<8, 1024>(b, n, m, 1024, coords, indices, st);
    }
  } else {
    if (workspace == nullptr) return BDM_ERR_NULL_POINTER;
    if (workspace_bytes < sizeof(float) * (size_t)b * n) return BDM_ERR_WORKSPACE_TOO_SMALL;
    fps_global_kernel<<<b, 1024, 0, st>>>(n, m, coords, static_cast<float *>(workspace), indices);
    e = cudaGetLastError();
  }
  return e == cudaSuccess ? BDM_OK : (int)e;
}

extern "C" int bdm_gather_features(int b, int c, int n, int m, const float *features,
                                   const int *indices, float *out, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && m >= 0 && b <= 65535);
  if (b == 0 || c == 0 || m == 0) return BDM_OK;
  BDM_CHECK_PTR(features); BDM_CHECK_PTR(indices); BDM_CHECK_PTR(out);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  gather_kernel<<<dim3(ceil_div(m, 128), min(c, 1024), b), 128, 0, st>>>(c, n, m, features, indices, out);
  BDM_RETURN_LAUNCH_STATUS();
}

extern "C" int bdm_gather_features_grad(int b, int c, int n, int m, const float *grad_y,
                                        const int *indices, float *grad_x, bdm_stream_t stream) {
  using namespace bdm;
  BDM_CHECK_SIZE(b >= 0 && c >= 0 && n >= 0 && m >= 0 && b <= 65535);
  if (b == 0 || c == 0 || n == 0) return BDM_OK;
  BDM_CHECK_PTR(grad_x);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)b * c * n, st);
  if (m > 0) {
    BDM_CHECK_PTR(grad_y); BDM_CHECK_PTR(indices);
    gather_grad_kernel<<<dim3(ceil_div(m, 128), min(c, 1024), b), 128, 0, st>>>(c, n, m, grad_y, indices,
                                                                               grad_x);
  }
  BDM_RETURN_LAUNCH_STATUS();
}
