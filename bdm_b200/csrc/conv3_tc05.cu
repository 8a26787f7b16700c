// conv3_tc05.cu -- the dense 3x3x3 / stride 1 / zero-padded convolution of a PVConv block (reference:
// experiments/model/pvcnn/modules/pvconv.py:75-88, the second nn.Conv3d of `voxel_layers`, run by cuDNN in TF32 under
// torch's default conv policy) as an implicit GEMM on Blackwell's 5th-generation tensor cores: tcgen05.mma with the
// accumulators in tensor memory, both operands fetched by 1-D TMA bulk copies, a persistent warp-specialised CTA per SM.
//
// Precision.  cuDNN's TF32 kernels keep 11 significant bits of every operand and accumulate in fp32.  Here the
// operands are fp16 -- also 11 significant bits -- after a power-of-two scaling that rules out overflow (weights:
// max|w| -> [512, 1024); activations: a bound derived from the GroupNorm affine that produced them), products are
// exact and accumulation is fp32 in TMEM.  tests/test_conv3_tc05_gpu.py checks the result against float64 and
// requires an error no larger than cuDNN's TF32 result has.
//
// The flat padded grid.  A voxel (x, y, z) of an R^3 grid sits at flat position p = (x*Q + y)*Q + z, Q = R + 1, of a
// (R+1)^3 volume whose extra plane / row / column hold zeros: the +1 neighbour of the last voxel of a row and the -1
// neighbour of the first voxel of the next row are the SAME zero position.  A tap (dx, dy, dz) is then the constant
// row offset dx*Q^2 + dy*Q + dz for every output position, so the A operand of one tap is simply a window of 128
// consecutive rows of the activation array, shifted.  Activations live in global memory as fp16 "chunk planes"
// xh[C/8][rows][8]: for each group of 8 channels a flat array of 16-byte rows.  A slab of consecutive rows of one plane
// is contiguous (one bulk copy) and lands in shared memory as the tensor core's canonical K-major no-swizzle
// layout: a core matrix = 8 consecutive rows x 16 bytes, SBO = 128, LBO = slab rows * 16; shifting the window by one
// row is +16 bytes on the descriptor's start address (verified on the device: tools/probe/tc05_probe_shift.cu).
// The price is (R+1)^2/R^2 - 1 padded output rows that are computed and dropped (6 % at R=32, 13 % at R=16).
//
// Rows of sample b: [guard G zeros][(R+1)^3 positions]...; sample stride S, see conv3_geometry().  Pad positions and
// guards are zero because the buffer is zero-filled when it is created and producers only ever write real voxels.
//
// Kernel.  A unit = 256 consecutive flat positions (two M=128 accumulators) x all N = Cout channels.
//   warp 0     one thread: A producer.  Per (dx, 64-channel chunk): the slab of 256 + 2Q + 2 rows that covers the 9
//              (dy, dz) windows, KC/8 bulk copies into a 3-stage ring.
//   warp 1     one thread: W producer.  Weights are pre-arranged (conv3_tc05_prep) as the exact shared-memory image of
//              each (dx, chunk, tap group) stage; one bulk copy per stage into a ring.
//   warp 2     issues the MMAs (2 tiles x taps x K/16 per stage) and commits stages back to the producers: the warp
//              runs the loops together (descriptor words in uniform registers), one elected lane issues.
//   warp 3     TMEM allocation.
//   warps 4-11 epilogue: thread = output row = TMEM lane; scale, + bias, GroupNorm statistics of the result
//              (per-group sum / sum of squares, reduced per unit, no atomics), channels-last fp32 rows of the real
//              voxels written with 256-bit stores.  Accumulators are double-buffered (2 x 2 x N TMEM columns), so the
//              epilogue of unit i overlaps the MMAs of unit i+1.
// For the FIRST convolution of a block the operand is a freshly voxelized cloud (conv3_fill_planes_kernel): mostly zero
// rows.  The writer leaves one occupancy bit per row; every role derives the same 9-bit mask of non-empty (dx, dy)
// windows per unit and the loads and MMAs of the empty ones are skipped (exact: those products are zeros).
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "voxel_plan.cuh"

namespace bdm {
namespace cv3 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t a = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory operand descriptor, no swizzle, K-major: lbo = bytes between core matrices adjacent along K,
// sbo = along M / N
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::f16, fp16 operands (both K-major), fp32 accumulate, M = 128
__host__ __device__ constexpr uint32_t instr_desc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
#define BDM_CV3_TMEM_LD32(r, taddr)                                                                                \
  asm volatile(                                                                                                    \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                    \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                    \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),            \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),      \
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),    \
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])     \
      : "r"(taddr))
// true in exactly one (always the same) lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float swish_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

constexpr int kUnitRows = 256;        // output positions per unit (2 accumulators of 128)
constexpr int kHeaderBytes = 256;     // prepared-weight header: f32 out_scale, act_scale, w_scale, 1/w_scale, u32 amax bits (dynamic scaling)
constexpr int kAStages = 3;
constexpr int kThreads = 384;         // 12 warps
constexpr int kEpiThreads = 256;

struct Geometry {
  int r, q, p;                // resolution, Q = r+1, P = Q^2
  int guard;                  // zero rows in front of every sample (>= P + Q + 1, multiple of 8)
  int units;                  // units per sample
  int units_pair;             // ... of the tap-pairing kernel instances (255 outputs per unit)
  long long sample_rows;      // row stride between samples
  long long total_rows;       // rows of one chunk plane
  int slab_rows;              // rows of an A stage: 256 + 2Q + 2
};
__host__ __device__ inline Geometry conv3_geometry(int b, int r) {
  Geometry g;
  g.r = r; g.q = r + 1; g.p = g.q * g.q;
  g.guard = (g.p + g.q + 1 + 7) / 8 * 8;
  const int last_valid = ((r - 1) * g.q + (r - 1)) * g.q + (r - 1);
  g.units = (last_valid + 1 + kUnitRows - 1) / kUnitRows;
  g.units_pair = (last_valid + 1 + kUnitRows - 2) / (kUnitRows - 1);
  g.sample_rows = ((long long)g.guard + (long long)g.q * g.p + 7) / 8 * 8;
  g.slab_rows = kUnitRows + 2 * g.q + 2;
  // the last unit of the last sample reads up to units*256 + P + Q + 1 rows past its sample's first position
  g.total_rows = (long long)g.guard + (long long)b * g.sample_rows + (long long)(last_valid + 1 + 2 * kUnitRows) + g.guard + 8;
  return g;
}

// -------------------------------------------------------------------------------------------------------------
// weights: f32[cout][cin][27] -> header + fp16 stage images.  Stage s = ((dx*NKC + kc)*NG + grp); inside a stage:
// [tap j of the group][chunk c of KC/8][row n of cout][8 halves] = w[n][kc*KC + 8c + e][dx*9 + grp*TG + j] * wscale
// -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pow2_scale_to(float amax, float target_exp, float *inv) {   // amax * s in [2^t, 2^(t+1))
  int e = (int)target_exp;
  if (amax > 0.0f && amax < INFINITY) e = (int)((__float_as_uint(amax) >> 23) & 255u) - 127;
  const int se = min(max((int)target_exp - e, -60), 60);
  *inv = __uint_as_float((uint32_t)(127 - se) << 23);
  return __uint_as_float((uint32_t)(127 + se) << 23);
}

// one CTA: max|w|, and the activation scale from the GroupNorm affine that produces the conv's input:
// |swish(gn(x))| <= max|gamma| * sqrt(group elements) + max|beta|  (sum of squared normalised values = count).
__global__ void __launch_bounds__(1024)
conv3_prep_header_kernel(size_t nw, const float *__restrict__ w, int c_in, const float *__restrict__ gamma,
                         const float *__restrict__ beta, float group_elems, float *__restrict__ header) {
  __shared__ float red[3][32];
  float mw = 0.0f, mg = 0.0f, mb = 0.0f;
  for (size_t i = threadIdx.x; i < nw; i += blockDim.x) mw = fmaxf(mw, fabsf(__ldg(w + i)));
  for (int i = threadIdx.x; i < c_in; i += blockDim.x) {
    mg = fmaxf(mg, gamma != nullptr ? fabsf(gamma[i]) : 1.0f);
    mb = fmaxf(mb, beta != nullptr ? fabsf(beta[i]) : 0.0f);
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, d));
    mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, d));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, d));
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = mw; red[1][threadIdx.x >> 5] = mg; red[2][threadIdx.x >> 5] = mb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
      mw = fmaxf(mw, red[0][i]); mg = fmaxf(mg, red[1][i]); mb = fmaxf(mb, red[2][i]);
    }
    float inv_w, inv_a = 1.0f, sa = 1.0f;
    const float sw = pow2_scale_to(mw, 9.0f, &inv_w);
    const float bound = mg * sqrtf(group_elems) + mb;
    if (bound > 32768.0f && bound < INFINITY) {          // scale down by a power of two so that bound * sa <= 32768
      const int e = (int)((__float_as_uint(bound) >> 23) & 255u) - 127;   // bound in [2^e, 2^(e+1))
      const int se = min(e + 1 - 15, 60);
      sa = __uint_as_float((uint32_t)(127 - se) << 23);
      inv_a = __uint_as_float((uint32_t)(127 + se) << 23);
    }
    header[0] = inv_w * inv_a;   // out_scale: accumulator -> convolution result
    header[1] = sa;              // act_scale: applied by the producer of xh
    header[2] = sw;
    header[3] = inv_w;
  }
}

__global__ void __launch_bounds__(256)
conv3_prep_weights_kernel(int c_in, int c_out, int kc_size, int tg, int pair, const float *__restrict__ w,
                          const float *__restrict__ header, __half *__restrict__ out) {
  const float sw = header[2];
  const int nkc = c_in / kc_size, ng = 9 / tg, chunks = kc_size / 8;
  const size_t total = (size_t)27 * c_in * c_out;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t rest = i;
    int n, c, j, grp, kc, dx;
    const int e = (int)(rest % 8); rest /= 8;
    if (pair) {
      // stage (dx, kc, dy): [chunk][2 c_out rows: dz = 0 then dz = +1][8 halves], then [chunk][c_out rows: dz = -1][8 halves]
      const int stage_elems = 3 * c_out * chunks;            // 16-byte rows per stage
      const int in_stage = (int)(rest % stage_elems); rest /= stage_elems;
      if (in_stage < 2 * c_out * chunks) {
        c = in_stage / (2 * c_out);
        const int row = in_stage - c * 2 * c_out;
        j = row < c_out ? 1 : 2;
        n = row < c_out ? row : row - c_out;
      } else {
        const int q = in_stage - 2 * c_out * chunks;
        c = q / c_out; n = q - c * c_out; j = 0;
      }
      grp = (int)(rest % 3); rest /= 3;
      kc = (int)(rest % nkc); rest /= nkc;
      dx = (int)rest;
      tg = 3;
    } else {
      n = (int)(rest % c_out); rest /= c_out;
      c = (int)(rest % chunks); rest /= chunks;
      j = (int)(rest % tg); rest /= tg;
      grp = (int)(rest % ng); rest /= ng;
      kc = (int)(rest % nkc); rest /= nkc;
      dx = (int)rest;
    }
    const int ci = kc * kc_size + c * 8 + e;
    const int tap = dx * 9 + grp * tg + j;
    out[i] = __float2half_rn(__ldg(w + ((size_t)n * c_in + ci) * 27 + tap) * sw);
  }
}

// -------------------------------------------------------------------------------------------------------------
// GroupNorm(+conv bias)+Swish of a channels-last grid x f32[b][r^3][c] with producer-made statistics, written as
// the fp16 chunk planes the convolution reads.  Thread = 8 channels of one voxel (two 128-bit loads, one 128-bit
// store); a warp reads 1 KB of consecutive voxels.
// -------------------------------------------------------------------------------------------------------------
constexpr int kApplyThreads = 256;
// group-level producer statistics f64[b][blocks][groups][2] -> per-channel table (the group's sums in its first channel's
// slot, zeros in the others); see groupnorm.cu
__device__ __forceinline__ void fold_group_partials(const double2 *__restrict__ gp, int b, int blocks, int groups, int c, int cg,
                                                    double2 *slice /* [256] */, double2 *chan /* [c] */) {
  const int t = threadIdx.x;
  const int nsl = 256 / groups, g = t % groups, sl = t / groups;
  double S1 = 0.0, S2 = 0.0;
  if (sl < nsl) {
    for (int blk = sl; blk < blocks; blk += nsl) {
      const double2 v = gp[((size_t)b * blocks + blk) * groups + g];
      S1 += v.x; S2 += v.y;
    }
  }
  slice[t] = make_double2(S1, S2);
  __syncthreads();
  if (t < c) {
    double A1 = 0.0, A2 = 0.0;
    if (t % cg == 0) {
      for (int k = 0; k < nsl; ++k) { A1 += slice[k * groups + t / cg].x; A2 += slice[k * groups + t / cg].y; }
    }
    chan[t] = make_double2(A1, A2);
  }
  __syncthreads();
}
__global__ void __launch_bounds__(kApplyThreads)
gn_apply_half_planar_kernel(int c, int r, int groups, int nchunks, int ntiles, float eps, int swish,
                            const float *__restrict__ x, const float *__restrict__ conv_bias,
                            const float *__restrict__ gamma, const float *__restrict__ beta,
                            const double2 *__restrict__ partials, const float *__restrict__ header,
                            __half *__restrict__ xh, int guard, long long sample_rows, long long total_rows) {
  __shared__ double2 slice[kApplyThreads];
  __shared__ double2 chan[256];
  __shared__ double2 grp[32];
  __shared__ float2 ab[256];
  // back to front: the convolution that produced x wrote it front to back, so the end of x is what L2 still holds -- and
  // the convolution that reads xh next starts at the front, which this kernel then writes last
  const int b = gridDim.y - 1 - blockIdx.y, tile = gridDim.x - 1 - blockIdx.x, t = threadIdx.x;
  const int cg = c / groups;
  const long long s = (long long)r * r * r;
  if (nchunks < 0) {
    fold_group_partials(partials, b, -nchunks, groups, c, cg, slice, chan);     // group-level statistics (bias included)
  } else {
  {
    const int nsl = kApplyThreads / c, tc = t % c, sl = t / c;
    double S1 = 0.0, S2 = 0.0;
    for (int ch = sl; ch < nchunks; ch += nsl) {
      const double2 v = partials[((size_t)b * nchunks + ch) * c + tc];
      S1 += v.x; S2 += v.y;
    }
    slice[t] = make_double2(S1, S2);
  }
  __syncthreads();
  if (t < c) {
    double S1 = 0.0, S2 = 0.0;
    for (int sl = 0; sl < kApplyThreads / c; ++sl) { S1 += slice[sl * c + t].x; S2 += slice[sl * c + t].y; }
    const double tt = conv_bias != nullptr ? (double)conv_bias[t] : 0.0;   // statistics are of the bias-less tensor
    const double ds = (double)s;
    chan[t] = make_double2(S1 + ds * tt, S2 + 2.0 * tt * S1 + ds * tt * tt);
  }
  __syncthreads();
  }
  if (t < groups) {
    double S1 = 0.0, S2 = 0.0;
    for (int j = 0; j < cg; ++j) { S1 += chan[t * cg + j].x; S2 += chan[t * cg + j].y; }
    const double n = (double)cg * (double)s;
    const double mean = S1 / n;
    const double var = fmax(S2 / n - mean * mean, 0.0);
    grp[t] = make_double2(mean, 1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if (t < c) {
    const double2 g = grp[t / cg];
    const float ga = gamma != nullptr ? gamma[t] : 1.0f;
    const float be = beta != nullptr ? beta[t] : 0.0f;
    const float cb = conv_bias != nullptr ? conv_bias[t] : 0.0f;
    const float A = (float)g.y * ga;
    ab[t] = make_float2(A, (float)((double)be + ((double)cb - g.x) * (double)A));
  }
  __syncthreads();
  const float act_scale = __ldg(header + 1);
  // A pass = rpp voxels x C channels = 256 float4.  Load side: thread = 4 channels of one voxel, so a warp reads 512
  // contiguous bytes of the fp32 rows; its 4 halves (8 bytes) go to shared memory as [chunk plane][voxel][8 halves].
  // Store side: thread = 16 bytes of (chunk plane, voxel), consecutive voxels in consecutive lanes: contiguous runs of
  // the fp16 planes.  (A first version loaded 8 channels per thread -- two half-used sectors per request -- and ran at
  // 4 TB/s against 6 TB/s for the fp32 apply kernel.)
  const int c4 = c >> 2, c8 = c >> 3, rpp = kApplyThreads / c4;
  const int qd = t % c4, r0 = t / c4;
  const int hp = kApplyThreads / 2;                // (plane, voxel) pairs per pass: rpp * c8
  const int jw = (t % hp) / rpp, vw = t % rpp;    // store side (threads t < hp)
  const float2 k0 = ab[4 * qd], k1 = ab[4 * qd + 1], k2 = ab[4 * qd + 2], k3 = ab[4 * qd + 3];
  const int q = r + 1;
  const int sh = 31 - __clz(r);                 // r is a power of two
  long long per = (s + ntiles - 1) / ntiles;
  per = (per + rpp - 1) / rpp * rpp;
  const long long lo = min((long long)tile * per, s), hi = min(lo + per, s);
  const float *px = x + (size_t)b * s * c + 4 * qd;
  __half *plane = xh + ((size_t)jw * total_rows + (size_t)guard + (size_t)b * sample_rows) * 8;
  constexpr int UNR = 8;
  __shared__ uint2 stage[UNR][2 * (kApplyThreads / 2 + 32)];   // per pass: [plane][rpp + 1] x 16 bytes, as 8-byte halves
  auto convert = [&](const float4 &u) {
    float a0 = fmaf(u.x, k0.x, k0.y), a1 = fmaf(u.y, k1.x, k1.y), a2 = fmaf(u.z, k2.x, k2.y), a3 = fmaf(u.w, k3.x, k3.y);
    if (swish) { a0 = swish_fast(a0); a1 = swish_fast(a1); a2 = swish_fast(a2); a3 = swish_fast(a3); }
    const __half2 h0 = __floats2half2_rn(a0 * act_scale, a1 * act_scale), h1 = __floats2half2_rn(a2 * act_scale, a3 * act_scale);
    return make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
  };
  auto dest = [&](long long row) {
    const int vz = (int)row & (r - 1), vy = ((int)row >> sh) & (r - 1), vx = (int)row >> (2 * sh);
    return plane + (size_t)(((long long)vx * q + vy) * q + vz) * 8;
  };
  const int st_slot = 2 * ((qd >> 1) * (rpp + 1) + r0) + (qd & 1);    // 8-byte slot of this thread's 4 channels
  const int ld_slot = 2 * (jw * (rpp + 1) + vw);                      // 16-byte slot read back by the store side
  const long long step = (long long)UNR * rpp;
  long long base = lo;                             // first voxel of the current group of UNR passes
  float4 u[UNR], un[UNR];
  auto load = [&](long long first, float4 (&a)[UNR]) {
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const long long row = first + (long long)k * rpp + r0;
      if (row < hi) a[k] = ld_stream_f4(px + (size_t)row * c);
    }
  };
  if (base < hi) load(base, u);
  while (base < hi) {                              // uniform over the CTA
    const long long nbase = base + step;
    if (nbase < hi) load(nbase, un);               // next group in flight while this one is converted and stored
#pragma unroll
    for (int k = 0; k < UNR; ++k)
      if (base + (long long)k * rpp + r0 < hi) stage[k][st_slot] = convert(u[k]);
    __syncthreads();
    if (t < hp) {
#pragma unroll
      for (int k = 0; k < UNR; ++k) {
        const long long row = base + (long long)k * rpp + vw;
        if (row < hi) *reinterpret_cast<uint4 *>(dest(row)) = *reinterpret_cast<const uint4 *>(&stage[k][ld_slot]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < UNR; ++k) u[k] = un[k];
    base = nbase;
  }
}

// Warp-level version for C = 32 / 64 / 128 (C4 = C/4 <= 32 float4 per voxel): a warp owns tiles of 32 consecutive voxels.
// Load: C4 instructions, each 512 contiguous bytes of the fp32 rows (lane = (voxel, quad) pair, the quad -- hence the
// coefficients -- fixed per lane); convert; 8-byte stores into the warp's private [chunk plane][33 voxels][16 bytes]
// slab; then C/8 instructions of lane = voxel: a 16-byte shared load and a 16-byte global store, 512 contiguous bytes of
// one fp16 plane per instruction.  No block-wide barrier after the prologue.
template <int C4>
__global__ void __launch_bounds__(kApplyThreads)
gn_apply_half_planar_warp_kernel(int r, int groups, int nchunks, int ntiles, float eps, int swish,
                                 const float *__restrict__ x, const float *__restrict__ conv_bias,
                                 const float *__restrict__ gamma, const float *__restrict__ beta,
                                 const double2 *__restrict__ partials, const float *__restrict__ header,
                                 __half *__restrict__ xh, int guard, long long sample_rows, long long total_rows) {
  constexpr int c = 4 * C4, C8 = C4 / 2, VPI = 32 / C4;      // channels, chunk planes, voxels per load instruction
  __shared__ double2 slice[kApplyThreads];
  __shared__ double2 chan[c];
  __shared__ double2 grp[32];
  __shared__ float2 ab[c];
  extern __shared__ __align__(16) unsigned char dyn[];       // [8 warps][C8][33][16 bytes]
  // back to front: the convolution that produced x wrote it front to back, so the end of x is what L2 still holds -- and
  // the convolution that reads xh next starts at the front, which this kernel then writes last
  const int b = gridDim.y - 1 - blockIdx.y, tile = gridDim.x - 1 - blockIdx.x, t = threadIdx.x;
  const int cg = c / groups;
  const long long s = (long long)r * r * r;
  if (nchunks < 0) {
    fold_group_partials(partials, b, -nchunks, groups, c, cg, slice, chan);     // group-level statistics (bias included)
  } else {
  {
    const int nsl = kApplyThreads / c, tc = t % c, sl = t / c;
    double S1 = 0.0, S2 = 0.0;
    for (int ch = sl; ch < nchunks; ch += nsl) {
      const double2 v = partials[((size_t)b * nchunks + ch) * c + tc];
      S1 += v.x; S2 += v.y;
    }
    slice[t] = make_double2(S1, S2);
  }
  __syncthreads();
  if (t < c) {
    double S1 = 0.0, S2 = 0.0;
    for (int sl = 0; sl < kApplyThreads / c; ++sl) { S1 += slice[sl * c + t].x; S2 += slice[sl * c + t].y; }
    const double tt = conv_bias != nullptr ? (double)conv_bias[t] : 0.0;
    const double ds = (double)s;
    chan[t] = make_double2(S1 + ds * tt, S2 + 2.0 * tt * S1 + ds * tt * tt);
  }
  __syncthreads();
  }
  if (t < groups) {
    double S1 = 0.0, S2 = 0.0;
    for (int j = 0; j < cg; ++j) { S1 += chan[t * cg + j].x; S2 += chan[t * cg + j].y; }
    const double n = (double)cg * (double)s;
    const double mean = S1 / n;
    const double var = fmax(S2 / n - mean * mean, 0.0);
    grp[t] = make_double2(mean, 1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if (t < c) {
    const double2 g = grp[t / cg];
    const float ga = gamma != nullptr ? gamma[t] : 1.0f;
    const float be = beta != nullptr ? beta[t] : 0.0f;
    const float cb = conv_bias != nullptr ? conv_bias[t] : 0.0f;
    const float A = (float)g.y * ga;
    ab[t] = make_float2(A, (float)((double)be + ((double)cb - g.x) * (double)A));
  }
  __syncthreads();
  const float act_scale = __ldg(header + 1);
  const int warp = t >> 5, lane = t & 31;
  const int qd = lane % C4, vsub = lane / C4;                  // this lane's quad of channels and voxel within a load
  const float2 k0 = ab[4 * qd], k1 = ab[4 * qd + 1], k2 = ab[4 * qd + 2], k3 = ab[4 * qd + 3];
  const int q = r + 1;
  const int sh = 31 - __clz(r);
  uint2 *slab = reinterpret_cast<uint2 *>(dyn) + (size_t)warp * C8 * 33 * 2;
  long long per = (s + ntiles - 1) / ntiles;
  per = (per + 31) / 32 * 32;
  const long long lo = min((long long)tile * per, s), hi = min(lo + per, s);
  const float *px = x + (size_t)b * s * c;
  __half *planes = xh + ((size_t)guard + (size_t)b * sample_rows) * 8;
  for (long long v0 = lo + 32LL * warp; v0 < hi; v0 += 32LL * (kApplyThreads / 32)) {
    constexpr int KB = C4 < 16 ? C4 : 16;                    // load instructions in flight per lane (64 registers)
#pragma unroll
    for (int kb = 0; kb < C4; kb += KB) {
      float4 u[KB];
#pragma unroll
      for (int i = 0; i < KB; ++i) {
        const long long v = v0 + (kb + i) * VPI + vsub;
        if (v < hi) u[i] = ld_stream_f4(px + (size_t)v * c + 4 * qd);
      }
#pragma unroll
      for (int i = 0; i < KB; ++i) {
        float a0 = fmaf(u[i].x, k0.x, k0.y), a1 = fmaf(u[i].y, k1.x, k1.y), a2 = fmaf(u[i].z, k2.x, k2.y), a3 = fmaf(u[i].w, k3.x, k3.y);
        if (swish) { a0 = swish_fast(a0); a1 = swish_fast(a1); a2 = swish_fast(a2); a3 = swish_fast(a3); }
        const __half2 h0 = __floats2half2_rn(a0 * act_scale, a1 * act_scale), h1 = __floats2half2_rn(a2 * act_scale, a3 * act_scale);
        slab[((qd >> 1) * 33 + (kb + i) * VPI + vsub) * 2 + (qd & 1)] =
            make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
      }
    }
    __syncwarp();
    const long long v = v0 + lane;
    if (v < hi) {
      const int vz = (int)v & (r - 1), vy = ((int)v >> sh) & (r - 1), vx = (int)v >> (2 * sh);
      __half *dst = planes + (size_t)(((long long)vx * q + vy) * q + vz) * 8;
#pragma unroll
      for (int j = 0; j < C8; ++j)
        *reinterpret_cast<uint4 *>(dst + (size_t)j * total_rows * 8) = *reinterpret_cast<const uint4 *>(&slab[(j * 33 + lane) * 2]);
    }
    __syncwarp();
  }
}

// -------------------------------------------------------------------------------------------------------------
// A freshly voxelized cloud as the convolution's operand (the FIRST Conv3d of a PVConv block, modules/pvconv.py:75-76,
// 91-97): per-occupied-voxel averages (bdm_avg_voxelize_compact, f32[b][c][n]) + the voxel plan's occupancy bitmask ->
// every real voxel row of the fp16 chunk planes (zeros for empty voxels; occupancy changes from call to call, so
// all of them are rewritten).  No normalisation precedes this convolution, so the activation scale is dynamic:
// max|average| is reduced first (conv3_amax_kernel, one atomicMax per warp), every thread derives the power-of-two
// scale from it, and one thread records it in the prepared header for the convolution's epilogue.
// -------------------------------------------------------------------------------------------------------------
__global__ void conv3_amax_kernel(size_t n4, const float4 *__restrict__ x, unsigned *__restrict__ amax_bits) {
  float m = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(amax_bits, __float_as_uint(m));
}

// one warp per 32-voxel occupancy word (lane = voxel), looping over the chunk planes: 512-byte stores
__global__ void __launch_bounds__(256)
conv3_fill_planes_kernel(int c, int n, int r, VoxAuxLayout L, const unsigned char *__restrict__ plan_ws,
                         const float *__restrict__ compact, float *__restrict__ header, __half *__restrict__ xh, int guard,
                         long long sample_rows, long long total_rows, uint32_t *__restrict__ occ_bits, int occ_words) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int word_idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const float amax = __uint_as_float(reinterpret_cast<const unsigned *>(header)[4]);
  float inv_a;
  const float sa = pow2_scale_to(amax, 13.0f, &inv_a);
  if (blockIdx.x == 0 && b == 0 && threadIdx.x == 0) {
    header[0] = header[3] * inv_a;
    header[1] = sa;
  }
  if (word_idx >= L.nw) return;
  const unsigned char *ws = plan_ws + (size_t)b * L.stride;
  const uint32_t word = __ldg(reinterpret_cast<const uint32_t *>(ws + L.bitmask) + word_idx);
  const int slot = (int)__ldg(reinterpret_cast<const uint16_t *>(ws + L.obase) + word_idx) + __popc(word & ((1u << lane) - 1u));
  const bool occ = (word >> lane) & 1u;
  const int v = word_idx * 32 + lane;
  if (v >= r * r * r) return;
  const int sh = 31 - __clz(r), q = r + 1;
  const int vz = v & (r - 1), vy = (v >> sh) & (r - 1), vx = v >> (2 * sh);
  const int p = (vx * q + vy) * q + vz;
  const size_t row = (size_t)guard + (size_t)b * sample_rows + (size_t)p;
  if (occ_bits != nullptr && occ)   // one bit per non-zero flat row (zeroed by the launcher): lets the convolution skip all-zero windows
    atomicOr(occ_bits + (size_t)b * occ_words + ((guard + p) >> 5), 1u << ((guard + p) & 31));
  const float *src = compact + (size_t)b * c * n + slot;
  const int c8 = c >> 3;
  for (int j = 0; j < c8; ++j) {
    uint4 h = make_uint4(0u, 0u, 0u, 0u);
    if (occ) {
      float a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = __ldg(src + (size_t)(8 * j + e) * n) * sa;
      const __half2 h0 = __floats2half2_rn(a[0], a[1]), h1 = __floats2half2_rn(a[2], a[3]);
      const __half2 h2 = __floats2half2_rn(a[4], a[5]), h3 = __floats2half2_rn(a[6], a[7]);
      h = make_uint4(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1),
                     *reinterpret_cast<const uint32_t *>(&h2), *reinterpret_cast<const uint32_t *>(&h3));
    }
    *reinterpret_cast<uint4 *>(xh + ((size_t)j * total_rows + row) * 8) = h;
  }
}

// -------------------------------------------------------------------------------------------------------------
// the convolution
// -------------------------------------------------------------------------------------------------------------
// Occupancy of the flat rows of one sample, one bit per row (row index includes the guard): set by
// conv3_fill_planes_kernel for the occupied voxels of a freshly voxelized cloud.  window_mask() -> bit dx*3+dy is set iff
// the 258-row window [unit row 0 + (dx-1)*P + (dy-1)*Q - 1, +258) that the three dz taps of (dx, dy) read holds any
// non-zero row.  Executed by a whole (converged) warp: the 9 windows x up to 10 words are spread over the lanes and
// combined with one REDUX.  occ == NULL: everything is treated as occupied.
constexpr int kOccSlackRows = 1024;     // rows past a sample's stride that a unit's windows may reach into
__host__ __device__ inline int occ_words_per_sample(const Geometry &g) { return (int)((g.sample_rows + kOccSlackRows + 31) / 32); }
__device__ __forceinline__ uint32_t window_mask(const uint32_t *__restrict__ occ, const Geometry &geo, int smp, int u, int lane,
                                                int uout = kUnitRows) {
  if (occ == nullptr) return 0x1FFu;
  const uint32_t *o = occ + (size_t)smp * occ_words_per_sample(geo);
  uint32_t m = 0;
#pragma unroll
  for (int rnd = 0; rnd < 3; ++rnd) {
    const int idx = rnd * 32 + lane;             // (window, word) pair
    if (idx < 90) {
      const int wdw = idx / 10, k = idx - wdw * 10;
      const int dx = wdw / 3, dy = wdw - dx * 3;
      const int first = geo.guard + u * uout + (dx - 1) * geo.p + (dy - 1) * geo.q - 1;           // >= 0: guard >= P + Q + 1
      const int last = first + kUnitRows + 1;                                                      // inclusive
      const int w = (first >> 5) + k;
      if (w <= (last >> 5)) {
        uint32_t bits = __ldg(o + w);
        if (w == (first >> 5)) bits &= 0xffffffffu << (first & 31);
        if (w == (last >> 5)) bits &= 0xffffffffu >> (31 - (last & 31));
        if (bits != 0u) m |= 1u << wdw;
      }
    }
  }
  return __reduce_or_sync(0xffffffffu, m);
}

// PAIR (N <= 64; a weight stage holds TG / 3 = 1 or 3 (dx, dy) rows of taps -- 3 for the 32-channel instance, whose
// 6 KB rows were too small a unit of work for the MMA warp's per-stage waits): the dz = 0 and dz = +1 taps of a (dx, dy) pair share ONE activation window -- B = [W(dz=0); W(dz=+1)]
// is a 2N-row operand and one MMA of N' = 2N fills two column blocks, `main` and `side` -- and the dz = -1 tap is a
// second MMA of N columns into `main` with the window shifted by one row, as before.  Per (dx, dy), tile and K-step that
// is 8 + 6 KB of shared-memory operands instead of 3 x 6 KB: the kernel is bound by exactly that traffic.  The side
// block belongs one row further down, out[m] = main[m] + side[m+1]: the epilogue takes it from the next lane (one
// shuffle per column) and, for a warp's last lane, from a row the next warp leaves in shared memory; a unit therefore
// yields rows 0..254 of its 256.
template <int N, int KC, int TG, bool PAIR = false>
struct Cfg {
  static_assert(!PAIR || ((TG == 3 || TG == 9) && N <= 64), "tap pairing: whole (dx, dy) rows of taps per stage, 8N <= 512 TMEM columns");
  static constexpr int kUOut = PAIR ? kUnitRows - 1 : kUnitRows;      // output rows per unit
  static constexpr int kDCols = PAIR ? 2 * N : N;                    // TMEM columns per tile
  static constexpr int kChunks = KC / 8;
  static constexpr int kK16 = KC / 16;
  static constexpr int kNG = 9 / TG;
  static constexpr int kTapBytes = N * KC * 2;
  static constexpr int kWStageBytes = TG * kTapBytes;
  static constexpr int kWStages = kWStageBytes <= 16 * 1024 ? 4 : 3;
  static constexpr int kTmemCols = 4 * kDCols;     // 2 buffers x 2 tiles x columns per tile
  static constexpr int kNumBars = 2 * kAStages + 2 * kWStages + 4;
};

template <int N, int KC, int TG, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
conv3_tc05_kernel(int b, int c_in, Geometry geo, int a_stage_bytes, const __half *__restrict__ xh,
                  const unsigned char *__restrict__ wprep, const float *__restrict__ bias,
                  float *__restrict__ out, double *__restrict__ unit_stats, const uint32_t *__restrict__ occ) {
  using C = Cfg<N, KC, TG, PAIR>;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkc = c_in / KC;
  const int units = PAIR ? geo.units_pair : geo.units;
  const int total_units = b * units;
  const int slab_rows = geo.slab_rows;

  unsigned char *a_smem = smem;
  unsigned char *w_smem = smem + kAStages * a_stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(w_smem + C::kWStages * C::kWStageBytes);
  uint64_t *a_full = bars, *a_empty = a_full + kAStages, *w_full = a_empty + kAStages, *w_empty = w_full + C::kWStages;
  uint64_t *t_full = w_empty + C::kWStages, *t_empty = t_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + C::kNumBars);
  float *red = reinterpret_cast<float *>(tmem_slot + 4);        // [2][8 warps][16]
  float *edge = red + 2 * 8 * 16;                               // PAIR: [8 warps][N], the side row of every warp's first lane
  volatile uint32_t *unit_empty = tmem_slot + 2;                // [2]: no MMA was issued for the unit in this accumulator buffer

  if (threadIdx.x == 0) {
    for (int s = 0; s < kAStages; ++s) { bar_init(a_full + s, 1); bar_init(a_empty + s, 1); }
    for (int s = 0; s < C::kWStages; ++s) { bar_init(w_full + s, 1); bar_init(w_empty + s, 1); }
    for (int s = 0; s < 2; ++s) { bar_init(t_full + s, 1); bar_init(t_empty + s, kEpiThreads); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== A producer =====================
    // (the warp runs the unit loop together for window_mask(); lane 0 does the waiting and the copies)
    {
      int sa = 0; uint32_t pa = 0;
      const uint32_t slab_bytes = (uint32_t)slab_rows * 16u;
      for (int g = blockIdx.x; g < total_units; g += gridDim.x) {
        const int smp = g / units, u = g - smp * units;
        const uint32_t wm = window_mask(occ, geo, smp, u, lane, C::kUOut);
        const long long row_base = (long long)geo.guard + (long long)smp * geo.sample_rows + (long long)u * C::kUOut;
        for (int dx = 0; dx < 3; ++dx) {
          if (((wm >> (3 * dx)) & 7u) == 0u) continue;          // an all-zero slab: nothing to multiply
          const long long src_row = row_base + (long long)(dx - 1) * geo.p - geo.q - 1;
          for (int kc = 0; kc < nkc; ++kc) {
            if (lane == 0) {
              bar_wait(a_empty + sa, pa ^ 1);
              bar_expect_tx(a_full + sa, C::kChunks * slab_bytes);
              const uint32_t dst = smem_u32(a_smem + sa * a_stage_bytes);
#pragma unroll
              for (int j = 0; j < C::kChunks; ++j)
                tma_load(dst + j * slab_bytes, xh + ((size_t)(kc * C::kChunks + j) * geo.total_rows + (size_t)src_row) * 8,
                         slab_bytes, a_full + sa);
            }
            if (++sa == kAStages) { sa = 0; pa ^= 1; }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== W producer =====================
    {
      int sw = 0; uint32_t pw = 0;
      const unsigned char *wsrc = wprep + kHeaderBytes;
      for (int g = blockIdx.x; g < total_units; g += gridDim.x) {
        const int smp = g / units, u = g - smp * units;
        const uint32_t wm = window_mask(occ, geo, smp, u, lane, C::kUOut);
        for (int dx = 0; dx < 3; ++dx) {
          const uint32_t sm = (wm >> (3 * dx)) & 7u;
          if (sm == 0u) continue;
          for (int kc = 0; kc < nkc; ++kc) {
            for (int grp = 0; grp < C::kNG; ++grp) {
              // taps of a stage: TG = 9 -> the whole dx plane, 3 -> one dy row, 1 -> one tap (dy = grp / 3)
              const uint32_t need = TG == 9 ? sm : (TG == 3 ? (sm >> grp) & 1u : (sm >> (grp / 3)) & 1u);
              if (need == 0u) continue;
              if (lane == 0) {
                const int s = (dx * nkc + kc) * C::kNG + grp;
                bar_wait(w_empty + sw, pw ^ 1);
                bar_expect_tx(w_full + sw, C::kWStageBytes);
                tma_load(smem_u32(w_smem + sw * C::kWStageBytes), wsrc + (size_t)s * C::kWStageBytes, C::kWStageBytes, w_full + sw);
              }
              if (++sw == C::kWStages) { sw = 0; pw ^= 1; }
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer =====================
    // One thread issues every MMA of the CTA, so the instructions per MMA matter (a first version that rebuilt both
    // 64-bit descriptors from scratch spent ~87 clocks per MMA whatever its shape).  A descriptor's upper word is
    // constant; its lower word is (address >> 4) | (LBO >> 4) << 16, and since an activation row is 16 bytes a window
    // shift of n rows is simply +n on that word.  Taps and k-steps are unrolled: their offsets are immediates or one
    // of a few registers.
    // The whole warp runs the (uniform) loops so that addresses and descriptor words live in uniform registers; one
    // elected lane issues the MMAs and the commits (the per-lane version paid a register -> uniform-register move and
    // a convergence loop in front of every MMA).
    {
      constexpr uint32_t idesc = instr_desc(N);
      constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);           // SBO = 128 bytes, descriptor version 1
      int sa = 0, sw = 0; uint32_t pa = 0, pw = 0;
      const uint32_t a_lbo16 = (uint32_t)slab_rows, b_lbo16 = (uint32_t)N;      // LBO in 16-byte units
      const uint32_t kstep = 2u * a_lbo16;                               // rows (16-byte units) per K=16 step
      const uint32_t q1 = (uint32_t)geo.q, q2 = 2u * (uint32_t)geo.q;
      int it = 0;
      for (int g = blockIdx.x; g < total_units; g += gridDim.x, ++it) {
        const int buf = it & 1;
        const int smp = g / units, u = g - smp * units;
        const uint32_t wm = window_mask(occ, geo, smp, u, lane, C::kUOut);      // windows of all-zero rows are skipped, loads and MMAs
        bar_wait(t_empty + buf, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem + buf * 2 * C::kDCols;
        uint32_t acc = 0;                                             // 0 until the unit's first MMA has been issued
        for (int dx = 0; dx < 3; ++dx) {
          const uint32_t sm = (wm >> (3 * dx)) & 7u;
          if (sm == 0u) continue;
          for (int kc = 0; kc < nkc; ++kc) {
            bar_wait(a_full + sa, pa);
            tc_fence_after();
            const uint32_t a_lo = (smem_u32(a_smem + sa * a_stage_bytes) >> 4) | (a_lbo16 << 16);
#pragma unroll
            for (int grp = 0; grp < C::kNG; ++grp) {
              const uint32_t need = TG == 9 ? sm : (TG == 3 ? (sm >> grp) & 1u : (sm >> (grp / 3)) & 1u);
              if (need == 0u) continue;
              bar_wait(w_full + sw, pw);
              tc_fence_after();
              const uint32_t b_lo = (smem_u32(w_smem + sw * C::kWStageBytes) >> 4) | (b_lbo16 << 16);
              __syncwarp();
              if (elect_one()) {
                uint32_t first = acc;                                  // accumulate flag of the next tap's k = 0 MMAs
                if constexpr (PAIR) {
                  // stage = TG / 3 rows of taps (dx, dy), each [pair image: 2N rows = W(dz=0) | W(dz=+1)][single image: W(dz=-1)]
                  constexpr uint32_t idesc_pair = instr_desc(2 * N);
#pragma unroll
                  for (int sg = 0; sg < TG / 3; ++sg) {
                    const int dy = grp * (TG / 3) + sg;
                    if (((sm >> dy) & 1u) == 0u) continue;
                    const uint32_t a_row = a_lo + (dy == 0 ? 0u : (dy == 1 ? q1 : q2));
                    const uint32_t bg_lo = b_lo + (uint32_t)(sg * ((3 * C::kTapBytes) >> 4));
                    const uint32_t bp_lo = (bg_lo & 0xffffu) | ((uint32_t)(2 * N) << 16);        // LBO = 2N rows of 16 bytes
                    const uint32_t bs_lo = bg_lo + (uint32_t)((2 * C::kTapBytes) >> 4);
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
#pragma unroll
                      for (int k = 0; k < C::kK16; ++k) {       // dz = 0 (main) and dz = +1 (side): window at row offset +1
                        const uint32_t da_lo = a_row + 1u + (uint32_t)(t * 128) + (uint32_t)k * kstep;
                        const uint32_t db_lo = bp_lo + (uint32_t)(k * 2 * 2 * N);
                        const uint32_t accumulate = k == 0 ? first : 1u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                            "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                            ::"r"(d0 + t * C::kDCols), "r"(da_lo), "r"(desc_hi), "r"(db_lo), "r"(idesc_pair), "r"(accumulate) : "memory");
                      }
#pragma unroll
                      for (int k = 0; k < C::kK16; ++k) {       // dz = -1 into main: window at row offset 0
                        const uint32_t da_lo = a_row + (uint32_t)(t * 128) + (uint32_t)k * kstep;
                        const uint32_t db_lo = bs_lo + (uint32_t)(k * 2 * N);
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                            "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                            ::"r"(d0 + t * C::kDCols), "r"(da_lo), "r"(desc_hi), "r"(db_lo), "r"(idesc), "r"(1u) : "memory");
                      }
                    }
                    first = 1u;
                  }
                } else {
#pragma unroll
                for (int j = 0; j < TG; ++j) {
                  const int tap = grp * TG + j;
                  const int dy = tap / 3, dz = tap - dy * 3;
                  if (((sm >> dy) & 1u) == 0u) continue;               // (only TG = 9 stages mix dy rows)
                  const uint32_t a_tap = a_lo + (dy == 0 ? 0u : (dy == 1 ? q1 : q2)) + (uint32_t)dz;
#pragma unroll
                  for (int t = 0; t < 2; ++t) {
#pragma unroll
                    for (int k = 0; k < C::kK16; ++k) {
                      const uint32_t da_lo = a_tap + (uint32_t)(t * 128) + (uint32_t)k * kstep;
                      const uint32_t db_lo = b_lo + (uint32_t)(j * (C::kTapBytes >> 4) + k * 2 * N);
                      const uint32_t accumulate = k == 0 ? first : 1u;
                      asm volatile(
                          "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
                          "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %2};\n\t"
                          "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                          ::"r"(d0 + t * N), "r"(da_lo), "r"(desc_hi), "r"(db_lo), "r"(idesc), "r"(accumulate) : "memory");
                    }
                  }
                  first = 1u;
                }
                }
                umma_commit(w_empty + sw);
              }
              __syncwarp();
              acc = 1;                                                 // a needed stage always issues at least one tap
              if (++sw == C::kWStages) { sw = 0; pw ^= 1; }
            }
            if (elect_one()) umma_commit(a_empty + sa);
            __syncwarp();
            if (++sa == kAStages) { sa = 0; pa ^= 1; }
          }
        }
        if (elect_one()) {
          unit_empty[buf] = acc == 0u ? 1u : 0u;
          __threadfence_block();
          if (acc != 0u) umma_commit(t_full + buf);
          else bar_arrive(t_full + buf);                               // nothing was multiplied: the result is the bias
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int e = warp - 4, t = e >> 2, qd = e & 3;    // tile, TMEM lane quarter (== warp % 4)
    const float out_scale = __ldg(reinterpret_cast<const float *>(wprep));
    const int r = geo.r, q = geo.q;
    const size_t s3 = (size_t)r * r * r;
    constexpr int CG = N / 8;                           // channels per GroupNorm group (8 groups)
    int it = 0;
    for (int g = blockIdx.x; g < total_units; g += gridDim.x, ++it) {
      const int buf = it & 1;
      const int smp = g / units, u = g - smp * units;
      const int m = t * 128 + qd * 32 + lane;
      const int p = u * C::kUOut + m;
      const int z = p % q, xy = p / q, y = xy % q, x = xy / q;
      const bool valid = x < r && y < r && z < r && m < C::kUOut;
      float *dst = out + ((size_t)smp * s3 + ((size_t)x * r + y) * r + z) * N;
      float gs1[8], gs2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) gs1[i] = gs2[i] = 0.0f;
      bar_wait(t_full + buf, (it >> 1) & 1);
      tc_fence_after();
      const bool empty = unit_empty[buf] != 0u;       // all of the unit's windows were zero rows: accumulators were not written
      const uint32_t taddr = tmem + ((uint32_t)(qd * 32) << 16) + buf * 2 * C::kDCols + t * C::kDCols;
      if constexpr (PAIR) {
        // the side row of this warp's first lane is what the previous warp's last lane adds to its main row
        if (!empty) {
#pragma unroll
          for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t ss[32];
            BDM_CV3_TMEM_LD32(ss, taddr + N + c0);
            tmem_wait_ld();
            if (lane == 0) {
#pragma unroll
              for (int i4 = 0; i4 < 8; ++i4)
                *reinterpret_cast<uint4 *>(edge + e * N + c0 + 4 * i4) = make_uint4(ss[4 * i4], ss[4 * i4 + 1], ss[4 * i4 + 2], ss[4 * i4 + 3]);
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
      const float *edge_next = edge + (e < 7 ? e + 1 : 7) * N;
#pragma unroll
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t rr[32];
        if (!empty) {
          BDM_CV3_TMEM_LD32(rr, taddr + c0);
          if constexpr (PAIR) {
            uint32_t ss[32];
            BDM_CV3_TMEM_LD32(ss, taddr + N + c0);
            tmem_wait_ld();
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const float4 ev = *reinterpret_cast<const float4 *>(edge_next + c0 + 4 * i4);     // same address in every lane
              const float en[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float sd = __shfl_down_sync(0xffffffffu, __uint_as_float(ss[4 * i4 + i]), 1);
                sd = lane == 31 ? en[i] : sd;
                rr[4 * i4 + i] = __float_as_uint(__uint_as_float(rr[4 * i4 + i]) + sd);
              }
            }
          } else {
            tmem_wait_ld();
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) rr[i] = 0u;
        }
        float f[32];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 bb = bias != nullptr ? __ldg(reinterpret_cast<const float4 *>(bias + c0) + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
          f[4 * i4] = fmaf(__uint_as_float(rr[4 * i4]), out_scale, bb.x);
          f[4 * i4 + 1] = fmaf(__uint_as_float(rr[4 * i4 + 1]), out_scale, bb.y);
          f[4 * i4 + 2] = fmaf(__uint_as_float(rr[4 * i4 + 2]), out_scale, bb.z);
          f[4 * i4 + 3] = fmaf(__uint_as_float(rr[4 * i4 + 3]), out_scale, bb.w);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float fv = valid ? f[i] : 0.0f;
          gs1[(c0 + i) / CG] += fv;
          gs2[(c0 + i) / CG] = fmaf(fv, fv, gs2[(c0 + i) / CG]);
        }
        if (valid) {   // 256-bit stores: every lane writes whole 32-byte sectors of its row (128-bit stores at this 4N-byte
                       // lane stride make two partial-sector requests per sector; the store floor is 160 us at C=64, R=32)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + c0 + 8 * i), "f"(f[8 * i]),
                         "f"(f[8 * i + 1]), "f"(f[8 * i + 2]), "f"(f[8 * i + 3]), "f"(f[8 * i + 4]), "f"(f[8 * i + 5]),
                         "f"(f[8 * i + 6]), "f"(f[8 * i + 7]) : "memory");
        }
      }
      tc_fence_before();
      bar_arrive(t_empty + buf);
      if (PAIR && unit_stats == nullptr) asm volatile("bar.sync 1, 256;" ::: "memory");     // `edge` is rewritten by the next unit
      if (unit_stats != nullptr) {
        // 16 values (sum, sum of squares of 8 groups) over 32 lanes as a halving butterfly: each step a lane hands half of
        // what it still holds to its partner, so 8 + 4 + 2 + 1 + 1 = 16 shuffles instead of 16 x 5; fixed order
        float v8[8], v4[4], v2[2], v1;
        {
          const bool up = (lane & 16) != 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {    // value index k = 2 * group + (0: sum, 1: sum of squares); k < 8 <-> groups 0..3
            const float lo_v = (i & 1) ? gs2[i >> 1] : gs1[i >> 1], hi_v = (i & 1) ? gs2[4 + (i >> 1)] : gs1[4 + (i >> 1)];
            v8[i] = (up ? hi_v : lo_v) + __shfl_xor_sync(0xffffffffu, up ? lo_v : hi_v, 16);
          }
        }
        {
          const bool up = (lane & 8) != 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) v4[i] = (up ? v8[i + 4] : v8[i]) + __shfl_xor_sync(0xffffffffu, up ? v8[i] : v8[i + 4], 8);
        }
        {
          const bool up = (lane & 4) != 0;
#pragma unroll
          for (int i = 0; i < 2; ++i) v2[i] = (up ? v4[i + 2] : v4[i]) + __shfl_xor_sync(0xffffffffu, up ? v4[i] : v4[i + 2], 4);
        }
        {
          const bool up = (lane & 2) != 0;
          v1 = (up ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, up ? v2[0] : v2[1], 2);
        }
        v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
        float *rb = red + (buf * 8 + e) * 16;
        if ((lane & 1) == 0) rb[((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)] = v1;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (e == 0 && lane < 16) {
          double a = 0.0;
#pragma unroll
          for (int w = 0; w < 8; ++w) a += (double)red[(buf * 8 + w) * 16 + lane];
          unit_stats[(size_t)g * 16 + lane] = a;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::kTmemCols));
  }
}

// unit partials f64[b][units][8 groups][2] -> the GroupNorm kernels' producer-statistics format f64[b][1][c][2]:
// the group's sums sit in its first channel's slot, the other channels hold zeros (the consumer adds the
// channels of a group).  One warp per value, lanes stride the units, fixed shuffle tree: deterministic.
__global__ void __launch_bounds__(512)
conv3_stats_fold_kernel(int units, int c, const double *__restrict__ unit_stats, double2 *__restrict__ stats) {
  __shared__ double sums[16];
  const int b = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double a = 0.0;
  for (int u = lane; u < units; u += 32) a += unit_stats[((size_t)b * units + u) * 16 + w];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
  if (lane == 0) sums[w] = a;
  __syncthreads();
  const int cg = c / 8;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const int grp = ch / cg;
    stats[(size_t)b * c + ch] = ch % cg == 0 ? make_double2(sums[2 * grp], sums[2 * grp + 1]) : make_double2(0.0, 0.0);
  }
}

static inline bool supported(int c_in, int c_out, int r) {
  const bool pow2 = r >= 4 && r <= 32 && (r & (r - 1)) == 0;   // r = 64: the three activation slabs no longer fit 227 KB
  return pow2 && (c_out == 32 || c_out == 64 || c_out == 128) && (c_in == 32 || (c_in >= 64 && c_in % 64 == 0 && c_in <= 512));
}
static inline int kc_of(int c_in) { return c_in == 32 ? 32 : 64; }
static inline int tg_of(int c_in, int c_out) {      // taps per weight stage: keep a stage between 8 and 24 KB
  const int tap_bytes = c_out * kc_of(c_in) * 2;
  return tap_bytes <= 2048 ? 9 : (tap_bytes <= 8192 ? 3 : 1);
}

template <int N, int KC, int TG, bool PAIR = false>
static int launch_conv(int b, int c_in, const Geometry &geo, const __half *xh, const unsigned char *wprep, const float *bias,
                       float *out, double *unit_stats, const uint32_t *occ, cudaStream_t st) {
  using C = Cfg<N, KC, TG, PAIR>;
  const int a_stage_bytes = (int)align_up((size_t)C::kChunks * geo.slab_rows * 16, 128);
  const size_t smem_bytes = (size_t)kAStages * a_stage_bytes + (size_t)C::kWStages * C::kWStageBytes + C::kNumBars * 8 + 16 +
                            2 * 8 * 16 * sizeof(float) + (PAIR ? 8 * N * sizeof(float) : 0);
  if (smem_bytes > 227 * 1024) return BDM_ERR_BAD_SIZE;
  cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(conv3_tc05_kernel<N, KC, TG, PAIR>), smem_bytes);
  if (e != cudaSuccess) return (int)e;
  const int total_units = b * (PAIR ? geo.units_pair : geo.units);
  const int grid = total_units < sm_count() ? total_units : sm_count();
  conv3_tc05_kernel<N, KC, TG, PAIR><<<grid, kThreads, smem_bytes, st>>>(b, c_in, geo, a_stage_bytes, xh, wprep, bias, out, unit_stats, occ);
  BDM_RETURN_LAUNCH_STATUS();
}

// tap pairing (see Cfg): on for c_out <= 64 with 64-channel... any chunking; BDM_CONV3_PAIR=0 keeps one MMA per tap
static bool pair_mode(int c_in, int c_out) {
  static const bool on = [] {
    const char *e = std::getenv("BDM_CONV3_PAIR");
    return e == nullptr || e[0] != '0';
  }();
  (void)c_in;
  return on && c_out <= 64;
}

}  // namespace cv3
}  // namespace bdm

using namespace bdm;

extern "C" int bdm_conv3_tc05_supported(int c_in, int c_out, int r) { return cv3::supported(c_in, c_out, r) ? 1 : 0; }

/* rows of one fp16 chunk plane of the padded activation array for b samples at resolution r */
extern "C" long long bdm_conv3_tc05_plane_rows(int b, int r) {
  if (b <= 0 || r <= 0) return 0;
  return cv3::conv3_geometry(b, r).total_rows;
}
/* units (blocks of the per-unit group statistics) the convolution c_in -> c_out works in at resolution r */
extern "C" int bdm_conv3_tc05_units(int c_in, int c_out, int r) {
  if (r <= 0) return 0;
  const cv3::Geometry g = cv3::conv3_geometry(1, r);
  return cv3::pair_mode(c_in, c_out) ? g.units_pair : g.units;
}
extern "C" size_t bdm_conv3_tc05_weight_bytes(int c_in, int c_out) {
  if (c_in <= 0 || c_out <= 0) return 0;
  return cv3::kHeaderBytes + (size_t)27 * c_in * c_out * 2;
}

extern "C" int bdm_conv3_tc05_prepare(int c_in, int c_out, const float *weight, const float *gamma, const float *beta,
                                      long long group_elems, void *prepared, size_t prepared_bytes, bdm_stream_t stream) {
  BDM_CHECK_SIZE(c_in > 0 && c_out > 0 && group_elems > 0);
  BDM_CHECK_SIZE(c_in == 32 || c_in % 64 == 0);
  BDM_CHECK_PTR(weight); BDM_CHECK_PTR(prepared);
  if (prepared_bytes < bdm_conv3_tc05_weight_bytes(c_in, c_out)) return BDM_ERR_WORKSPACE_TOO_SMALL;
  if ((reinterpret_cast<uintptr_t>(prepared) & 255) != 0) return BDM_ERR_MISALIGNED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float *header = static_cast<float *>(prepared);
  cv3::conv3_prep_header_kernel<<<1, 1024, 0, st>>>((size_t)27 * c_in * c_out, weight, c_in, gamma, beta, (float)group_elems, header);
  const size_t total = (size_t)27 * c_in * c_out;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  cv3::conv3_prep_weights_kernel<<<blocks, 256, 0, st>>>(
      c_in, c_out, cv3::kc_of(c_in), cv3::pair_mode(c_in, c_out) ? 3 : cv3::tg_of(c_in, c_out), cv3::pair_mode(c_in, c_out) ? 1 : 0, weight, header,
      reinterpret_cast<__half *>(static_cast<unsigned char *>(prepared) + cv3::kHeaderBytes));
  BDM_RETURN_LAUNCH_STATUS();
}

extern "C" int bdm_groupnorm_swish_half_planar(int b, int c, int r, int groups, float eps, int swish, const float *x,
                                               const float *conv_bias, const float *gamma, const float *beta,
                                               const double *partials, int chunks, const void *prepared, void *xh,
                                               long long plane_rows, bdm_stream_t stream) {
  BDM_CHECK_SIZE(b >= 0 && b <= 65535 && c >= 8 && c % 8 == 0 && c <= 256 && 256 % c == 0 && groups >= 1 && groups <= 32 &&
                 c % groups == 0 && chunks != 0 && r >= 2 && r <= 64 && (r & (r - 1)) == 0);
  BDM_CHECK_SIZE(chunks > 0 || conv_bias == nullptr);      // chunks < 0: -chunks blocks of group partials f64[b][blocks][groups][2]
  if (b == 0) return BDM_OK;
  BDM_CHECK_PTR(x); BDM_CHECK_PTR(partials); BDM_CHECK_PTR(prepared); BDM_CHECK_PTR(xh);
  const cv3::Geometry geo = cv3::conv3_geometry(b, r);
  BDM_CHECK_SIZE(plane_rows == geo.total_rows);
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(xh)) & 15) != 0) return BDM_ERR_MISALIGNED;
  const long long s = (long long)r * r * r;
  const int rpp = cv3::kApplyThreads / (c / 8);
  long long want = ((long long)8 * sm_count() + b - 1) / b;
  long long maxt = (s + 4LL * rpp - 1) / (4LL * rpp);
  const int ntiles = (int)(want < maxt ? (want < 1 ? 1 : want) : (maxt < 1 ? 1 : maxt));
  if (c == 32 || c == 64 || c == 128) {   // warp-level kernel: no block barriers in the streaming loop
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const double2 *pp = reinterpret_cast<const double2 *>(partials);
    const float *hd = static_cast<const float *>(prepared);
    __half *xp = static_cast<__half *>(xh);
    long long wt = ((long long)6 * sm_count() + b - 1) / b;
    long long mt = (s + 255) / 256;
    const int nt = (int)(wt < mt ? (wt < 1 ? 1 : wt) : (mt < 1 ? 1 : mt));
    const size_t smem = (size_t)8 * (c / 8) * 33 * 16;
    cudaError_t e = cudaSuccess;
    if (c == 32) {
      cv3::gn_apply_half_planar_warp_kernel<8><<<dim3(nt, b), cv3::kApplyThreads, smem, st>>>(
          r, groups, chunks, nt, eps, swish, x, conv_bias, gamma, beta, pp, hd, xp, geo.guard, geo.sample_rows, geo.total_rows);
    } else if (c == 64) {
      cv3::gn_apply_half_planar_warp_kernel<16><<<dim3(nt, b), cv3::kApplyThreads, smem, st>>>(
          r, groups, chunks, nt, eps, swish, x, conv_bias, gamma, beta, pp, hd, xp, geo.guard, geo.sample_rows, geo.total_rows);
    } else {
      e = ensure_dynamic_smem(reinterpret_cast<const void *>(cv3::gn_apply_half_planar_warp_kernel<32>), smem);
      if (e != cudaSuccess) return (int)e;
      cv3::gn_apply_half_planar_warp_kernel<32><<<dim3(nt, b), cv3::kApplyThreads, smem, st>>>(
          r, groups, chunks, nt, eps, swish, x, conv_bias, gamma, beta, pp, hd, xp, geo.guard, geo.sample_rows, geo.total_rows);
    }
    BDM_RETURN_LAUNCH_STATUS();
  }
  cv3::gn_apply_half_planar_kernel<<<dim3(ntiles, b), cv3::kApplyThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      c, r, groups, chunks, ntiles, eps, swish, x, conv_bias, gamma, beta, reinterpret_cast<const double2 *>(partials),
      static_cast<const float *>(prepared), static_cast<__half *>(xh), geo.guard, geo.sample_rows, geo.total_rows);
  BDM_RETURN_LAUNCH_STATUS();
}

/* compact f32[b][c][n] (bdm_avg_voxelize_compact) + the voxel plan (bdm_voxel_plan's workspace for the same b, n, r) ->
 * xh; `prepared` (of the convolution that will read xh) receives the dynamic activation scale.  amax_ready != 0: word 4
 * of `prepared` already holds the bit pattern of max|compact| (bdm_avg_voxelize_compact_amax wrote it).
 * occ (or NULL): u32[b][bdm_conv3_tc05_occ_words(r)] receives one bit per non-zero row of every sample; handed to
 * bdm_conv3_tc05 it lets the convolution skip the windows that hold only zeros (a third of them at r = 32). */
extern "C" int bdm_conv3_tc05_occ_words(int r) {
  if (r <= 0) return 0;
  return cv3::occ_words_per_sample(cv3::conv3_geometry(1, r));
}
extern "C" int bdm_conv3_tc05_fill_planes(int b, int c, int n, int r, const float *compact, const void *plan_workspace,
                                          size_t plan_workspace_bytes, void *prepared, void *xh, long long plane_rows,
                                          int amax_ready, unsigned *occ, bdm_stream_t stream) {
  BDM_CHECK_SIZE(b >= 0 && b <= 65535 && c >= 8 && c % 8 == 0 && n >= 1 && r >= 4 && r <= 32 && (r & (r - 1)) == 0);
  BDM_CHECK_SIZE(vox_fast_path(n, r * r * r) && ((size_t)c * n) % 4 == 0);
  if (b == 0) return BDM_OK;
  BDM_CHECK_PTR(compact); BDM_CHECK_PTR(prepared); BDM_CHECK_PTR(xh);
  const cv3::Geometry geo = cv3::conv3_geometry(b, r);
  BDM_CHECK_SIZE(plane_rows == geo.total_rows);
  const VoxAuxLayout L = vox_aux_layout(n, r * r * r);
  int rc = check_workspace(L, b, plan_workspace, plan_workspace_bytes);
  if (rc != BDM_OK) return rc;
  if (((reinterpret_cast<uintptr_t>(compact) | reinterpret_cast<uintptr_t>(xh) | reinterpret_cast<uintptr_t>(prepared)) & 15) != 0)
    return BDM_ERR_MISALIGNED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float *header = static_cast<float *>(prepared);
  if (!amax_ready) {   // otherwise bdm_avg_voxelize_compact_amax already left max|average| in the header's word 4
    cudaMemsetAsync(header + 4, 0, 4, st);
    cv3::conv3_amax_kernel<<<2 * sm_count(), 256, 0, st>>>((size_t)b * c * n / 4, reinterpret_cast<const float4 *>(compact),
                                                         reinterpret_cast<unsigned *>(header) + 4);
  }
  const int occ_words = cv3::occ_words_per_sample(geo);
  if (occ != nullptr) cudaMemsetAsync(occ, 0, sizeof(unsigned) * (size_t)b * occ_words, st);
  cv3::conv3_fill_planes_kernel<<<dim3((L.nw + 7) / 8, b), 256, 0, st>>>(
      c, n, r, L, static_cast<const unsigned char *>(plan_workspace), compact, header, static_cast<__half *>(xh), geo.guard,
      geo.sample_rows, geo.total_rows, occ, occ_words);
  BDM_RETURN_LAUNCH_STATUS();
}

/* out f32[b][r^3][c_out] (channels last) = conv3x3x3(xh) + bias; stats (or NULL) f64[b][1][c_out][2] receives the
 * per-group (sum, sum of squares) of the result in the layout bdm_groupnorm_act_cl(precomputed_chunks = 1) takes
 * (8 groups); workspace: b * units * 16 doubles when stats != NULL. */
extern "C" size_t bdm_conv3_tc05_workspace_bytes(int b, int r) {
  if (b <= 0 || r <= 0) return 16;
  return (size_t)b * cv3::conv3_geometry(b, r).units_pair * 16 * sizeof(double);     // units_pair >= units
}
extern "C" int bdm_conv3_tc05(int b, int c_in, int c_out, int r, const void *xh, long long plane_rows, const void *prepared,
                              const float *bias, float *out, double *stats, void *workspace, size_t workspace_bytes,
                              const unsigned *occ, bdm_stream_t stream) {
  BDM_CHECK_SIZE(b >= 0 && cv3::supported(c_in, c_out, r));
  if (b == 0) return BDM_OK;
  BDM_CHECK_PTR(xh); BDM_CHECK_PTR(prepared); BDM_CHECK_PTR(out);
  const cv3::Geometry geo = cv3::conv3_geometry(b, r);
  BDM_CHECK_SIZE(plane_rows == geo.total_rows && (long long)b * geo.units_pair < 0x7fffffffLL);
  if (((reinterpret_cast<uintptr_t>(xh) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(prepared)) & 15) != 0)
    return BDM_ERR_MISALIGNED;
  double *unit_stats = nullptr;
  const bool fold = stats != nullptr && workspace != nullptr;
  if (fold) {
    if (workspace_bytes < bdm_conv3_tc05_workspace_bytes(b, r)) return BDM_ERR_WORKSPACE_TOO_SMALL;
    unit_stats = static_cast<double *>(workspace);
  } else if (stats != nullptr) {
    unit_stats = stats;          // the caller takes the per-unit group partials f64[b][units][8][2] as they are
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __half *x = static_cast<const __half *>(xh);
  const unsigned char *wp = static_cast<const unsigned char *>(prepared);
  int rc;
  const int kc = cv3::kc_of(c_in);
  const bool pair = cv3::pair_mode(c_in, c_out);
  const int units = pair ? geo.units_pair : geo.units;
  if (pair && c_out == 32 && kc == 32) rc = cv3::launch_conv<32, 32, 9, true>(b, c_in, geo, x, wp, bias, out, unit_stats, occ, st);
  else if (pair && c_out == 32) rc = cv3::launch_conv<32, 64, 3, true>(b, c_in, geo, x, wp, bias, out, unit_stats, occ, st);
  else if (pair && kc == 32) rc = cv3::launch_conv<64, 32, 3, true>(b, c_in, geo, x, wp, bias, out, unit_stats, occ, st);
  else if (pair) rc = cv3::launch_conv<64, 64, 3, true>(b, c_in, geo, x, wp, bias, out, unit_stats, occ, st);
  else if (c_out == 32 && kc == 32) rc = cv3::launch_conv<32, 32, 9>(b, c_in, geo, x, wp, bias, out, unit_stats, occ, st);
  else if (c_out == 32) rc = cv3::launch_conv<32, 64, 3>(b, c_in, geo, x, wp, bias, out, unit_stats, occ, st);
  else if (c_out == 64 && kc == 32) rc = cv3::launch_conv<64, 32, 3>(b, c_in, geo, x, wp, bias, out, unit_stats, occ, st);
  else if (c_out == 64) rc = cv3::launch_conv<64, 64, 3>(b, c_in, geo, x, wp, bias, out, unit_stats, occ, st);
  else if (kc == 32) rc = cv3::launch_conv<128, 32, 3>(b, c_in, geo, x, wp, bias, out, unit_stats, occ, st);
  else rc = cv3::launch_conv<128, 64, 1>(b, c_in, geo, x, wp, bias, out, unit_stats, occ, st);
  if (rc != BDM_OK) return rc;
  if (fold) {
    cv3::conv3_stats_fold_kernel<<<b, 512, 0, st>>>(units, c_out, unit_stats, reinterpret_cast<double2 *>(stats));
    BDM_RETURN_LAUNCH_STATUS();
  }
  return BDM_OK;
}
