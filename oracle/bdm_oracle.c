/*
 * bdm_oracle.c -- CPU restatement of the reference's point-voxel hot-path ops.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA path in bdm_b200/csrc; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path never routes through it and fails loudly if the CUDA library is missing.
 *
 * Every function restates, in plain scalar C, the algorithm of one reference CUDA kernel and cites
 * the reference file:line it follows (paths relative to
 * /root/reference/experiments/model/pvcnn/modules/functional/src/; the PVD copy under
 * experiments/pvd/modules/functional/src/ is byte-identical).
 *
 * Floating-point contract.  The reference is compiled by nvcc with the default -fmad=true, so
 * `a*a + b*b + c*c` contracts to  fma(c, c, fma(a, a, b*b))  (first product fused into the second,
 * which is rounded on its own; SASS: FMUL, FFMA, FFMA -- see DESIGN.md "fp contract").  This file is
 * compiled with -ffp-contract=off and spells every contraction out with fmaf(), so the integer
 * outputs that depend on those distances (ball query, FPS, 3-NN indices) are bit-exact
 * restatements, not approximations.
 *
 * Parity pin: tests/golden/ref_*.npz are outputs of the reference's own CUDA kernels (oracle/_ref,
 * built from /root/reference by oracle/build_ref.py) run on a B200 by tests/golden/make_golden.py;
 * tests/test_oracle_golden.py checks this file against them.
 *
 * Threading: OpenMP over the batch dimension (shapes are independent in every op).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__)
#define ORC_CLONES __attribute__((target_clones("fma", "default")))
#else
#define ORC_CLONES
#endif

/* squared distance with the reference's contraction: t = dy*dy; t = fma(dx,dx,t); fma(dz,dz,t) */
static inline float sqdist_ref(float dx, float dy, float dz) {
  float t = dy * dy;
  t = fmaf(dx, dx, t);
  return fmaf(dz, dz, t);
}

/* ------------------------------------------------------------------------------------------------
 * avg_voxelize forward   (voxelization/vox.cu:18-34 grid_stats_kernel, :48-72 avg_voxelize_kernel,
 *                         host wrapper voxelization/vox.cpp:17-43: ind/out/cnt are zero-initialised)
 *   ind[b,i]   = x*r*r + y*r + z
 *   cnt[b,v]   = #points in voxel v
 *   out[b,c,v] = sum_i fl(feat[b,c,i] * fl(1/cnt))          (atomic order unspecified in the reference;
 *                                                             here: ascending point index)
 * ---------------------------------------------------------------------------------------------- */
ORC_CLONES
void orc_avg_voxelize_forward(int b, int c, int n, int r, const int *coords, const float *feat,
                              int *ind, int *cnt, float *out) {
  const int r2 = r * r, r3 = r2 * r;
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const int *co = coords + (size_t)bi * 3 * n;
    int *in = ind + (size_t)bi * n;
    int *cn = cnt + (size_t)bi * r3;
    const float *f = feat + (size_t)bi * c * n;
    float *o = out + (size_t)bi * c * r3;
    memset(cn, 0, sizeof(int) * (size_t)r3);
    memset(o, 0, sizeof(float) * (size_t)c * r3);
    for (int i = 0; i < n; ++i) { /* vox.cu:28-33 */
      in[i] = co[i] * r2 + co[i + n] * r + co[i + n + n];
      cn[in[i]] += 1;
    }
    for (int j = 0; j < c; ++j) { /* vox.cu:60-71 */
      for (int i = 0; i < n; ++i) {
        const int pos = in[i];
        const int cur = cn[pos];
        if (cur > 0) {
          /* `1.0 / static_cast<float>(cur_cnt)` is a double division rounded to float; that equals
             the correctly rounded float division (53 >= 2*24+2), which is what nvcc emits. */
          const float inv = 1.0f / (float)cur;
          o[(size_t)j * r3 + pos] += f[(size_t)j * n + i] * inv;
        }
      }
    }
  }
}

/* avg_voxelize backward  (vox.cu:86-110):  grad_x[b,c,i] = grad_y[b,c,ind[i]] * fl(1/cnt) */
ORC_CLONES
void orc_avg_voxelize_backward(int b, int c, int n, int s, const int *ind, const int *cnt,
                               const float *grad_y, float *grad_x) {
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const int *in = ind + (size_t)bi * n;
    const int *cn = cnt + (size_t)bi * s;
    const float *gy = grad_y + (size_t)bi * c * s;
    float *gx = grad_x + (size_t)bi * c * n;
    for (int j = 0; j < c; ++j)
      for (int i = 0; i < n; ++i) {
        const int pos = in[i];
        const int cur = cn[pos];
        float v = 0.0f;
        if (cur > 0) v = 0.0f + gy[(size_t)j * s + pos] * (1.0f / (float)cur);
        gx[(size_t)j * n + i] = v;
      }
  }
}

/* ------------------------------------------------------------------------------------------------
 * trilinear_devoxelize forward  (interpolate/trilinear_devox.cu:21-105; wrapper .cpp:18-55)
 *   weights  (x_d*y_d)*z_d, left to right                            (:52-59)
 *   indices  idx000 + {z: +1 iff zd1>0, y: +r iff yd1>0, x: +r2 iff xd1>0}   (:61-75)
 *   out = w000*f000 + w001*f001 + ... + w111*f111, left-assoc, nvcc contraction:
 *         acc = w001*f001; acc = fma(w000,f000,acc); acc = fma(w010,f010,acc); ... 111  (:96-103)
 *   inds/wgts [b,8,n] written only when is_training (:77-94), corner order 000,001,...,111
 * ---------------------------------------------------------------------------------------------- */
ORC_CLONES
void orc_trilinear_devoxelize_forward(int b, int c, int n, int r, int is_training,
                                      const float *coords, const float *feat, int *inds,
                                      float *wgts, float *outs) {
  const int r2 = r * r, r3 = r2 * r;
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const float *co = coords + (size_t)bi * 3 * n;
    const float *f = feat + (size_t)bi * c * r3;
    float *o = outs + (size_t)bi * c * n;
    for (int i = 0; i < n; ++i) {
      const float x = co[i], y = co[i + n], z = co[i + n + n];
      const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
      const float xd1 = x - xl, yd1 = y - yl, zd1 = z - zl;
      const float xd0 = 1.0f - xd1, yd0 = 1.0f - yd1, zd0 = 1.0f - zd1;
      float w[8];
      w[0] = xd0 * yd0 * zd0; w[1] = xd0 * yd0 * zd1; w[2] = xd0 * yd1 * zd0; w[3] = xd0 * yd1 * zd1;
      w[4] = xd1 * yd0 * zd0; w[5] = xd1 * yd0 * zd1; w[6] = xd1 * yd1 * zd0; w[7] = xd1 * yd1 * zd1;
      const int xlo = (int)xl, ylo = (int)yl, zlo = (int)zl;
      const int xh = (xd1 > 0) ? -1 : 0, yh = (yd1 > 0) ? -1 : 0, zh = (zd1 > 0) ? 1 : 0;
      int id[8];
      id[0] = xlo * r2 + ylo * r + zlo;
      id[1] = id[0] + zh;
      id[2] = id[0] + (yh & r);
      id[3] = id[2] + zh;
      id[4] = id[0] + (xh & r2);
      id[5] = id[4] + zh;
      id[6] = id[4] + (yh & r);
      id[7] = id[6] + zh;
      if (is_training) {
        int *in = inds + (size_t)bi * 8 * n;
        float *wg = wgts + (size_t)bi * 8 * n;
        for (int k = 0; k < 8; ++k) { wg[i + (size_t)n * k] = w[k]; in[i + (size_t)n * k] = id[k]; }
      }
      for (int j = 0; j < c; ++j) {
        const float *fj = f + (size_t)j * r3;
        float acc = w[1] * fj[id[1]];
        acc = fmaf(w[0], fj[id[0]], acc);
        for (int k = 2; k < 8; ++k) acc = fmaf(w[k], fj[id[k]], acc);
        o[(size_t)j * n + i] = acc;
      }
    }
  }
}

/* trilinear_devoxelize backward (trilinear_devox.cu:119-162): 8 scatter-adds of fl(w*g) per (point,ch);
   atomic order unspecified in the reference; here ascending point index, corner order 000..111. */
ORC_CLONES
void orc_trilinear_devoxelize_backward(int b, int c, int n, int r3, const int *inds,
                                       const float *wgts, const float *grad_y, float *grad_x) {
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const int *in = inds + (size_t)bi * 8 * n;
    const float *wg = wgts + (size_t)bi * 8 * n;
    const float *gy = grad_y + (size_t)bi * c * n;
    float *gx = grad_x + (size_t)bi * c * r3;
    memset(gx, 0, sizeof(float) * (size_t)c * r3);
    for (int j = 0; j < c; ++j)
      for (int i = 0; i < n; ++i) {
        const float g = gy[(size_t)j * n + i];
        for (int k = 0; k < 8; ++k)
          gx[(size_t)j * r3 + in[i + (size_t)n * k]] += wg[i + (size_t)n * k] * g;
      }
  }
}

/* ------------------------------------------------------------------------------------------------
 * gather  (sampling/sampling.cu:17-31):  out[b,c,j] = feat[b,c,idx[b,j]]
 * ---------------------------------------------------------------------------------------------- */
void orc_gather_features_forward(int b, int c, int n, int m, const float *feat, const int *idx,
                                 float *out) {
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float *f = feat + ((size_t)bi * c + l) * n;
      float *o = out + ((size_t)bi * c + l) * m;
      const int *ix = idx + (size_t)bi * m;
      for (int j = 0; j < m; ++j) o[j] = f[ix[j]];
    }
}

/* gather backward (sampling.cu:52-66): grad_x[b,c,idx[b,j]] += grad_y[b,c,j]  (ascending j here) */
void orc_gather_features_backward(int b, int c, int n, int m, const float *grad_y, const int *idx,
                                  float *grad_x) {
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float *gy = grad_y + ((size_t)bi * c + l) * m;
      float *gx = grad_x + ((size_t)bi * c + l) * n;
      const int *ix = idx + (size_t)bi * m;
      memset(gx, 0, sizeof(float) * (size_t)n);
      for (int j = 0; j < m; ++j) gx[ix[j]] += gy[j];
    }
}

/* ------------------------------------------------------------------------------------------------
 * furthest point sampling  (sampling/sampling.cu:86-167; wrapper sampling.cpp:43-58)
 *   The kernel is launched with exactly 512 threads (sampling.cu:171) and its result depends on
 *   that: thread t scans k = t, t+512, ... keeping the FIRST strict maximum (:141-144); the
 *   9-level shared-memory tree keeps the LOWER slot on ties (`dists[i1] < dists[i2]`, :149-159).
 *   Both are simulated literally.  distances start at 1e38f (sampling.cpp:53-54); idx[0] = 0.
 * ---------------------------------------------------------------------------------------------- */
ORC_CLONES
void orc_furthest_point_sampling(int b, int n, int m, const float *coords, int *indices) {
  enum { BS = 512 };
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const float *co = coords + (size_t)bi * 3 * n;
    int *out = indices + (size_t)bi * m;
    for (int j = 0; j < m; ++j) out[j] = 0; /* torch::zeros */
    if (m <= 0) continue;
    float *dist = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int k = 0; k < n; ++k) dist[k] = 1e38f;
    float dists[BS];
    int dists_i[BS];
    int old = 0;
    out[0] = old;
    for (int j = 1; j < m; ++j) {
      const float x1 = co[old], y1 = co[old + n], z1 = co[old + n + n];
      for (int t = 0; t < BS; ++t) {
        int besti = 0;
        float best = -1.0f;
        for (int k = t; k < n; k += BS) {
          const float td = dist[k];
          const float d = sqdist_ref(co[k] - x1, co[k + n] - y1, co[k + n + n] - z1);
          const float d2 = (d < td) ? d : td; /* min(d, td) */
          if (d2 != td) dist[k] = d2;
          if (d2 > best) { best = d2; besti = k; }
        }
        dists[t] = best;
        dists_i[t] = besti;
      }
      for (int u = 0; (1 << u) < BS; ++u)
        for (int t = 0; t < (BS >> (u + 1)); ++t) {
          const int i1 = (t * 2) << u, i2 = (t * 2 + 1) << u;
          if (dists[i1] < dists[i2]) { dists[i1] = dists[i2]; dists_i[i1] = dists_i[i2]; }
        }
      old = dists_i[0];
      out[j] = old;
    }
    free(dist);
  }
}

/* ------------------------------------------------------------------------------------------------
 * ball query  (ball_query/ball_query.cu:19-50; wrapper ball_query.cpp:6-30: r2 = radius*radius in
 *   fp32 on the host, output zero-initialised).  d2 = (c - p) squared with the contraction above;
 *   hit iff d2 < r2; first hit fills the whole row, c-th hit overwrites slot c; stop at u hits.
 * ---------------------------------------------------------------------------------------------- */
ORC_CLONES
void orc_ball_query(int b, int n, int m, float r2, int u, const float *centers, const float *points,
                    int *neighbors) {
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const float *pc = points + (size_t)bi * 3 * n;
    const float *cc = centers + (size_t)bi * 3 * m;
    int *nb = neighbors + (size_t)bi * m * u;
    for (int j = 0; j < m; ++j) {
      int *row = nb + (size_t)j * u;
      for (int v = 0; v < u; ++v) row[v] = 0;
      const float cx = cc[j], cy = cc[j + m], cz = cc[j + m + m];
      for (int k = 0, cnt = 0; k < n && cnt < u; ++k) {
        const float d2 = sqdist_ref(cx - pc[k], cy - pc[k + n], cz - pc[k + n + n]);
        if (d2 < r2) {
          if (cnt == 0)
            for (int v = 0; v < u; ++v) row[v] = k;
          row[cnt] = k;
          ++cnt;
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * grouping  (grouping/grouping.cu:18-36):  out[b,c,m,u] = feat[b,c,idx[b,m,u]]
 * ---------------------------------------------------------------------------------------------- */
void orc_grouping_forward(int b, int c, int n, int m, int u, const float *feat, const int *idx,
                          float *out) {
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const int *ix = idx + (size_t)bi * m * u;
    for (int l = 0; l < c; ++l) {
      const float *f = feat + ((size_t)bi * c + l) * n;
      float *o = out + ((size_t)bi * c + l) * m * u;
      for (int q = 0; q < m * u; ++q) o[q] = f[ix[q]];
    }
  }
}

/* grouping backward (grouping.cu:58-77): grad_x[b,c,idx[b,m,u]] += grad_y[b,c,m,u] (ascending here) */
void orc_grouping_backward(int b, int c, int n, int m, int u, const float *grad_y, const int *idx,
                           float *grad_x) {
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const int *ix = idx + (size_t)bi * m * u;
    for (int l = 0; l < c; ++l) {
      const float *gy = grad_y + ((size_t)bi * c + l) * m * u;
      float *gx = grad_x + ((size_t)bi * c + l) * n;
      memset(gx, 0, sizeof(float) * (size_t)n);
      for (int q = 0; q < m * u; ++q) gx[ix[q]] += gy[q];
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * three nearest neighbours + interpolation
 *   search  (interpolate/neighbor_interpolate.cu:20-75): d = (u - x) squared, contraction as above;
 *           bests are doubles initialised to 1e40 with index 0; strict-< insertion 2 -> 1 -> 0
 *           (:44-59); clamp to [1e-10f, 1e10f] in double (:61-63); pair products in double rounded to
 *           float (:64-66); inv = 1.0f / ((d0d1 + d0d2) + d1d2) (:67); weights (:68-72).
 *   interp  (:90-116): out = f[i1]*w1 + f[i2]*w2 + f[i3]*w3 -> fma(f3,w3, fma(f1,w1, f2*w2)).
 * ---------------------------------------------------------------------------------------------- */
ORC_CLONES
void orc_three_nn(int b, int n, int m, const float *points, const float *centers, float *weights,
                  int *indices) {
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const float *pc = points + (size_t)bi * 3 * n;
    const float *cc = centers + (size_t)bi * 3 * m;
    float *w = weights + (size_t)bi * 3 * n;
    int *ix = indices + (size_t)bi * 3 * n;
    for (int j = 0; j < n; ++j) {
      const float ux = pc[j], uy = pc[j + n], uz = pc[j + n + n];
      double best0 = 1e40, best1 = 1e40, best2 = 1e40;
      int i0 = 0, i1 = 0, i2 = 0;
      for (int k = 0; k < m; ++k) {
        const float d = sqdist_ref(ux - cc[k], uy - cc[k + m], uz - cc[k + m + m]);
        if (d < best2) {
          best2 = d; i2 = k;
          if (d < best1) {
            best2 = best1; i2 = i1; best1 = d; i1 = k;
            if (d < best0) { best1 = best0; i1 = i0; best0 = d; i0 = k; }
          }
        }
      }
      const double lo = (double)1e-10f, hi = (double)1e10f;
      best0 = fmax(fmin(hi, best0), lo);
      best1 = fmax(fmin(hi, best1), lo);
      best2 = fmax(fmin(hi, best2), lo);
      const float d0d1 = (float)(best0 * best1);
      const float d0d2 = (float)(best0 * best2);
      const float d1d2 = (float)(best1 * best2);
      const float inv = 1.0f / (d0d1 + d0d2 + d1d2);
      w[j] = d1d2 * inv;          ix[j] = i0;
      w[j + n] = d0d2 * inv;      ix[j + n] = i1;
      w[j + n + n] = d0d1 * inv;  ix[j + n + n] = i2;
    }
  }
}

ORC_CLONES
void orc_three_interpolate(int b, int c, int m, int n, const float *feat, const int *indices,
                           const float *weights, float *out) {
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const float *w = weights + (size_t)bi * 3 * n;
    const int *ix = indices + (size_t)bi * 3 * n;
    for (int l = 0; l < c; ++l) {
      const float *f = feat + ((size_t)bi * c + l) * m;
      float *o = out + ((size_t)bi * c + l) * n;
      for (int j = 0; j < n; ++j) {
        float acc = f[ix[j + n]] * w[j + n];
        acc = fmaf(f[ix[j]], w[j], acc);
        o[j] = fmaf(f[ix[j + n + n]], w[j + n + n], acc);
      }
    }
  }
}

/* three_nn interpolate backward (neighbor_interpolate.cu:145-170): 3 scatter-adds of fl(g*w) */
ORC_CLONES
void orc_three_interpolate_backward(int b, int c, int n, int m, const float *grad_y,
                                    const int *indices, const float *weights, float *grad_x) {
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const float *w = weights + (size_t)bi * 3 * n;
    const int *ix = indices + (size_t)bi * 3 * n;
    for (int l = 0; l < c; ++l) {
      const float *gy = grad_y + ((size_t)bi * c + l) * n;
      float *gx = grad_x + ((size_t)bi * c + l) * m;
      memset(gx, 0, sizeof(float) * (size_t)m);
      for (int j = 0; j < n; ++j) {
        gx[ix[j]] += gy[j] * w[j];
        gx[ix[j + n]] += gy[j] * w[j + n];
        gx[ix[j + n + n]] += gy[j] * w[j + n + n];
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Projection conditioning  (experiments/model/projection_model.py:127-157 surface_projection).
 * The arithmetic lives in pytorch3d (un-vendored, version unpinned; README.md:73 installs a 0.7.x
 * wheel): PointsRasterizer(radius=0.0075, points_per_pixel=1, bin_size=0) == naive rasteriser.
 * PARITY UNPINNED for this function: it restates pytorch3d's published algorithm
 * (rasterize_points naive: pixel centre NDC = -1 + (2*(S-1-i)+1)/S with +X left / +Y up, skip z<0,
 * hit iff dx^2+dy^2 < radius^2, K=1 keeps the smallest z, earlier point wins z ties) and the
 * PerspectiveCameras NDC projection  view = X R + T,  ndc = f * view.xy / view.z + pp.
 *
 *   points [b,n,3]; R [b,3,3] row-vector convention; T [b,3]; focal [b,2]; pp [b,2]
 *   zbuf_idx [b,H,W]  = winning point index (per-sample, 0..n-1) or -1
 *   out [b,n,C]       = feature vector of a pixel the point wins (lowest pixel index when it wins
 *                       several; the reference's order there is unspecified), zeros otherwise.
 *   feat [b,C,H,W]
 * ---------------------------------------------------------------------------------------------- */
void orc_project_points(int b, int n, const float *points, const float *R, const float *T,
                        const float *focal, const float *pp, float *ndc /* [b,n,3] */) {
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const float *r = R + (size_t)bi * 9, *t = T + (size_t)bi * 3;
    const float fx = focal[bi * 2], fy = focal[bi * 2 + 1], px = pp[bi * 2], py = pp[bi * 2 + 1];
    for (int i = 0; i < n; ++i) {
      const float *p = points + ((size_t)bi * n + i) * 3;
      float v[3];
      for (int k = 0; k < 3; ++k) {
        float acc = p[0] * r[0 * 3 + k];
        acc = fmaf(p[1], r[1 * 3 + k], acc);
        acc = fmaf(p[2], r[2 * 3 + k], acc);
        v[k] = acc + t[k];
      }
      float *o = ndc + ((size_t)bi * n + i) * 3;
      o[0] = fmaf(fx, v[0] / v[2], px);
      o[1] = fmaf(fy, v[1] / v[2], py);
      o[2] = v[2];
    }
  }
}

void orc_rasterize_points(int b, int n, int H, int W, float radius, const float *ndc,
                          int *zbuf_idx /* [b,H,W] */) {
  const float r2 = radius * radius;
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    int *zi = zbuf_idx + (size_t)bi * H * W;
    float *zd = (float *)malloc(sizeof(float) * (size_t)H * W);
    for (int q = 0; q < H * W; ++q) { zi[q] = -1; zd[q] = 0.0f; }
    for (int i = 0; i < n; ++i) { /* ascending i + strict '<' on z  =>  earlier point wins ties */
      const float *p = ndc + ((size_t)bi * n + i) * 3;
      const float px = p[0], py = p[1], pz = p[2];
      if (!(pz >= 0.0f)) continue;
      /* candidate window (pure cull, generous margin); the exact strict test is below */
      const int ky = (int)ceilf(radius * (float)H * 0.5f) + 2, kx = (int)ceilf(radius * (float)W * 0.5f) + 2;
      const float ycf = (float)(H - 1) - ((py + 1.0f) * (float)H - 1.0f) * 0.5f;
      const float xcf = (float)(W - 1) - ((px + 1.0f) * (float)W - 1.0f) * 0.5f;
      if (!(ycf > -1e6f && ycf < 1e6f && xcf > -1e6f && xcf < 1e6f)) continue;
      int y0 = (int)floorf(ycf) - ky, y1 = (int)floorf(ycf) + ky + 1;
      int x0 = (int)floorf(xcf) - kx, x1 = (int)floorf(xcf) + kx + 1;
      if (y0 < 0) y0 = 0;
      if (x0 < 0) x0 = 0;
      if (y1 > H - 1) y1 = H - 1;
      if (x1 > W - 1) x1 = W - 1;
      for (int yi = y0; yi <= y1; ++yi) {
        const float yf = -1.0f + (2.0f * (float)(H - 1 - yi) + 1.0f) / (float)H;
        const float dy = yf - py;
        for (int xi = x0; xi <= x1; ++xi) {
          const float xf = -1.0f + (2.0f * (float)(W - 1 - xi) + 1.0f) / (float)W;
          const float dx = xf - px;
          const float d2 = fmaf(dx, dx, dy * dy); /* dx*dx + dy*dy as nvcc contracts it */
          if (d2 < r2) {
            const int q = yi * W + xi;
            if (zi[q] < 0 || pz < zd[q]) { zi[q] = i; zd[q] = pz; }
          }
        }
      }
    }
    free(zd);
  }
}

void orc_splat_features(int b, int n, int C, int H, int W, const int *zbuf_idx, const float *feat,
                        float *out /* [b,n,C] */) {
#pragma omp parallel for schedule(dynamic)
  for (int bi = 0; bi < b; ++bi) {
    const int *zi = zbuf_idx + (size_t)bi * H * W;
    float *o = out + (size_t)bi * n * C;
    memset(o, 0, sizeof(float) * (size_t)n * C);
    char *done = (char *)calloc((size_t)(n > 0 ? n : 1), 1);
    for (int q = 0; q < H * W; ++q) {
      const int i = zi[q];
      if (i < 0 || done[i]) continue;
      done[i] = 1;
      for (int ch = 0; ch < C; ++ch) o[(size_t)i * C + ch] = feat[((size_t)bi * C + ch) * H * W + q];
    }
    free(done);
  }
}

/* ------------------------------------------------------------------------------------------------
 * Evaluation nearest neighbour (fp64)
 *   direct form  (pytorch3d.loss.chamfer_distance via knn_points, called at
 *                 experiments/evaluation/evaluation_cd.py:125; PARITY UNPINNED, published algorithm:
 *                 cham = mean_i min_j |x_i-y_j|^2 + mean_j min_i |.|^2, lowest index on ties)
 *   expansion form (experiments/evaluation/evaluation_f1.py:90-98: -2ab + |a|^2 + |b|^2, clamp 1e-12)
 * ---------------------------------------------------------------------------------------------- */
void orc_nn_direct_f64(int b, int n, int m, const double *src /*[b,n,3]*/,
                       const double *tgt /*[b,m,3]*/, double *dist /*[b,n]*/, int *idx /*[b,n]*/) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int i = 0; i < n; ++i) {
      const double *s = src + ((size_t)bi * n + i) * 3;
      const double *t = tgt + (size_t)bi * m * 3;
      double best = INFINITY;
      int besti = 0;
      for (int j = 0; j < m; ++j) {
        const double dx = s[0] - t[j * 3], dy = s[1] - t[j * 3 + 1], dz = s[2] - t[j * 3 + 2];
        const double d = dx * dx + dy * dy + dz * dz;
        if (d < best) { best = d; besti = j; }
      }
      dist[(size_t)bi * n + i] = best;
      idx[(size_t)bi * n + i] = besti;
    }
}

void orc_nn_expanded_f64(int b, int n, int m, const double *src, const double *tgt, double *dist) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int bi = 0; bi < b; ++bi)
    for (int i = 0; i < n; ++i) {
      const double *s = src + ((size_t)bi * n + i) * 3;
      const double *t = tgt + (size_t)bi * m * 3;
      const double ss = s[0] * s[0] + s[1] * s[1] + s[2] * s[2];
      double best = INFINITY;
      for (int j = 0; j < m; ++j) {
        const double *q = t + j * 3;
        const double ab = s[0] * q[0] + s[1] * q[1] + s[2] * q[2];
        const double tt = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
        double d = -2.0 * ab;
        d += ss;
        d += tt;
        if (d < 1e-12) d = 1e-12;
        if (d < best) best = d;
      }
      dist[(size_t)bi * n + i] = best;
    }
}
